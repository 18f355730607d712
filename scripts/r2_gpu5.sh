#!/bin/bash
mkdir -p gpurun_out
: > gpurun_out/r2_intra_ab.log
for v in old ta tb; do
  echo "== $v" >> gpurun_out/r2_intra_ab.log
  X266_B200_LIB=$PWD/tools/libx266_$v.so timeout 300 python scripts/time_intra_modes.py 2>&1 | grep "i % 35\|mode  0\|mode  2\|mode  6\|mode 22\|mode 26\|mode 30" | head -7 >> gpurun_out/r2_intra_ab.log
done
python - <<'PY' >> gpurun_out/r2_intra_ab.log
import torch
x = torch.empty(1 << 30, dtype=torch.uint8, device="cuda")
for fn, name, b in ((lambda: x.fill_(3), "fill (write only)", 1), (lambda: x[: 1 << 29].copy_(x[1 << 29:]), "copy (read + write)", 1)):
    for _ in range(3): fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): fn()
    e1.record(); torch.cuda.synchronize()
    print(f"{name}: {(1 << 30) * 10 / e0.elapsed_time(e1) / 1e6:.0f} GB/s")
PY
cat gpurun_out/r2_intra_ab.log
