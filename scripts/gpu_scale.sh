#!/bin/bash
# 1/2/4/8-GPU sweep on one box (under gpurun --gpus 8).
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/gpus_scale.txt
timeout 300 python bench.py --impl reference --gpus 1 --steps 3 --warmup 1 > gpurun_out/scale_ref.log 2>&1
timeout 300 python bench.py --gpus 1 --steps 20 --warmup 3 --no-secondary > gpurun_out/scale_n1.log 2> gpurun_out/scale_n1.err
for N in 2 4 8; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2953$N \
      bench.py --gpus $N --steps 20 --warmup 3 --no-secondary > gpurun_out/scale_n$N.log 2> gpurun_out/scale_n$N.err
done
python - <<'PY'
import json
base=None
for n in (1,2,4,8):
    try:
        d=json.loads([l for l in open(f'gpurun_out/scale_n{n}.log') if l.startswith('{')][-1])
    except Exception as e:
        print(n, 'failed', e); continue
    base = base or d['value']
    print(f"N={n}: {d['value']/1e9:.3f} G blocks/s  x{d['value']/base:.2f}  ms/step={d['ms_per_step']:.3f}  frac={d['roofline']['frac']:.3f}  e2e={d['e2e']['value']/1e6:.1f} M blocks/s")
print(open('gpurun_out/scale_ref.log').read()[:600])
PY
