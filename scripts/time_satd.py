#!/usr/bin/env python
"""Times the two SATD batch kernels (tensor-core default vs CUDA-core) on 16.8M candidates."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import x266_b200 as xb
dev = torch.device("cuda:0")
n = 1 << 24
st = torch.cuda.current_stream().cuda_stream
for rng, name in ((256, "9-bit"), (32768, "full int16")):
    d = torch.randint(-rng + 1 if rng == 256 else -rng, rng, (n, 64), device=dev, dtype=torch.int16)
    o = {v: torch.empty(n, device=dev, dtype=torch.int32) for v in (0, 1, 5)}
    for v in (0, 1, 5):
        xb.tune(2, v)
        for _ in range(3):
            xb.xSatd8x8BatchDev(d.data_ptr(), o[v].data_ptr(), n, st)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            xb.xSatd8x8BatchDev(d.data_ptr(), o[v].data_ptr(), n, st)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        print(f"satd batch { {0: 'ring, 3 stages', 1: 'cuda-core', 5: 'ring, 4 stages'}[v]:14s} {name:10s}: {ms:.3f} ms  {n/ms/1e6:.2f} G cand/s  {n*132/ms/1e6:.0f} GB/s  {n*132/ms/1e6/6459.3*100:.1f}% of measured HBM", flush=True)
    print("  all equal:", all(torch.equal(o[0], o[k]) for k in (1, 5)))
xb.tune(2, 0)
