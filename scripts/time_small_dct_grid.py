#!/usr/bin/env python
"""Persistent-grid size (CTAs per SM) sweep of the DCT8 / DCT4 kernels on 1 Gi samples (tune keys 10 / 11)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, x266_b200 as xb
dev = torch.device("cuda:0"); ns = 1 << 30
src = torch.randint(-255, 256, (ns,), device=dev, dtype=torch.int16); dst = torch.empty_like(src)
st = torch.cuda.current_stream().cuda_stream
for key, log2n, sh, vals in ((10, 3, (2, 9), (2, 3, 4, 6)), (11, 2, (1, 8), (4, 6, 8, 10, 12, 16))):
    for v in vals:
        xb.tune(key, v)
        for _ in range(3): xb.xDctNBatchDev(log2n, src.data_ptr(), dst.data_ptr(), ns >> (2 * log2n), sh[0], sh[1], st)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10): xb.xDctNBatchDev(log2n, src.data_ptr(), dst.data_ptr(), ns >> (2 * log2n), sh[0], sh[1], st)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        print(f"dct{1 << log2n} ctas/SM {v:2d}: {ms:.3f} ms  {ns * 4 / ms / 1e6:.0f} GB/s", flush=True)
    xb.tune(key, 0)
