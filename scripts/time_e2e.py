#!/usr/bin/env python
"""e2e (host pointers, pinned) throughput of xDct32Batch vs pipeline chunk size."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import x266_b200 as xb
torch.cuda.set_device(0)
n = (int(sys.argv[1]) if len(sys.argv) > 1 else 8) * 32400
hin = torch.randint(-1023, 1024, (n, 32, 32), dtype=torch.int16).pin_memory()
hout = torch.empty_like(hin).pin_memory()
a, b = hin.numpy(), hout.numpy()
for chunk in ((8192, 16384, 32768, 65536, 131072) if len(sys.argv) > 1 else (1024, 2048, 4096, 8192, 16384, 32768, 65536)):
    xb.tune(4, chunk)
    for _ in range(2):
        xb.xDct32Batch(a, 6, 11, out=b)
    t = time.perf_counter()
    for _ in range(5):
        xb.xDct32Batch(a, 6, 11, out=b)
    dt = (time.perf_counter() - t) / 5
    print(f"chunk {chunk:6d} blocks: {n / dt / 1e6:6.2f} M blocks/s  ({n * 2048 / dt / 1e9:5.1f} GB/s each way)", flush=True)
xb.tune(4, 0)
# pure copies for reference
d = torch.empty((n, 32, 32), dtype=torch.int16, device="cuda")
torch.cuda.synchronize(); t = time.perf_counter()
for _ in range(5): d.copy_(hin, non_blocking=True)
torch.cuda.synchronize(); dt = (time.perf_counter() - t) / 5
print(f"H2D alone: {n * 2048 / dt / 1e9:5.1f} GB/s")
t = time.perf_counter()
for _ in range(5): hout.copy_(d, non_blocking=True)
torch.cuda.synchronize(); dt = (time.perf_counter() - t) / 5
print(f"D2H alone: {n * 2048 / dt / 1e9:5.1f} GB/s")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
d2 = torch.empty_like(d)
torch.cuda.synchronize(); t = time.perf_counter()
for _ in range(5):
    with torch.cuda.stream(s1): d.copy_(hin, non_blocking=True)
    with torch.cuda.stream(s2): hout.copy_(d2, non_blocking=True)
torch.cuda.synchronize(); dt = (time.perf_counter() - t) / 5
print(f"H2D + D2H concurrently: {n * 2048 / dt / 1e9:5.1f} GB/s each way")
