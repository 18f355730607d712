#!/usr/bin/env python
"""Read-only / write-only / copy HBM bandwidth on this GPU with plain torch ops (context for write-dominated kernels)."""
import torch
dev = torch.device("cuda:0")
n = 1 << 30     # 4 GiB of int32
a = torch.empty(n, dtype=torch.int32, device=dev)
b = torch.empty(n, dtype=torch.int32, device=dev)
def t(fn, reps=10):
    for _ in range(3): fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
ms = t(lambda: a.fill_(7)); print(f"write-only (fill_ 4 GiB): {ms:.3f} ms  {n * 4 / ms / 1e6:.0f} GB/s")
ms = t(lambda: b.copy_(a)); print(f"copy (4 GiB -> 4 GiB):    {ms:.3f} ms  {n * 8 / ms / 1e6:.0f} GB/s (read+write)")
ms = t(lambda: a.sum());    print(f"read-only (sum 4 GiB):    {ms:.3f} ms  {n * 4 / ms / 1e6:.0f} GB/s")
