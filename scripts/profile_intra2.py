#!/usr/bin/env python
"""ncu target: the mode-major intra kernel on fractional modes only, and the generic kernel on one fractional mode."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import x266_b200 as xb
torch.cuda.set_device(0)
n = 1 << 18
g = torch.Generator(device="cuda"); g.manual_seed(1)
refs = torch.randint(0, 256, (n, 129), device="cuda", generator=g, dtype=torch.uint8)
pred = torch.empty((n, 1024), device="cuda", dtype=torch.uint8)
full = (1 << 35) - 1
frac = full & ~((1 << 0) | (1 << 1) | (1 << 2) | (1 << 10) | (1 << 18) | (1 << 26) | (1 << 34))
nb = n // 28
for _ in range(2):
    xb.xIntra32PredModesDev(refs.data_ptr(), nb, frac, pred.data_ptr(), 0)
    xb.xIntra32PredDev(refs.data_ptr(), torch.full((n,), 30, device="cuda", dtype=torch.uint8).data_ptr(), pred.data_ptr(), n, 0)
torch.cuda.synchronize()
