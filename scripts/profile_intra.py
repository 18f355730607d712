#!/usr/bin/env python
"""ncu driver: a few launches of intra32_kernel on resident data."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import x266_b200 as xb
dev = torch.device("cuda:0")
npred = 1 << 19
refs = torch.randint(0, 256, (npred, 129), device=dev, dtype=torch.uint8)
modes = (torch.arange(npred, device=dev) % 35).to(torch.uint8)
pred = torch.empty((npred, 1024), device=dev, dtype=torch.uint8)
st = torch.cuda.current_stream().cuda_stream
for _ in range(3):
    xb.xIntra32PredDev(refs.data_ptr(), modes.data_ptr(), pred.data_ptr(), npred, st)
torch.cuda.synchronize()
print("done")
