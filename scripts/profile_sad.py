#!/usr/bin/env python
"""ncu driver: SAD full search on config 3 with the u32 surface, the u16 surface and no surface (one launch each, in that order)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import x266_b200 as xb
dev = torch.device("cuda:0")
w, h, rg = 1920, 1080, 32
cur = torch.randint(0, 256, (h, w), device=dev, dtype=torch.uint8)
refp = torch.randint(0, 256, (h + 64, w + 64), device=dev, dtype=torch.uint8)
nb = 240 * 135
cost = torch.empty((nb, 65, 65), device=dev, dtype=torch.int32)
cost16 = torch.empty((nb, 65, 65), device=dev, dtype=torch.int16)
best = torch.empty((nb, 3), device=dev, dtype=torch.int32)
st = torch.cuda.current_stream().cuda_stream
xb.xSad8x8SearchDev(cur.data_ptr(), refp.data_ptr(), w + 64, w, h, rg, 0, nb, cost.data_ptr(), best.data_ptr(), st)
xb.xSad8x8SearchU16Dev(cur.data_ptr(), refp.data_ptr(), w + 64, w, h, rg, 0, nb, cost16.data_ptr(), best.data_ptr(), st)
xb.xSad8x8SearchDev(cur.data_ptr(), refp.data_ptr(), w + 64, w, h, rg, 0, nb, 0, best.data_ptr(), st)
torch.cuda.synchronize()
print("done")
