#!/usr/bin/env python
"""SATD full search v3, config 3: accumulate / add forms (xGpuTune key 6): 1 = IDP.2A per word (shipped in round 1), 3 = the same with the
transform's plain sums issued as IMAD on the FMA pipe, 0 / 2 = packed-add forms."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import x266_b200 as xb
dev = torch.device("cuda:0")
w, h, rg = 1920, 1080, 32
g = torch.Generator(device=dev); g.manual_seed(3)
cur = torch.randint(0, 256, (h, w), device=dev, generator=g, dtype=torch.uint8)
refp = torch.randint(0, 256, (h + 64, w + 64), device=dev, generator=g, dtype=torch.uint8)
nb = 240 * 135
cost = torch.empty((nb, 65, 65), device=dev, dtype=torch.int32)
best = torch.empty((nb, 3), device=dev, dtype=torch.int32)
st = torch.cuda.current_stream().cuda_stream
ref_out = None
for form in (1, 3, 1, 3, 0, 2):
    xb.tune(6, form)
    for _ in range(3):
        xb.xSatd8x8SearchDev(cur.data_ptr(), refp.data_ptr(), w + 64, w, h, rg, 0, nb, cost.data_ptr(), best.data_ptr(), st)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        xb.xSatd8x8SearchDev(cur.data_ptr(), refp.data_ptr(), w + 64, w, h, rg, 0, nb, cost.data_ptr(), best.data_ptr(), st)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    same = True if ref_out is None else bool(torch.equal(cost, ref_out[0]) and torch.equal(best, ref_out[1]))
    if ref_out is None:
        ref_out = (cost.clone(), best.clone())
    print(f"form {form}: {ms:.4f} ms/frame  {nb * 4225 / ms / 1e6:.1f} G cand/s  identical to form 1: {same}", flush=True)
xb.tune(6, 1)
