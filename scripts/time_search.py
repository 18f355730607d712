#!/usr/bin/env python
"""Times the full-search kernels on config 3 (1920x1080, +-32): tune(1, v) v=0 v3, 1 v1 (one CTA per block, any range)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import x266_b200 as xb
dev = torch.device("cuda:0")
w, h, rg = 1920, 1080, 32
cur = torch.randint(0, 256, (h, w), device=dev, dtype=torch.uint8)
refp = torch.randint(0, 256, (h + 64, w + 64), device=dev, dtype=torch.uint8)
nb = 240 * 135
cost = torch.empty((nb, 65, 65), device=dev, dtype=torch.int32)
best = torch.empty((nb, 3), device=dev, dtype=torch.int32)
st = torch.cuda.current_stream().cuda_stream
outs = {}
for v1 in (1, 0):
    xb.tune(1, v1)
    for with_cost in (True, False):
        c = cost.data_ptr() if with_cost else 0
        for _ in range(2):
            xb.xSatd8x8SearchDev(cur.data_ptr(), refp.data_ptr(), w + 64, w, h, rg, 0, nb, c, best.data_ptr(), st)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            xb.xSatd8x8SearchDev(cur.data_ptr(), refp.data_ptr(), w + 64, w, h, rg, 0, nb, c, best.data_ptr(), st)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
        print(f"search {('v3','v1')[v1]} cost_surface={with_cost}: {ms:.3f} ms/frame  {nb*4225/ms/1e6:.1f} G cand/s", flush=True)
        if with_cost: outs[v1] = (cost.clone(), best.clone())
print("v1 == v3:", torch.equal(outs[0][0], outs[1][0]), torch.equal(outs[0][1], outs[1][1]))
xb.tune(1, 0)
# plain SAD full search (N4): tune(7, 1) = first-generation kernel, 0 = position-tile kernel
for sv1, with_cost in ((1, True), (0, True), (0, False)):
    xb.tune(7, sv1)
    c = cost.data_ptr() if with_cost else 0
    for _ in range(2):
        xb.xSad8x8SearchDev(cur.data_ptr(), refp.data_ptr(), w + 64, w, h, rg, 0, nb, c, best.data_ptr(), st)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        xb.xSad8x8SearchDev(cur.data_ptr(), refp.data_ptr(), w + 64, w, h, rg, 0, nb, c, best.data_ptr(), st)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    print(f"SAD search {('v2','v1')[sv1]} cost_surface={with_cost}: {ms:.3f} ms/frame  {nb*4225/ms/1e6:.1f} G cand/s", flush=True)
    if with_cost: outs[('sad', sv1)] = (cost.clone(), best.clone())
print("SAD v1 == v2:", torch.equal(outs[('sad', 0)][0], outs[('sad', 1)][0]), torch.equal(outs[('sad', 0)][1], outs[('sad', 1)][1]))
xb.tune(7, 0)
