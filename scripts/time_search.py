#!/usr/bin/env python
"""Times the full-search kernels on config 3 (1920x1080, +-32).
tune(1, v): 0 = v3 (position tiles), 1 = v1 (one CTA per block, any range); tune(7, v): SAD search generation; u16 = the ...U16Dev entry points (16-bit cost surface)."""
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
cost16 = torch.empty((nb, 65, 65), device=dev, dtype=torch.int16)
best = torch.empty((nb, 3), device=dev, dtype=torch.int32)
st = torch.cuda.current_stream().cuda_stream
REPS = 10


def timed(fn, c):
    for _ in range(2):
        fn(cur.data_ptr(), refp.data_ptr(), w + 64, w, h, rg, 0, nb, c, best.data_ptr(), st)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(REPS):
        fn(cur.data_ptr(), refp.data_ptr(), w + 64, w, h, rg, 0, nb, c, best.data_ptr(), st)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / REPS


outs = {}
for v1, var in ((1, 0), (0, 0)):
    xb.tune(1, v1)
    for with_cost in (True, False):
        ms = timed(xb.xSatd8x8SearchDev, cost.data_ptr() if with_cost else 0)
        print(f"search {('v3','v1')[v1]} VAR={var} cost_surface={with_cost}: {ms:.3f} ms/frame  {nb*4225/ms/1e6:.1f} G cand/s", flush=True)
        if with_cost: outs[(v1, var)] = (cost.clone(), best.clone())
    if v1 == 0:
        ms = timed(xb.xSatd8x8SearchU16Dev, cost16.data_ptr())
        print(f"search v3 VAR={var} u16 cost surface: {ms:.3f} ms/frame  {nb*4225/ms/1e6:.1f} G cand/s", flush=True)
        print("  u16 == u32:", torch.equal(cost16.to(torch.int32), outs[(0, var)][0]), torch.equal(best, outs[(0, var)][1]))
for var in (0,):
    print(f"v1 == v3 VAR={var}:", torch.equal(outs[(0, var)][0], outs[(1, 0)][0]), torch.equal(outs[(0, var)][1], outs[(1, 0)][1]))
xb.tune(1, 0)
# plain SAD full search (N4): tune(7, 1) = first-generation kernel, 0 = position-tile kernel
for sv1, with_cost in ((1, True), (0, True), (0, False)):
    xb.tune(7, sv1)
    ms = timed(xb.xSad8x8SearchDev, cost.data_ptr() if with_cost else 0)
    print(f"SAD search {('v2','v1')[sv1]} cost_surface={with_cost}: {ms:.3f} ms/frame  {nb*4225/ms/1e6:.1f} G cand/s", flush=True)
    if with_cost: outs[('sad', sv1)] = (cost.clone(), best.clone())
print("SAD v1 == v2:", torch.equal(outs[('sad', 0)][0], outs[('sad', 1)][0]), torch.equal(outs[('sad', 0)][1], outs[('sad', 1)][1]))
xb.tune(7, 0)
ms = timed(xb.xSad8x8SearchU16Dev, cost16.data_ptr())
print(f"SAD search v2 u16 cost surface: {ms:.3f} ms/frame  {nb*4225/ms/1e6:.1f} G cand/s", flush=True)
print("  u16 == u32:", torch.equal(cost16.to(torch.int32), outs[('sad', 0)][0]), torch.equal(best, outs[('sad', 0)][1]))
