#!/usr/bin/env python
"""intra32 predictor throughput per mode class: 1 Mi predictions of ONE mode each, the mode-interleaved bench input (i % 35), and the
same multiset of modes sorted -- where the 0.75 of the HBM roofline goes."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import x266_b200 as xb

torch.cuda.set_device(0)
n = 1 << 20
g = torch.Generator(device="cuda"); g.manual_seed(1)
refs = torch.randint(0, 256, (n, 129), device="cuda", generator=g, dtype=torch.uint8)
pred = torch.empty((n, 1024), device="cuda", dtype=torch.uint8)
st = torch.cuda.current_stream().cuda_stream
peak = 6459.3


def timed(modes, reps=10):
    for _ in range(2):
        xb.xIntra32PredDev(refs.data_ptr(), modes.data_ptr(), pred.data_ptr(), n, st)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        xb.xIntra32PredDev(refs.data_ptr(), modes.data_ptr(), pred.data_ptr(), n, st)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def show(label, modes):
    ms = timed(modes)
    print(f"{label:34s} {ms:.4f} ms  {n / ms / 1e6:6.2f} G pred/s  {n * 1154 / ms / 1e6 / peak:.3f} of HBM", flush=True)


inter = (torch.arange(n, device="cuda") % 35).to(torch.uint8)
show("mode = i % 35 (bench input)", inter)
show("same modes, sorted", torch.sort(inter)[0].contiguous())
for m, name in ((0, "planar"), (1, "DC"), (2, "copy diag +32 hor"), (6, "frac hor +"), (10, "pure horizontal"), (14, "frac hor -"),
                (18, "copy diag -32"), (22, "frac ver -"), (26, "pure vertical"), (30, "frac ver +"), (34, "copy diag +32 ver")):
    show(f"all mode {m:2d} ({name})", torch.full((n,), m, device="cuda", dtype=torch.uint8))
for v in sys.argv[1:]:
    k, val = v.split("=")
    xb.tune(int(k), int(val))
    show(f"tune {k}={val}: i % 35", inter)
    xb.tune(int(k), 0)
# mode-major entry: all 35 modes of 29960 blocks = 1 048 600 predictions
nb = n // 35 + 1
big = torch.empty((nb, 35, 1024), device="cuda", dtype=torch.uint8)
full = (1 << 35) - 1
for label, mask in (("all 35 modes", full), ("fractional modes only", full & ~((1 << 0) | (1 << 1) | (1 << 2) | (1 << 10) | (1 << 18) | (1 << 26) | (1 << 34))),
                    ("mode 30 only", 1 << 30)):
    nm = bin(mask).count("1")
    for _ in range(2):
        xb.xIntra32PredModesDev(refs.data_ptr(), nb, mask, big.data_ptr(), st)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        xb.xIntra32PredModesDev(refs.data_ptr(), nb, mask, big.data_ptr(), st)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    byts = nb * (129 + nm * 1024)
    print(f"xIntra32PredModes, {label:22s} {ms:.4f} ms  {nb * nm / ms / 1e6:6.2f} G pred/s  {byts / ms / 1e6 / peak:.3f} of HBM", flush=True)
