#!/usr/bin/env python
"""Single-launch latency of the DCT32 kernels versus batch size (config 2 = 2040 blocks)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import x266_b200 as xb
dev = torch.device("cuda:0")
src = torch.randint(-1023, 1024, (1 << 16, 32, 32), device=dev, dtype=torch.int16)
dst = torch.empty_like(src)
st = torch.cuda.current_stream().cuda_stream
L = xb.lib()
for variant, name in ((xb.DCT_IMMA, "imma"), (xb.DCT_BFLY, "bfly")):
    xb.set_dct_variant(variant)
    for n in (8, 296, 2040, 8160, 32400):
        for _ in range(20):
            L.xDct32BatchDev(src.data_ptr(), dst.data_ptr(), n, 6, 11, st)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t = time.perf_counter()
        e0.record()
        for _ in range(200):
            L.xDct32BatchDev(src.data_ptr(), dst.data_ptr(), n, 6, 11, st)
        e1.record()
        host = (time.perf_counter() - t) / 200 * 1e6
        torch.cuda.synchronize()
        print(f"{name} n={n:6d}: {e0.elapsed_time(e1) / 200 * 1e3:7.2f} us per launch (host issue {host:5.2f} us/call)", flush=True)
xb.set_dct_variant(xb.DCT_AUTO)
# the other entry points at a small size: anything far above the ~5 us host issue time has a prologue problem
def lat(name, fn):
    for _ in range(20): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(200): fn()
    e1.record(); torch.cuda.synchronize()
    print(f"{name}: {e0.elapsed_time(e1) / 200 * 1e3:7.2f} us per launch", flush=True)
d = torch.randint(-255, 256, (4096, 64), device=dev, dtype=torch.int16)
o = torch.empty(4096, device=dev, dtype=torch.int32)
lat("satd batch n=4096", lambda: L.xSatd8x8BatchDev(d.data_ptr(), o.data_ptr(), 4096, st))
for log2n in (4, 3, 2):
    lat(f"dct{1 << log2n} 64Ki samples", lambda: L.xDctNBatchDev(log2n, src.data_ptr(), dst.data_ptr(), 65536 >> (2 * log2n), log2n - 1, log2n + 6, st))
lat("idct32 n=64", lambda: L.xIdct32BatchDev(src.data_ptr(), dst.data_ptr(), 64, 7, 12, st))
refs = torch.randint(0, 256, (256, 129), device=dev, dtype=torch.uint8)
modes = (torch.arange(256, device=dev) % 35).to(torch.uint8)
pred = torch.empty((256, 1024), device=dev, dtype=torch.uint8)
lat("intra32 n=256", lambda: L.xIntra32PredDev(refs.data_ptr(), modes.data_ptr(), pred.data_ptr(), 256, st))
curb = torch.randint(0, 256, (64, 1024), device=dev, dtype=torch.uint8)
cost = torch.empty((64, 35), device=dev, dtype=torch.int32); bm = torch.empty(64, device=dev, dtype=torch.int32)
lat("intra decide n=64", lambda: L.xIntra32DecideDev(curb.data_ptr(), refs.data_ptr(), cost.data_ptr(), bm.data_ptr(), 64, st))
