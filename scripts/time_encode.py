#!/usr/bin/env python
"""Timing of the closed intra block loop (csrc/encode.cu) on one 8K frame of blocks (32400): the fused kernel, its two halves
(decide alone, recon channel alone) and the quantiser stub.  CUDA events on the launch stream."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import x266_b200 as xb

torch.cuda.set_device(0)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 32400
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
g = torch.Generator(device="cuda"); g.manual_seed(1)
cur = torch.randint(0, 256, (n, 1024), device="cuda", generator=g, dtype=torch.uint8)
refs = torch.randint(0, 256, (n, 129), device="cuda", generator=g, dtype=torch.uint8)
level = torch.empty((n, 1024), device="cuda", dtype=torch.int16)
recon = torch.empty((n, 1024), device="cuda", dtype=torch.uint8)
best = torch.empty(n, device="cuda", dtype=torch.int32)
cost = torch.empty((n, 35), device="cuda", dtype=torch.int32)
st = torch.cuda.current_stream().cuda_stream


def timed(fn, reps=reps):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


peak = 6459.3
ms = timed(lambda: xb.xIntra32EncodeBlockDev(cur.data_ptr(), refs.data_ptr(), n, 27, level.data_ptr(), recon.data_ptr(), best.data_ptr(), cost.data_ptr(), st))
bytes_blk = 1024 + 129 + 2048 + 1024 + 4 + 140
print(f"fused encode (decide+recon), {n} blocks: {ms:.3f} ms = {n / ms / 1e3:.2f} M blocks/s, {n * bytes_blk / ms / 1e6:.0f} GB/s algorithmic = {n * bytes_blk / ms / 1e6 / peak:.3f} of HBM")
ms = timed(lambda: xb.xIntra32EncodeBlockDev(cur.data_ptr(), refs.data_ptr(), n, 27, level.data_ptr(), recon.data_ptr(), best.data_ptr(), 0, st))
print(f"fused encode without the cost table:    {ms:.3f} ms = {n / ms / 1e3:.2f} M blocks/s")
ms_d = timed(lambda: xb.xIntra32DecideDev(cur.data_ptr(), refs.data_ptr(), cost.data_ptr(), best.data_ptr(), n, st))
print(f"decide alone (xIntra32DecideDev):       {ms_d:.3f} ms = {n / ms_d / 1e3:.2f} M blocks/s")
modes = best.to(torch.uint8)
ms_r = timed(lambda: xb.xIntra32ReconDev(cur.data_ptr(), refs.data_ptr(), modes.data_ptr(), n, 27, level.data_ptr(), recon.data_ptr(), st))
rb = 1024 + 130 + 2048 + 1024
print(f"recon channel alone (xIntra32ReconDev): {ms_r:.3f} ms = {n / ms_r / 1e3:.2f} M blocks/s, {n * rb / ms_r / 1e6:.0f} GB/s = {n * rb / ms_r / 1e6 / peak:.3f} of HBM")
print(f"decide + recon as two launches:         {ms_d + ms_r:.3f} ms")
nc = 1 << 28
c = torch.randint(-2000, 2000, (nc,), device="cuda", generator=g, dtype=torch.int16)
lv = torch.empty_like(c); dq = torch.empty_like(c)
ms = timed(lambda: xb.xQuantDequantDev(c.data_ptr(), lv.data_ptr(), dq.data_ptr(), nc, 27, st))
print(f"quant+dequant stub, {nc} coefficients:  {ms:.3f} ms = {nc * 6 / ms / 1e6:.0f} GB/s = {nc * 6 / ms / 1e6 / peak:.3f} of HBM")
