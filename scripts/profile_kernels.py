#!/usr/bin/env python
"""Small driver for ncu: a few launches of each kernel on resident data (used by scripts/gpu_profile.sh)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import x266_b200 as xb

dev = torch.device("cuda:0")
torch.cuda.set_device(0)
frames = int(sys.argv[1]) if len(sys.argv) > 1 else 16
n = frames * 32400
src = (torch.randint(0, 1024, (n, 32, 32), device=dev, dtype=torch.int16) - torch.randint(0, 1024, (n, 32, 32), device=dev, dtype=torch.int16))
dst = torch.empty_like(src)
st = torch.cuda.current_stream().cuda_stream
for v in (xb.DCT_IMMA, xb.DCT_BFLY):
    xb.set_dct_variant(v)
    for _ in range(2):
        xb.xDct32BatchDev(src.data_ptr(), dst.data_ptr(), n, 6, 11, st)
for log2n, (s1, s2) in ((2, (1, 8)), (3, (2, 9)), (4, (3, 10))):
    xb.xDctNBatchDev(log2n, src.data_ptr(), dst.data_ptr(), (n * 1024) >> (2 * log2n), s1, s2, st)
npred = 1 << 18
refs = torch.randint(0, 256, (npred, 129), device=dev, dtype=torch.uint8)
modes = (torch.arange(npred, device=dev) % 35).to(torch.uint8)
pred = torch.empty((npred, 1024), device=dev, dtype=torch.uint8)
xb.xIntra32PredDev(refs.data_ptr(), modes.data_ptr(), pred.data_ptr(), npred, st)
nc = 1 << 22
d = torch.randint(-255, 256, (nc, 64), device=dev, dtype=torch.int16)
o = torch.empty(nc, device=dev, dtype=torch.int32)
for v in (0, 1):
    xb.tune(2, v)
    for _ in range(2):
        xb.xSatd8x8BatchDev(d.data_ptr(), o.data_ptr(), nc, st)
xb.tune(2, 0)
w, h, rg = 1920, 1080, 32
cur = torch.randint(0, 256, (h, w), device=dev, dtype=torch.uint8)
refp = torch.randint(0, 256, (h + 64, w + 64), device=dev, dtype=torch.uint8)
nb = 240 * 135
cost = torch.empty((nb, 65, 65), device=dev, dtype=torch.int32)
best = torch.empty((nb, 3), device=dev, dtype=torch.int32)
for _ in range(2):
    xb.xSatd8x8SearchDev(cur.data_ptr(), refp.data_ptr(), w + 64, w, h, rg, 0, nb, cost.data_ptr(), best.data_ptr(), st)
xb.tune(1, 1)
xb.xSatd8x8SearchDev(cur.data_ptr(), refp.data_ptr(), w + 64, w, h, rg, 0, nb, cost.data_ptr(), best.data_ptr(), st)
xb.tune(1, 0)
for _ in range(2):
    xb.xSad8x8SearchDev(cur.data_ptr(), refp.data_ptr(), w + 64, w, h, rg, 0, nb, cost.data_ptr(), best.data_ptr(), st)
# 16-bit cost surfaces (the u16 instantiations) and the argmin-only form of the SAD search
xb.xSatd8x8SearchU16Dev(cur.data_ptr(), refp.data_ptr(), w + 64, w, h, rg, 0, nb, cost.data_ptr(), best.data_ptr(), st)
xb.xSad8x8SearchU16Dev(cur.data_ptr(), refp.data_ptr(), w + 64, w, h, rg, 0, nb, cost.data_ptr(), best.data_ptr(), st)
xb.xSad8x8SearchDev(cur.data_ptr(), refp.data_ptr(), w + 64, w, h, rg, 0, nb, 0, best.data_ptr(), st)
# round 2: inverse, fused residual + DCT32 on tiled frames, mode-major predictor, decision, the closed block loop and its Recon leg, quantiser stub
xb.xIdct32BatchDev(src.data_ptr(), dst.data_ptr(), n, 7, 12, st)
fw, fh = 7680, 4320 * 2 // 32 * 32
nt = (fw // 16) * (fh // 16)
ct = torch.randint(0, 256, (nt * 512,), device=dev, dtype=torch.uint8)
pt = torch.randint(0, 256, (nt * 512,), device=dev, dtype=torch.uint8)
xb.xFrameResiDct32Dev(ct.data_ptr(), pt.data_ptr(), fw, fh, dst.data_ptr(), 4, 11, st)
xb.xIntra32PredModesDev(refs.data_ptr(), npred // 35, (1 << 35) - 1, pred.data_ptr(), st)
nblk = 32400
ecur = torch.randint(0, 256, (nblk, 1024), device=dev, dtype=torch.uint8)
ecost = torch.empty((nblk, 35), device=dev, dtype=torch.int32)
ebest = torch.empty(nblk, device=dev, dtype=torch.int32)
elev = torch.empty((nblk, 1024), device=dev, dtype=torch.int16)
erec = torch.empty((nblk, 1024), device=dev, dtype=torch.uint8)
xb.xIntra32DecideDev(ecur.data_ptr(), refs.data_ptr(), ecost.data_ptr(), ebest.data_ptr(), nblk, st)
xb.xIntra32EncodeBlockDev(ecur.data_ptr(), refs.data_ptr(), nblk, 27, elev.data_ptr(), erec.data_ptr(), ebest.data_ptr(), 0, st)
xb.xIntra32ReconDev(ecur.data_ptr(), refs.data_ptr(), ebest.to(torch.uint8).data_ptr(), nblk, 27, elev.data_ptr(), erec.data_ptr(), st)
xb.xQuantDequantDev(src.data_ptr(), dst.data_ptr(), 0, n * 1024, 27, st)
torch.cuda.synchronize()
print("profile driver done")
