#!/usr/bin/env python
"""Small invocation of every kernel (all variants) for compute-sanitizer (memcheck / racecheck / initcheck)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import x266_b200 as xb
from oracle import Oracle
o = Oracle()
ok = True
def check(name, cond):
    global ok
    print(("ok   " if cond else "FAIL ") + name, flush=True)
    ok = ok and bool(cond)
x = o.residual(301 * 1024, 1, 2)
want = o.dct(x.reshape(-1, 32, 32), 5, 4, 11, threads=4).ravel()
for v, name in ((xb.DCT_BFLY, "bfly"), (xb.DCT_IMMA, "imma direct")):
    xb.set_dct_variant(v)
    check("dct32 " + name, np.array_equal(xb.xDct32Batch(x, 4, 11), want))
xb.set_dct_variant(xb.DCT_IMMA)
for cfg in (0, 3, 7, 12):
    xb.tune(0, cfg)
    check(f"dct32 imma cfg {cfg}", np.array_equal(xb.xDct32Batch(x, 4, 11), want))
xb.tune(0, -1); xb.set_dct_variant(xb.DCT_AUTO)
for log2n, sh in ((2, (1, 8)), (3, (2, 9)), (4, (3, 10))):
    n = 1 << log2n
    xs = o.residual(1237 * n * n, 2, 2)
    for cc in (0, 1):
        xb.tune(3, cc)
        check(f"dct{n} cuda_core={cc}", np.array_equal(xb.xDctNBatch(log2n, xs, *sh), o.dct(xs.reshape(-1, n, n), log2n, *sh).ravel()))
xb.tune(3, 0)
check("partialButterfly32", np.array_equal(xb.partialButterfly32(x[:77 * 32], 4, 77), o.partial(x[:77 * 32], 4, 77)))
d = o.residual(1003 * 64, 3, 2)
for v in (0, 1, 5):
    xb.tune(2, v)
    check(f"satd batch variant={v}", np.array_equal(xb.xSatd8x8Batch(d), o.satd(d)))
xb.tune(2, 0)
rng = np.random.default_rng(0)
for R, (w, h) in ((3, (40, 24)), (8, (200, 24)), (32, (200, 16))):
    cur = rng.integers(0, 256, (h, w)).astype(np.uint8)
    refp = rng.integers(0, 256, (h + 2 * R, w + 2 * R)).astype(np.uint8)
    wc, wb = o.satd_search(cur, refp, R, 0, (w // 8) * (h // 8))
    for v1 in (0, 1):
        xb.tune(1, v1)
        c, b = xb.xSatd8x8Search(cur, refp, R)
        check(f"search R={R} mode={v1}", np.array_equal(c, wc) and np.array_equal(b, wb))
    xb.tune(1, 0)
    for form in (0, 2, 1):                       # accumulate forms of the v3 kernel (1 = shipped)
        xb.tune(6, form)
        c, b = xb.xSatd8x8Search(cur, refp, R, 3, 13)
        check(f"search v3 R={R} form={form} sub-range", np.array_equal(c, wc[3:13]) and np.array_equal(b, wb[3:13]))
xb.tune(1, 0)
refs = rng.integers(0, 256, (70, 129)).astype(np.uint8)
modes = np.tile(np.arange(35, dtype=np.uint8), 2)
for swar in (0, 1):                              # 0 = tensor-core angular path (shipped), 1 = CUDA-core SWAR interpolation
    xb.tune(8, swar)
    pred = xb.xIntra32Pred(refs, modes)
    check(f"intra32 swar={swar}", all(np.array_equal(pred[i], o.intra32(refs[i, :64], refs[i, 64:], int(modes[i]))) for i in range(70)))
xb.tune(8, 0)
w, h = 96, 64
fr = lambda: o.conv_input_fmt(rng.integers(0, 256, (h, w)).astype(np.uint8), rng.integers(0, 256, (h // 2, w // 2)).astype(np.uint8), rng.integers(0, 256, (h // 2, w // 2)).astype(np.uint8))
a, b = fr(), fr()
check("frame resi dct32", np.array_equal(xb.xFrameResiDct32(a, b, w, h, 4, 11), o.frame_resi_dct32(a, b, w, h, 4, 11)))
ci = o.residual(9 * 1024, 5, 2).reshape(-1, 32, 32)
check("idct32", np.array_equal(xb.xIdct32Batch(ci, 7, 12), o.idct(ci, 5, 7, 12)))
cb = rng.integers(0, 256, (6, 32, 32)).astype(np.uint8)
rb = rng.integers(0, 256, (6, 129)).astype(np.uint8)
cst, bst = xb.xIntra32Decide(cb, rb)
check("intra decide", all(np.array_equal(cst[i], o.intra32_decide(cb[i], rb[i, :64], rb[i, 64:])[0]) for i in range(6)))
# round 2: mode-major predictor, fused block loop, Recon channel, quantiser stub, tiled search, host converters, pageable staging
pm = xb.xIntra32PredModes(rb[:3])
check("intra32 mode-major", all(np.array_equal(pm[bi, m], o.intra32(rb[bi, :64], rb[bi, 64:], m)) for bi in range(3) for m in range(35)))
lv, rc, bm, cs = xb.xIntra32EncodeBlock(cb.reshape(6, 1024), rb, 27)
good = True
for i in range(6):
    wl, wr, wcost, wbm = o.intra32_encode(cb[i], rb[i, :64], rb[i, 64:], 27)
    good &= bool(bm[i] == wbm and np.array_equal(lv[i], wl) and np.array_equal(rc[i], wr) and np.array_equal(cs[i], wcost))
check("intra32 encode (fused block loop)", good)
lv, rc = xb.xIntra32Recon(cb.reshape(6, 1024), rb, np.array([0, 1, 2, 12, 22, 33], np.uint8), 20)
good = True
for i, m in enumerate((0, 1, 2, 12, 22, 33)):
    wl, wr = o.intra32_recon(cb[i], rb[i, :64], rb[i, 64:], m, 20)
    good &= bool(np.array_equal(lv[i], wl) and np.array_equal(rc[i], wr))
check("intra32 recon", good)
tw, th, tr = 48, 32, 8
tc_, tf_ = rng.integers(0, 256, (th, tw)).astype(np.uint8), rng.integers(0, 256, (th, tw)).astype(np.uint8)
zc = np.zeros((th // 2, tw // 2), np.uint8)
gc, gb = xb.xSatd8x8SearchTiled(o.conv_input_fmt(tc_, zc, zc), o.conv_input_fmt(tf_, zc, zc), tw, th, tr)
wc, wb = o.satd_search(tc_, np.pad(tf_, tr, mode="edge"), tr, 0, (tw // 8) * (th // 8))
check("tiled search", np.array_equal(gc, wc) and np.array_equal(gb, wb))
Y, U, V = rng.integers(0, 256, (32, 48)).astype(np.uint8), rng.integers(0, 256, (16, 24)).astype(np.uint8), rng.integers(0, 256, (16, 24)).astype(np.uint8)
ht = np.zeros(3 * 2 * 512, np.uint8)
xb.xConvInputFmt(ht, Y, U, V, 48, 48, 32)
hY, hU, hV = np.zeros_like(Y), np.zeros_like(U), np.zeros_like(V)
xb.xConvOutput420(ht, hY, 48, hU, hV, 24, 48, 32)
check("host converters round trip", np.array_equal(hY, Y) and np.array_equal(hU, U) and np.array_equal(hV, V))
xb.tune(4, 64)                                     # many small chunks through the pageable staging ring
xs2 = o.residual(700 * 1024, 9, 1)
check("pageable staging ring", np.array_equal(xb.xDct32Batch(xs2, 6, 11), o.dct(xs2.reshape(-1, 32, 32), 5, 6, 11, threads=4).ravel()))
xb.tune(4, 0)
tt = rng.integers(0, 256, (11, 32, 32)).astype(np.uint8)
check("transpose32", np.array_equal(xb.xTranspose32x32Batch(tt), tt.transpose(0, 2, 1)))
sa, sb2 = rng.integers(0, 256, (65, 65)).astype(np.uint8), rng.integers(0, 256, (65, 65)).astype(np.uint8)
check("sad region", xb.sad(sa, sb2) == o.sad(sa, sb2))
scur = rng.integers(0, 256, (24, 40)).astype(np.uint8); sref = rng.integers(0, 256, (24 + 16, 40 + 16)).astype(np.uint8)
wc, wb = o.sad_search(scur, sref, 8, 0, 15)
for v1 in (0, 1):
    xb.tune(7, v1)
    c, b = xb.xSad8x8Search(scur, sref, 8)
    check(f"sad search v1={v1}", np.array_equal(c, wc) and np.array_equal(b, wb))
xb.tune(7, 0)
scur = rng.integers(0, 256, (16, 200)).astype(np.uint8); sref = rng.integers(0, 256, (16 + 64, 200 + 64)).astype(np.uint8)
c, b = xb.xSad8x8Search(scur, sref, 32)
wc, wb = o.sad_search(scur, sref, 32, 0, 50)
check("sad search v2 R=32", np.array_equal(c, wc) and np.array_equal(b, wb))
# 16-bit cost surfaces (both searches, R = 32 tile kernels and the any-range kernels), misaligned planes (byte-load prologue)
c, b = xb.xSad8x8Search(scur, sref, 32, u16=True)
check("sad search v2 R=32 u16 surface", np.array_equal(c, wc) and np.array_equal(b, wb))
c, b = xb.xSatd8x8Search(scur, sref, 32, u16=True)
wc2, wb2 = o.satd_search(scur, sref, 32, 0, 50)
check("satd search v3 R=32 u16 surface", np.array_equal(c, wc2) and np.array_equal(b, wb2))
s3c, s3r = rng.integers(0, 256, (16, 24)).astype(np.uint8), rng.integers(0, 256, (16 + 6, 24 + 6)).astype(np.uint8)
c, b = xb.xSatd8x8Search(s3c, s3r, 3, u16=True)
wc3, wb3 = o.satd_search(s3c, s3r, 3, 0, 6)
check("satd search any-range R=3 u16 surface", np.array_equal(c, wc3) and np.array_equal(b, wb3))
c, b = xb.xSad8x8Search(s3c, s3r, 3, u16=True)
wc3, wb3 = o.sad_search(s3c, s3r, 3, 0, 6)
check("sad search any-range R=3 u16 surface", np.array_equal(c, wc3) and np.array_equal(b, wb3))
xs8 = o.residual(1237 * 64 + 64 * 5, 4, 2)           # DCT8: 1242 blocks = 77 full units + a ragged one of 10
check("dct8 ragged tail", np.array_equal(xb.xDctNBatch(3, xs8, 2, 9), o.dct(xs8.reshape(-1, 8, 8), 3, 2, 9).ravel()))
print("ALL OK" if ok else "SOME FAILED")
sys.exit(0 if ok else 1)
