#!/usr/bin/env python
"""Times the secondary kernels: DCT 4/8/16, one-pass partialButterfly32, intra32 (GB/s of algorithmic bytes)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import x266_b200 as xb
dev = torch.device("cuda:0")
st = torch.cuda.current_stream().cuda_stream
def timeit(fn, reps=10):
    for _ in range(3): fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
nsamp = 1 << 30          # 1 Gi samples = 2 GiB in + 2 GiB out
src = (torch.randint(0, 1024, (nsamp,), device=dev, dtype=torch.int16) - torch.randint(0, 1024, (nsamp,), device=dev, dtype=torch.int16))
dst = torch.empty_like(src)
for log2n, (s1, s2) in ((2, (1, 8)), (3, (2, 9)), (4, (3, 10)), (5, (4, 11))):
    nb = nsamp >> (2 * log2n)
    ms = timeit(lambda: xb.xDctNBatchDev(log2n, src.data_ptr(), dst.data_ptr(), nb, s1, s2, st))
    print(f"dct{1 << log2n:<2d}: {ms:7.3f} ms  {nb / ms / 1e6:9.2f} G blocks/s  {nsamp * 4 / ms / 1e6:7.0f} GB/s  {nsamp * 4 / ms / 1e6 / 6459.3 * 100:5.1f}% of measured HBM", flush=True)
nb32 = nsamp >> 10
ms = timeit(lambda: xb.xIdct32BatchDev(src.data_ptr(), dst.data_ptr(), nb32, 7, 12, st))
print(f"idct32: {ms:7.3f} ms  {nb32 / ms / 1e6:9.2f} G blocks/s  {nsamp * 4 / ms / 1e6:7.0f} GB/s  {nsamp * 4 / ms / 1e6 / 6459.3 * 100:5.1f}% of measured HBM", flush=True)
line = 1 << 22
ms = timeit(lambda: xb.xPartialButterfly32Dev(src.data_ptr(), dst.data_ptr(), 4, line, st))
print(f"partialButterfly32 line={line}: {ms:7.3f} ms  {line * 128 / ms / 1e6:7.0f} GB/s", flush=True)
n = 1 << 20
refs = torch.randint(0, 256, (n, 129), device=dev, dtype=torch.uint8)
modes = (torch.arange(n, device=dev) % 35).to(torch.uint8)
pred = torch.empty((n, 1024), device=dev, dtype=torch.uint8)
ms = timeit(lambda: xb.xIntra32PredDev(refs.data_ptr(), modes.data_ptr(), pred.data_ptr(), n, st))
print(f"intra32 n={n}: {ms:7.3f} ms  {n / ms / 1e6:7.3f} G pred/s  {n * (1024 + 130) / ms / 1e6:7.0f} GB/s  {n * 1154 / ms / 1e6 / 6459.3 * 100:5.1f}% of measured HBM", flush=True)
nd = 1 << 17
curb = torch.randint(0, 256, (nd, 1024), device=dev, dtype=torch.uint8)
costs = torch.empty((nd, 35), device=dev, dtype=torch.int32)
bestm = torch.empty((nd,), device=dev, dtype=torch.int32)
ms = timeit(lambda: xb.xIntra32DecideDev(curb.data_ptr(), refs.data_ptr(), costs.data_ptr(), bestm.data_ptr(), nd, st), reps=3)
print(f"intra32 decide (35 modes x 16 SATD8x8 per block) n={nd}: {ms:7.3f} ms  {nd / ms / 1e3:7.2f} M blocks/s  {nd * 35 * 16 / ms / 1e6:7.2f} G (mode,8x8) SATDs/s", flush=True)
# fused residual + DCT32 from tiled frames: one launch over 8 stacked 8K luma frames (7680 x 34816)
w, h = 7680, 4352 * 8
ntile = (w // 16) * (h // 16)
cur = torch.randint(0, 256, (ntile * 512,), device=dev, dtype=torch.uint8)
prd = torch.randint(0, 256, (ntile * 512,), device=dev, dtype=torch.uint8)
coef = torch.empty(((w // 32) * (h // 32) * 1024,), device=dev, dtype=torch.int16)
ms = timeit(lambda: xb.xFrameResiDct32Dev(cur.data_ptr(), prd.data_ptr(), w, h, coef.data_ptr(), 4, 11, st))
nblk = (w // 32) * (h // 32)
print(f"frame residual+dct32 (tiled u8 cur/pred -> coef), {nblk} blocks: {ms:7.3f} ms  {nblk / ms / 1e6:6.3f} G blocks/s  "
      f"{nblk * 4096 / ms / 1e6:7.0f} GB/s useful (2 KB luma in + 2 KB coef out per block; the tiles' chroma/info halves are not touched)  "
      f"{nblk * 4096 / ms / 1e6 / 6459.3 * 100:5.1f}% of measured HBM", flush=True)
