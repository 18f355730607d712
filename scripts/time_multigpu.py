#!/usr/bin/env python
"""xDct32BatchMultiGpu (one process, one host thread per GPU, chunks claimed from one shared counter) on all visible GPUs, pinned host
buffers; per-GPU chunk counts show how the claiming follows the host links."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import x266_b200 as xb

frames = int(sys.argv[1]) if len(sys.argv) > 1 else 16
have = torch.cuda.device_count()
for g in (1, 2, 4, 8):
    if g > have:
        break
    n = g * frames * 32400
    hin = torch.randint(-1023, 1024, (n, 32, 32), dtype=torch.int16).pin_memory()
    hout = torch.empty_like(hin).pin_memory()
    a, b = hin.numpy(), hout.numpy()
    xb.xDct32BatchMultiGpu(a, 6, 11, n_gpus=g, out=b)
    t = time.perf_counter()
    reps = 3
    for _ in range(reps):
        xb.xDct32BatchMultiGpu(a, 6, 11, n_gpus=g, out=b)
    dt = (time.perf_counter() - t) / reps
    # spot check against a single-GPU call
    chk = xb.xDct32Batch(a[:1000], 6, 11)
    ok = bool(np.array_equal(chk, b[:1000])) and bool(np.array_equal(xb.xDct32Batch(a[-1000:], 6, 11), b[-1000:]))
    print(f"{g} GPUs, {frames} frames per GPU: {n / dt / 1e6:7.2f} M blocks/s  ({n * 2048 / dt / 1e9:5.1f} GB/s each way)  matches single-GPU: {ok}", flush=True)
    del hin, hout
