#!/usr/bin/env python
"""e2e throughput of xDct32Batch by kind of caller memory (N=1): pinned, pageable (staged through the library's pinned
ring: thread count / chunk size / non-temporal copies), pageable handed to the driver, cudaHostRegister per call, and
registered once with xGpuHostRegister.  Pageable buffers are posix_memalign(4096)-style numpy arrays, as
src/x266.cpp:505,647-649 would hand them over."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import x266_b200 as xb

torch.cuda.set_device(0)
frames = int(sys.argv[1]) if len(sys.argv) > 1 else 16
n = frames * 32400
rng = np.random.default_rng(1)


def aligned(nbytes):
    raw = np.empty(nbytes + 4096, np.uint8)
    off = (-raw.ctypes.data) % 4096
    return raw[off:off + nbytes]


a = aligned(n * 2048).view(np.int16)
a[:] = rng.integers(-1023, 1024, a.size, dtype=np.int16)
b = aligned(n * 2048).view(np.int16)
b[:] = 0
pin_a = torch.from_numpy(a.copy()).pin_memory()
pin_b = torch.empty_like(pin_a).pin_memory()


def run(label, src, dst, reps=4):
    for _ in range(2):
        xb.xDct32Batch(src, 6, 11, out=dst)
    t = time.perf_counter()
    for _ in range(reps):
        xb.xDct32Batch(src, 6, 11, out=dst)
    dt = (time.perf_counter() - t) / reps
    print(f"{label:58s} {n / dt / 1e6:7.2f} M blocks/s  ({n * 2048 / dt / 1e9:5.1f} GB/s each way)", flush=True)
    return n / dt


print(f"## {frames} frames per call, host cpus {len(os.sched_getaffinity(0))}")
ref = run("pinned (cudaHostAlloc) in+out", pin_a.numpy(), pin_b.numpy())
want = pin_b.numpy().copy()
NT = {3: "nt both ways", 2: "nt to caller only", 1: "nt to pinned only", 0: "memcpy both"}
for thr in (1, 2, 4, 8, 16):
    if thr > len(os.sched_getaffinity(0)):
        break
    xb.tune(13, thr)
    for nt in (3, 0):
        xb.tune(14, nt)
        run(f"pageable staged: {thr:2d} threads, {NT[nt]}", a, b)
assert np.array_equal(b, want), "staged result differs"
# cache-resident ring: small chunks + cached stores into the pinned slots, so the DMA engine reads the staged bytes from the
# last-level cache instead of DRAM (and the D2H bytes are read back from it)
for thr in (8, 16):
    if thr > len(os.sched_getaffinity(0)):
        break
    xb.tune(13, thr)
    for nt in (3, 2, 0):
        xb.tune(14, nt)
        for chunk in (256, 512, 1024, 2048, 4096, 16384):
            xb.tune(4, chunk)
            run(f"pageable staged: {thr:2d} thr, {NT[nt]:18s} chunk {chunk:5d} ({chunk * 2048 >> 20 or chunk * 2 / 1024} MiB)", a, b)
assert np.array_equal(b, want), "staged result differs"
xb.tune(14, 3)
xb.tune(13, 0)
xb.tune(4, 0)
b[:] = 0
xb.tune(12, 1)
run("pageable handed to the driver (cudaMemcpyAsync stages)", a, b, reps=2)
assert np.array_equal(b, want)
b[:] = 0
xb.tune(12, 2)
run("pageable, cudaHostRegister + unregister per call", a, b, reps=2)
assert np.array_equal(b, want)
xb.tune(12, 0)
t = time.perf_counter()
xb.host_register(a)
xb.host_register(b)
print(f"xGpuHostRegister of 2 x {n * 2048 / 1e9:.2f} GB: {(time.perf_counter() - t) * 1e3:.1f} ms")
b[:] = 0
run("pageable registered once (xGpuHostRegister)", a, b)
assert np.array_equal(b, want)
xb.host_unregister(a)
xb.host_unregister(b)
# mixed: pageable in, pinned out
run("pageable in -> pinned out", a, pin_b.numpy())
run("pinned in -> pageable out", pin_a.numpy(), b)
assert np.array_equal(b, want)
print("all variants bit-identical to the pinned path")
