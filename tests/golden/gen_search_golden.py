#!/usr/bin/env python
"""Mints tests/golden/search_kat.npz: config 3 through the UNMODIFIED reference satd8x8 (oracle/_ref =
src_tb/satd.c compiled in place) -- every candidate's cost is ref satd8x8(cur - ref(mv)); the window convention and the
argmin rule are ours (tests/search_frames.py), as the reference has no search loop.  Runs only where /root/reference exists.
  best[32400][3]      (cost, mvx, mvy) of every 8x8 block of the 1920x1080 frame, +-32
  sample[...]         block indices of the sampled full cost surfaces
  sample_fnv[...]     FNV-1a-64 of each sampled u32 cost surface [65][65]
  fnv_cur / fnv_ref   FNV-1a-64 of the generated planes (pins the generator)"""
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
sys.path.insert(0, os.path.join(HERE, ".."))
from oracle import Ref  # noqa: E402
from search_frames import argmin_rule, config3_frames, fnv1a64, sample_blocks  # noqa: E402

R = 32


def main():
    ref = Ref()
    cur, refp = config3_frames(rng=R)
    h, w = cur.shape
    bw, nblk, side = w // 8, (w // 8) * (h // 8), 2 * R + 1
    best = np.empty((nblk, 3), np.int32)
    sample = sample_blocks(nblk)
    sample_fnv = {}
    t0 = time.time()
    for by in range(h // 8):
        band = refp[by * 8: by * 8 + 2 * R + 8]                                    # rows the block row can reach
        win = np.lib.stride_tricks.sliding_window_view(band, (8, 8))               # [65][w+2R-7][8][8]
        for bx0 in range(0, bw, 40):
            blocks = range(bx0, min(bx0 + 40, bw))
            d = np.empty((len(blocks), side, side, 8, 8), np.int16)
            for j, bx in enumerate(blocks):
                c = cur[by * 8: by * 8 + 8, bx * 8: bx * 8 + 8].astype(np.int16)
                d[j] = c[None, None] - win[:, bx * 8: bx * 8 + side].astype(np.int16)
            cost = ref.satd(d.reshape(-1), threads=os.cpu_count() or 1).reshape(len(blocks), side, side)
            for j, bx in enumerate(blocks):
                b = by * bw + bx
                best[b] = argmin_rule(cost[j], R)
                if b in sample:
                    sample_fnv[b] = fnv1a64(cost[j].astype(np.uint32))
        if by % 10 == 0:
            print(f"row {by}/{h // 8}  {time.time() - t0:.0f}s", flush=True)
    np.savez_compressed(os.path.join(HERE, "search_kat.npz"), best=best, sample=np.array(sample, np.int64),
                        sample_fnv=np.array([sample_fnv[b] for b in sample], np.uint64),
                        fnv_cur=np.uint64(fnv1a64(cur)), fnv_ref=np.uint64(fnv1a64(refp)))
    interior = [by * bw + bx for by in range(8, 127, 17) for bx in range(8, 232, 23)]
    print("interior blocks at (+5,-3):", all(best[b, 1] == 5 and best[b, 2] == -3 for b in interior))
    print("done", time.time() - t0, "s")


if __name__ == "__main__":
    main()
