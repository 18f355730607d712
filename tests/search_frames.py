"""Config 3 of SURVEY.md 8(d), literally: the frames of the full-search known-answer test.
cur  = 1920x1080 luma, low byte of splitmix64(266) draws (one draw per pixel, raster order)
ref  = cur translated by the global motion (+5, -3) plus noise ((z >> 8) % 9) - 4 with z from splitmix64(267), clamped,
       then edge-replicated by the search range.  Interior blocks must therefore find mv = (+5, -3)."""
import numpy as np

GAMMA = np.uint64(0x9E3779B97F4A7C15)


def splitmix64(seed, n):
    with np.errstate(over="ignore"):
        z = np.uint64(seed) + GAMMA * np.arange(1, n + 1, dtype=np.uint64)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        return z ^ (z >> np.uint64(31))


def fnv1a64(buf):
    """FNV-1a-64 over the raw bytes, vectorised per byte position is not possible: plain loop in chunks."""
    h = 1469598103934665603
    for b in np.ascontiguousarray(buf).view(np.uint8).ravel().tolist():
        h = ((h ^ b) * 1099511628211) & 0xFFFFFFFFFFFFFFFF
    return h


def config3_frames(w=1920, h=1080, rng=32, dx=5, dy=-3):
    cur = (splitmix64(266, w * h) & np.uint64(0xFF)).astype(np.uint8).reshape(h, w)
    z = splitmix64(267, w * h).reshape(h, w)
    ys = np.clip(np.arange(h) - dy, 0, h - 1)
    xs = np.clip(np.arange(w) - dx, 0, w - 1)
    noise = ((z >> np.uint64(8)) % np.uint64(9)).astype(np.int32) - 4
    ref = np.clip(cur[ys][:, xs].astype(np.int32) + noise, 0, 255).astype(np.uint8)
    return cur, np.pad(ref, rng, mode="edge")


def sample_blocks(nblk, count=256, seed=1):
    return sorted(set(int(v) for v in np.random.default_rng(seed).integers(0, nblk, count)))


def argmin_rule(cost, rng):
    """lowest cost; ties -> smallest mvx^2 + mvy^2; then raster order (mvy, then mvx).  cost: [side][side] -> (cost, mvx, mvy)"""
    side = 2 * rng + 1
    my, mx = np.mgrid[0:side, 0:side]
    d2 = (mx - rng) ** 2 + (my - rng) ** 2
    key = (cost.astype(np.int64) << 40) | (d2.astype(np.int64) << 24) | (my.astype(np.int64) << 12) | mx.astype(np.int64)
    k = int(key.min())
    return k >> 40, (k & 0xFFF) - rng, ((k >> 12) & 0xFFF) - rng
