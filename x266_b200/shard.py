"""Sharding rule for the multi-GPU path: units (32x32 blocks, SATD candidates, 8x8 search
blocks) are independent, so rank r of R owns the contiguous range [r*N/R, (r+1)*N/R).  Every unit is a
multiple of 64 bytes, so every shard keeps the 16-byte alignment the kernels require.  No data-path
collective exists; a gather is only needed if a caller wants one rank to hold the whole frame."""


def shard_range(n_units, rank, world):
    if world < 1 or not (0 <= rank < world):
        raise ValueError("bad rank/world")
    return (n_units * rank) // world, (n_units * (rank + 1)) // world


def shard_sizes(n_units, world):
    return [shard_range(n_units, r, world)[1] - shard_range(n_units, r, world)[0] for r in range(world)]
