"""CPU suite, part 3: the N>1 host logic on world_size-2 gloo -- shard ranges are a partition, shards are
processed independently (the oracle stands in for the device kernel here: no GPU), and the optional frame
re-assembly gather reproduces the single-rank result."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from x266_b200.shard import shard_range, shard_sizes


def test_shard_ranges_partition():
    for n in (0, 1, 7, 2040, 32400, 2073600):
        for world in (1, 2, 3, 4, 8):
            rs = [shard_range(n, r, world) for r in range(world)]
            assert rs[0][0] == 0 and rs[-1][1] == n
            assert all(rs[i][1] == rs[i + 1][0] for i in range(world - 1))
            assert max(shard_sizes(n, world)) - min(shard_sizes(n, world)) <= 1


def _worker(rank, world, port, n_blocks, out_q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from oracle import Oracle
    o = Oracle()
    x = o.residual(n_blocks * 1024, 266, 1).reshape(n_blocks, 32, 32)      # same global batch on every rank
    lo, hi = shard_range(n_blocks, rank, world)
    mine = o.dct(x[lo:hi], 5, 6, 11)                                       # rank-local work, no exchange
    # timing protocol of bench.py: barrier, then max over ranks
    dist.barrier()
    t = torch.tensor([float(rank + 1)], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    assert t.item() == world
    # optional re-assembly: gather the (possibly ragged) coefficient slabs on every rank
    sizes = shard_sizes(n_blocks, world)
    pad = max(sizes)
    buf = torch.zeros((pad, 32, 32), dtype=torch.int16)
    buf[: hi - lo] = torch.from_numpy(mine)
    raw = buf.view(torch.uint8)                                            # gloo has no int16: ship the bytes
    parts = [torch.zeros_like(raw) for _ in range(world)]
    dist.all_gather(parts, raw)
    full = torch.cat([p.view(torch.int16)[:s] for p, s in zip(parts, sizes)]).numpy()
    if rank == 0:
        out_q.put(bool(np.array_equal(full, o.dct(x, 5, 6, 11))))
    dist.destroy_process_group()


def test_world2_gloo_shard_and_gather():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, 37, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert q.get(timeout=10) is True
