#!/usr/bin/env python
"""Host<->device link ceiling of this box: pure concurrent H2D + D2H copies from pinned memory, no kernel.

This is the roofline of the END-TO-END number (bench.py `e2e`): xDct32Batch with host buffers moves 2 KiB in and
2 KiB out per block, so blocks/s <= (GB/s each way) / 2048.

  one process:   python scripts/time_link_ceiling.py                (sweeps 1/2/4/8 of the visible GPUs, one thread)
  N processes:   torchrun --nproc-per-node N scripts/time_link_ceiling.py --per-rank   (each rank its own GPU, same instant)

Prints one JSON line per measurement; `each_way_GBps_total` is the aggregate over the GPUs taking part.
"""
import argparse
import glob
import json
import os
import subprocess
import sys
import time

import torch


def topo():
    info = {"cpus_allowed": len(os.sched_getaffinity(0))}
    nodes = {}
    for d in sorted(glob.glob("/sys/devices/system/node/node[0-9]*")):
        try:
            nodes[os.path.basename(d)] = open(os.path.join(d, "cpulist")).read().strip()
        except OSError:
            pass
    info["numa_nodes"] = nodes
    try:
        out = subprocess.run(["nvidia-smi", "--query-gpu=index,pci.bus_id,pcie.link.gen.current,pcie.link.width.current", "--format=csv,noheader"],
                             capture_output=True, text=True, timeout=20).stdout.strip().splitlines()
        gpus = []
        for line in out:
            idx, bus, gen, width = [v.strip() for v in line.split(",")]
            bdf = bus.lower()[4:] if len(bus) > 12 else bus.lower()
            node = None
            try:
                node = int(open(f"/sys/bus/pci/devices/{bdf}/numa_node").read())
            except (OSError, ValueError):
                pass
            gpus.append({"gpu": int(idx), "bdf": bdf, "gen": gen, "width": width, "numa_node": node})
        info["gpus"] = gpus
    except Exception as e:          # diagnostics only
        info["gpus"] = str(e)
    return info


def measure(devs, mb, reps, direction, seconds_out=None):
    """direction: 'h2d', 'd2h' or 'both'; returns GB/s each way, total over devs (wall clock around a full sync)."""
    n = mb << 20
    bufs = []
    for d in devs:
        torch.cuda.set_device(d)
        hin = torch.empty(n, dtype=torch.uint8, pin_memory=True)
        hout = torch.empty(n, dtype=torch.uint8, pin_memory=True)
        hin.fill_(1)
        din = torch.empty(n, dtype=torch.uint8, device=f"cuda:{d}")
        dout = torch.ones(n, dtype=torch.uint8, device=f"cuda:{d}")
        bufs.append((d, hin, hout, din, dout, torch.cuda.Stream(device=d), torch.cuda.Stream(device=d)))

    def go(k):
        for _ in range(k):
            for d, hin, hout, din, dout, s1, s2 in bufs:
                if direction in ("h2d", "both"):
                    with torch.cuda.stream(s1):
                        din.copy_(hin, non_blocking=True)
                if direction in ("d2h", "both"):
                    with torch.cuda.stream(s2):
                        hout.copy_(dout, non_blocking=True)
        for d in devs:
            torch.cuda.synchronize(d)

    go(2)
    t = time.perf_counter()
    go(reps)
    dt = time.perf_counter() - t
    if seconds_out is not None:
        seconds_out.append(dt)
    return len(devs) * n * reps / dt / 1e9


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--mb", type=int, default=512)
    ap.add_argument("--reps", type=int, default=8)
    ap.add_argument("--per-rank", action="store_true")
    args = ap.parse_args()
    if args.per_rank:
        import torch.distributed as dist
        rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
        torch.cuda.set_device(local)
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        for direction in ("h2d", "d2h", "both"):
            dist.barrier()
            torch.cuda.synchronize()
            secs = []
            v = measure([local], args.mb, args.reps, direction, secs)
            t = torch.tensor([v], dtype=torch.float64, device="cuda")
            mn = t.clone()
            mx = torch.tensor([secs[0]], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
            dist.all_reduce(mn, op=dist.ReduceOp.MIN)
            dist.all_reduce(mx, op=dist.ReduceOp.MAX)
            if rank == 0:
                # equal work per rank finishes with the slowest rank: `equal_split` is what a statically sharded job can get;
                # `sum_of_rank_rates` (each rank over its own duration) is what a dynamically balanced one could approach
                eq = world * (args.mb << 20) * args.reps / float(mx.item()) / 1e9
                print(json.dumps({"form": "one process per GPU", "gpus": world, "direction": direction,
                                  "each_way_GBps_total_equal_split": eq, "sum_of_rank_rates_GBps": float(t.item()),
                                  "slowest_rank_GBps": float(mn.item()),
                                  "e2e_blocks_per_s_ceiling": eq * 1e9 / 2048 if direction == "both" else None}), flush=True)
        dist.destroy_process_group()
        return
    print(json.dumps({"topology": topo()}), flush=True)
    have = torch.cuda.device_count()
    for g in (1, 2, 4, 8):
        if g > have:
            break
        for direction in ("h2d", "d2h", "both"):
            v = measure(list(range(g)), args.mb, args.reps, direction)
            print(json.dumps({"form": "one process", "gpus": g, "direction": direction, "each_way_GBps_total": v,
                              "e2e_blocks_per_s_ceiling": v * 1e9 / 2048 if direction == "both" else None}), flush=True)
    if have >= 2:
        # which GPUs share a host link: every GPU alone, then pairs (0, k)
        for d in range(have):
            print(json.dumps({"form": "single", "gpu": d, "both_GBps": measure([d], args.mb, 4, "both")}), flush=True)
        for d in range(1, have):
            print(json.dumps({"form": "pair", "gpus": [0, d], "both_GBps_total": measure([0, d], args.mb, 4, "both")}), flush=True)


if __name__ == "__main__":
    sys.exit(main())
