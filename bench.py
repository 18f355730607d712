#!/usr/bin/env python
"""bench.py -- headline benchmark of the x266 hot path on B200 (contract: see README / DESIGN.md).

Default workload (BASELINE.json configs[4]): 7680x4320 frames of 11-bit synthetic
residuals (32 400 32x32 blocks per frame), forward 2-D DCT32 with shifts 6/11, FRAMES frames resident in
HBM per GPU.  A "step" is one pass of the DCT over the whole resident batch of this rank.

  value     = blocks/s, device-resident in -> device-resident out, CUDA events, max over ranks
  e2e       = the same metric through the host-pointer C-ABI call xDct32Batch() with pinned host buffers
              (H2D + kernel + D2H inside the timed region); `e2e.roofline` is that figure against the host-link
              ceiling measured in the same run (pure concurrent H2D+D2H copies on all ranks); `e2e.pageable` is the
              same call on plain page-aligned malloc memory, as src/x266.cpp:505,647-649 would hand it over
  roofline  = algorithmic 4096 B/block / kernel time vs the measured HBM copy bandwidth
  cpu_baseline = the unmodified reference C (oracle/_ref) on this box's host cores, bounded sample

`--workload config3` (1080p +-32 SATD full search, one frame per GPU) and `--workload config4` (one 4K frame of mixed
4/8/16/32 transforms + search, sharded over the ranks: strong scaling) make those configurations the headline line.
`--impl reference` times the reference's own CPU implementation (oracle/_ref, all host threads) on the
same config / metric.  Launch: `python bench.py --gpus N --steps K --warmup W` (N>1 via torchrun).
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

BLOCKS_PER_FRAME = 32400            # 7680x4320 / 32x32
SHIFTS = (6, 11)                    # 10-bit video: shift_1st = log2N-1+(bitDepth-8) = 6, shift_2nd = 11
METRICS = {"config5": ("dct32_blocks_per_s", "blocks/s"),
           "config3": ("satd8x8_search_candidates_per_s", "candidates/s"),
           "config4": ("config4_4k_frames_per_s", "frames/s")}
SEARCH_BYTES_1080P = 551903296      # cur + padded ref + u32 cost surface


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def workload_name(args):
    if args.workload == "config3":
        return "config3: 1920x1080 8-bit luma per GPU, 8x8 SATD full search +-32 (32400 blocks x 4225 candidates), u32 cost surface + argmin"
    if args.workload == "config4":
        return ("config4: ONE 3840x2176 frame, mixed 4/8/16/32 forward transforms (8160 regions, seeded partition) + +-32 SATD search "
                "argmins (129600 blocks), every size class and the block rows sharded over the ranks")
    f = args.frames
    return f"config5: 7680x4320 11-bit residuals, DCT32 shifts 6/11, {f} frames ({f * BLOCKS_PER_FRAME} blocks) resident per GPU"


class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.p = None
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                       "-lms", "20"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.p = None

    def stop(self):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.05)
        self.p.terminate()
        try:
            out = self.p.communicate(timeout=5)[0]
        except Exception:
            self.p.kill()
            out = ""
        sm, mx, reasons, pw = [], [], set(), []
        for line in out.strip().splitlines():
            f = [v.strip() for v in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2])); pw.append(float(f[3]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "power_w_max": max(pw), "samples": len(sm),
                "reasons": sorted(reasons)}


def host_threads():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def aligned_empty(nbytes, align=4096):
    """page-aligned pageable memory, as _aligned_malloc(..., 4096) gives the reference's caller (src/x266.cpp:505)"""
    raw = np.empty(nbytes + align, np.uint8)
    off = (-raw.ctypes.data) % align
    return raw[off:off + nbytes]


# ---------------------------------------------------------------------------------------------------------------
# CPU arm: the reference's own C (oracle/_ref), all host threads
# ---------------------------------------------------------------------------------------------------------------
class CpuRef:
    """Constructed once; the timed calls below run only the reference routine on preallocated, pre-touched buffers."""

    def __init__(self):
        from oracle import Oracle, Ref, have_ref, build
        build()
        self.orc = Oracle()
        self.ref = Ref() if have_ref() else None
        self.kind = "reference" if self.ref is not None else "port"

    def dct32(self, x, out, threads):
        if self.ref is not None:
            return self.ref.dct32(x, *SHIFTS, threads=threads, out=out)
        return self.orc.dct(x, 5, *SHIFTS, threads=threads, out=out)

    def dct32_rate(self, x, threads, reps):
        x = x.reshape(-1, 32, 32)
        out = np.zeros_like(x)                              # pre-touched
        best = None
        for _ in range(reps):
            t = time.perf_counter()
            self.dct32(x, out, threads)
            dt = time.perf_counter() - t
            best = dt if best is None else min(best, dt)
        return x.shape[0] / best, out

    def satd(self, d, threads, out=None):
        return self.ref.satd(d, threads=threads, out=out) if self.ref is not None else self.orc.satd(d, threads=threads)

    def satd_rate(self, d, threads, reps):
        out = np.zeros(d.size // 64, np.int32)
        best = None
        for _ in range(reps):
            t = time.perf_counter()
            got = self.satd(d, threads, out)
            dt = time.perf_counter() - t
            best = dt if best is None else min(best, dt)
        return (d.size // 64) / best, got


def cpu_reference_o3_rate(n_blocks, threads):
    """Second row of BASELINE.md section 2: the same reference C built with -O3 -march=native (built where the
    reference tree lives; the GPU box may have another CPU, so it runs in a subprocess and may legitimately fail)."""
    code = (
        "import sys, time; sys.path.insert(0, %r)\n"
        "import numpy as np\n"
        "from oracle import Oracle, Ref\n"
        "o = Oracle(); r = Ref('o3'); x = o.residual(%d * 1024, 266, 1).reshape(-1, 32, 32); y = np.zeros_like(x)\n"
        "best = 1e9\n"
        "for _ in range(2):\n"
        "    t = time.perf_counter(); r.dct32(x, %d, %d, threads=%d, out=y); best = min(best, time.perf_counter() - t)\n"
        "print(%d / best)\n" % (ROOT, n_blocks, SHIFTS[0], SHIFTS[1], threads, n_blocks))
    try:
        out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=120)
        return float(out.stdout.strip().splitlines()[-1]) if out.returncode == 0 else None
    except Exception:
        return None


def run_reference(args):
    """--impl reference: the reference's own CPU implementation of the path, all host threads, bounded sample per step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    threads = host_threads()
    metric, unit = METRICS[args.workload]
    cpu = CpuRef()
    if args.workload == "config5":
        n = args.ref_frames * BLOCKS_PER_FRAME
        x = cpu.orc.residual(n * 1024, 266, 1).reshape(-1, 32, 32)
        out = np.zeros_like(x)
        fn = lambda: cpu.dct32(x, out, threads)                                        # noqa: E731
        units = n
        sample = f"{args.ref_frames} frames ({n} blocks) of the workload per step, gcc -O2, {threads} pthreads over contiguous ranges"
    else:
        # the reference has no search loop and no small transforms: its per-candidate routine satd8x8 (src_tb/satd.c:31-118) on
        # precomputed differences is the CPU arm of both search workloads (forming the differences is not even counted)
        n = 1 << 22
        d = cpu.orc.residual(n * 64, 266, 0)
        out = np.zeros(n, np.int32)
        fn = lambda: cpu.satd(d, threads, out)                                          # noqa: E731
        units = n if args.workload == "config3" else n / (129600 * 4225)                # config4: frames' worth of candidates
        sample = (f"{n} precomputed 8x8 differences per step through the reference satd8x8, gcc -O2, {threads} pthreads"
                  + ("" if args.workload == "config3" else "; expressed in frames of 547.56 M candidates (transforms not counted)"))
    for _ in range(args.warmup):
        fn()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        fn()
    dt = time.perf_counter() - t0
    value = units * args.steps / dt
    line = {
        "impl": "reference", "metric": metric, "value": value, "unit": unit, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True,
        "scaling": "strong" if args.workload == "config4" else "weak",
        "vs_baseline": None, "dtype": "int16 (int32 accumulate)", "data": "synthetic",
        "config": {"workload": workload_name(args), "sample": sample},
        "cpu_baseline": {"value": value, "unit": unit, "cores": threads, "kind": cpu.kind, "sample": sample},
        "e2e": {"value": value, "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)
    return 0


# ---------------------------------------------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------------------------------------------
class Env:
    """process group, device, timing helpers shared by the workloads"""

    def __init__(self, args):
        import torch
        import torch.distributed as dist
        import x266_b200 as xb
        self.torch, self.dist, self.xb, self.args = torch, dist, xb, args
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        if not torch.cuda.is_available():
            raise SystemExit("bench.py: no CUDA device -- the x266_b200 hot path has no CPU fallback")
        torch.cuda.set_device(self.local)
        self.dev = torch.device("cuda", self.local)
        self.host_grp = None
        if self.world > 1:
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            dist.init_process_group("nccl", device_id=self.dev)
            self.host_grp = dist.new_group(backend="gloo")       # host-side barrier: waiting ranks sleep instead of spinning a kernel
        xb.lib()                                                 # fail loudly if the CUDA library is missing
        self.stream = torch.cuda.current_stream()
        self.st = self.stream.cuda_stream
        self.gen = torch.Generator(device=self.dev)
        self.gen.manual_seed(266 + self.rank)
        self.peak, self.peak_src = measured_peaks()
        self.all_cpus = os.sched_getaffinity(0)
        self.numa = self.bind_near_gpu()
        # host copy threads of the pageable path: the ranks of one box share its cores
        self.copy_threads = max(1, min(8, host_threads() // (2 * self.world))) if self.world > 1 else 0
        xb.tune(13, self.copy_threads)

    def bind_near_gpu(self):
        """NUMA placement: run this rank (and so its pinned allocations and copy threads) on the cores next to its GPU, if the box
        exposes more than one node and the GPU's node has cores we are allowed to use.  Best effort; returns what was done."""
        try:
            props = self.torch.cuda.get_device_properties(self.local)
            bdf = f"{props.pci_domain_id:04x}:{props.pci_bus_id:02x}:{props.pci_device_id:02x}.0"
            node = int(open(f"/sys/bus/pci/devices/{bdf}/numa_node").read())
            nodes = [d for d in os.listdir("/sys/devices/system/node") if d.startswith("node") and d[4:].isdigit()]
            if node < 0 or len(nodes) < 2:
                return {"gpu_numa_node": node, "nodes": len(nodes), "bound": False}
            cpus = set()
            for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
                a, _, b = part.partition("-")
                cpus.update(range(int(a), int(b or a) + 1))
            mine = cpus & os.sched_getaffinity(0)
            if not mine:
                return {"gpu_numa_node": node, "nodes": len(nodes), "bound": False, "why": "no allowed cpu on that node"}
            os.sched_setaffinity(0, mine)
            return {"gpu_numa_node": node, "nodes": len(nodes), "bound": True, "cpus": len(mine)}
        except Exception as e:
            return {"bound": False, "why": str(e)[:80]}

    def unbind(self):
        """back to every allowed core (the CPU baseline uses all of them)"""
        os.sched_setaffinity(0, self.all_cpus)

    def barrier(self):
        self.torch.cuda.synchronize()
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def host_barrier(self):
        if self.world > 1:
            self.dist.barrier(group=self.host_grp)

    def max_over_ranks(self, v):
        if self.world == 1:
            return v
        t = self.torch.tensor([v], dtype=self.torch.float64, device=self.dev)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(self, v):
        if self.world == 1:
            return v
        t = self.torch.tensor([v], dtype=self.torch.float64, device=self.dev)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM)
        return float(t.item())

    def timed(self, fn, reps, warm=2):
        """ms per call: CUDA events on the launch stream, barrier on both sides, max over ranks"""
        torch = self.torch
        for _ in range(warm):
            fn()
        self.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(self.stream)
        for _ in range(reps):
            fn()
        e1.record(self.stream)
        torch.cuda.synchronize()
        return self.max_over_ranks(e0.elapsed_time(e1) / reps)

    def timed_steps(self, step, steps, warmup):
        """the contract's timed region: W warm-up steps, K steps bracketed by barrier + synchronize, per-step events, max over ranks"""
        torch, xb = self.torch, self.xb
        sampler = ClockSampler(self.local) if self.rank == 0 else None
        for _ in range(warmup):
            step()
        self.barrier()
        l0 = xb.kernel_launches()
        evs = [torch.cuda.Event(enable_timing=True) for _ in range(steps + 1)]
        evs[0].record(self.stream)
        for i in range(steps):
            step()
            evs[i + 1].record(self.stream)
        torch.cuda.synchronize()
        launches = xb.kernel_launches() - l0
        clocks = sampler.stop() if sampler else None
        self.barrier()
        per_step = [evs[i].elapsed_time(evs[i + 1]) for i in range(steps)]
        return {"total_ms": self.max_over_ranks(evs[0].elapsed_time(evs[-1])), "avg_ms": self.max_over_ranks(sum(per_step) / len(per_step)),
                "best_ms": self.max_over_ranks(min(per_step)), "launches": int(launches), "clocks": clocks}

    def wall(self, fn, reps, warm=1):
        """seconds per call of a SYNCHRONOUS host-pointer entry point: wall clock between barriers, max over ranks"""
        for _ in range(warm):
            fn()
        self.barrier()
        t0 = time.perf_counter()
        for _ in range(reps):
            fn()
        self.torch.cuda.synchronize()
        dt = self.max_over_ranks(time.perf_counter() - t0)
        self.barrier()
        return dt / reps

    def pinned(self, shape, dtype):
        return self.torch.empty(shape, dtype=dtype, pin_memory=True)

    def link_ceiling(self, mb=512, reps=8, rounds=3):
        """GB/s each way, summed over ranks: concurrent pure H2D + D2H copies from pinned memory, every rank at the same time --
        what the host link gives a host-buffer call that moves as many bytes out as in (scripts/time_link_ceiling.py)"""
        torch = self.torch
        n = mb << 20
        hin, hout = self.pinned(n, torch.uint8), self.pinned(n, torch.uint8)
        hin.fill_(1)
        din = torch.empty(n, dtype=torch.uint8, device=self.dev)
        dout = torch.ones(n, dtype=torch.uint8, device=self.dev)
        s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()

        def go():
            with torch.cuda.stream(s1):
                din.copy_(hin, non_blocking=True)
            with torch.cuda.stream(s2):
                hout.copy_(dout, non_blocking=True)

        go()
        go()
        best = None
        for _ in range(rounds):                              # best of `rounds`: a ceiling must not be undercut by a cold first pass
            self.barrier()
            t0 = time.perf_counter()
            for _ in range(reps):
                go()
            torch.cuda.synchronize()
            dt = self.max_over_ranks(time.perf_counter() - t0)
            best = dt if best is None else min(best, dt)
        self.barrier()
        return self.world * n * reps / best / 1e9

    def hbm_entry(self, metric, units, bytes_per_unit, ms, note=None):
        gbs = units * bytes_per_unit / (ms * 1e-3) / 1e9            # per GPU
        e = {"metric": metric, "value": self.world * units / (ms * 1e-3), "n_gpus": self.world, "ms_per_launch": ms,
             "roofline": {"bound": "hbm", "achieved": gbs, "peak": self.peak, "unit": "GB/s", "frac": gbs / self.peak, "traffic": None,
                          "algorithmic_bytes_per_unit": bytes_per_unit}}
        if note:
            e["config"] = note
        return e


def search_entry(env, name, ms, cands_per_gpu, config, cpu=None, surface_bytes=4):
    torch = env.torch
    props = torch.cuda.get_device_properties(env.dev)
    clk_hz = getattr(props, "clock_rate", 1965000) * 1e3
    # the bound that applies: the integer ALU pipe issues one warp instruction per two cycles per SM sub-partition;
    # a SATD candidate needs 32 VIMNMX.S16x2 on it, a SAD candidate 16 VABSDIFF4 -- nothing else counted
    alu_peak = props.multi_processor_count * 4 * clk_hz / 2 * 32 / (32 if name.startswith("satd") else 16)
    # algorithmic bytes of a 1080p +-32 frame: both planes + the cost surface at its element size (+ 12 B per block of argmins)
    gbs = (SEARCH_BYTES_1080P - (4 - surface_bytes) * 136890000 + (0 if surface_bytes else 32400 * 12)) / (ms * 1e-3) / 1e9
    return {"metric": name, "value": env.world * cands_per_gpu / (ms * 1e-3), "n_gpus": env.world, "ms_per_frame": ms, "cpu_baseline": cpu,
            "alu_pipe_bound": {"peak": alu_peak, "unit": "candidates/s per GPU", "frac": cands_per_gpu / (ms * 1e-3) / alu_peak,
                               "model": "32 VIMNMX.S16x2 (SATD) / 16 VABSDIFF4 (SAD) per candidate, 1 ALU warp instruction per 2 cycles per sub-partition"},
            "config": config,
            "roofline": {"bound": "integer ALU pipe / shared memory (not HBM; SURVEY 8(d))", "achieved": gbs, "peak": env.peak, "unit": "GB/s",
                         "frac": gbs / env.peak, "traffic": None}}


def satd_cpu_baseline(env, cpu, d_dev, o_dev):
    """the reference satd8x8 on the host cores over a bounded sample of the device's differences; also checks the device result"""
    n_s = 1 << 22
    ds = d_dev[:n_s].cpu().numpy().reshape(-1)
    rate, want = cpu.satd_rate(ds, host_threads(), 3)
    return {"value": rate, "unit": "candidates/s", "cores": host_threads(), "kind": cpu.kind,
            "sample": f"{n_s} of the 16.8M differences, gcc -O2, best of 3",
            "gpu_output_bit_exact_on_sample": bool(np.array_equal(o_dev[:n_s].cpu().numpy(), want))}


class Config4:
    """ONE 3840x2176 frame: every 32x32 region split 1x32^2 / 4x16^2 / 16x8^2 / 64x4^2 by a seeded draw, the four size classes and the
    +-32 SATD search of the luma sharded over the ranks (strong scaling, no collective)"""

    def __init__(self, env, sp, dp):
        from x266_b200.shard import shard_range
        torch = env.torch
        self.env, self.sp, self.dp = env, sp, dp
        regions = (3840 // 32) * (2176 // 32)
        zc = np.random.default_rng(268).integers(0, 4, regions)
        self.w, self.h, self.rg = 3840, 2160, 32
        self.nb = (self.w // 8) * (self.h // 8)
        self.cur = torch.randint(0, 256, (self.h, self.w), device=env.dev, generator=env.gen, dtype=torch.uint8)
        self.ref = torch.randint(0, 256, (self.h + 64, self.w + 64), device=env.dev, generator=env.gen, dtype=torch.uint8)
        self.b0, self.b1 = shard_range(self.nb, env.rank, env.world)
        self.best = torch.empty((self.b1 - self.b0, 3), device=env.dev, dtype=torch.int32)
        self.shards = []
        self.samples = 0
        for cls, log2n in enumerate((5, 4, 3, 2)):
            nblk = int((zc == cls).sum()) * (1024 >> (2 * log2n))
            lo, hi = shard_range(nblk, env.rank, env.world)
            self.shards.append((log2n, lo, hi - lo))
            self.samples += nblk << (2 * log2n)

    def launch(self, st_dct, st_search):
        """the frame's work as plain launches: the four transform classes on st_dct, the search on st_search"""
        xb = self.env.xb
        for log2n, lo, cnt in self.shards:
            off = lo << (2 * log2n + 1)                                         # bytes: this rank's slice of the class array
            if log2n == 5:
                xb.xDct32BatchDev(self.sp + off, self.dp + off, cnt, 4, 11, st_dct)
            else:
                xb.xDctNBatchDev(log2n, self.sp + off, self.dp + off, cnt, log2n - 1, log2n + 6, st_dct)
        xb.xSatd8x8SearchDev(self.cur.data_ptr(), self.ref.data_ptr(), self.w + 64, self.w, self.h, self.rg, self.b0, self.b1, 0,
                             self.best.data_ptr(), st_search)

    def capture(self):
        """One CUDA graph per frame: the transforms and the search are independent, so inside the graph they sit on two branches (the four
        small transform launches hide under the search), and a replay costs one launch instead of eight (7 kernels + a memset).  That is how
        an encoder's frame loop would drive a launch-bound frame; falls back to plain launches if the capture is refused."""
        torch = self.env.torch
        self.graph = None
        try:
            self.launch(self.env.st, self.env.st)                               # warm: attributes, pools, occupancy caches
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            side = torch.cuda.Stream()
            l0 = self.env.xb.kernel_launches()
            with torch.cuda.graph(g):
                cap = torch.cuda.current_stream()
                side.wait_stream(cap)
                self.launch(side.cuda_stream, cap.cuda_stream)
                cap.wait_stream(side)
            self.kernels_per_frame = self.env.xb.kernel_launches() - l0      # kernels recorded into the graph = kernels per replay
            g.replay()
            torch.cuda.synchronize()
            self.graph = g
        except Exception as e:                                                  # noqa: BLE001 - measured either way, and said which
            self.graph_error = str(e)[:160]
            torch.cuda.synchronize()

    def run(self):
        if getattr(self, "graph", None) is not None:
            self.graph.replay()
        else:
            self.launch(self.env.st, self.env.st)

    def entry(self, ms):
        return {"metric": "config4_4k_frames_per_s", "value": 1e3 / ms, "n_gpus": self.env.world, "ms_per_frame": ms, "scaling": "strong",
                "config": "config4: one 3840x2176 frame, mixed 4/8/16/32 forward transforms (8160 regions) + +-32 SATD search argmins "
                          "(129600 blocks, 547.6 M candidates), every class and the block rows split over the ranks; "
                          + ("one CUDA graph replay per frame (transforms and search on two branches)" if getattr(self, "graph", None) is not None
                             else "plain launches (graph capture refused: " + getattr(self, "graph_error", "not attempted") + ")"),
                "roofline": {"bound": "integer ALU pipe (search dominates)", "achieved": self.nb * 4225 / (ms * 1e-3) / 1e9, "peak": None,
                             "unit": "G candidates/s", "frac": None, "traffic": None}}


def secondary_section(env, src, dst, n_blocks, cpu):
    """the other kernels of the path, same run.  Every rank runs the same per-GPU workload (independent units, no collective), times are
    max over ranks, values are whole-job aggregates."""
    torch, xb, st, dev, g, world, rank = env.torch, env.xb, env.st, env.dev, env.gen, env.world, env.rank
    sp, dp = src.data_ptr(), dst.data_ptr()
    secondary = []
    # sustained: >= 2 s of back-to-back launches of the headline kernel with the clocks sampled -- does the burst figure hold?
    sampler = ClockSampler(env.local) if rank == 0 else None
    reps = max(50, int(2200 / max(0.05, n_blocks * 4096 / (env.peak * 1e9) * 1e3)))
    ms = env.timed(lambda: xb.xDct32BatchDev(sp, dp, n_blocks, SHIFTS[0], SHIFTS[1], st), reps, warm=3)
    e = env.hbm_entry("dct32_sustained_blocks_per_s", n_blocks, 4096, ms, f"{reps} back-to-back launches = {reps * ms / 1e3:.2f} s of the headline kernel")
    e["clocks"] = sampler.stop() if sampler else None
    secondary.append(e)

    n_c = 1 << 24                                                     # 16.8M candidates = 2.1 GB of diffs per GPU
    d = torch.randint(-255, 256, (n_c, 64), device=dev, generator=g, dtype=torch.int16)
    o = torch.empty(n_c, device=dev, dtype=torch.int32)
    ms = env.timed(lambda: xb.xSatd8x8BatchDev(d.data_ptr(), o.data_ptr(), n_c, st), 10)
    secondary.append(env.hbm_entry("satd8x8_batch_candidates_per_s", n_c, 132, ms, "16.8M precomputed 9-bit 8x8 differences per GPU"))
    satd_cpu = None
    if rank == 0 and cpu is not None:
        # the reference has no search loop, so the same per-candidate rate is the CPU baseline of the full search
        satd_cpu = satd_cpu_baseline(env, cpu, d, o)
        secondary[-1]["cpu_baseline"] = satd_cpu
    env.host_barrier()
    del d, o
    # config 3: full search +-32 over one 1920x1080 frame per GPU, argmin + full u32 cost surface
    w, h, rg = 1920, 1080, 32
    cur = torch.randint(0, 256, (h, w), device=dev, generator=g, dtype=torch.uint8)
    refp = torch.randint(0, 256, (h + 2 * rg, w + 2 * rg), device=dev, generator=g, dtype=torch.uint8)
    nb = (w // 8) * (h // 8)
    cost = torch.empty((nb, 65, 65), device=dev, dtype=torch.int32)
    best = torch.empty((nb, 3), device=dev, dtype=torch.int32)
    for name, fn in (("satd8x8_search_candidates_per_s", xb.xSatd8x8SearchDev), ("sad8x8_search_candidates_per_s", xb.xSad8x8SearchDev)):
        ms = env.timed(lambda: fn(cur.data_ptr(), refp.data_ptr(), w + 2 * rg, w, h, rg, 0, nb, cost.data_ptr(), best.data_ptr(), st), 5)
        secondary.append(search_entry(env, name, ms, nb * 65 * 65, "config3: 1920x1080 per GPU, +-32, u32 cost surface + argmin",
                                      satd_cpu if name.startswith("satd") else None))
    # the same searches with the 16-bit cost surface (xS*SearchU16Dev; exact, half the surface bytes) and with the argmins only
    for name, fn16, fn in (("satd8x8_search", xb.xSatd8x8SearchU16Dev, xb.xSatd8x8SearchDev), ("sad8x8_search", xb.xSad8x8SearchU16Dev, xb.xSad8x8SearchDev)):
        ms = env.timed(lambda: fn16(cur.data_ptr(), refp.data_ptr(), w + 2 * rg, w, h, rg, 0, nb, cost.data_ptr(), best.data_ptr(), st), 5)
        secondary.append(search_entry(env, name + "_u16_surface_candidates_per_s", ms, nb * 65 * 65,
                                      "config3 frame, +-32, u16 cost surface + argmin", surface_bytes=2))
        ms = env.timed(lambda: fn(cur.data_ptr(), refp.data_ptr(), w + 2 * rg, w, h, rg, 0, nb, 0, best.data_ptr(), st), 5)
        secondary.append(search_entry(env, name + "_argmin_only_candidates_per_s", ms, nb * 65 * 65,
                                      "config3 frame, +-32, argmin triples only (no cost surface)", surface_bytes=0))
    del cur, refp, cost, best
    # config 2: one 1080p frame of residuals (2040 blocks) -- latency of a single launch
    ms = env.timed(lambda: xb.xDct32BatchDev(sp, dp, 2040, SHIFTS[0], SHIFTS[1], st), 200, warm=20)
    secondary.append({"metric": "dct32_1080p_frame_launch_latency_us", "value": ms * 1e3, "n_gpus": world, "higher_is_better": False,
                      "config": "config2: 2040 blocks, one launch, back-to-back launches on one stream (launch bound)",
                      "roofline": {"bound": "launch latency", "achieved": 2040 * 4096 / (ms * 1e-3) / 1e9, "peak": env.peak, "unit": "GB/s",
                                   "frac": 2040 * 4096 / (ms * 1e-3) / 1e9 / env.peak, "traffic": None}})
    c4 = Config4(env, sp, dp)
    assert c4.samples <= src.numel(), "config 4 class arrays must fit the resident buffers"
    c4.capture()
    secondary.append(c4.entry(env.timed(c4.run, 20)))
    del c4
    # config 4 flavour: the small transforms and the inverse on up to 1 Gi samples per GPU -- never past the resident buffers
    ns = 1 << 30
    while ns > src.numel():
        ns >>= 1
    assert ns <= src.numel() and ns <= dst.numel()
    for log2n, sh in ((4, (3, 10)), (3, (2, 9)), (2, (1, 8))):
        ms = env.timed(lambda: xb.xDctNBatchDev(log2n, sp, dp, ns >> (2 * log2n), sh[0], sh[1], st), 10)
        secondary.append(env.hbm_entry(f"dct{1 << log2n}_blocks_per_s", ns >> (2 * log2n), 4 << (2 * log2n), ms))
    ms = env.timed(lambda: xb.xIdct32BatchDev(sp, dp, ns >> 10, 7, 10, st), 10)
    secondary.append(env.hbm_entry("idct32_blocks_per_s", ns >> 10, 4096, ms, "parity unpinned (no inverse in the reference)"))
    # N2: fused residual + DCT32 straight from ref_block_t-tiled current / prediction frames (16 stacked 8K luma frames, one launch)
    fw, fh = 7680, 4320 * 16
    ntile = (fw // 16) * (fh // 16)
    if (fw // 32) * (fh // 32) * 1024 <= dst.numel():
        fcur = torch.randint(0, 256, (ntile * 512,), device=dev, generator=g, dtype=torch.uint8)
        fprd = torch.randint(0, 256, (ntile * 512,), device=dev, generator=g, dtype=torch.uint8)
        nfb = (fw // 32) * (fh // 32)
        ms = env.timed(lambda: xb.xFrameResiDct32Dev(fcur.data_ptr(), fprd.data_ptr(), fw, fh, dp, 4, 11, st), 10)
        secondary.append(env.hbm_entry("frame_resi_dct32_blocks_per_s", nfb, 4096, ms,
                                       "xFrameResiDct32: 16 stacked 8K frames of ref_block_t tiles (cur, pred) -> coefficients; 2 KB of luma in + 2 KB out per block"))
        del fcur, fprd
    npred = 1 << 20
    refs = torch.randint(0, 256, (npred, 129), device=dev, generator=g, dtype=torch.uint8)
    modes = (torch.arange(npred, device=dev) % 35).to(torch.uint8)
    pred = torch.empty((npred, 1024), device=dev, dtype=torch.uint8)
    sampler = ClockSampler(env.local) if rank == 0 else None
    ms = env.timed(lambda: xb.xIntra32PredDev(refs.data_ptr(), modes.data_ptr(), pred.data_ptr(), npred, st), 200)
    e = env.hbm_entry("intra32_predictions_per_s", npred, 1154, ms,
                      "mode = i % 35 (mode-interleaved), 200 launches; restatement pinned against every table / projection list of the BSV")
    e["clocks"] = sampler.stop() if sampler else None      # the kernel is issue / latency bound: its rate follows the SM clock
    secondary.append(e)
    nblk = npred // 35
    ms = env.timed(lambda: xb.xIntra32PredModesDev(refs.data_ptr(), nblk, (1 << 35) - 1, pred.data_ptr(), st), 20)
    secondary.append(env.hbm_entry("intra32_mode_major_predictions_per_s", nblk * 35, (129 + 35 * 1024) / 35.0, ms,
                                   "xIntra32PredModes: all 35 modes of 29959 blocks, a block's references staged once"))
    del modes, pred
    # N1 + N3: the closed block loop (decide -> predict -> residual -> DCT32 -> quant stub -> IDCT32 -> recon), one 8K frame of blocks
    nenc = BLOCKS_PER_FRAME
    ecur = torch.randint(0, 256, (nenc, 1024), device=dev, generator=g, dtype=torch.uint8)
    elev = torch.empty((nenc, 1024), device=dev, dtype=torch.int16)
    erec = torch.empty((nenc, 1024), device=dev, dtype=torch.uint8)
    ebest = torch.empty(nenc, device=dev, dtype=torch.int32)
    ms = env.timed(lambda: xb.xIntra32EncodeBlockDev(ecur.data_ptr(), refs.data_ptr(), nenc, 27, elev.data_ptr(), erec.data_ptr(), ebest.data_ptr(), 0, st), 10)
    e = env.hbm_entry("intra32_encode_blocks_per_s", nenc, 1024 + 129 + 2048 + 1024 + 4, ms,
                      "xIntra32EncodeBlock: fused 35-mode decision + reconstruction loop, qp 27; instruction-issue bound (profiles/r02_ncu_encode.md), "
                      "quantiser stub and inverse transform unpinned (not in the reference)")
    e["roofline"]["bound"] = "instruction issue (HBM fraction shown)"
    secondary.append(e)
    emodes = ebest.to(torch.uint8)
    ms = env.timed(lambda: xb.xIntra32ReconDev(ecur.data_ptr(), refs.data_ptr(), emodes.data_ptr(), nenc, 27, elev.data_ptr(), erec.data_ptr(), st), 10)
    secondary.append(env.hbm_entry("intra32_recon_blocks_per_s", nenc, 1024 + 130 + 2048 + 1024, ms, "xIntra32Recon: the Recon channel alone (mode given)"))
    del refs, ecur, elev, erec, ebest, emodes
    if world > 1:
        # optional frame re-assembly: every rank contributes one 8K frame of coefficients (66 MB) and
        # receives all of them -- the only collective in the repo, off the hot path, NCCL over NVLink/NVSwitch.
        # (bytes view: torch's NCCL binding has no int16)
        try:
            slab = dst[:BLOCKS_PER_FRAME].reshape(-1).view(torch.uint8)
            full = torch.empty(world * slab.numel(), dtype=torch.uint8, device=dev)
            ms = env.timed(lambda: env.dist.all_gather_into_tensor(full, slab), 10)
            ok = bool(torch.equal(full[rank * slab.numel():(rank + 1) * slab.numel()], slab))
            gbs = (world - 1) * slab.numel() / (ms * 1e-3) / 1e9
            secondary.append({"metric": "coef_frame_allgather_GBps_per_rank", "value": gbs, "n_gpus": world, "ms": ms,
                              "config": "ncclAllGather of one 8K coefficient frame per rank (optional re-assembly, not on the hot path)",
                              "own_slab_intact": ok,
                              "roofline": {"bound": "nvlink", "achieved": gbs, "peak": 770.0, "unit": "GB/s", "frac": gbs / 770.0, "traffic": None}})
        except Exception as e:          # optional metric: never fail the benchmark over it
            secondary.append({"metric": "coef_frame_allgather_GBps_per_rank", "value": None, "error": str(e)[:200],
                              "roofline": {"bound": "nvlink", "achieved": None, "peak": 770.0, "unit": "GB/s", "frac": None, "traffic": None}})
    return secondary


def e2e_dct32(env, src, dst):
    """the headline metric through the reference-facing call xDct32Batch with HOST buffers, copies inside the timed region"""
    torch, xb, args, world = env.torch, env.xb, env.args, env.world
    frames = args.e2e_frames if args.e2e_frames > 0 else 32          # 2 x 2.1 GB of pinned memory per rank, at every N
    while True:
        try:
            nb = frames * BLOCKS_PER_FRAME
            hin = env.pinned((nb, 32, 32), torch.int16)
            hout = env.pinned((nb, 32, 32), torch.int16)
            break
        except RuntimeError:
            if frames <= 4:
                raise
            frames //= 2
    hin.copy_(src[:nb].cpu())
    a, b = hin.numpy(), hout.numpy()
    steps = max(3, min(args.steps, 6))
    sec = env.wall(lambda: xb.xDct32Batch(a, *SHIFTS, out=b), steps, warm=2)
    value = world * nb / sec
    ok = bool(np.array_equal(b, dst[:nb].cpu().numpy()))
    ceiling = env.link_ceiling()
    gbs = value * 2048 / 1e9
    out = {"value": value, "unit": "blocks/s", "h2d_bytes_per_step": world * nb * 2048, "d2h_bytes_per_step": world * nb * 2048,
           "api": "xDct32Batch (host pointers, pinned)", "sample": f"{frames} frames per step per rank, {steps} steps",
           "matches_device_path": ok, "numa": env.numa,
           "roofline": {"bound": "host link", "achieved": gbs, "peak": ceiling, "unit": "GB/s each way, all ranks", "frac": gbs / ceiling,
                        "peak_source": "measured in this run: concurrent pinned H2D + D2H copies of 512 MiB on every rank at once, no kernel"}}
    # pageable caller memory (page-aligned malloc): staged through the library's pinned ring by its host copy threads
    pf = min(frames, 16)
    pn = pf * BLOCKS_PER_FRAME
    pa = aligned_empty(pn * 2048).view(np.int16)
    pb = aligned_empty(pn * 2048).view(np.int16)
    pa[:] = a.reshape(-1)[:pn * 1024]
    pb[:] = 0
    sec = env.wall(lambda: xb.xDct32Batch(pa, *SHIFTS, out=pb), 3, warm=1)
    pv = world * pn / sec
    out["pageable"] = {"value": pv, "unit": "blocks/s", "frac_of_pinned": pv / value, "matches_device_path": bool(np.array_equal(pb, b.reshape(-1)[:pn * 1024])),
                       "api": "xDct32Batch (host pointers, posix_memalign-style pageable buffers, staged ring)",
                       "host_copy_threads_per_rank": int(xb.host_copy_threads()), "sample": f"{pf} frames per step per rank, 3 steps"}
    # the same pageable buffers page-locked once with xGpuHostRegister -- what a caller that allocates its frame buffers once
    # (src/x266.cpp:647-649) does at start-up; registration time is reported, not inside the timed region
    try:
        t0 = time.perf_counter()
        xb.host_register(pa)
        xb.host_register(pb)
        reg_ms = (time.perf_counter() - t0) * 1e3
        try:
            pb[:] = 0
            sec = env.wall(lambda: xb.xDct32Batch(pa, *SHIFTS, out=pb), 3, warm=1)
            rv = world * pn / sec
            out["registered"] = {"value": rv, "unit": "blocks/s", "frac_of_pinned": rv / value,
                                 "matches_device_path": bool(np.array_equal(pb, b.reshape(-1)[:pn * 1024])),
                                 "api": "xGpuHostRegister once, then xDct32Batch on the same malloc'd buffers",
                                 "register_ms_once": reg_ms, "registered_bytes": int(2 * pn * 2048)}
        finally:
            xb.host_unregister(pa)
            xb.host_unregister(pb)
    except Exception as e:
        out["registered"] = {"value": None, "error": str(e)[:200]}
    del pa, pb
    if world > 1:
        # the native single-process form: ONE process, one host thread per GPU (xDct32BatchMultiGpu), the other ranks idle
        out["single_process_multi_gpu"] = None
        env.host_barrier()
        if env.rank == 0:
            try:
                per = max(4, frames // 2)
                tot = world * per * BLOCKS_PER_FRAME
                big_in = env.pinned((tot, 32, 32), torch.int16)
                big_out = env.pinned((tot, 32, 32), torch.int16)
                for r in range(world):
                    big_in[r * per * BLOCKS_PER_FRAME:(r + 1) * per * BLOCKS_PER_FRAME].copy_(hin[:per * BLOCKS_PER_FRAME])
                ai, ao = big_in.numpy(), big_out.numpy()
                xb.xDct32BatchMultiGpu(ai, *SHIFTS, n_gpus=world, out=ao)
                t0 = time.perf_counter()
                for _ in range(3):
                    xb.xDct32BatchMultiGpu(ai, *SHIFTS, n_gpus=world, out=ao)
                sec = (time.perf_counter() - t0) / 3
                torch.cuda.set_device(env.local)
                out["single_process_multi_gpu"] = {"value": tot / sec, "unit": "blocks/s", "api": "xDct32BatchMultiGpu (one process, one host thread per GPU, pinned)",
                                                   "sample": f"{per} frames per GPU per step, 3 steps",
                                                   "matches": bool(np.array_equal(ao[:per * BLOCKS_PER_FRAME], b[:per * BLOCKS_PER_FRAME]))}
                del big_in, big_out
            except Exception as e:
                out["single_process_multi_gpu"] = {"value": None, "error": str(e)[:200]}
        env.host_barrier()
    return out


def bench_config5(env):
    torch, xb, args, world, rank, dev = env.torch, env.xb, env.args, env.world, env.rank, env.dev
    from x266_b200.shard import shard_range
    xb.set_dct_variant({"auto": xb.DCT_AUTO, "bfly": xb.DCT_BFLY, "imma": xb.DCT_IMMA}[args.variant])
    # resident synthetic batch: this rank's shard (independent blocks, no collective on the path)
    lo, hi = shard_range(world * args.frames * BLOCKS_PER_FRAME, rank, world)      # weak scaling: global batch grows with N
    n_blocks = hi - lo
    g = env.gen
    src = (torch.randint(0, 1024, (n_blocks, 32, 32), device=dev, generator=g, dtype=torch.int16)
           - torch.randint(0, 1024, (n_blocks, 32, 32), device=dev, generator=g, dtype=torch.int16))
    dst = torch.empty_like(src)
    sp, dp, st = src.data_ptr(), dst.data_ptr(), env.st
    t = env.timed_steps(lambda: xb.xDct32BatchDev(sp, dp, n_blocks, SHIFTS[0], SHIFTS[1], st), args.steps, args.warmup)
    value = world * n_blocks * args.steps / (t["total_ms"] * 1e-3)
    e2e = e2e_dct32(env, src, dst)
    # snapshot of the device result for the CPU-baseline parity check (the secondary section reuses the buffers)
    n_cpu = min(args.cpu_frames * BLOCKS_PER_FRAME, n_blocks)
    xs = src[:n_cpu].cpu().numpy() if rank == 0 else None
    gpu_sample = dst[:n_cpu].cpu().numpy().reshape(-1) if rank == 0 else None
    cpu_ref = CpuRef() if rank == 0 else None
    env.unbind()
    secondary = [] if args.no_secondary else secondary_section(env, src, dst, n_blocks, cpu_ref)
    # CPU baseline: the reference C on this box's host cores, bounded sample, rank 0 (the other ranks sleep on a host barrier)
    cpu = None
    if rank == 0:
        threads = host_threads()
        rate, y = cpu_ref.dct32_rate(xs, threads, 2)
        rate1, _ = cpu_ref.dct32_rate(xs[: max(1, n_cpu // 8)], 1, 1)
        cpu = {"value": rate, "unit": "blocks/s", "cores": threads, "kind": cpu_ref.kind,
               "sample": f"{n_cpu // BLOCKS_PER_FRAME} frames ({n_cpu} blocks) of the workload, gcc -O2, best of 2, output buffer pre-touched",
               "single_thread_value": rate1, "o3_march_native_value": cpu_reference_o3_rate(n_cpu, threads) if world == 1 else None,
               "gpu_output_bit_exact_on_sample": bool(np.array_equal(y.reshape(-1), gpu_sample))}
    env.host_barrier()
    if rank != 0:
        return None
    achieved = n_blocks * 4096 / (t["avg_ms"] * 1e-3) / 1e9
    traffic, traffic_src = None, None
    tp = os.path.join(ROOT, "profiles", "dct32_traffic.json")
    if os.path.exists(tp):
        try:
            tj = json.load(open(tp))
            traffic = tj.get("dram_bytes_per_launch")
            traffic_src = f"NOT measured in this run: ncu --set full capture of the same launch, {tj.get('source', 'profiles/dct32_traffic.json')}"
        except Exception:
            traffic = None
    return {
        "metric": "dct32_blocks_per_s", "value": value, "unit": "blocks/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": t["total_ms"] / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "int16 (s8/u8 byte planes on int8 tensor cores, int32 accumulate)" if args.variant != "bfly" else "int16 (int32 accumulate)",
        "data": "synthetic",
        "config": {"workload": workload_name(args), "variant": args.variant, "l2_policy": "inputs (4.25 GB) larger than L2",
                   "timing": "CUDA events on the launch stream, max over ranks", "best_ms_per_step": t["best_ms"]},
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": env.peak, "unit": "GB/s", "frac": achieved / env.peak,
                     "traffic": traffic, "traffic_source": traffic_src, "peak_source": env.peak_src, "algorithmic_bytes_per_block": 4096,
                     "blocks_per_launch": n_blocks, "avg_launch_ms": t["avg_ms"]},
        "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": t["launches"], "clocks": t["clocks"], "secondary": secondary,
    }


def bench_config3(env):
    """headline = config 3: one 1080p frame per GPU, +-32 SATD full search, u32 cost surface + argmin"""
    torch, xb, args, world, rank, dev, g, st = env.torch, env.xb, env.args, env.world, env.rank, env.dev, env.gen, env.st
    w, h, rg = 1920, 1080, 32
    nb = (w // 8) * (h // 8)
    cands = nb * 65 * 65
    cur = torch.randint(0, 256, (h, w), device=dev, generator=g, dtype=torch.uint8)
    refp = torch.randint(0, 256, (h + 2 * rg, w + 2 * rg), device=dev, generator=g, dtype=torch.uint8)
    cost = torch.empty((nb, 65, 65), device=dev, dtype=torch.int32)
    best = torch.empty((nb, 3), device=dev, dtype=torch.int32)
    t = env.timed_steps(lambda: xb.xSatd8x8SearchDev(cur.data_ptr(), refp.data_ptr(), w + 2 * rg, w, h, rg, 0, nb, cost.data_ptr(), best.data_ptr(), st),
                        args.steps, args.warmup)
    value = world * cands * args.steps / (t["total_ms"] * 1e-3)
    # e2e: the host-pointer call with pinned planes and outputs
    hc, hr = env.pinned((h, w), torch.uint8), env.pinned((h + 2 * rg, w + 2 * rg), torch.uint8)
    hc.copy_(cur.cpu()); hr.copy_(refp.cpu())
    hcost, hbest = env.pinned((nb, 65, 65), torch.int32), env.pinned((nb, 3), torch.int32)
    L = xb.lib()

    def host_call():
        rc = L.xSatd8x8Search(hc.data_ptr(), hr.data_ptr(), w + 2 * rg, w, h, rg, 0, nb, hcost.data_ptr(), hbest.data_ptr())
        assert rc == 0, xb.last_error()

    sec = env.wall(host_call, 3, warm=1)
    e2e_ok = bool(torch.equal(hbest, best.cpu()) and torch.equal(hcost, cost.cpu()))
    # the same call with the 16-bit cost surface (xSatd8x8SearchU16): the D2H copy of the surface is what the end-to-end time consists of
    hcost16 = hcost.view(torch.int16).view(-1)[:nb * 65 * 65].view(nb, 65, 65)

    def host_call16():
        rc = L.xSatd8x8SearchU16(hc.data_ptr(), hr.data_ptr(), w + 2 * rg, w, h, rg, 0, nb, hcost16.data_ptr(), hbest.data_ptr())
        assert rc == 0, xb.last_error()

    sec16 = env.wall(host_call16, 3, warm=1)
    e2e16_ok = bool(torch.equal(hbest, best.cpu()) and torch.equal(hcost16.to(torch.int32), cost.cpu()))
    cpu = None
    env.unbind()
    if rank == 0:
        cr = CpuRef()
        n_s = 1 << 22
        rate, _ = cr.satd_rate(cr.orc.residual(n_s * 64, 266, 0), host_threads(), 3)
        cpu = {"value": rate, "unit": "candidates/s", "cores": host_threads(), "kind": cr.kind,
               "sample": f"{n_s} precomputed 8x8 differences through the reference satd8x8 (the reference has no search loop), gcc -O2, best of 3"}
    env.host_barrier()
    if rank != 0:
        return None
    e = search_entry(env, "satd8x8_search_candidates_per_s", t["avg_ms"], cands, workload_name(args))
    return {
        "metric": "satd8x8_search_candidates_per_s", "value": value, "unit": "candidates/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": t["total_ms"] / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u8 pixels, int16 Hadamard coefficients (packed 16x2), u32 costs", "data": "synthetic",
        "config": {"workload": workload_name(args), "l2_policy": "the 548 MB cost surface written per step is larger than L2; the 4 MB of pixels are reused 4225x by design",
                   "timing": "CUDA events on the launch stream, max over ranks", "best_ms_per_step": t["best_ms"]},
        "roofline": dict(e["roofline"], bound="hbm", note="not the binding resource: the kernel is integer-ALU bound, see alu_pipe_bound"),
        "alu_pipe_bound": e["alu_pipe_bound"], "cpu_baseline": cpu,
        "e2e": {"value": world * cands / sec, "unit": "candidates/s", "h2d_bytes_per_step": world * (hc.numel() + hr.numel()),
                "d2h_bytes_per_step": world * (hcost.numel() * 4 + hbest.numel() * 4), "api": "xSatd8x8Search (host pointers, pinned)",
                "matches_device_path": e2e_ok},
        "e2e_u16_surface": {"value": world * cands / sec16, "unit": "candidates/s", "h2d_bytes_per_step": world * (hc.numel() + hr.numel()),
                            "d2h_bytes_per_step": world * (hcost16.numel() * 2 + hbest.numel() * 4),
                            "api": "xSatd8x8SearchU16 (host pointers, pinned; 16-bit cost surface, exact)", "matches_device_path": e2e16_ok},
        "gpu_launches": t["launches"], "clocks": t["clocks"], "secondary": [],
    }


def bench_config4(env):
    """headline = config 4: ONE 4K frame sharded over the ranks (strong scaling)"""
    torch, xb, args, world, rank, dev = env.torch, env.xb, env.args, env.world, env.rank, env.dev
    n = 3840 * 2176
    src = (torch.randint(0, 256, (n,), device=dev, generator=env.gen, dtype=torch.int16)
           - torch.randint(0, 256, (n,), device=dev, generator=env.gen, dtype=torch.int16))
    dst = torch.empty_like(src)
    c4 = Config4(env, src.data_ptr(), dst.data_ptr())
    c4.capture()
    t = env.timed_steps(c4.run, args.steps, args.warmup)
    value = args.steps / (t["total_ms"] * 1e-3)
    # e2e: the same shard through the host-pointer calls (pinned): class arrays up, coefficients down, planes up, argmins down
    hsrc, hdst = env.pinned((n,), torch.int16), env.pinned((n,), torch.int16)
    hsrc.copy_(src.cpu())
    hc, hr = env.pinned((c4.h, c4.w), torch.uint8), env.pinned((c4.h + 64, c4.w + 64), torch.uint8)
    hc.copy_(c4.cur.cpu()); hr.copy_(c4.ref.cpu())
    hbest = env.pinned((c4.b1 - c4.b0, 3), torch.int32)
    L = xb.lib()
    h2d = d2h = 0
    for log2n, lo, cnt in c4.shards:
        h2d += cnt << (2 * log2n + 1)
        d2h += cnt << (2 * log2n + 1)
    h2d += hc.numel() + hr.numel()
    d2h += hbest.numel() * 4

    def host_call():
        for log2n, lo, cnt in c4.shards:
            off = lo << (2 * log2n + 1)
            rc = L.xDctNBatch(log2n, hsrc.data_ptr() + off, hdst.data_ptr() + off, cnt, log2n - 1, log2n + 6)
            assert rc == 0, xb.last_error()
        rc = L.xSatd8x8Search(hc.data_ptr(), hr.data_ptr(), c4.w + 64, c4.w, c4.h, 32, c4.b0, c4.b1, None, hbest.data_ptr())
        assert rc == 0, xb.last_error()

    sec = env.wall(host_call, 3, warm=1)
    e2e_ok = bool(torch.equal(hbest, c4.best.cpu()))
    h2d_all, d2h_all = env.sum_over_ranks(h2d), env.sum_over_ranks(d2h)
    cpu = None
    env.unbind()
    if rank == 0:
        cr = CpuRef()
        n_s = 1 << 22
        rate, _ = cr.satd_rate(cr.orc.residual(n_s * 64, 266, 0), host_threads(), 3)
        cpu = {"value": rate / (c4.nb * 4225), "unit": "frames/s", "cores": host_threads(), "kind": cr.kind,
               "sample": f"{n_s} precomputed differences through the reference satd8x8, expressed in frames of {c4.nb * 4225} candidates; "
                         "transforms and difference forming not counted (the reference has neither a search loop nor small transforms)"}
    env.host_barrier()
    if rank != 0:
        return None
    e = c4.entry(t["avg_ms"])
    return {
        "metric": "config4_4k_frames_per_s", "value": value, "unit": "frames/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": t["total_ms"] / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "int16 residuals / u8 pixels (int8 tensor cores for N>=8, packed int16 search)", "data": "synthetic",
        "config": {"workload": workload_name(args), "l2_policy": "one frame (33 MB of residuals + coefficients, 17 MB of pixels) is smaller than L2 by construction of "
                   "config 4; back-to-back steps therefore run L2-warm, as an encoder's frame loop would", "timing": "CUDA events on the launch stream, max over ranks",
                   "best_ms_per_step": t["best_ms"]},
        "roofline": dict(e["roofline"], bound="hbm", note="search-dominated: integer-ALU bound, G candidates/s reported"),
        "cpu_baseline": cpu,
        "e2e": {"value": 1.0 / sec, "unit": "frames/s", "h2d_bytes_per_step": int(h2d_all), "d2h_bytes_per_step": int(d2h_all),
                "api": "xDctNBatch x4 + xSatd8x8Search (host pointers, pinned)", "matches_device_path": e2e_ok},
        "gpu_launches": t["launches"] if c4.graph is None else int(args.steps * c4.kernels_per_frame), "clocks": t["clocks"], "secondary": [],
    }


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="config5", choices=sorted(METRICS))
    ap.add_argument("--frames", type=int, default=64, help="8K frames resident per GPU (config5)")
    ap.add_argument("--e2e-frames", type=int, default=0, help="frames per e2e step per rank (host buffers); 0 = 32")
    ap.add_argument("--ref-frames", type=int, default=8, help="frames per step of --impl reference")
    ap.add_argument("--cpu-frames", type=int, default=16, help="frames of the bounded cpu_baseline sample")
    ap.add_argument("--variant", default="auto", choices=["auto", "bfly", "imma"])
    ap.add_argument("--no-secondary", action="store_true", help="skip the secondary measurements")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        return run_reference(args)
    env = Env(args)
    line = {"config5": bench_config5, "config3": bench_config3, "config4": bench_config4}[args.workload](env)
    if env.rank == 0:
        print(json.dumps(line), flush=True)
    if env.world > 1:
        env.dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
