#!/usr/bin/env python
"""Prints the measurement table of DESIGN.md section 4 from a bench.py JSON line (one denominator: the line's own roofline.peak)."""
import json, sys
d = json.loads([l for l in open(sys.argv[1]) if l.startswith("{")][-1])
peak = d["roofline"]["peak"]
print(f"Source: `{sys.argv[1]}`; N = {d['n_gpus']}; HBM peak = {peak} GB/s ({d['roofline'].get('peak_source')}); clocks {d['clocks']}.\n")
print("| Kernel / entry | ms per launch | units/s (whole job) | algorithmic GB/s per GPU | frac of measured HBM |")
print("|---|---|---|---|---|")
r = d["roofline"]
print(f"| `dct32_imma` (headline, config 5, {r['blocks_per_launch']} blocks per launch) | {r['avg_launch_ms']:.3f} | {d['value']:.4g} blocks/s | {r['achieved']:.0f} | **{r['frac']:.3f}** |")
for s in d["secondary"]:
    rf = s["roofline"]
    ms = s.get("ms_per_launch", s.get("ms_per_frame", s.get("ms")))
    frac = f"{rf['frac']:.3f}" if rf.get("frac") is not None else "—"
    ach = f"{rf['achieved']:.0f}" if isinstance(rf.get("achieved"), (int, float)) and rf.get("unit") == "GB/s" else "—"
    extra = ""
    if "alu_pipe_bound" in s:
        extra = f" (integer-ALU floor {s['alu_pipe_bound']['frac']:.2f})"
    if s.get("clocks"):
        extra += f" [SM {s['clocks'].get('sm_mhz')} MHz, {','.join(s['clocks'].get('reasons') or []) or 'no throttle reason'}]"
    print(f"| `{s['metric']}` | {ms if ms is None else round(ms, 4)} | {s['value']:.4g} | {ach} | {frac}{extra} |")
e = d["e2e"]
print(f"\nEnd to end (`{e['api']}`, {e['sample']}): {e['value']:.4g} blocks/s = {e['roofline']['achieved']:.1f} GB/s each way of the {e['roofline']['peak']:.1f} the host link gives (frac {e['roofline']['frac']:.2f}); "
      f"pageable {e['pageable']['value']:.4g} ({e['pageable']['frac_of_pinned']:.2f} of pinned, {e['pageable']['host_copy_threads_per_rank']} copy threads); "
      f"registered once {e['registered']['value']:.4g} ({e['registered']['frac_of_pinned']:.2f}).")
c = d["cpu_baseline"]
if c:
    print(f"CPU baseline ({c['kind']}, {c['cores']} threads): {c['value']:.4g} blocks/s (single thread {c['single_thread_value']:.4g}, -O3 -march=native {c['o3_march_native_value']}); bit-exact on sample: {c['gpu_output_bit_exact_on_sample']}.")
