#!/usr/bin/env python
"""Turns an .ncu-rep into the compact per-kernel summary committed under profiles/ (run here, no GPU needed)."""
import csv, io, subprocess, sys

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
idx = {h: i for i, h in enumerate(hdr)}
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "sm__cycles_elapsed.avg.per_second",
        "smsp__inst_executed.sum"]
seen = {}
for r in rows[2:]:
    seen.setdefault(r[idx["Kernel Name"]].split("(")[0], []).append(r)
print(f"# ncu summary of {rep} (last captured launch of each kernel; --set full --clock-control none)\n")
for name, rs in seen.items():
    r = rs[-1]
    print(f"## {name}  (launches captured: {len(rs)})")
    for w in want:
        if w in idx:
            print(f"- {w} = {r[idx[w]]} {units[idx[w]]}")
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + name.replace("void ", "").split("::")[-1].split("<")[0] + "$" if "<" not in name else "regex:" + name.replace("void ", "").split("::")[-1].split("<")[0],
                          "--launch-skip", str(len(rs) - 1), "--launch-count", "1"], capture_output=True, text=True).stdout
    srows = list(csv.reader(io.StringIO(src)))
    if len(srows) > 2:
        sh = srows[1]
        si = {h: i for i, h in enumerate(sh)}
        stalls = [h for h in sh if h.startswith("stall_") and "Not Issued" not in h]
        tot = {s: 0 for s in stalls}
        ins = []
        for q in srows[2:]:
            if len(q) < len(sh) or not q[si["# Samples"]].isdigit():
                continue
            for s in stalls:
                tot[s] += int(q[si[s]] or 0)
            ins.append((int(q[si["# Samples"]] or 0), q[si["Source"]]))
        T = sum(tot.values()) or 1
        print("- stall samples: " + ", ".join(f"{s[6:]} {100 * v / T:.1f}%" for s, v in sorted(tot.items(), key=lambda kv: -kv[1])[:7]))
        uniq = []
        for n, t in sorted(ins, key=lambda d: -d[0]):
            if (n, t) not in uniq:
                uniq.append((n, t))
        print("- hottest SASS: " + " | ".join(f"{100 * n / T:.1f}% {t.strip()[:48]}" for n, t in uniq[:6]))
    print()
