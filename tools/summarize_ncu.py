#!/usr/bin/env python
"""Summarise an .ncu-rep (one kernel capture, --set full --import-source on) into a text file for profiles/.

    python tools/summarize_ncu.py gpurun_out/prof.ncu-rep profiles/r1_norm3.txt [steps_per_launch]
"""
import csv, io, re, subprocess, sys
from collections import defaultdict

rep, out = sys.argv[1], sys.argv[2]
steps = float(sys.argv[3]) if len(sys.argv) > 3 else None
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
h, units, vals = rows[0], rows[1], rows[2]
want = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum.per_second",
        "dram__bytes_write.sum.per_second", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.per_cycle_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__grid_size",
        "launch__block_size", "launch__shared_mem_per_block_dynamic", "smsp__inst_executed.sum",
        "sm__cycles_elapsed.max", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active"]
lines = [f"# ncu summary of {rep}", ""]
d = {n: (v, u) for n, v, u in zip(h, vals, units)}
for w in want:
    if w in d:
        lines.append(f"{w:70s} {d[w][0]} {d[w][1]}")
try:
    rd = float(d["dram__bytes_read.sum"][0].replace(",", "")); wr = float(d["dram__bytes_write.sum"][0].replace(",", ""))
    mult = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1}
    tot = rd * mult[d["dram__bytes_read.sum"][1]] + wr * mult[d["dram__bytes_write.sum"][1]]
    lines.append(f"{'dram traffic per launch (read+write)':70s} {tot/1e9:.3f} GB")
    if steps:
        lines.append(f"{'dram traffic per step':70s} {tot/steps:.2f} B  ({steps:.3g} steps per launch)")
except Exception as e:
    lines.append(f"traffic: {e}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
if len(rows) > 2:
    h = rows[1]; si = h.index("Source"); ei = h.index("Instructions Executed")
    stall_cols = [i for i, n in enumerate(h) if n.startswith("stall_") and "Not Issued" not in n]
    ops, stalls, tot = defaultdict(float), defaultdict(float), 0.0
    for r in rows[2:]:
        if len(r) <= ei: continue
        if r[ei] == "Instructions Executed": break  # a second kernel's section: the summary is of the first
        m = re.match(r"\s*(@!?U?P\w+\s+)?([A-Z0-9_]+)", r[si])
        if not m: continue
        n = float(r[ei] or 0); ops[m.group(2)] += n; tot += n
        for i in stall_cols: stalls[h[i]] += float(r[i] or 0)
    lines += ["", f"warp instructions executed: {tot:.4g}" + (f"  = {tot*32/steps:.1f} thread-instructions per step" if steps else "")]
    lines.append("instruction mix (share of executed warp instructions" + (", per step" if steps else "") + "):")
    for k, v in sorted(ops.items(), key=lambda kv: -kv[1])[:24]:
        lines.append(f"  {k:10s} {100*v/tot:5.1f} %" + (f"  {v*32/steps:6.1f}" if steps else ""))
    ts = sum(stalls.values()) or 1
    lines.append("warp stall sampling (all samples):")
    for k, v in sorted(stalls.items(), key=lambda kv: -kv[1])[:10]:
        lines.append(f"  {k:28s} {100*v/ts:5.1f} %")
open(out, "w").write("\n".join(lines) + "\n")
print("\n".join(lines[:40]))
