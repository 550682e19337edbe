#!/usr/bin/env python
"""Summarise an ncu report for profiles/: key raw metrics + instruction-count buckets.
    python tools/ncu_summary.py gpurun_out/prof.ncu-rep > profiles/<name>.txt"""
import collections, csv, io, math, subprocess, sys

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "launch__grid_size", "launch__block_size", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__cycles_elapsed.max",
        "smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio", "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio"]

def run(args):
    return subprocess.run(["ncu", "-i", sys.argv[1], *args], capture_output=True, text=True).stdout

rows = list(csv.reader(io.StringIO(run(["--page", "raw", "--csv"]))))
hdr = rows[0]
print(f"# ncu summary of {sys.argv[1]}")
for r in rows[2:]:
    name = r[hdr.index("Kernel Name")][:70] if "Kernel Name" in hdr else ""
    print(f"\n## kernel: {name}")
    for w in WANT:
        if w in hdr:
            print(f"{w:75s} {r[hdr.index(w)]:>18s} {rows[1][hdr.index(w)]}")
src = list(csv.reader(io.StringIO(run(["--page", "source", "--csv"]))))
try:
    h = src[1]; ie = h.index("Instructions Executed"); ss = h.index("# Samples")
    b = collections.Counter(); bs = collections.Counter(); bn = collections.Counter()
    tot = 0
    for r in src[2:]:
        if len(r) > ie and r[ie].isdigit():
            c = int(r[ie]); tot += c
            k = 0 if c == 0 else round(math.log10(c) * 2) / 2
            b[k] += c; bn[k] += 1; bs[k] += int(r[ss]) if r[ss].isdigit() else 0
    print(f"\n## SASS instructions by execution count (total {tot/1e9:.3f} G warp-instructions)")
    for k in sorted(b):
        if b[k]:
            print(f"executed ~10^{k:4.1f} times: {bn[k]:5d} SASS instrs, {b[k]/1e9:8.3f} G executed ({100*b[k]/tot:5.1f} %), {bs[k]} stall samples")
except Exception as e:
    print("source page unavailable:", e)
