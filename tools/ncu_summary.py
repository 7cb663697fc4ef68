#!/usr/bin/env python
"""Markdown table of the headline raw metrics of every kernel launch in an ncu report.   python tools/ncu_summary.py rep.ncu-rep"""
import csv, subprocess, sys

METRICS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
    "launch__shared_mem_per_block_dynamic", "smsp__inst_executed.sum", "lts__t_sector_hit_rate.pct",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
]
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines())); hdr = rows[0]; units = rows[1]; idx = {h: i for i, h in enumerate(hdr)}
ker = rows[2:]
names = [r[idx["Kernel Name"]].split("(")[0].replace("void ", "").replace("mrp::", "") for r in ker]
print("| metric | " + " | ".join(names) + " |")
print("|---|" + "---|" * len(names))
for m in METRICS:
    if m not in idx: continue
    print(f"| {m} [{units[idx[m]]}] | " + " | ".join(r[idx[m]] for r in ker) + " |")
