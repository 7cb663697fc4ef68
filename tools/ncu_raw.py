#!/usr/bin/env python
"""Headline raw metrics + stall breakdown of the first kernel in an ncu report.   python tools/ncu_raw.py rep [row]"""
import csv, subprocess, sys
rep = sys.argv[1]; row = int(sys.argv[2]) if len(sys.argv) > 2 else 0
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines())); hdr = rows[0]; idx = {h: i for i, h in enumerate(hdr)}; r = rows[2 + row]
print(r[idx["Kernel Name"]][:60])
for w in ["gpu__time_duration.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
          "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
          "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
          "l1tex__throughput.avg.pct_of_peak_sustained_active", "lts__throughput.avg.pct_of_peak_sustained_elapsed", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
          "smsp__inst_executed_op_local_ld.sum", "smsp__inst_executed_op_local_st.sum"]:
    print(" ", w, r[idx[w]] if w in idx else "NA")
for h in hdr:
    if "issue_stalled" in h and "per_issue_active" in h and float(r[idx[h]] or 0) > 0.15:
        print("    stall", h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""), r[idx[h]])
