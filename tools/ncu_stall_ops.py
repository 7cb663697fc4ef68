#!/usr/bin/env python
"""Per-opcode totals of one stall reason (default stall_wait) for a kernel.   python tools/ncu_stall_ops.py rep kernel [stall_wait]"""
import csv, subprocess, sys, collections
rep, kern = sys.argv[1:3]; col = sys.argv[3] if len(sys.argv) > 3 else "stall_wait"
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass", "--kernel-name", "regex:" + kern], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines())); hdr = None
acc = collections.Counter(); cnt = collections.Counter(); tot = 0; allsamp = 0
for r in rows:
    if r and r[0] == "Address": hdr = r; continue
    if hdr is None or len(r) < len(hdr): continue
    try: v = int(r[hdr.index(col)]); n = int(r[hdr.index("Instructions Executed")]); sm = int(r[hdr.index("# Samples")])
    except ValueError: continue
    src = r[hdr.index("Source")].split()
    op = next((t for t in src if not t.startswith("@")), "?").split(".")[0]
    acc[op] += v; cnt[op] += n; tot += v; allsamp += sm
print(f"{col}: {tot} of {allsamp} samples")
for op, v in acc.most_common(18): print(f"  {op:10s} {100*v/max(tot,1):5.1f}%  execs {cnt[op]:,}  per-exec {v/max(cnt[op],1)*1e3:.3f}e-3")
