#!/usr/bin/env python
"""Instruction / stall-sample share per source-line RANGE of one kernel.
    python tools/ncu_regions.py rep kernel file.cu 300-330 331-352 ...
"""
import csv, subprocess, sys
rep, kern, fn = sys.argv[1:4]
ranges = [tuple(map(int, a.split("-"))) for a in sys.argv[4:]]
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name", "regex:" + kern], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = None; cur = ""; data = []
for r in rows:
    if r and r[0] == "File Path": cur = r[1].split("/")[-1]
    if r and r[0] == "Line No": hdr = r; continue
    if hdr is None or not r or r[0] == "" or len(r) < 10: continue
    try: ln = int(r[0])
    except ValueError: continue
    try: data.append((cur, ln, int(r[hdr.index("Instructions Executed")]), int(r[hdr.index("# Samples")])))
    except ValueError: pass
ti = sum(d[2] for d in data); ts = sum(d[3] for d in data)
print(f"total warp-inst {ti:,} samples {ts:,}")
acc = 0
for lo, hi in ranges:
    i = sum(d[2] for d in data if d[0] == fn and lo <= d[1] <= hi); s = sum(d[3] for d in data if d[0] == fn and lo <= d[1] <= hi)
    acc += i
    print(f"{fn}:{lo}-{hi}: inst {100*i/ti:5.1f}% ({i:,})  samples {100*s/ts:5.1f}%")
oth = sum(d[2] for d in data if d[0] != fn); so = sum(d[3] for d in data if d[0] != fn)
print(f"other files: inst {100*oth/ti:5.1f}% samples {100*so/ts:5.1f}%")
