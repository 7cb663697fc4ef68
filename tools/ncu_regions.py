#!/usr/bin/env python
"""Instruction / stall-sample share of line ranges of one source file in an ncu report.
    python tools/ncu_regions.py rep.ncu-rep kernel file.cu name:lo-hi [name:lo-hi ...]   (lines of inlined headers go to 'inlined')"""
import csv, subprocess, sys

rep, kern, fname = sys.argv[1:4]
regions = []
for a in sys.argv[4:]:
    n, r = a.split(":")
    lo, hi = r.split("-")
    regions.append((n, int(lo), int(hi)))
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name", "regex:" + kern],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, cur = None, ""
acc = {}
for r in rows:
    if r and r[0] == "File Path":
        cur = r[1].split("/")[-1]
    if r and r[0] == "Line No":
        hdr = r
        continue
    if hdr is None or not r or len(r) < 10:
        continue
    try:
        ln = int(r[0]); inst = int(r[hdr.index("Instructions Executed")]); samp = int(r[hdr.index("# Samples")])
    except ValueError:
        continue
    key = "inlined:" + cur
    if cur == fname:
        key = "other"
        for n, lo, hi in regions:
            if lo <= ln <= hi:
                key = n
    a = acc.setdefault(key, [0, 0]); a[0] += inst; a[1] += samp
ti = sum(a[0] for a in acc.values()) or 1; ts = sum(a[1] for a in acc.values()) or 1
print(f"total warp-instructions {ti:,}, samples {ts:,}")
for k, (i, s) in sorted(acc.items(), key=lambda kv: -kv[1][0]):
    print(f"{k:40s} inst {i:>13,} {100*i/ti:5.1f}%   samples {100*s/ts:5.1f}%")
