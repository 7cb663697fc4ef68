#!/usr/bin/env python
"""Opcode histogram (warp-instructions and stall samples) of one kernel in an ncu report.
    python tools/ncu_opcodes.py rep kernel rows [top]"""
import csv, subprocess, sys
from collections import Counter
rep, kern, nrows = sys.argv[1], sys.argv[2], float(sys.argv[3])
top = int(sys.argv[4]) if len(sys.argv) > 4 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name", "regex:" + kern], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = None; cnt = Counter(); samp = Counter()
for r in rows:
    if r and r[0] == 'Line No': hdr = r; continue
    if hdr is None or len(r) < 8 or r[0] != '' or r[2] in ('...', '-'): continue
    try: n = int(r[hdr.index('Instructions Executed')]); st = int(r[hdr.index('# Samples')])
    except ValueError: continue
    t = r[3].split()
    op = t[1] if t[0].startswith('@') else t[0]
    cnt[op] += n; samp[op] += st
tot = sum(cnt.values()); ts = sum(samp.values()) or 1
print(f"total warp-inst {tot:,}; thread-inst per row {tot*32/nrows:.0f}")
for k, v in cnt.most_common(top):
    print(f"{k:22s} {100*v/tot:5.1f}%  per-row {v*32/nrows:7.1f}  samples {100*samp[k]/ts:5.1f}%")
