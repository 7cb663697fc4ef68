#!/usr/bin/env python
"""Top SASS instructions of a kernel by stall samples, with the dominant stall reasons.   python tools/ncu_top_sass.py rep kernel [n]"""
import csv, subprocess, sys
rep, kern = sys.argv[1:3]; top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass", "--kernel-name", "regex:" + kern], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines())); hdr = None; recs = []
for r in rows:
    if r and r[0] == "Address": hdr = r; continue
    if hdr is None or len(r) < len(hdr): continue
    try: sm = int(r[hdr.index("# Samples")])
    except ValueError: continue
    recs.append((sm, r))
stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_")]
tot = sum(s for s, _ in recs) or 1
idx = {id(r): i for i, (_, r) in enumerate(recs)}
for sm, r in sorted(recs, key=lambda x: -x[0])[:top]:
    st = sorted(((int(r[i] or 0), hdr[i][6:]) for i in stall_cols), reverse=True)[:3]
    i = idx[id(r)]
    prev = recs[i - 1][1][hdr.index("Source")][:50] if i else ""
    print(f"{100*sm/tot:5.1f}%  {r[hdr.index('Source')][:70]:70s} | {' '.join(f'{n}:{v}' for v, n in st if v)} | prev: {prev}")
