#!/usr/bin/env python
"""Summarise an ncu report per CUDA source line: instructions executed and stall samples.

    python tools/ncu_lines.py gpurun_out/prof.ncu-rep k_book [top_n]
"""
import csv
import subprocess
import sys


def main():
    rep, kern = sys.argv[1], sys.argv[2]
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name", "regex:" + kern],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr = None
    lines = []
    fname = ""
    for r in rows:
        if r and r[0] == "File Path":
            fname = r[1].split("/")[-1]
        if r and r[0] == "Line No":
            hdr = r
            continue
        if hdr is None or not r or r[0] == "" or len(r) < 10:
            continue
        try:
            ln = int(r[0])
        except ValueError:
            continue
        i_inst = hdr.index("Instructions Executed")
        i_samp = hdr.index("# Samples")
        try:
            lines.append((fname, ln, r[1], int(r[i_inst]), int(r[i_samp])))
        except ValueError:
            pass
    tot_i = sum(l[3] for l in lines) or 1
    tot_s = sum(l[4] for l in lines) or 1
    print(f"{kern}: total warp-instructions {tot_i:,}  samples {tot_s:,}")
    key = 4 if "--by-samples" in sys.argv else 3
    for l in sorted(lines, key=lambda x: -x[key])[:top]:
        print(f"{l[0]}:{l[1]:4d} inst {100*l[3]/tot_i:5.1f}%  samp {100*l[4]/tot_s:5.1f}%  | {l[2].strip()[:110]}")


if __name__ == "__main__":
    main()
