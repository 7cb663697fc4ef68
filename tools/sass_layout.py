#!/usr/bin/env python
"""Static code layout of a kernel: instruction index of every barrier / loop-closing branch, with the source line.
   python tools/sass_layout.py <object-or-so> <cubin-name-substring> <function-substring>"""
import re, subprocess, sys, tempfile, os, glob
obj, cub, fn = sys.argv[1:4]
d = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=d, capture_output=True)
f = [x for x in glob.glob(d + "/*.cubin") if cub in os.path.basename(x)][0]
txt = subprocess.run(["nvdisasm", "-g", "-c", f], capture_output=True, text=True).stdout
cur = None; name = None; idx = 0; labels = {}; rows = []
for ln in txt.splitlines():
    m = re.match(r"\s*\.text\.(\S+):", ln)
    if m: name = m.group(1); idx = 0; continue
    if name is None or fn not in name: continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m: cur = (os.path.basename(m.group(1)), int(m.group(2))); continue
    m = re.match(r"\s*(\.L_x_\d+):", ln)
    if m: labels[m.group(1)] = idx; continue
    if re.match(r"\s+/\*[0-9a-f]{4,5}\*/", ln):
        rows.append((idx, cur, ln.strip())); idx += 1
print("instructions:", idx, "=", idx * 16 / 1024, "KB")
for i, c, l in rows:
    m = re.search(r"BRA\S*\s+(?:\w+,\s*)?`\((\.L_x_\d+)\)", l)
    if m and labels.get(m.group(1), 1 << 30) < i:
        print(f"{i:5d} line {c[1] if c else 0:4d}  loop back to {labels[m.group(1)]:5d}  (body {i - labels[m.group(1)]} instr)")
    elif "BAR." in l or "EXIT" in l:
        print(f"{i:5d} line {c[1] if c else 0:4d}  {l[9:70]}")
