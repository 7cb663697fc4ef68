"""Experiment builds with -DMRP_WIN_TIMERS: per-role phase timers of k_windows (cycles per chunk).  python tools/win_timers.py [rows]"""
import ctypes, sys
import torch
import mrp_b200
from mrp_b200 import synth
rows = int(sys.argv[1]) if len(sys.argv) > 1 else 8_000_000
dev = torch.device("cuda:0")
b = synth.make_batch([synth.SynthSpec(n_rows=rows, seed=1)], device=dev)
eng = mrp_b200.Engine(mrp_b200.default_params(), device=0)
eng.bind_device(b)
lib = ctypes.CDLL(mrp_b200.LIB_PATH)
buf = (ctypes.c_ulonglong * 36)()
for _ in range(3):
    eng.run()
torch.cuda.synchronize()
lib.mrp_fg_debug_win_timers(buf)
eng.run()
torch.cuda.synchronize()
lib.mrp_fg_debug_win_timers(buf)
names = {0: ["top->wait", "mbar_wait", "load+math", "totals+scan", "bar1", "cross", "bar3", "stores", "chunk barrier"],
         1: ["top", "gather", "finish", "plan(order0)", "chunk barrier", "plan(order1)"], 2: ["top", "compute", "bar3", "stores", "chunk barrier"]}
for r, nm in names.items():
    it = buf[r * 12 + 10] or 1
    vals = [buf[r * 12 + i] / it for i in range(len(nm))]
    print(["producer", "consumer", "trade"][r], f"iters={it}", " ".join(f"{n}={v:.0f}" for n, v in zip(nm, vals)), f"sum={sum(vals):.0f}")
