#!/usr/bin/env python
"""Run one build of the reference's features_to_csv pipeline (oracle/_ref/libmrp_refpipe.so = unmodified reference,
libmrp_gpupipe.so = the same with the INTEGRATION.md §3 patch: GpuFeatureProcessor -> C ABI -> CUDA) on the seeded synthetic MBO day.

    python tools/run_pipe.py <lib.so> <project_root> [--date 20250505] [--events 400000] [--seed 24301] [--interval-ns 500000000]
Prints one JSON line: {"rc", "seconds", "dir"}.  TEST / BENCH INFRASTRUCTURE (it executes oracle/_ref)."""
import argparse
import ctypes
import json
import os
import time

ap = argparse.ArgumentParser()
ap.add_argument("lib")
ap.add_argument("root")
ap.add_argument("--date", default="20250505")
ap.add_argument("--events", type=int, default=400000)
ap.add_argument("--seed", type=int, default=24301)
ap.add_argument("--interval-ns", type=int, default=500_000_000)
a = ap.parse_args()
os.environ["MRP_REF_PROJECT_ROOT"] = a.root
os.environ["MRP_SYN_EVENTS"] = str(a.events)
os.environ["MRP_SYN_SEED"] = str(a.seed)
os.makedirs(a.root, exist_ok=True)
lib = ctypes.CDLL(os.path.abspath(a.lib))
lib.mrp_pipe_run.argtypes = [ctypes.c_char_p, ctypes.c_char_p, ctypes.c_char_p, ctypes.c_uint64]
t0 = time.perf_counter()
rc = lib.mrp_pipe_run(a.date.encode(), b"SPY", b"ES", a.interval_ns)
sec = time.perf_counter() - t0
secs = "%f" % (a.interval_ns / 1e9)
print(json.dumps({"rc": rc, "seconds": sec, "dir": os.path.join(a.root, "data", f"output_Snapshot{secs}_Window30000_Events1500", a.date)}))
