#!/usr/bin/env python
"""Host-side CSV emitter: rows/s of microregime::CsvWriter (per-row sink path and threaded bulk path) beside
the reference's own CsvWriter compiled unmodified (oracle/_ref/libmrp_refcsv.so; present only where /root/reference was).
    python tools/csv_bench.py [rows]
Output: one JSON line.  The three writers produce byte-identical files (tests/test_host_csv.py)."""
import json, os, sys, tempfile, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import mrp_b200  # noqa: E402
from mrp_b200 import hostlib  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 400_000
rng = np.random.default_rng(7)
raw = rng.standard_normal((19, n)) * 1e3
norm = rng.standard_normal((19, n))
grid = (np.arange(n, dtype=np.uint64) * 500_000_000) + np.uint64(1_700_000_000_000_000_000)
out = {"rows": n, "bytes_per_row_pair": None, "threads": os.cpu_count()}
with tempfile.TemporaryDirectory(dir="/dev/shm" if os.path.isdir("/dev/shm") else None) as d:   # memory-backed: measure the formatter, not the disk
    for name, kw in (("adaptor_per_row", dict(bulk=False)), ("adaptor_bulk_threaded", dict(bulk=True))):
        dt = 1e30
        for rep in range(3):   # best of three: the first call pays the page faults of fresh buffers
            t0 = time.time()
            p = hostlib.write_csv(os.path.join(d, name + str(rep)), "base", 500_000_000, "20250505", "SYN", grid, raw, norm, row_ts=grid, **kw)
            dt = min(dt, time.time() - t0)
        sz = sum(os.path.getsize(os.path.join(p, f)) for f in os.listdir(p))
        out[name] = {"rows_per_s": n / dt, "seconds": dt}
        out["bytes_per_row_pair"] = sz / n
    try:
        from oracle import binding as ob
        if os.path.exists(ob.REF_CSV_SO):
            dt = 1e30
            for rep in range(2):
                t0 = time.time()
                ob.ref_write_csv(os.path.join(d, "ref" + str(rep)), "base", 500_000_000, "20250505", "SYN", grid, grid, raw, norm)
                dt = min(dt, time.time() - t0)
            out["reference_csvwriter"] = {"rows_per_s": n / dt, "seconds": dt}
    except Exception as e:  # the compiled reference does not travel to every box
        out["reference_csvwriter"] = {"unavailable": str(e)[:80]}
print(json.dumps(out))
