#!/usr/bin/env python
"""validation/validate.py — diff the GPU feature stage against the reference's own CSV files.

For one synthetic instrument-day (the shapes of BASELINE.json configs[0]/[1]) this harness
  1. runs the reference: the UNMODIFIED FeatureEngine + FeatureProcessor + FeatureNormalizer compiled into
     oracle/_ref/libmrp_ref.so, formatted by the reference's OWN CsvWriter (oracle/_ref/libmrp_refcsv.so) -> <out>/reference/...
     (falls back to the C restatement oracle/libmrp_oracle.so + the product's CsvWriter when oracle/_ref is absent, and says so);
  2. runs the product: microregime::GpuFeatureProcessor (C++20 adaptor) staging ticks, trades and snapshots in event
     order, the CUDA pass behind the C ABI, microregime::CsvWriter -> <out>/gpu/...;
  3. diffs the four CSV files column by column: header and timestamps identical, NaN / inf cells identical,
     integer-valued columns bit-exact, raw floating-point columns within 1e-9 relative (tiny absolute floor for cells
     that cross zero), z-scores within 1e-9 * max(1,|z|) plus the conditioning allowance of SURVEY.md §7 H1.
Exit status 0 = every column within tolerance.  It is a checker: nothing here is on the product path.

    python validation/validate.py --rows 46600 --out /tmp/mrp_validate
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import mrp_b200  # noqa: E402
from mrp_b200 import hostlib, synth  # noqa: E402
from oracle import binding as ob  # noqa: E402

INT_COLUMNS = ("midprice", "market_depth")   # quarter-tick / integer valued: bit-exact
RAW_RTOL, RAW_ATOL_SCALE, Z_TOL = 1e-9, 1e-13, 1e-9


def read_csv(path):
    import pandas as pd

    df = pd.read_csv(path, dtype={"instrument": str}, float_precision="round_trip")
    return df


def diff_file(ref_path, gpu_path, kind, raw_ref=None, window=30000):
    a, b = read_csv(ref_path), read_csv(gpu_path)
    report, ok = [], True
    if list(a.columns) != list(b.columns) or len(a) != len(b):
        return [{"column": "<layout>", "ok": False, "note": "header or row count differs"}], False
    if not (a["timestamp_ns"].equals(b["timestamp_ns"]) and a["instrument"].equals(b["instrument"])):
        return [{"column": "<keys>", "ok": False, "note": "timestamp_ns / instrument differ"}], False
    names = list(mrp_b200.FEATURES)
    R = np.stack([a[n].to_numpy(np.float64) for n in names])
    G = np.stack([b[n].to_numpy(np.float64) for n in names])
    if kind == "norm":
        # conditioning-aware: against the exactly summed definition AND against the reference itself,
        # allowing the reference's own drift cell by cell (tests/helpers.py:norm_check)
        from helpers import norm_check

        X = np.stack([raw_ref[n].to_numpy(np.float64) for n in names])
        wa, wb, drift, _ = norm_check(G, X, R, np.array([0, R.shape[1]]), window, Z_TOL)
    for c, name in enumerate(names):
        r, g = R[c], G[c]
        pat = np.array_equal(np.isnan(r), np.isnan(g)) and np.array_equal(np.isinf(r), np.isinf(g)) and \
            np.array_equal(r[np.isinf(r)], g[np.isinf(r)])
        fin = np.isfinite(r) & np.isfinite(g)
        d = np.abs(g[fin] - r[fin])
        entry = {"column": name, "nonfinite_cells_match": bool(pat), "bit_exact": bool(np.array_equal(r, g, equal_nan=True)),
                 "max_abs": float(d.max()) if d.size else 0.0}
        if kind == "raw":
            scale = float(np.abs(r[fin]).max()) if fin.any() else 0.0
            tol = RAW_RTOL * np.abs(r[fin]) + RAW_ATOL_SCALE * max(scale, 1e-300)
            entry["worst_err_over_tol"] = float((d / tol).max()) if d.size else 0.0
            good = pat and entry["worst_err_over_tol"] <= 1.0 and (entry["bit_exact"] or name not in INT_COLUMNS)
        else:
            entry["worst_err_over_tol"] = float(max(wa[c], wb[c]))
            entry["reference_drift_vs_exact"] = float(drift[c])
            good = pat and wa[c] <= 1.0 and wb[c] <= 1.0
        entry["ok"] = bool(good)
        ok &= bool(good)
        report.append(entry)
    return report, ok


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", type=int, default=46600, help="snapshots of the instrument-day (46 600 = one 6.5 h session at 500 ms)")
    ap.add_argument("--seed", type=int, default=0x5EED)
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "validate"))
    ap.add_argument("--bulk", action="store_true", help="use the threaded CSV column path instead of the per-row sink")
    ap.add_argument("--cold", action="store_true", help="cold start: windows grow from empty, ragged cursors")
    args = ap.parse_args()

    spec = synth.SynthSpec(n_rows=args.rows, seed=args.seed, variable_cursors=args.cold,
                           warm_ticks=0 if args.cold else 2000, warm_trades=0 if args.cold else 1600)
    batch = synth.to_numpy(synth.make_batch([spec]))
    hv = mrp_b200.HostView(batch)
    interval, date, first_grid = 500_000_000, "20250505", 1_746_451_900_000_000_000
    grid = (np.uint64(first_grid) + np.arange(args.rows, dtype=np.uint64) * np.uint64(interval)).astype(np.uint64)

    # ---- reference side
    t0 = time.perf_counter()
    if ob.ref_available(10) and ob.ref_csv_available():
        ref = ob.run_reference(hv, batch["trade_size"], use_book=True)
        ref_dir = ob.ref_write_csv(os.path.join(args.out, "reference"), "future_ES", interval, date, "ES", grid, batch["ts_ns"],
                                   ref["raw"], ref["norm"])
        ref_kind = "unmodified reference (oracle/_ref) + the reference's own CsvWriter"
    else:
        ref = ob.run_oracle(hv, nthreads=os.cpu_count() or 1)
        ref_dir = hostlib.write_csv(os.path.join(args.out, "reference"), "future_ES", interval, date, "ES", grid, ref["raw"],
                                    ref["norm"], row_ts=batch["ts_ns"])
        ref_kind = "C restatement (oracle/libmrp_oracle.so); oracle/_ref not present on this machine"
    t_ref = time.perf_counter() - t0

    # ---- product side: C++ adaptor -> C ABI -> CUDA -> CsvWriter
    t0 = time.perf_counter()
    gpu_dir = hostlib.features_to_csv(hv, os.path.join(args.out, "gpu"), "future_ES", date, "ES", first_grid, interval, bulk=args.bulk)
    t_gpu = time.perf_counter() - t0

    raw_ref = read_csv(os.path.join(ref_dir, "future_ES_raw.csv"))
    rep_raw, ok_raw = diff_file(os.path.join(ref_dir, "future_ES_raw.csv"), os.path.join(gpu_dir, "future_ES_raw.csv"), "raw")
    rep_norm, ok_norm = diff_file(os.path.join(ref_dir, "future_ES_norm.csv"), os.path.join(gpu_dir, "future_ES_norm.csv"), "norm", raw_ref)
    for kind, rep in (("raw", rep_raw), ("norm", rep_norm)):
        print(f"--- {kind}: column | bit-exact | max |d| | err/tol | ok")
        for e in rep:
            print(f"{e['column']:26s} {str(e.get('bit_exact', '')):5s} {e.get('max_abs', 0):10.3e} {e.get('worst_err_over_tol', 0):9.3g} {'ok' if e['ok'] else 'FAIL'}")
    summary = {"rows": args.rows, "reference": ref_kind, "reference_seconds": t_ref, "gpu_seconds_incl_csv": t_gpu,
               "raw_ok": ok_raw, "norm_ok": ok_norm, "raw": rep_raw, "norm": rep_norm,
               "bit_exact_raw_columns": [e["column"] for e in rep_raw if e.get("bit_exact")]}
    os.makedirs(args.out, exist_ok=True)
    with open(os.path.join(args.out, "report.json"), "w") as f:
        json.dump(summary, f, indent=1)
    print(f"reference: {ref_kind}: {t_ref:.2f} s;  gpu path incl. staging + CSV: {t_gpu:.2f} s;  "
          f"raw {'OK' if ok_raw else 'FAIL'}, norm {'OK' if ok_norm else 'FAIL'}")
    return 0 if (ok_raw and ok_norm) else 1


if __name__ == "__main__":
    sys.exit(main())
