#!/usr/bin/env python
"""Generate tests/golden/*.npz from the UNMODIFIED reference compiled into oracle/_ref/ (oracle/Makefile).

Run in the build container, where /root/reference exists:
    python tests/golden/make_golden.py

The reference ships no golden vectors for this path, so these fixtures are outputs of the
reference's own FeatureEngine + FeatureProcessor + FeatureNormalizer (driven by oracle/ref_shim.cpp through
the reference's real OrderBookManager, use_book=1) on seeded synthetic streams.  Each fixture holds the
generator spec (the inputs are a pure function of it, micro-regime-propogation_b200/synth.py), a checksum of
the generated inputs, and the reference's raw / normalised / direction outputs for the listed rows.
"""
from __future__ import annotations

import dataclasses
import hashlib
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

import mrp_b200  # noqa: E402
from mrp_b200 import synth  # noqa: E402
from oracle import binding as ob  # noqa: E402

# name -> (spec, rows kept)   rows=None keeps every row
CASES = {
    # warm windows, L = 10, the shape of BASELINE.json configs[0]/[1]
    "l10_warm": (synth.SynthSpec(n_rows=700, seed=101), None),
    # cold start: windows grow from empty, first rows have no trades (0/0 = NaN), ragged cursors
    "l10_cold_ragged": (synth.SynthSpec(n_rows=500, seed=102, warm_ticks=0, warm_trades=0, variable_cursors=True), None),
    # deep book (configs[3]); compiled from the DEPTH_LEVELS=50 header patch
    "l50_warm": (synth.SynthSpec(n_rows=300, levels=50, seed=103), None),
    # longer than WINDOW_SIZE = 30000: the normaliser evicts; keep the rows around the first eviction and the tail
    "l10_evict": (synth.SynthSpec(n_rows=31200, seed=104, max_trades_per_row=2),
                  np.r_[0:40, 29960:30060, 31100:31200]),
}


def input_digest(batch) -> str:
    h = hashlib.sha256()
    for k in sorted(batch):
        if k == "trade_size":
            continue
        h.update(k.encode())
        h.update(np.ascontiguousarray(batch[k]).tobytes())
    return h.hexdigest()


def main():
    for name, (spec, rows) in CASES.items():
        batch = synth.to_numpy(synth.make_batch([spec]))
        hv = mrp_b200.HostView(batch)
        if not ob.ref_available(spec.levels):
            raise SystemExit("oracle/_ref is not built: run `make -C oracle` where /root/reference exists")
        ref = ob.run_reference(hv, batch["trade_size"], use_book=True)
        # the reference's own FeatureEngine must reproduce the staged rolling volumes exactly
        assert np.array_equal(ref["snap_buy"], hv.arrays["buy_vol"]) and np.array_equal(ref["snap_sell"], hv.arrays["sell_vol"])
        idx = np.arange(spec.n_rows) if rows is None else np.asarray(rows)
        np.savez_compressed(
            os.path.join(HERE, name + ".npz"),
            spec=json.dumps(dataclasses.asdict(spec)), digest=input_digest(batch), rows=idx.astype(np.int64),
            raw=ref["raw"][:, idx], norm=ref["norm"][:, idx], bid_dir=ref["bid_dir"][:, idx], ask_dir=ref["ask_dir"][:, idx])
        print(name, spec.n_rows, "rows ->", len(idx), "kept;", os.path.getsize(os.path.join(HERE, name + ".npz")) >> 10, "KiB")


def main_csv():
    """Golden CSV files: the reference's own CsvWriter (features_to_csv.cpp) formatting the reference's own features."""
    import shutil
    import tempfile

    spec = synth.SynthSpec(n_rows=64, seed=105, warm_trades=0, max_trades_per_row=1, warm_ticks=3)
    batch = synth.to_numpy(synth.make_batch([spec]))
    hv = mrp_b200.HostView(batch)
    ref = ob.run_reference(hv, batch["trade_size"], use_book=True)
    raw, norm = ref["raw"].copy(), ref["norm"].copy()
    # a few special values the formatter must spell like the reference's stream does
    raw[7, 5], raw[7, 6], raw[8, 5], raw[9, 5], raw[10, 5] = np.inf, -np.inf, -0.0, 4.9e-324, 1.7976931348623157e308
    raw[11, 6], raw[12, 6] = np.nan, np.copysign(np.nan, -1.0)
    norm[3, 9], norm[4, 9] = np.copysign(np.nan, -1.0), 1e-300
    grid = (1_746_453_700_000_000_000 + np.arange(spec.n_rows, dtype=np.uint64) * np.uint64(500_000_000)).astype(np.uint64)
    out = os.path.join(HERE, "csv")
    os.makedirs(out, exist_ok=True)
    tmp = tempfile.mkdtemp()
    try:
        d = ob.ref_write_csv(tmp, "base_SYN", 500_000_000, "20250505", "SYN", grid, batch["ts_ns"], raw, norm)
        for f in ("base_SYN_raw.csv", "base_SYN_norm.csv"):
            shutil.copy(os.path.join(d, f), os.path.join(out, f))
        np.savez_compressed(os.path.join(out, "values.npz"), raw=raw, norm=norm, grid=grid, row_ts=batch["ts_ns"],
                            rel_dir=os.path.relpath(d, tmp))
        print("csv golden:", os.path.relpath(d, tmp), os.path.getsize(os.path.join(out, "base_SYN_raw.csv")), "bytes")
    finally:
        shutil.rmtree(tmp)


if __name__ == "__main__":
    main()
    main_csv()
