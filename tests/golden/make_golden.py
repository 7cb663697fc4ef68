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


if __name__ == "__main__":
    main()
