"""Pins the oracle restatement (oracle/fg_oracle.c) to the UNMODIFIED reference compiled into oracle/_ref/
(oracle/Makefile + oracle/ref_shim.cpp) on fresh seeded inputs — beyond the committed golden rows.
Skipped when oracle/_ref was never built (it is built wherever /root/reference exists and travels to the GPU box)."""
import numpy as np
import pytest

import mrp_b200
from mrp_b200 import synth
from oracle import binding as ob

CASES = [
    ("l10", synth.SynthSpec(n_rows=2500, seed=3), False),
    ("l10_book", synth.SynthSpec(n_rows=1200, seed=31), True),
    ("l10_cold", synth.SynthSpec(n_rows=1500, seed=4, warm_ticks=0, warm_trades=0, variable_cursors=True), True),
    ("l10_sparse_trades", synth.SynthSpec(n_rows=1500, seed=6, warm_trades=3, max_trades_per_row=1), False),
    ("l50", synth.SynthSpec(n_rows=900, levels=50, seed=5), True),
]


@pytest.mark.parametrize("name,spec,use_book", CASES, ids=[c[0] for c in CASES])
def test_restatement_is_bit_exact(name, spec, use_book):
    if not ob.ref_available(spec.levels):
        pytest.skip("oracle/_ref not built (needs /root/reference at build time)")
    batch = synth.to_numpy(synth.make_batch([spec]))
    hv = mrp_b200.HostView(batch)
    p = mrp_b200.default_params(depth_levels=spec.levels, emit_dirs=1, emit_counts=1)
    o = ob.run_oracle(hv, p, nthreads=4)
    r = ob.run_reference(hv, batch["trade_size"], use_book=use_book)
    assert np.array_equal(o["bid_dir"], r["bid_dir"]) and np.array_equal(o["ask_dir"], r["ask_dir"])
    for key in ("raw", "norm"):
        for c, fname in enumerate(mrp_b200.FEATURES):
            assert np.array_equal(o[key][c], r[key][c], equal_nan=True), (key, fname)
    if use_book:
        # the reference's FeatureEngine rebuilt the staged rolling volumes itself (feature_engine.cpp:95-110)
        assert np.array_equal(r["snap_buy"], hv.arrays["buy_vol"]) and np.array_equal(r["snap_sell"], hv.arrays["sell_vol"])


def test_multi_stream_batches_are_independent():
    if not ob.ref_available(10):
        pytest.skip("oracle/_ref not built")
    specs = [synth.SynthSpec(n_rows=n, seed=40 + i) for i, n in enumerate((1, 2, 333, 800))]
    batch = synth.to_numpy(synth.make_batch(specs))
    hv = mrp_b200.HostView(batch)
    o = ob.run_oracle(hv, nthreads=2)
    r = ob.run_reference(hv, batch["trade_size"], use_book=False, nthreads=3)
    assert np.array_equal(o["raw"], r["raw"], equal_nan=True) and np.array_equal(o["norm"], r["norm"], equal_nan=True)
    # and each stream equals the same stream run alone
    for i, sp in enumerate(specs):
        alone = ob.run_oracle(mrp_b200.HostView(synth.to_numpy(synth.make_batch([sp]))), nthreads=1)
        a, b = hv.arrays["row_offsets"][i], hv.arrays["row_offsets"][i + 1]
        assert np.array_equal(o["raw"][:, a:b], alone["raw"], equal_nan=True)
        assert np.array_equal(o["norm"][:, a:b], alone["norm"], equal_nan=True)
