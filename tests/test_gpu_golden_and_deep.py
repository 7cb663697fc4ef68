"""GPU against the committed golden outputs of the UNMODIFIED reference (tests/golden/*.npz), and the deep-book
configuration (L = 50, BASELINE.json configs[3]) against the oracle / the L = 50 build of the reference."""
import numpy as np
import pytest

import mrp_b200
from mrp_b200 import synth
from oracle import binding as ob

from helpers import golden_cases, load_golden, norm_check, norm_errors, raw_errors, run_gpu
from test_gpu_parity import _check

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", golden_cases())
def test_gpu_matches_reference_golden(name):
    spec, batch, gold = load_golden(name)
    p = mrp_b200.default_params(depth_levels=spec.levels, emit_dirs=1, emit_counts=1)
    hv, got = run_gpu(batch, p)
    rows = gold["rows"]
    assert np.array_equal(got["bid_dir"][:, rows], gold["bid_dir"]) and np.array_equal(got["ask_dir"][:, rows], gold["ask_dir"])
    for col in ("midprice", "market_depth"):
        c = mrp_b200.FEATURES.index(col)
        assert np.array_equal(got["raw"][c, rows], gold["raw"][c]), col
    re = raw_errors(got["raw"][:, rows], gold["raw"])
    assert (re <= 1.0).all(), re
    # z-scores: conditioning-aware against the exactly summed definition and the reference's own drift (helpers.norm_check);
    # the fixture stores a row subset, so the history comes from the oracle (bit-exact against the reference, tests/test_oracle_vs_reference.py)
    ref = ob.run_oracle(hv, p)
    assert np.array_equal(ref["norm"][:, rows], gold["norm"], equal_nan=True)
    wa, wb, drift, frac = norm_check(got["norm"], ref["raw"], ref["norm"], hv.arrays["row_offsets"], p.norm_window, 1e-9)
    assert (wa <= 1.0).all() and (wb <= 1.0).all(), (wa, wb)
    ze = norm_errors(got["norm"][:, rows], gold["norm"])
    print("golden %s: plain max|dz|/max(1,|z|) vs the reference %.3g (reference drift %.3g)" % (name, np.nanmax(ze), drift.max()))


def test_deep_book_50_levels():
    p = mrp_b200.default_params(depth_levels=50, emit_dirs=1, emit_counts=1)
    batch = synth.to_numpy(synth.make_batch([synth.SynthSpec(n_rows=5000, levels=50, seed=51, variable_cursors=True)]))
    got, ref = _check(batch, params=p)
    if ob.ref_available(50):   # the reference itself, compiled with DEPTH_LEVELS = 50
        r = ob.run_reference(mrp_b200.HostView(batch), batch["trade_size"], use_book=False)
        assert np.array_equal(r["raw"], ref["raw"], equal_nan=True)


@pytest.mark.parametrize("levels", [2, 3, 4, 7, 16])
def test_odd_depths(levels):
    p = mrp_b200.default_params(depth_levels=levels, stress_levels=min(5, levels), emit_dirs=1, emit_counts=1)
    # with 2-4 levels the generator's occasional empty deep levels empty a whole side: log(0) = -inf, NaN poison — compared cell for cell
    _check(synth.to_numpy(synth.make_batch([synth.SynthSpec(n_rows=1500, levels=levels, seed=60 + levels)])), params=p, min_frac=0.2)


def test_other_window_parameters():
    p = mrp_b200.default_params(mid_window=900, dir_window=700, spread_divisor=700, ewm_span=300, norm_window=1200, return_lag=7,
                                emit_dirs=1, emit_counts=1)
    _check(synth.to_numpy(synth.make_batch([synth.SynthSpec(n_rows=4000, seed=71, dir_window=700, variable_cursors=True)])), params=p)
    p = mrp_b200.default_params(mid_window=3000, dir_window=2500, norm_window=500, emit_dirs=1, emit_counts=1)   # 8192-entry ring
    _check(synth.to_numpy(synth.make_batch([synth.SynthSpec(n_rows=3000, seed=72, dir_window=2500, warm_ticks=4000, warm_trades=3000)])), params=p)
