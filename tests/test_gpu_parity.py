"""GPU parity: the CUDA path (through the C ABI) against the oracle on the same seeded inputs."""
import numpy as np
import pytest

import mrp_b200
from mrp_b200 import synth
from oracle import binding as ob

from helpers import (make_host_batch, nan_pattern_equal, norm_check, norm_errors, raw_errors, run_gpu)

pytestmark = pytest.mark.gpu


def _check(batch, params=None, z_tol=1e-9, mode="upload", min_frac=0.95):
    p = params or mrp_b200.default_params(emit_dirs=1, emit_counts=1)
    hv, got = run_gpu(batch, p, mode)
    ref = ob.run_oracle(hv, p)
    # integer book-delta and count columns: bit-exact
    assert np.array_equal(got["bid_dir"], ref["bid_dir"])
    assert np.array_equal(got["ask_dir"], ref["ask_dir"])
    for i, name in enumerate(mrp_b200.COUNTS):
        assert np.array_equal(got["counts"][i], ref["counts"][i]), f"count column {name}"
    # integer-valued feature columns: bit-exact
    for name in ("midprice", "market_depth"):
        c = mrp_b200.FEATURES.index(name)
        assert np.array_equal(got["raw"][c], ref["raw"][c]), name
    assert nan_pattern_equal(got["raw"], ref["raw"])
    re = raw_errors(got["raw"], ref["raw"])
    assert (re <= 1.0).all(), {mrp_b200.FEATURES[i]: re[i] for i in np.nonzero(re > 1.0)[0]}
    exact_cols = [mrp_b200.FEATURES[c] for c in range(19) if np.array_equal(got["raw"][c], ref["raw"][c], equal_nan=True)]
    print("raw: %d/19 columns bit-exact; worst error/tolerance %.3g (%s)" % (len(exact_cols), re.max(), mrp_b200.FEATURES[int(re.argmax())]))
    assert nan_pattern_equal(got["norm"], ref["norm"])
    # z-scores: conditioning-aware, against the exactly-summed definition and against the reference itself
    wa, wb, drift, frac = norm_check(got["norm"], ref["raw"], ref["norm"], hv.arrays["row_offsets"], p.norm_window, z_tol)
    assert frac > min_frac
    assert (wa <= 1.0).all(), {mrp_b200.FEATURES[i]: wa[i] for i in np.nonzero(wa > 1.0)[0]}
    assert (wb <= 1.0).all(), {mrp_b200.FEATURES[i]: (wb[i], drift[i]) for i in np.nonzero(wb > 1.0)[0]}
    zr = norm_errors(got["norm"], ref["norm"])
    print("worst z ratio vs exact %.3g, vs reference %.3g; reference drift %.3g; plain max|dz| vs reference %.3g"
          % (wa.max(), wb.max(), drift.max(), zr.max()))
    return got, ref


def test_single_stream_default():
    _check(make_host_batch([synth.SynthSpec(n_rows=6000, seed=11)]))


def test_cold_start_variable_cursors():
    # windows grow from empty; first rows have no trades (aggressor_bias = 0/0 = NaN)
    _check(make_host_batch([synth.SynthSpec(n_rows=3000, seed=12, warm_ticks=0, warm_trades=0, variable_cursors=True)]))


def test_multi_stream_device_pointers():
    specs = [synth.SynthSpec(n_rows=n, seed=20 + i, variable_cursors=bool(i & 1)) for i, n in enumerate((1, 2, 300, 2500, 1025))]
    _check(make_host_batch(specs), mode="device")


def test_normaliser_window_eviction():
    # longer than WINDOW_SIZE = 30000 rows so the rolling z-score window evicts
    _check(make_host_batch([synth.SynthSpec(n_rows=33500, seed=13)]))


def test_coscheduled_book_share_matches_the_default_run(monkeypatch):
    """MRP_FG_CORUN (experiment switch, default off): a share of the book pass's segments runs as the staging-free k_book form on a
    low-priority stream beside k_windows.  Same arithmetic per row: every column is byte-identical to the default run except the
    two row-axis EMAs (ofi, liquidity_stress), whose block scans are cut into chunks of another length (last-bit differences)."""
    batch = synth.to_numpy(synth.make_batch([synth.SynthSpec(n_rows=9000, seed=77), synth.SynthSpec(n_rows=4000, seed=78, variable_cursors=True)]))
    p = mrp_b200.default_params(norm_window=2000, emit_dirs=1, emit_counts=1)
    _, one = run_gpu(batch, p)
    monkeypatch.setenv("MRP_FG_CORUN", "0.5")
    _, two = run_gpu(batch, p)
    for k in ("bid_dir", "ask_dir", "counts"):
        assert np.array_equal(one[k], two[k]), k
    emas = {mrp_b200.FEATURES.index("ofi"), mrp_b200.FEATURES.index("liquidity_stress")}
    for c in range(mrp_b200.NUM_FEATURES):
        if c in emas:
            scale = np.nanmax(np.abs(one["raw"][c]))
            assert np.allclose(one["raw"][c], two["raw"][c], rtol=1e-12, atol=1e-15 * scale, equal_nan=True), mrp_b200.FEATURES[c]
            assert np.allclose(one["norm"][c], two["norm"][c], rtol=1e-7, atol=1e-9, equal_nan=True), mrp_b200.FEATURES[c]
        else:
            assert np.array_equal(one["raw"][c], two["raw"][c], equal_nan=True), mrp_b200.FEATURES[c]
            assert np.array_equal(one["norm"][c], two["norm"][c], equal_nan=True), mrp_b200.FEATURES[c]
