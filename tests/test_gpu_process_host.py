"""GPU: the pipelined host-to-host entry point (mrp_fg_process_host: time chunks with halos on two alternating lanes)
against the single-shot upload / run / download sequence and the oracle."""
import numpy as np
import pytest

import mrp_b200
from mrp_b200 import synth
from oracle import binding as ob

from helpers import norm_check, raw_errors, run_gpu, z_consistency

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("chunk", [3000, 7001, 50000])
def test_chunked_pipeline_matches_single_shot(chunk):
    p = mrp_b200.default_params(norm_window=2000)
    specs = [synth.SynthSpec(n_rows=n, seed=140 + i, variable_cursors=bool(i & 1), warm_ticks=0 if i == 2 else 2000,
                             warm_trades=0 if i == 2 else 1600) for i, n in enumerate((21000, 9000, 5000))]
    batch = synth.to_numpy(synth.make_batch(specs))
    hv, one = run_gpu(batch, p)
    eng = mrp_b200.Engine(p, device=0)
    raw, norm = eng.process_host(hv, chunk_rows=chunk)
    raw2, norm2 = eng.process_host(hv, chunk_rows=chunk)      # lanes are reused: second call must give the same bytes
    eng.close()
    assert np.array_equal(raw, raw2, equal_nan=True) and np.array_equal(norm, norm2, equal_nan=True)
    ref = ob.run_oracle(hv, p)
    assert np.array_equal(np.isnan(raw), np.isnan(ref["raw"]))
    re = raw_errors(raw, ref["raw"])
    assert (re <= 1.0).all(), re
    # row-local columns are bit-identical; the sliding-window and EMA columns depend (within rounding) on where a segment
    # or a chunk starts, so they agree to ~1e-13 with the single-shot run — and both are within tolerance of the oracle
    F = mrp_b200.FEATURES.index
    for name in ("midprice", "log_spread", "market_depth", "depth_imbalance", "price_gap", "signed_volume_pressure",
                 "order_arrival_rate", "lob_slope", "reversal_rate", "aggressor_bias", "shannon_entropy"):
        assert np.array_equal(raw[F(name)], one["raw"][F(name)], equal_nan=True), name
    for c in range(19):
        scale = np.nanmax(np.abs(one["raw"][c]))
        assert np.allclose(raw[c], one["raw"][c], rtol=1e-10, atol=1e-13 * scale, equal_nan=True), mrp_b200.FEATURES[c]
    # z-scores: conditioning-aware, against the exactly summed definition and the oracle (as in test_gpu_parity)
    wa, wb, drift, frac = norm_check(norm, ref["raw"], ref["norm"], hv.arrays["row_offsets"], p.norm_window, 1e-9)
    assert (wa <= 1.0).all() and (wb <= 1.0).all(), (wa, wb)
    # chunked against single-shot: same z-scores from raw values that agree to 1e-10 (asserted above): conditioning-aware
    worst, frac = z_consistency(norm, one["norm"], one["raw"], hv.arrays["row_offsets"], p.norm_window, raw_rel=1e-10)
    assert worst <= 1.0 and frac > 0.9, (worst, frac)


def test_default_window_and_chunk():
    hv = mrp_b200.HostView(synth.to_numpy(synth.make_batch([synth.SynthSpec(n_rows=120_000, seed=150)])))
    eng = mrp_b200.Engine(mrp_b200.default_params(), device=0)
    eng.upload(hv)
    eng.run()
    one = eng.download()
    raw, norm = eng.process_host(hv, chunk_rows=40_000)      # halo of mrp_fg_halo_rows() = 30 416 rows in front of every chunk
    eng.close()
    assert np.array_equal(raw[0], one["raw"][0])
    for c in range(19):
        assert np.allclose(raw[c], one["raw"][c], rtol=1e-10, atol=1e-13 * np.abs(one["raw"][c]).max(), equal_nan=True), c
    worst, frac = z_consistency(norm, one["norm"], one["raw"], hv.arrays["row_offsets"], 30000, raw_rel=1e-10)
    assert worst <= 1.0 and frac > 0.9, (worst, frac)


def test_poisoned_columns_stay_poisoned_across_chunks():
    """feature_normalizer.cpp:37-51: a NaN raw value turns the running sums NaN FOR EVER — also for rows thousands of chunks later.
    An empty bid side at row 700 (price 0 -> log(0) = -inf in log_spread, lob_slope via the book) must poison every later chunk's
    z-scores exactly as the single-shot run (and the oracle) have it."""
    p = mrp_b200.default_params(norm_window=1000)
    batch = synth.to_numpy(synth.make_batch([synth.SynthSpec(n_rows=12000, seed=160)]))
    batch["bid_px"] = batch["bid_px"].copy()
    batch["bid_px"][:, 700] = np.nan          # a NaN book row: midprice, log_spread, lob_slope, price_gap ... go NaN at row 700
    batch["ask_px"] = batch["ask_px"].copy()
    batch["ask_px"][0, 2500] = np.inf          # an infinity: +-inf sums, NaN once it is evicted 1000 rows later
    hv, one = run_gpu(batch, p)
    ref = ob.run_oracle(hv, p)
    assert np.array_equal(np.isnan(one["norm"]), np.isnan(ref["norm"]))
    eng = mrp_b200.Engine(p, device=0)
    raw, norm = eng.process_host(hv, chunk_rows=1500)     # halo = 1000 + 384 + 32 rows: chunks from row ~3000 on never see row 700
    eng.close()
    assert np.array_equal(np.isnan(norm), np.isnan(ref["norm"])), (np.isnan(norm).sum(), np.isnan(ref["norm"]).sum())
    assert np.array_equal(np.isinf(norm), np.isinf(ref["norm"]))
    poisoned = np.isnan(ref["norm"][:, -1])
    assert poisoned.sum() >= 3                             # several columns are NaN to the end of the stream
    fin = np.isfinite(ref["norm"])
    assert np.allclose(norm[fin], one["norm"][fin], rtol=0, atol=1e-6)


def test_one_output_may_be_left_on_the_device():
    """raw == NULL or norm == NULL: that matrix is not copied back; the other one is byte-identical to the two-output call."""
    hv = mrp_b200.HostView(synth.to_numpy(synth.make_batch([synth.SynthSpec(n_rows=30_000, seed=171)])))
    p = mrp_b200.default_params(norm_window=4000)
    eng = mrp_b200.Engine(p, device=0)
    raw, norm = eng.process_host(hv, chunk_rows=8000)
    only = np.full_like(norm, -7.0)
    eng.process_host(hv, chunk_rows=8000, norm_ptr=only.ctypes.data)
    assert np.array_equal(only, norm, equal_nan=True)
    only[:] = -7.0
    eng.process_host(hv, chunk_rows=8000, raw_ptr=only.ctypes.data)
    assert np.array_equal(only, raw, equal_nan=True)
    with pytest.raises(mrp_b200.MrpError):
        eng.lib.mrp_fg_process_host.restype = int
        eng._check(eng.lib.mrp_fg_process_host(eng.ctx, None, None, None, 0))
    eng.close()
