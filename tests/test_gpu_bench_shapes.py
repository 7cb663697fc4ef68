"""GPU parity AT THE SHAPES THE BENCHMARK RUNS (SURVEY.md §8d "Parity runs").

Segment length and normaliser tile are functions of the batch size (api.cu: setup_shape / pick_tile_rows), so the
50 M-row benchmark runs 15 360-row segments and 2 000-row tiles.  These tests pin those two values and compare with
the compiled, unmodified reference (oracle/_ref, when it travelled with the repository; the C restatement otherwise):

  * one full instrument-day of 46 600 rows (BASELINE.json configs[0] / [2]);
  * the first 1 M rows of configs[1] (L = 10) and of configs[3] (L = 50), every row, the checker restarted
    one normaliser window + warm-up earlier for each range so that all host cores work side by side;
  * norm_mode = FAITHFUL: bit-equal z-scores wherever the raw column is bit-equal.

Every test appends its per-column numbers (plain max|dz|/max(1,|z|) against the checker, kappa, the checker's own
drift from the exactly summed definition) to the parity report (helpers.parity_record).
"""
import os

import numpy as np
import pytest

import mrp_b200
from mrp_b200 import synth
from oracle import binding as ob

from helpers import (ORACLE_WARM_ROWS, column_report, make_host_batch, nan_pattern_equal, norm_check, oracle_on_windows,
                     parity_record, raw_errors, run_gpu)

pytestmark = pytest.mark.gpu


@pytest.fixture
def bench_shape(monkeypatch):
    """The code path of the 50 M-row benchmark: 15 360-row segments (what the wave-aware choice in setup_shape picks for 50 M rows on
    148 SMs), 2 000-row normaliser tiles."""
    monkeypatch.setenv("MRP_FG_SEGMENT_ROWS", "15360")
    monkeypatch.setenv("MRP_FG_NORM_TILE", "2000")
    yield


def _reference_runner(levels):
    """(runner, kind): the compiled reference when it is here, else the restatement."""
    cores = os.cpu_count() or 1
    if ob.ref_available(levels):
        return (lambda hv, p: ob.run_reference(hv, None, use_book=False, nthreads=cores)), "reference"
    return (lambda hv, p: ob.run_oracle(hv, p, nthreads=cores)), "port"


def _compare_ranges(name, batch, got, windows, p, runner, kind, count_windows=2):
    """GPU output of the whole stream against the checker on `windows`; returns the merged per-column report."""
    refs = oracle_on_windows(batch, windows, p, runner=runner)
    worst = {}
    for (a, b), ref in zip(windows, refs):
        g_raw, g_norm = got["raw"][:, a:b], got["norm"][:, a:b]
        assert np.array_equal(got["bid_dir"][:, a:b], ref["bid_dir"]) and np.array_equal(got["ask_dir"][:, a:b], ref["ask_dir"]), (a, b)
        for col in ("midprice", "market_depth"):
            c = mrp_b200.FEATURES.index(col)
            assert np.array_equal(g_raw[c], ref["raw"][c]), (col, a, b)
        assert nan_pattern_equal(g_raw, ref["raw"]) and nan_pattern_equal(g_norm, ref["norm"]), (a, b)
        re = raw_errors(g_raw, ref["raw"])
        assert (re <= 1.0).all(), {mrp_b200.FEATURES[i]: re[i] for i in np.nonzero(re > 1.0)[0]}
        # conditioning-aware z check on the range, with the rows in front of it as history
        skip = ref["skip"]
        hist_gpu = np.concatenate([ref["norm_with_warm"][:, :skip], g_norm], axis=1)   # only the range's own rows are judged
        n_hist = ref["raw_with_warm"].shape[1]
        wa, wb, drift, frac = norm_check(hist_gpu, ref["raw_with_warm"], ref["norm_with_warm"], [0, n_hist], p.norm_window, 1e-9, first_row=skip)
        assert (wa <= 1.0).all(), ((a, b), {mrp_b200.FEATURES[i]: wa[i] for i in np.nonzero(wa > 1.0)[0]})
        assert (wb <= 1.0).all(), ((a, b), {mrp_b200.FEATURES[i]: (wb[i], drift[i]) for i in np.nonzero(wb > 1.0)[0]})
        rep = column_report(g_raw, g_norm, ref["raw"], ref["norm"], ref["raw_with_warm"], ref["norm_with_warm"], skip, p.norm_window)
        for col, r in rep.items():
            w = worst.setdefault(col, dict(r))
            for k, v in r.items():
                if isinstance(v, bool):
                    w[k] = w[k] and v
                elif k == "cells_compared":
                    w[k] = min(w[k], v)
                else:
                    w[k] = max(w[k], v)
    # integer window counts: bit-exact against the restatement (the compiled reference does not export them)
    cw = windows[:: max(1, len(windows) // count_windows)][:count_windows]
    for (a, b), ref in zip(cw, oracle_on_windows(batch, cw, p)):
        for i, cname in enumerate(mrp_b200.COUNTS):
            if cname == "norm_n":
                want = np.minimum(np.arange(a, b) + 1, p.norm_window)
                assert np.array_equal(got["counts"][i][a:b], want), cname
            else:
                assert np.array_equal(got["counts"][i][a:b], ref["counts"][i]), (cname, a, b)
    parity_record(name, checker=kind, rows=int(got["raw"].shape[1]), ranges=len(windows), warm_rows=ORACLE_WARM_ROWS,
                  segment_rows=os.environ.get("MRP_FG_SEGMENT_ROWS"), norm_tile=os.environ.get("MRP_FG_NORM_TILE"), columns=worst)
    zmax = max(r["z_max_abs_over_max1"] for r in worst.values())
    print("%s: checker=%s, %d/19 raw columns bit-exact, worst raw rel %.3g, worst plain |dz|/max(1,|z|) %.3g (checker drift %.3g)"
          % (name, kind, sum(r["raw_bit_exact"] for r in worst.values()), max(r["raw_max_rel"] for r in worst.values()), zmax,
             max(r["z_ref_vs_exact"] for r in worst.values())))
    return worst


def _ranges(n, parts):
    edges = np.linspace(0, n, parts + 1).astype(np.int64)
    return [(int(edges[i]), int(edges[i + 1])) for i in range(parts)]


def test_full_instrument_day_46600(bench_shape):
    """configs[0] / configs[2] shape: (6.5 h * 3600 - 100 s) / 0.5 s = 46 600 rows (dual_feature_pipeline.cpp:61,118)."""
    p = mrp_b200.default_params(emit_dirs=1, emit_counts=1)
    batch = make_host_batch([synth.SynthSpec(n_rows=46600, seed=0x5EED)])
    hv, got = run_gpu(batch, p)
    runner, kind = _reference_runner(10)
    _compare_ranges("instrument_day_46600", batch, got, [(0, 46600)], p, runner, kind, count_windows=1)


@pytest.mark.parametrize("levels", [10, 50], ids=["config1_L10", "config3_L50"])
def test_first_million_rows(levels, bench_shape):
    """First 1 M rows of the 50 M x 10-level (configs[1]) and 20 M x 50-level (configs[3]) instrument-days."""
    n = 1_000_000
    p = mrp_b200.default_params(depth_levels=levels, emit_dirs=1, emit_counts=1)
    batch = make_host_batch([synth.SynthSpec(n_rows=n, levels=levels, seed=0x5EED)])
    hv, got = run_gpu(batch, p)
    runner, kind = _reference_runner(levels)
    parts = max(8, min(64, os.cpu_count() or 8))
    _compare_ranges("first_1M_rows_L%d" % levels, batch, got, _ranges(n, parts), p, runner, kind)


def test_faithful_mode_is_bit_equal(bench_shape):
    """norm_mode = FAITHFUL re-enacts feature_normalizer.cpp:37-51,61-69 in its own rounding order: wherever the raw column is
    bit-equal to the checker's, the z-scores must be bit-equal too (window eviction included: 33 500 rows)."""
    p = mrp_b200.default_params(norm_mode=mrp_b200.NORM_FAITHFUL, emit_dirs=1, emit_counts=1)
    batch = make_host_batch([synth.SynthSpec(n_rows=33500, seed=13), synth.SynthSpec(n_rows=700, seed=14, warm_trades=0)])
    hv, got = run_gpu(batch, p)
    runner, kind = _reference_runner(10)
    ref = runner(hv, p)
    exact_cols = [c for c in range(19) if np.array_equal(got["raw"][c], ref["raw"][c], equal_nan=True)]
    assert len(exact_cols) >= 8, [mrp_b200.FEATURES[c] for c in exact_cols]
    for c in exact_cols:
        assert np.array_equal(got["norm"][c], ref["norm"][c], equal_nan=True), mrp_b200.FEATURES[c]
    # the other columns: the chain runs on raw values that differ in the last bits, nothing more
    others = [c for c in range(19) if c not in exact_cols]
    wa, wb, drift, frac = norm_check(got["norm"], ref["raw"], ref["norm"], hv.arrays["row_offsets"], p.norm_window, 1e-9)
    assert (wb[others] <= 1.0).all(), {mrp_b200.FEATURES[i]: wb[i] for i in others if wb[i] > 1.0}
    parity_record("faithful_mode", checker=kind, rows=int(hv.n_rows), raw_bit_exact_columns=[mrp_b200.FEATURES[c] for c in exact_cols],
                  z_bit_exact_columns=[mrp_b200.FEATURES[c] for c in exact_cols])
    print("faithful: %d columns raw AND z bit-equal to the %s: %s" % (len(exact_cols), kind, [mrp_b200.FEATURES[c] for c in exact_cols]))
