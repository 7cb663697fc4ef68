"""The oracle (oracle/fg_oracle.c) against the committed golden vectors — outputs of the UNMODIFIED reference
(tests/golden/make_golden.py).  Bit-exact on every column: the restatement runs the reference's own loops
in the reference's own order."""
import numpy as np
import pytest

import mrp_b200
from oracle import binding as ob

from helpers import golden_cases, load_golden


def test_fixtures_exist():
    assert {"l10_warm", "l10_cold_ragged", "l50_warm", "l10_evict"} <= set(golden_cases())


@pytest.mark.parametrize("name", golden_cases())
def test_oracle_matches_golden(name):
    spec, batch, gold = load_golden(name)
    hv = mrp_b200.HostView(batch)
    p = mrp_b200.default_params(depth_levels=spec.levels, emit_dirs=1, emit_counts=1)
    out = ob.run_oracle(hv, p, nthreads=4)
    rows = gold["rows"]
    for key in ("bid_dir", "ask_dir"):
        assert np.array_equal(out[key][:, rows], gold[key]), key
    for key in ("raw", "norm"):
        for c, fname in enumerate(mrp_b200.FEATURES):
            assert np.array_equal(out[key][c, rows], gold[key][c], equal_nan=True), (key, fname)


def test_oracle_thread_count_does_not_change_results():
    spec, batch, _ = load_golden("l10_cold_ragged")
    hv = mrp_b200.HostView(batch)
    a = ob.run_oracle(hv, nthreads=1)
    b = ob.run_oracle(hv, nthreads=7)
    for k in ("raw", "norm", "counts"):
        assert np.array_equal(a[k], b[k], equal_nan=True)


def test_reference_quirks_pinned():
    """Known-answer checks of the reference behaviours SURVEY.md §7 H5 lists, on the golden outputs."""
    spec, batch, gold = load_golden("l10_cold_ragged")
    F = mrp_b200.FEATURES.index
    raw, norm = gold["raw"], gold["norm"]
    # row 0: window of one row -> variance 0 -> std 1 -> z = 0 for every finite column (feature_normalizer.cpp:62-68)
    fin0 = np.isfinite(raw[:, 0])
    assert (norm[fin0, 0] == 0.0).all()
    # no trades yet: aggressor_bias = 0/0 = NaN (feature_processor.cpp:242-246); NaN then poisons the running sums forever
    empty = batch["trade_cursor"] == 0
    assert np.array_equal(np.isnan(raw[F("aggressor_bias")]), empty)
    if empty.any():
        assert np.isnan(norm[F("aggressor_bias"), np.argmax(empty):]).all()
    # log_return is 0 until 21 earlier rows exist (feature_processor.cpp:37-42, feature_normalizer.hpp:19-24)
    assert (raw[F("log_return"), :21] == 0.0).all()
    # first snapshot: every populated level's size rose from an empty book -> direction +1 (order_book.cpp:115,168,208-239)
    sz = batch["bid_sz"][:, 0]
    assert np.array_equal(gold["bid_dir"][:, 0], (sz > 0).astype(np.int8))
    # order_arrival_rate is 0 on the first row (cache_.last_arrival_time_ns == 0, feature_processor.cpp:143-152)
    assert raw[F("order_arrival_rate"), 0] == 0.0
    # market_depth is integer valued
    md = raw[F("market_depth")]
    assert np.array_equal(md, np.round(md))
