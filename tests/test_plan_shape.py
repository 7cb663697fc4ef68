"""Host arithmetic behind the C ABI (no GPU): how a batch is cut into segments / normaliser tiles and how mrp_fg_process_host cuts
a stream into chunks.  These lengths decide the wave count of the segment kernels, the normaliser's look-back and the head / tail of
the host pipeline — the shapes the bench lines were measured at are pinned here."""
import numpy as np
import pytest

import mrp_b200


def _ro(rows_per_stream):
    return np.concatenate([[0], np.cumsum(rows_per_stream)]).astype(np.int64)


def test_benchmarked_shapes(monkeypatch):
    monkeypatch.delenv("MRP_FG_SEGMENT_ROWS", raising=False)
    monkeypatch.delenv("MRP_FG_NORM_TILE", raising=False)
    # BASELINE.json configs[1]: 50 M rows on 148 SMs -> 15 360-row segments = 3256 CTAs = 22 full waves, 2000-row tiles
    seg, tile = mrp_b200.plan_shape(_ro([50_000_000]), sm_count=148)
    assert (seg, tile) == (15360, 2000)
    assert -(-50_000_000 // seg) == 22 * 148
    # configs[2]: 32 instrument-days of 46 600 rows -> 9 segments per day = 288 CTAs = two waves (2304-row segments were five)
    seg, tile = mrp_b200.plan_shape(_ro([46_600] * 32), sm_count=148)
    assert (seg, tile) == (5632, 2000)
    assert 32 * -(-46_600 // seg) == 288
    # a small batch keeps enough tiles for one round of 148 SMs x 2 CTAs
    seg, tile = mrp_b200.plan_shape(_ro([200_000]), sm_count=148)
    assert tile == 500 and seg % 512 == 0


def test_shape_rules():
    p = mrp_b200.default_params()
    for rows in ([33_500], [70_000, 33_000, 50_001], [8_000_000], [1_500_000] * 32):
        seg, tile = mrp_b200.plan_shape(_ro(rows), p, sm_count=148)
        assert 2048 <= seg <= 16384 and seg % 512 == 0          # whole chunks of the book pass
        assert p.norm_window % tile == 0 and tile % 4 == 0         # a window is a whole number of tiles
    # fewer multiprocessors, fewer and longer segments; 0 means a B200
    assert mrp_b200.plan_shape(_ro([46_600] * 32), sm_count=0) == mrp_b200.plan_shape(_ro([46_600] * 32), sm_count=148)
    assert mrp_b200.plan_shape(_ro([46_600] * 32), sm_count=32)[0] >= mrp_b200.plan_shape(_ro([46_600] * 32), sm_count=148)[0]
    with pytest.raises(mrp_b200.MrpError):
        mrp_b200.plan_shape(np.array([0, 10, 5], np.int64))


def test_the_environment_override_is_reported(monkeypatch):
    monkeypatch.setenv("MRP_FG_SEGMENT_ROWS", "257")
    monkeypatch.setenv("MRP_FG_NORM_TILE", "1500")
    assert mrp_b200.plan_shape(_ro([100_000])) == (257, 1500)


def test_host_chunk_schedule():
    C = 2 << 20
    ch = mrp_b200.plan_host_chunks(50_000_000)               # the e2e leg of the bench
    assert sum(ch) == 50_000_000 and max(ch) == C
    assert ch[:3] == [C // 8, C // 4, C // 2]                  # the first upload overlaps nothing: short first chunks
    assert ch[-1] <= C // 8 and all(a >= b for a, b in zip(ch[-5:], ch[-4:]))   # ... and the last download: a halving tail
    # a stream in the middle of a call is cut plainly
    mid = mrp_b200.plan_host_chunks(5 * C + 17, first_stream=False, last_stream=False)
    assert mid == [C] * 5 + [17]
    # short streams, explicit chunk lengths
    assert mrp_b200.plan_host_chunks(0) == []
    assert sum(mrp_b200.plan_host_chunks(30_000, 8000)) == 30_000
    assert sum(mrp_b200.plan_host_chunks(1000, 50_000)) == 1000
    for rows, c in ((1, 1), (9, 8), (100, 7), (123_457, 3000)):
        got = mrp_b200.plan_host_chunks(rows, c)
        assert sum(got) == rows and min(got) >= 1 and max(got) <= c
