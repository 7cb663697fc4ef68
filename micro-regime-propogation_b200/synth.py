"""Synthetic limit-order-book snapshot streams (test / benchmark input, not product logic).

Counter-based and integer-only: every value is a pure function of (seed, tag, index) through a
64-bit mixing hash evaluated with wrapping int64 arithmetic, plus exact integer cumulative sums.
The same code runs on torch CPU tensors (parity tests, golden fixtures) and on CUDA tensors
(the 50 M-row benchmark, where inputs never cross PCIe) and gives bit-identical arrays.

The arrays follow include/mrp_fg.h (mrp_fg_soa_view).  Semantics mirrored from the reference:
  * tick axis  = what FeatureEngine::UpdateMidpriceAndSpread receives every 50 ms
                 (data_ingestion/src/core/feature_engine.cpp:130-150): mid = (bid+ask)/2, spread = ask-bid;
  * trade axis = what FeatureEngine::update_trade receives (feature_engine.cpp:87-110): dir = +-1, a size;
                 rolling_buy/sell_volume = exact sums over the last 1500 trades;
  * snapshot rows: top-L book at the last tick before the snapshot, level i one tick further out,
                 a suffix of deep levels occasionally empty (price 0, size 0: order_book.cpp:119-147 zero-fills);
  * cursors    = number of ticks / trades pushed before the snapshot.
Prices are integer multiples of 0.25 (ES tick), mids of 0.125, so they are exact in binary64.
"""
from __future__ import annotations

import dataclasses
from typing import Dict, List, Optional

import torch

_M63 = (1 << 63) - 1


def _s64(x: int) -> int:
    x &= (1 << 64) - 1
    return x - (1 << 64) if x >= (1 << 63) else x


_C0 = _s64(0x9E3779B97F4A7C15)
_C1 = _s64(0xBF58476D1CE4E5B9)
_C2 = _s64(0x94D049BB133111EB)


def _lsr(x: torch.Tensor, s: int) -> torch.Tensor:
    return (x >> s) & ((1 << (64 - s)) - 1)


def mix64(x: torch.Tensor) -> torch.Tensor:
    """splitmix64 finaliser on wrapping int64; returns a non-negative 63-bit value."""
    x = x + _C0
    x = (x ^ _lsr(x, 30)) * _C1
    x = (x ^ _lsr(x, 27)) * _C2
    x = x ^ _lsr(x, 31)
    return x & _M63


def _h(seed: int, tag: int, idx: torch.Tensor) -> torch.Tensor:
    return mix64(idx * _s64(0xD1342543DE82EF95) + _s64(seed * 0x632BE59BD9B4E019 + tag * 0x2545F4914F6CDD1D))


def _tri(idx: torch.Tensor, period: int, amp: int) -> torch.Tensor:
    """integer triangle wave in [0, amp] with the given period."""
    ph = idx % (2 * period)
    ph = torch.where(ph < period, ph, 2 * period - ph)
    return (ph * amp) // period


@dataclasses.dataclass
class SynthSpec:
    n_rows: int
    levels: int = 10
    ticks_per_row: int = 10       # m: SNAPSHOT_INTERVAL_NS / 50 ms (common_constants.hpp:11, dual_feature_pipeline.cpp:48)
    seed: int = 0x5EED
    warm_ticks: int = 2000        # ticks pushed before the first snapshot (100 s of 50 ms ticks, dual_feature_pipeline.cpp:61)
    warm_trades: int = 1600       # trades before the first snapshot (0 => aggressor_bias of row 0 is 0/0 = NaN)
    variable_cursors: bool = False  # ticks per row uniform in [0, 2m] instead of exactly m
    max_trades_per_row: int = 4   # trades per row uniform in [0, max]; mean q = max/2
    dir_window: int = 1500        # ROLLING_WINDOW for the rolling buy/sell volume (feature_engine.cpp:103)
    base_ticks: int = 20000       # centre of the price band in ticks of 0.25 (5000.00)
    band: int = 4000              # half-width of the reflecting band, in ticks


def make_stream(spec: SynthSpec, device: str | torch.device = "cpu") -> Dict[str, torch.Tensor]:
    """One stream.  Level-major book arrays have shape [L, n_rows]."""
    dev = torch.device(device)
    i64 = torch.int64
    N, L, m, seed = spec.n_rows, spec.levels, spec.ticks_per_row, spec.seed
    rows = torch.arange(N, dtype=i64, device=dev)

    # ---- cursors
    if spec.variable_cursors:
        per = _h(seed, 1, rows) % (2 * m + 1)
        per[0] = 0
        kcur = spec.warm_ticks + torch.cumsum(per, 0)
    else:
        kcur = spec.warm_ticks + m * rows
    K = int(kcur[-1].item()) if N else 0
    K = max(K, 1)
    per_tr = _h(seed, 2, rows) % (spec.max_trades_per_row + 1)
    jcur = spec.warm_trades + torch.cumsum(per_tr, 0)
    T = int(jcur[-1].item()) if N else 0

    # ---- tick axis: reflected integer random walk of the best bid (in ticks), regime-dependent activity
    k = torch.arange(K, dtype=i64, device=dev)
    hk = _h(seed, 3, k)
    thr = 100 + _tri(k, 150_000, 500)                   # P(move) between ~10% and ~59%, slow regimes
    move = (hk % 1024) < thr
    sign = torch.where(((hk >> 20) & 1) == 1, 1, -1)
    step = torch.where(move, sign, torch.zeros_like(sign))
    walk = torch.cumsum(step, 0)
    del hk, thr, move, sign, step
    B2 = 2 * spec.band
    ph = (walk + spec.band) % (2 * B2)
    best_bid_t = spec.base_ticks - spec.band + torch.where(ph < B2, ph, 2 * B2 - ph)
    del walk, ph
    hs = _h(seed, 4, k) % 8
    spr_t = torch.where(hs < 6, torch.ones_like(hs), hs - 4)   # 1 (75%), 2, 3 ticks
    del hs
    mid = (2 * best_bid_t + spr_t).to(torch.float64) * 0.125
    spread = spr_t.to(torch.float64) * 0.25

    # ---- snapshot rows: book state at the last tick before the snapshot
    at = torch.clamp(kcur - 1, min=0)
    bb = best_bid_t[at]
    sp = spr_t[at]
    del best_bid_t, spr_t, k
    bid_px = torch.empty((L, N), dtype=torch.float64, device=dev)
    ask_px = torch.empty((L, N), dtype=torch.float64, device=dev)
    bid_sz = torch.empty((L, N), dtype=torch.int32, device=dev)
    ask_sz = torch.empty((L, N), dtype=torch.int32, device=dev)
    hz = _h(seed, 5, rows)
    zb = torch.where(hz % 64 == 0, 1 + (hz >> 8) % 3, torch.zeros_like(hz))        # empty deep bid levels
    za = torch.where((hz >> 16) % 64 == 0, 1 + (hz >> 24) % 3, torch.zeros_like(hz))
    for i in range(L):
        hold = 1 + (i % 3)                         # level i's size changes every `hold` rows -> dir == 0 runs
        sb = 1 + _h(seed, 100 + i, rows // hold) % 200
        sa = 1 + _h(seed, 300 + i, rows // hold) % 200
        eb = (L - i) <= zb
        ea = (L - i) <= za
        bid_px[i] = torch.where(eb, 0.0, (bb - i).to(torch.float64) * 0.25)
        ask_px[i] = torch.where(ea, 0.0, (bb + sp + i).to(torch.float64) * 0.25)
        bid_sz[i] = torch.where(eb, 0, sb).to(torch.int32)
        ask_sz[i] = torch.where(ea, 0, sa).to(torch.int32)
    del hz, zb, za, bb, sp, at

    ts = 1_746_453_700_000_000_000 + rows * 500_000_000 + _h(seed, 6, rows) % 20_000_000
    adds = (_h(seed, 7, rows) % 300).to(torch.int32)

    # ---- trade axis
    j = torch.arange(T, dtype=i64, device=dev)
    hj = _h(seed, 8, j)
    pthr = 358 + _tri(j, 37_000, 308)               # P(buy) between ~0.35 and ~0.65
    tdir = torch.where((hj % 1024) < pthr, 1, -1).to(torch.int8)
    tsize = (1 + (hj >> 12) % 50)
    cb = torch.cumsum(torch.where(tdir > 0, tsize, torch.zeros_like(tsize)), 0)
    cs = torch.cumsum(torch.where(tdir < 0, tsize, torch.zeros_like(tsize)), 0)
    zero = torch.zeros(1, dtype=i64, device=dev)
    cb = torch.cat([zero, cb])
    cs = torch.cat([zero, cs])
    lo = torch.clamp(jcur - spec.dir_window, min=0)
    buy = (cb[jcur] - cb[lo]).to(torch.float64)
    sell = (cs[jcur] - cs[lo]).to(torch.float64)
    del cb, cs, lo, hj, pthr, j

    return {
        "ts_ns": ts, "bid_px": bid_px, "ask_px": ask_px, "bid_sz": bid_sz, "ask_sz": ask_sz,
        "buy_vol": buy, "sell_vol": sell, "adds": adds,
        "tick_cursor": kcur.to(torch.int32), "trade_cursor": jcur.to(torch.int32),
        "mid": mid, "spread": spread, "trade_dir": tdir, "trade_size": tsize.to(torch.float64),
    }


def _blocked_sum(fn, n: int, dev, block: int = 1 << 24) -> int:
    """sum_{i<n} fn(i) with i handed to fn in blocks (int64 tensors on dev); returns a Python int."""
    total = 0
    for a in range(0, n, block):
        idx = torch.arange(a, min(a + block, n), dtype=torch.int64, device=dev)
        total += int(fn(idx).sum().item())
    return total


def make_stream_range(spec: SynthSpec, r0: int, r1: int, device: str | torch.device = "cpu") -> Dict[str, object]:
    """The rows [r0, r1) of make_stream(spec) and the ticks / trades that arrive during those rows (for r0 = 0: from the
    stream's first tick / trade), bit-identical to slicing the whole stream — without materialising it: the generator is
    counter-based, so only the running sums at the range start (the price walk, the trade count, the buy / sell volume
    prefixes) are accumulated block by block.  Cursors stay GLOBAL (counts since the stream start).  Returns the arrays plus
    'row0', 'tick0', 'trade0' = global index of the first row / tick / trade held.  Regular cursors only."""
    if spec.variable_cursors:
        raise ValueError("make_stream_range needs regular cursors (ticks_per_row fixed)")
    dev = torch.device(device)
    i64 = torch.int64
    N, L, m, seed = spec.n_rows, spec.levels, spec.ticks_per_row, spec.seed
    assert 0 <= r0 < r1 <= N
    rows = torch.arange(r0, r1, dtype=i64, device=dev)
    n = r1 - r0
    kcur = spec.warm_ticks + m * rows
    # ticks pushed before row r0's predecessor snapshot ... the shard holds the ticks [t0, t1) with t0 = cursor of row r0 - 1
    t0 = 0 if r0 == 0 else spec.warm_ticks + m * (r0 - 1)
    t1 = spec.warm_ticks + m * (r1 - 1) if r1 < N else max(spec.warm_ticks + m * (N - 1), 1)
    per_tr = _h(seed, 2, rows) % (spec.max_trades_per_row + 1)
    j_before = spec.warm_trades + _blocked_sum(lambda i: _h(seed, 2, i) % (spec.max_trades_per_row + 1), r0, dev)   # cursor of row r0 - 1 (r0 > 0)
    jcur = j_before + torch.cumsum(per_tr, 0)
    if r0 == 0:
        j0 = 0
    else:
        j0 = j_before
    T_total_here = int(jcur[-1].item())
    j1 = T_total_here   # trades up to the last row's cursor (== the whole stream's trade count when r1 == N)

    def step_of(k):
        hk = _h(seed, 3, k)
        thr = 100 + _tri(k, 150_000, 500)
        move = (hk % 1024) < thr
        sign = torch.where(((hk >> 20) & 1) == 1, 1, -1)
        return torch.where(move, sign, torch.zeros_like(sign))

    # the book at row r needs the tick before its cursor: ticks from min(t0, cursor(r0) - 1) on
    ta = min(t0, max(int(kcur[0].item()) - 1, 0))
    walk0 = _blocked_sum(step_of, ta, dev)                                  # the walk before tick ta
    k = torch.arange(ta, t1, dtype=i64, device=dev)
    walk = walk0 + torch.cumsum(step_of(k), 0)
    B2 = 2 * spec.band
    ph = (walk + spec.band) % (2 * B2)
    best_bid_t = spec.base_ticks - spec.band + torch.where(ph < B2, ph, 2 * B2 - ph)
    del walk, ph
    hs = _h(seed, 4, k) % 8
    spr_t = torch.where(hs < 6, torch.ones_like(hs), hs - 4)
    del hs
    mid = (2 * best_bid_t + spr_t).to(torch.float64) * 0.125
    spread = spr_t.to(torch.float64) * 0.25
    at = torch.clamp(kcur - 1, min=0) - ta
    bb = best_bid_t[at]
    sp = spr_t[at]
    del best_bid_t, spr_t, k
    mid, spread = mid[t0 - ta:], spread[t0 - ta:]
    bid_px = torch.empty((L, n), dtype=torch.float64, device=dev)
    ask_px = torch.empty((L, n), dtype=torch.float64, device=dev)
    bid_sz = torch.empty((L, n), dtype=torch.int32, device=dev)
    ask_sz = torch.empty((L, n), dtype=torch.int32, device=dev)
    hz = _h(seed, 5, rows)
    zb = torch.where(hz % 64 == 0, 1 + (hz >> 8) % 3, torch.zeros_like(hz))
    za = torch.where((hz >> 16) % 64 == 0, 1 + (hz >> 24) % 3, torch.zeros_like(hz))
    for i in range(L):
        hold = 1 + (i % 3)
        sb = 1 + _h(seed, 100 + i, rows // hold) % 200
        sa = 1 + _h(seed, 300 + i, rows // hold) % 200
        eb = (L - i) <= zb
        ea = (L - i) <= za
        bid_px[i] = torch.where(eb, 0.0, (bb - i).to(torch.float64) * 0.25)
        ask_px[i] = torch.where(ea, 0.0, (bb + sp + i).to(torch.float64) * 0.25)
        bid_sz[i] = torch.where(eb, 0, sb).to(torch.int32)
        ask_sz[i] = torch.where(ea, 0, sa).to(torch.int32)
    del hz, zb, za, bb, sp, at
    ts = 1_746_453_700_000_000_000 + rows * 500_000_000 + _h(seed, 6, rows) % 20_000_000
    adds = (_h(seed, 7, rows) % 300).to(torch.int32)

    # trades: the volume windows reach dir_window trades back from the first cursor
    def trade_of(j):
        hj = _h(seed, 8, j)
        pthr = 358 + _tri(j, 37_000, 308)
        d = torch.where((hj % 1024) < pthr, 1, -1)
        return d, (1 + (hj >> 12) % 50)

    ja = max(0, min(j0, int(jcur[0].item()) - spec.dir_window))
    j = torch.arange(ja, j1, dtype=i64, device=dev)
    d, sz = trade_of(j)
    zero = torch.zeros(1, dtype=i64, device=dev)
    cb = torch.cat([zero, torch.cumsum(torch.where(d > 0, sz, torch.zeros_like(sz)), 0)])   # relative prefixes: only differences are used
    cs = torch.cat([zero, torch.cumsum(torch.where(d < 0, sz, torch.zeros_like(sz)), 0)])
    lo = torch.clamp(jcur - spec.dir_window, min=0)
    buy = (cb[jcur - ja] - cb[lo - ja]).to(torch.float64)
    sell = (cs[jcur - ja] - cs[lo - ja]).to(torch.float64)
    tdir = d.to(torch.int8)[j0 - ja:]
    tsize = sz.to(torch.float64)[j0 - ja:]
    return {
        "ts_ns": ts, "bid_px": bid_px, "ask_px": ask_px, "bid_sz": bid_sz, "ask_sz": ask_sz,
        "buy_vol": buy, "sell_vol": sell, "adds": adds,
        "tick_cursor": kcur, "trade_cursor": jcur,            # GLOBAL, int64 (a 400 M-row stream reaches 4e9 ticks)
        "mid": mid, "spread": spread, "trade_dir": tdir, "trade_size": tsize,
        "row0": r0, "tick0": t0, "trade0": j0, "n_rows_total": N,
    }


_ROW_KEYS_1D = ("ts_ns", "buy_vol", "sell_vol", "adds", "tick_cursor", "trade_cursor")
_ROW_KEYS_2D = ("bid_px", "ask_px", "bid_sz", "ask_sz")
_TICK_KEYS = ("mid", "spread")
_TRADE_KEYS = ("trade_dir", "trade_size")


def make_batch(specs: List[SynthSpec], device: str | torch.device = "cpu") -> Dict[str, object]:
    """Concatenate independent streams (instrument-days) into one batch + host offset lists."""
    streams = [make_stream(s, device) for s in specs]
    return concat_streams(streams)


def concat_streams(streams: List[Dict[str, torch.Tensor]]) -> Dict[str, object]:
    out: Dict[str, object] = {}
    for key in _ROW_KEYS_1D + _TICK_KEYS + _TRADE_KEYS:
        out[key] = torch.cat([s[key] for s in streams]) if len(streams) > 1 else streams[0][key]
    for key in _ROW_KEYS_2D:
        out[key] = (torch.cat([s[key] for s in streams], dim=1) if len(streams) > 1 else streams[0][key]).contiguous()
    ro, to, tro = [0], [0], [0]
    for s in streams:
        ro.append(ro[-1] + s["ts_ns"].numel())
        to.append(to[-1] + s["mid"].numel())
        tro.append(tro[-1] + s["trade_dir"].numel())
    out["row_offsets"], out["tick_offsets"], out["trade_offsets"] = ro, to, tro
    return out


def to_numpy(batch: Dict[str, object]) -> Dict[str, object]:
    import numpy as np

    res: Dict[str, object] = {}
    for k, v in batch.items():
        if isinstance(v, torch.Tensor):
            a = v.detach().cpu().numpy()
            if k == "ts_ns":
                a = a.astype(np.uint64)
            elif k in ("tick_cursor", "trade_cursor"):
                a = a.astype(np.uint32)
            res[k] = np.ascontiguousarray(a)
        else:
            res[k] = np.asarray(v, dtype=np.int64)
    return res
