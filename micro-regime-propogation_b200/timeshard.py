"""Time-sharding of ONE long stream across ranks (BASELINE.json configs[4]: a 400 M-snapshot stream over 8 B200).

Rank g owns the snapshot rows [r0, r1) of the stream and the ticks / trades that arrive during those rows.  Every state
of the feature path has finite memory (30 000 rows for the z-score window, 21 rows for the log return,
1 row for the book-delta directions and the arrival rate, 1800 ticks, 1500 trades, EMAs that forget to < 1e-17 in 372 rows),
so a rank reproduces its rows exactly from its own slice plus a HALO taken from its left neighbour:

    rows   : the last  H = mrp_fg_halo_rows() = norm_window + EMA warm-up + 32  rows of the neighbour (inputs, not features: recomputing the
             neighbour's last 30 000 raw feature rows costs 0.06 % extra work at 50 M rows per GPU and needs no second
             exchange step and no extra kernel entry point)
    ticks  : from (tick cursor of the halo's first row) - max(mid_window, dir_window + 1) - 2
    trades : from (trade cursor of the halo's first row) - dir_window - 2

The halo travels peer to peer (torch.distributed send / recv: NCCL over NVLink on the GPU box, gloo in the CPU tests);
the extended slice is then handed to the engine as an ordinary stream whose cursors are rebased to the slice, and the
first `skip` rows of the outputs are dropped.  The only collective is the all-reduce of the global column moments.
This module is host-side plumbing: it contains no feature arithmetic.
"""
from __future__ import annotations

import math
from typing import Dict, Optional

import torch

ROW_KEYS_1D = ("ts_ns", "buy_vol", "sell_vol", "adds", "tick_cursor", "trade_cursor")
ROW_KEYS_2D = ("bid_px", "ask_px", "bid_sz", "ask_sz")
TICK_KEYS = ("mid", "spread")
TRADE_KEYS = ("trade_dir",)


def halo_rows(norm_window: int = 30000, params=None) -> int:
    """Rows of history a shard needs from its left neighbour: the engine's own rule (mrp_fg_halo_rows_for — the ONE definition)."""
    from . import cabi

    p = params or cabi.default_params(norm_window=norm_window)
    return cabi.halo_rows(p)


def shard_bounds(n_rows: int, world: int, rank: int):
    """Contiguous, near-equal row ranges."""
    base, rem = divmod(n_rows, world)
    r0 = rank * base + min(rank, rem)
    return r0, r0 + base + (1 if rank < rem else 0)


def shard_of(batch: Dict[str, object], r0: int, r1: int) -> Dict[str, object]:
    """Cut the rows [r0, r1) of a single-stream batch (global cursors kept) together with the ticks / trades that
    arrive during those rows.  Used by the tests and the benchmark to play the role of "what this rank holds"."""
    kc, jc = batch["tick_cursor"], batch["trade_cursor"]
    n = int(kc.shape[0])
    t0 = int(kc[r0 - 1]) if r0 > 0 else 0
    j0 = int(jc[r0 - 1]) if r0 > 0 else 0
    t1 = int(kc[r1 - 1]) if r1 < n else int(batch["mid"].shape[0])
    j1 = int(jc[r1 - 1]) if r1 < n else int(batch["trade_dir"].shape[0])
    out: Dict[str, object] = {"row0": r0, "tick0": t0, "trade0": j0, "n_rows_total": n}
    for k in ROW_KEYS_1D:
        out[k] = batch[k][r0:r1]
    for k in ROW_KEYS_2D:
        out[k] = batch[k][:, r0:r1]
    for k in TICK_KEYS:
        out[k] = batch[k][t0:t1]
    for k in TRADE_KEYS:
        out[k] = batch[k][j0:j1]
    return out


def tail_for_right(shard: Dict[str, object], H: int, mid_window: int = 1800, dir_window: int = 1500) -> Dict[str, object]:
    """What this rank sends to its right neighbour: its last H rows and the ticks / trades those rows can see."""
    n = int(shard["tick_cursor"].shape[0])
    if n < H:
        raise ValueError(f"a time shard must hold at least {H} rows (has {n}): the halo comes from ONE neighbour")
    a = n - H
    wmax = max(mid_window, dir_window + 1)
    t_need = max(0, int(shard["tick_cursor"][a]) - wmax - 2)
    j_need = max(0, int(shard["trade_cursor"][a]) - dir_window - 2)
    if (t_need < shard["tick0"] and shard["tick0"] > 0) or (j_need < shard["trade0"] and shard["trade0"] > 0):
        raise ValueError("the shard does not hold enough ticks / trades before its halo rows; use longer shards")
    t_need, j_need = max(t_need, shard["tick0"]), max(j_need, shard["trade0"])
    out: Dict[str, object] = {"row0": shard["row0"] + a, "tick0": t_need, "trade0": j_need}
    for k in ROW_KEYS_1D:
        out[k] = shard[k][a:].contiguous()
    for k in ROW_KEYS_2D:
        out[k] = shard[k][:, a:].contiguous()
    for k in TICK_KEYS:
        out[k] = shard[k][t_need - shard["tick0"]:].contiguous()
    for k in TRADE_KEYS:
        out[k] = shard[k][j_need - shard["trade0"]:].contiguous()
    return out


def extend(shard: Dict[str, object], halo: Optional[Dict[str, object]]):
    """halo + shard as ONE engine stream: cursors rebased to the slice, host offsets, and the rows to skip on output."""
    if halo is None:
        ext = dict(shard)
        skip = 0
    else:
        assert halo["row0"] + int(halo["tick_cursor"].shape[0]) == shard["row0"], "halo and shard rows are not adjacent"
        ext = {"row0": halo["row0"], "tick0": halo["tick0"], "trade0": halo["trade0"]}
        for k in ROW_KEYS_1D:
            ext[k] = torch.cat([halo[k], shard[k]])
        for k in ROW_KEYS_2D:
            ext[k] = torch.cat([halo[k], shard[k]], dim=1).contiguous()
        for k in TICK_KEYS + TRADE_KEYS:
            ext[k] = torch.cat([halo[k], shard[k]])
        skip = int(halo["tick_cursor"].shape[0])
    out = {k: ext[k] for k in ROW_KEYS_1D + ROW_KEYS_2D + TICK_KEYS + TRADE_KEYS}
    out["tick_cursor"] = (ext["tick_cursor"].to(torch.int64) - ext["tick0"]).to(ext["tick_cursor"].dtype)
    out["trade_cursor"] = (ext["trade_cursor"].to(torch.int64) - ext["trade0"]).to(ext["trade_cursor"].dtype)
    out["row_offsets"] = [0, int(out["tick_cursor"].shape[0])]
    out["tick_offsets"] = [0, int(out["mid"].shape[0])]
    out["trade_offsets"] = [0, int(out["trade_dir"].shape[0])]
    return out, skip


# ---- the same through the C ABI (mrp_fg_halo_query / _export / _import): the halo travels as ONE packed device buffer
def layout_with_headroom(shard: Dict[str, object], lead) -> "tuple[Dict[str, object], int]":
    """This rank's shard as ONE engine stream with room in front for the left neighbour's tail (`lead`: a HaloInfo, or None on
    rank 0): own rows from row lead.n_rows on, own ticks / trades behind lead.n_ticks / lead.n_trades, own cursors counted
    from the first HALO tick / trade.  Returns (batch of torch tensors for Engine.bind_device, rows to skip on output)."""
    H, K, T = (int(lead.n_rows), int(lead.n_ticks), int(lead.n_trades)) if lead is not None else (0, 0, 0)
    dev = shard["mid"].device
    out: Dict[str, object] = {}

    def front(t: torch.Tensor, lead_n: int) -> torch.Tensor:
        if lead_n == 0:
            return t.contiguous()
        shape = list(t.shape)
        shape[-1] += lead_n
        o = torch.zeros(shape, dtype=t.dtype, device=dev)
        o[..., lead_n:] = t
        return o

    for k in ("ts_ns", "buy_vol", "sell_vol", "adds") + ROW_KEYS_2D:
        out[k] = front(shard[k], H)
    out["tick_cursor"] = front((shard["tick_cursor"].to(torch.int64) - shard["tick0"] + K).to(torch.int32), H)
    out["trade_cursor"] = front((shard["trade_cursor"].to(torch.int64) - shard["trade0"] + T).to(torch.int32), H)
    for k in TICK_KEYS:
        out[k] = front(shard[k], K)
    for k in TRADE_KEYS:
        out[k] = front(shard[k], T)
    out["row_offsets"] = [0, int(out["ts_ns"].shape[0])]
    out["tick_offsets"] = [0, int(out["mid"].shape[0])]
    out["trade_offsets"] = [0, int(out["trade_dir"].shape[0])]
    return out, H


class TimeShard:
    """One rank of a stream split by time, driven through the C ABI.  setup() binds the shard with headroom and agrees the halo
    sizes with the neighbours (a chain of W - 1 tiny host messages, once); exchange() moves the packed tail to the right
    neighbour (one NCCL / gloo message per step) and unpacks the left neighbour's into the headroom."""

    def __init__(self, engine, shard: Dict[str, object], rank: int, world: int):
        import torch.distributed as dist

        from . import cabi

        self.engine, self.rank, self.world = engine, rank, world
        dev = shard["mid"].device
        self.lead = None
        if rank > 0:
            meta = torch.empty(4, dtype=torch.int64, device=dev)
            dist.recv(meta, rank - 1)
            self.lead = cabi.HaloInfo(*[int(x) for x in meta.tolist()])
        self.batch, self.skip = layout_with_headroom(shard, self.lead)
        self.row_base = int(shard["row0"]) - self.skip      # absolute stream row of the bound stream's first row
        engine.bind_device(self.batch)
        self.tail = None
        self.send_buf = self.recv_buf = None
        if rank + 1 < world:
            self.tail = engine.halo_query(torch.cuda.current_stream().cuda_stream if dev.type == "cuda" else 0)
            dist.send(torch.tensor([self.tail.n_rows, self.tail.n_ticks, self.tail.n_trades, self.tail.bytes], dtype=torch.int64, device=dev), rank + 1)
            self.send_buf = torch.empty(int(self.tail.bytes), dtype=torch.uint8, device=dev)
        if self.lead is not None:
            self.recv_buf = torch.empty(int(self.lead.bytes), dtype=torch.uint8, device=dev)

    def exchange(self) -> int:
        """Pack, ship, unpack: returns the bytes this rank sent."""
        import torch.distributed as dist

        st = torch.cuda.current_stream().cuda_stream
        ops = []
        if self.tail is not None:
            self.engine.halo_export(self.send_buf.data_ptr(), self.send_buf.numel(), st)
            ops.append(dist.P2POp(dist.isend, self.send_buf, self.rank + 1))
        if self.lead is not None:
            ops.append(dist.P2POp(dist.irecv, self.recv_buf, self.rank - 1))
        if ops:
            for w in dist.batch_isend_irecv(ops):
                w.wait()
        if self.lead is not None:
            self.engine.halo_import(self.recv_buf.data_ptr(), self.lead, st)
        return int(self.tail.bytes) if self.tail is not None else 0

    def carry_poison(self) -> bool:
        """After Engine.run: a NaN (or an evicted infinity) poisons the reference's running sums for ever, also across shard
        boundaries (feature_normalizer.cpp:37-51).  All-gathers every shard's first non-finite rows (57 integers per rank) and
        rewrites this shard's z-scores for what its predecessors saw.  Returns whether anything had to be rewritten."""
        import numpy as np
        import torch.distributed as dist

        dev = self.batch["mid"].device
        mine = torch.from_numpy(self.engine.poison_state([self.row_base])).to(dev)
        allst = [torch.empty_like(mine) for _ in range(self.world)]
        dist.all_gather(allst, mine)
        if self.rank == 0:
            return False
        before = torch.stack(allst[:self.rank]).amin(dim=0).cpu().numpy()
        if (before == np.iinfo(np.int64).max).all():
            return False
        self.engine.poison_apply(before, [self.row_base])
        return True

    @property
    def own_rows(self):
        return self.skip, int(self.batch["row_offsets"][1])


_META = ("row0", "tick0", "trade0")
_ALL = ROW_KEYS_1D + ROW_KEYS_2D + TICK_KEYS + TRADE_KEYS


def exchange_halo(shard: Dict[str, object], rank: int, world: int, H: int, mid_window: int = 1800, dir_window: int = 1500):
    """Neighbour exchange: every rank but the last sends its tail to rank + 1; every rank but the first receives.
    Returns the received halo (None on rank 0).  Uses the default process group (NCCL -> NVLink P2P, or gloo)."""
    import torch.distributed as dist

    dev = shard["mid"].device
    ops, recv = [], None
    if rank + 1 < world:
        tail = tail_for_right(shard, H, mid_window, dir_window)
        meta = torch.tensor([tail[m] for m in _META] + [tail[k].shape[-1] for k in _ALL], dtype=torch.int64, device=dev)
        dist.send(meta, rank + 1)
        for k in _ALL:
            ops.append(dist.P2POp(dist.isend, tail[k], rank + 1))
    if rank > 0:
        meta = torch.empty(len(_META) + len(_ALL), dtype=torch.int64, device=dev)
        dist.recv(meta, rank - 1)
        meta = [int(x) for x in meta.tolist()]
        recv = {m: meta[i] for i, m in enumerate(_META)}
        L = shard["bid_px"].shape[0]
        for i, k in enumerate(_ALL):
            n = meta[len(_META) + i]
            shape = (L, n) if k in ROW_KEYS_2D else (n,)
            recv[k] = torch.empty(shape, dtype=shard[k].dtype, device=dev)
            ops.append(dist.P2POp(dist.irecv, recv[k], rank - 1))
    if ops:
        for w in dist.batch_isend_irecv(ops):
            w.wait()
    return recv
