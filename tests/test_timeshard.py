"""Time-sharding of one stream across ranks (host-side logic, gloo, world_size 2, CPU): halo sizing, neighbour exchange,
cursor rebasing and the all-reduce of the global column moments.  Each rank runs the CHECKER (oracle) on its halo-extended
slice; its own rows must reproduce the whole-stream run."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import mrp_b200
from mrp_b200 import synth, timeshard
from oracle import binding as ob

NORM_WINDOW = 2000
N_ROWS = 9000


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _np_batch(ext):
    out = {}
    for k, v in ext.items():
        if isinstance(v, torch.Tensor):
            a = v.numpy()
            if k == "ts_ns":
                a = a.astype(np.uint64)
            elif k in ("tick_cursor", "trade_cursor"):
                a = a.astype(np.uint32)
            out[k] = np.ascontiguousarray(a)
        else:
            out[k] = np.asarray(v, dtype=np.int64)
    return out


def _worker(rank, world, port, variable, results):
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from helpers import z_consistency

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        p = mrp_b200.default_params(norm_window=NORM_WINDOW)
        spec = synth.SynthSpec(n_rows=N_ROWS, seed=77, variable_cursors=variable)
        full = synth.make_batch([spec])                              # every rank can regenerate the stream (counter-based)
        full = {k: v for k, v in full.items() if k != "trade_size"}
        H = timeshard.halo_rows(NORM_WINDOW)
        r0, r1 = timeshard.shard_bounds(N_ROWS, world, rank)
        shard = timeshard.shard_of(full, r0, r1)                     # ... but only ITS rows / ticks / trades are used below
        halo = timeshard.exchange_halo(shard, rank, world, H)
        assert (halo is None) == (rank == 0)
        ext, skip = timeshard.extend(shard, halo)
        assert skip == (0 if rank == 0 else H)
        got = ob.run_oracle(mrp_b200.HostView(_np_batch(ext)), p, nthreads=2, want_dirs=True, want_counts=False)
        ref = ob.run_oracle(mrp_b200.HostView(synth.to_numpy(full)), p, nthreads=2, want_dirs=True, want_counts=False)
        own_raw, own_norm = got["raw"][:, skip:], got["norm"][:, skip:]
        ref_raw, ref_norm = ref["raw"][:, r0:r1], ref["norm"][:, r0:r1]
        F = mrp_b200.FEATURES.index
        ema = {F("ofi"), F("liquidity_stress")}
        for c in range(19):
            if c in ema:   # warmed over >= 512 halo rows: agrees to ~1e-17 of the column scale
                assert np.allclose(own_raw[c], ref_raw[c], rtol=1e-12, atol=1e-15 * np.abs(ref_raw[c]).max()), mrp_b200.FEATURES[c]
            else:
                assert np.array_equal(own_raw[c], ref_raw[c], equal_nan=True), mrp_b200.FEATURES[c]
        assert np.array_equal(got["bid_dir"][:, skip:], ref["bid_dir"][:, r0:r1])
        # the reference's own running-sum drift differs with the start row: conditioning-aware, each side's drift allowed for
        spliced = ref["norm"].copy()
        spliced[:, r0:r1] = own_norm
        worst, frac = z_consistency(spliced, ref["norm"], ref["raw"], [0, N_ROWS], NORM_WINDOW, allow_drift=True)
        assert worst <= 1.0 and frac > 0.9, (worst, frac)
        # global column moments: one all-reduce of 19 x {n, sum, sum^2}
        mom = torch.from_numpy(np.stack([np.full(19, r1 - r0, np.float64), own_raw.sum(1), (own_raw ** 2).sum(1)], 1))
        dist.all_reduce(mom)
        glob = np.stack([np.full(19, N_ROWS, np.float64), ref["raw"].sum(1), (ref["raw"] ** 2).sum(1)], 1)
        assert np.allclose(mom.numpy(), glob, rtol=1e-10, equal_nan=True)
        results[rank] = "ok"
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("variable", [False, True], ids=["regular_cursors", "ragged_cursors"])
def test_two_rank_time_shards_reproduce_the_whole_stream(variable):
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    results = ctx.Manager().dict()
    procs = [ctx.Process(target=_worker, args=(r, world, port, variable, results)) for r in range(world)]
    for pr in procs:
        pr.start()
    for pr in procs:
        pr.join(300)
    assert all(pr.exitcode == 0 for pr in procs), [pr.exitcode for pr in procs]
    assert dict(results) == {0: "ok", 1: "ok"}


def test_halo_geometry():
    # 0.9^372 < 1e-17, rounded up to whole 384-row chunks of the book pass: the engine's rule, one definition for every caller
    assert timeshard.halo_rows() == mrp_b200.halo_rows() == 30000 + 384 + 32
    assert timeshard.halo_rows(2000) == mrp_b200.halo_rows(mrp_b200.default_params(norm_window=2000)) == 2000 + 384 + 32
    assert [timeshard.shard_bounds(10, 3, r) for r in range(3)] == [(0, 4), (4, 7), (7, 10)]
    full = synth.make_batch([synth.SynthSpec(n_rows=600, seed=3)])
    sh = timeshard.shard_of(full, 200, 400)
    with pytest.raises(ValueError):
        timeshard.tail_for_right(sh, 500)                  # a shard shorter than the halo cannot serve its neighbour
