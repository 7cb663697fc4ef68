"""GPU: two time shards of one stream (halo-extended slices, rebased cursors) run one after the other on the same B200
must reproduce the whole-stream GPU run on their own rows, and the oracle."""
import numpy as np
import pytest
import torch

import mrp_b200
from mrp_b200 import synth, timeshard
from oracle import binding as ob

from helpers import raw_errors

pytestmark = pytest.mark.gpu


def _run_device(batch, p):
    eng = mrp_b200.Engine(p, device=0)
    dev = {k: (v.cuda().contiguous() if isinstance(v, torch.Tensor) else v) for k, v in batch.items()}
    eng.bind_device(dev)
    eng.run()
    out = eng.download()
    eng.close()
    return out


@pytest.mark.parametrize("variable", [False, True])
def test_time_shards_match_whole_stream(variable):
    W, N, world = 2000, 9000, 2
    p = mrp_b200.default_params(norm_window=W, emit_dirs=1, emit_counts=1)
    full = synth.make_batch([synth.SynthSpec(n_rows=N, seed=91, variable_cursors=variable)])
    full.pop("trade_size")
    whole = _run_device(full, p)
    ref = ob.run_oracle(mrp_b200.HostView(synth.to_numpy(full)), p)
    H = timeshard.halo_rows(W)
    shards = [timeshard.shard_of(full, *timeshard.shard_bounds(N, world, r)) for r in range(world)]
    for r in range(world):
        r0, r1 = timeshard.shard_bounds(N, world, r)
        halo = timeshard.tail_for_right(shards[r - 1], H) if r > 0 else None   # what the left neighbour would send
        ext, skip = timeshard.extend(shards[r], halo)
        got = _run_device(ext, p)
        assert np.array_equal(got["bid_dir"][:, skip:], whole["bid_dir"][:, r0:r1])
        assert np.array_equal(got["counts"][:, skip:], whole["counts"][:, r0:r1])
        re = raw_errors(got["raw"][:, skip:], ref["raw"][:, r0:r1])
        assert (re <= 1.0).all(), re
        dz = np.abs(got["norm"][:, skip:] - whole["norm"][:, r0:r1]) / np.maximum(1.0, np.abs(whole["norm"][:, r0:r1]))
        assert np.nanmax(dz) < 1e-9, np.nanmax(dz)


def test_boundary_rows_at_the_real_window(monkeypatch):
    """configs[4] at the reference's own 30 000-row normaliser window: the +-2 000 rows round the shard
    boundary, against the checker restarted more than one window + warm-up earlier, and against the whole-stream run."""
    import os

    from helpers import ORACLE_WARM_ROWS, column_report, nan_pattern_equal, norm_check, oracle_on_windows, parity_record

    monkeypatch.setenv("MRP_FG_SEGMENT_ROWS", "15360")
    monkeypatch.setenv("MRP_FG_NORM_TILE", "2000")
    W, N, world = 30000, 140_000, 2
    p = mrp_b200.default_params(emit_dirs=1, emit_counts=1)
    full = synth.make_batch([synth.SynthSpec(n_rows=N, seed=0x5EED)])
    full.pop("trade_size")
    whole = _run_device(full, p)
    H = timeshard.halo_rows(W)
    shards = [timeshard.shard_of(full, *timeshard.shard_bounds(N, world, r)) for r in range(world)]
    r0, r1 = timeshard.shard_bounds(N, world, 1)
    ext, skip = timeshard.extend(shards[1], timeshard.tail_for_right(shards[0], H))
    assert skip == H
    got = _run_device(ext, p)
    lo, hi = r0 - 2000, r0 + 2000
    cores = os.cpu_count() or 1
    kind = "reference" if ob.ref_available(10) else "port"
    runner = (lambda hv, pp: ob.run_reference(hv, None, use_book=False, nthreads=cores)) if kind == "reference" else None
    ref = oracle_on_windows(synth.to_numpy(full), [(lo, hi)], p, warm=2 * W + 2048, runner=runner)[0]
    # rows before the boundary come from shard 0 == the whole-stream run; rows after it from the halo-extended shard 1
    g_raw = np.concatenate([whole["raw"][:, lo:r0], got["raw"][:, skip:skip + hi - r0]], axis=1)
    g_norm = np.concatenate([whole["norm"][:, lo:r0], got["norm"][:, skip:skip + hi - r0]], axis=1)
    assert np.array_equal(got["bid_dir"][:, skip:skip + hi - r0], ref["bid_dir"][:, r0 - lo:])
    assert np.array_equal(got["counts"][:, skip:skip + hi - r0], whole["counts"][:, r0:hi])
    assert nan_pattern_equal(g_raw, ref["raw"]) and nan_pattern_equal(g_norm, ref["norm"])
    re = raw_errors(g_raw, ref["raw"])
    assert (re <= 1.0).all(), re
    sk = ref["skip"]
    hist = np.concatenate([ref["norm_with_warm"][:, :sk], g_norm], axis=1)
    wa, wb, drift, frac = norm_check(hist, ref["raw_with_warm"], ref["norm_with_warm"], [0, hist.shape[1]], W, 1e-9, first_row=sk)
    assert (wa <= 1.0).all() and (wb <= 1.0).all(), (wa, wb)
    # and the shard against the whole-stream GPU run on its own rows
    dz = np.abs(got["norm"][:, skip:] - whole["norm"][:, r0:r1]) / np.maximum(1.0, np.abs(whole["norm"][:, r0:r1]))
    assert np.nanmax(dz) < 1e-9, np.nanmax(dz)
    rep = column_report(g_raw, g_norm, ref["raw"], ref["norm"], ref["raw_with_warm"], ref["norm_with_warm"], sk, W)
    parity_record("timeshard_boundary_W30000", checker=kind, rows=N, boundary=r0, compared_rows=[lo, hi], halo_rows=H,
                  shard_vs_whole_max_dz=float(np.nanmax(dz)), columns=rep)


@pytest.mark.parametrize("variable", [False, True])
def test_halo_through_the_c_abi(variable):
    """mrp_fg_halo_query / _export / _import: shard 0's packed tail lands in shard 1's headroom; shard 1's own rows equal the
    whole-stream run (what a C++ host with cudaMemcpyPeerAsync or ncclSend / ncclRecv between the two calls gets)."""
    W, N, world = 2000, 9000, 2
    p = mrp_b200.default_params(norm_window=W, emit_dirs=1, emit_counts=1)
    full = synth.make_batch([synth.SynthSpec(n_rows=N, seed=93, variable_cursors=variable)])
    full.pop("trade_size")
    whole = _run_device(full, p)
    shards = [timeshard.shard_of(full, *timeshard.shard_bounds(N, world, r)) for r in range(world)]
    shards = [{k: (v.cuda() if isinstance(v, torch.Tensor) else v) for k, v in sh.items()} for sh in shards]
    e0 = mrp_b200.Engine(p, device=0)
    b0, skip0 = timeshard.layout_with_headroom(shards[0], None)
    e0.bind_device(b0)
    info = e0.halo_query()
    assert info.n_rows == mrp_b200.halo_rows(p) == e0.halo_rows()
    buf = torch.empty(int(info.bytes), dtype=torch.uint8, device="cuda")
    e0.halo_export(buf.data_ptr(), buf.numel())
    e1 = mrp_b200.Engine(p, device=0)
    b1, skip1 = timeshard.layout_with_headroom(shards[1], info)
    e1.bind_device(b1)
    e1.halo_import(buf.data_ptr(), info)
    e1.run()
    got = e1.download()
    r0, r1 = timeshard.shard_bounds(N, world, 1)
    assert skip1 == info.n_rows and got["raw"].shape[1] - skip1 == r1 - r0
    assert np.array_equal(got["bid_dir"][:, skip1:], whole["bid_dir"][:, r0:r1])
    assert np.array_equal(got["counts"][:, skip1:], whole["counts"][:, r0:r1])
    scale = np.nanmax(np.abs(whole["raw"]), axis=1, keepdims=True)
    assert np.allclose(got["raw"][:, skip1:], whole["raw"][:, r0:r1], rtol=1e-10, atol=1e-13 * scale.max(), equal_nan=True)
    dz = np.abs(got["norm"][:, skip1:] - whole["norm"][:, r0:r1]) / np.maximum(1.0, np.abs(whole["norm"][:, r0:r1]))
    assert np.nanmax(dz) < 1e-9, np.nanmax(dz)
    # own-row moments: the halo rows stay out (their features are recomputed by the left neighbour too)
    m_own = e1.moments(rows=(skip1, got["raw"].shape[1]))
    assert np.allclose(m_own[:, 0], np.isfinite(whole["raw"][:, r0:r1]).sum(1))
    # the import must refuse a halo that does not fit
    small = mrp_b200.Engine(p, device=0)
    small.bind_device(b0)
    bad = mrp_b200.HaloInfo(10 ** 9, info.n_ticks, info.n_trades, info.bytes)
    with pytest.raises(mrp_b200.MrpError):
        small.halo_import(buf.data_ptr(), bad)
    for e in (e0, e1, small):
        e.close()


def test_poison_travels_between_shards():
    """mrp_fg_poison_state / mrp_fg_poison_apply: a NaN in shard 0 poisons shard 1's z-scores as in the whole-stream run."""
    W, N, world = 2000, 12000, 2
    p = mrp_b200.default_params(norm_window=W)
    full = synth.make_batch([synth.SynthSpec(n_rows=N, seed=94)])
    full.pop("trade_size")
    full["ask_px"][0, 300] = float("nan")
    whole = _run_device(full, p)
    assert np.isnan(whole["norm"][:, -1]).sum() >= 2
    shards = [timeshard.shard_of(full, *timeshard.shard_bounds(N, world, r)) for r in range(world)]
    shards = [{k: (v.cuda() if isinstance(v, torch.Tensor) else v) for k, v in sh.items()} for sh in shards]
    e0 = mrp_b200.Engine(p, device=0)
    b0, _ = timeshard.layout_with_headroom(shards[0], None)
    e0.bind_device(b0)
    e0.run()
    info = e0.halo_query()
    buf = torch.empty(int(info.bytes), dtype=torch.uint8, device="cuda")
    e0.halo_export(buf.data_ptr(), buf.numel())
    state0 = e0.poison_state([0])
    e1 = mrp_b200.Engine(p, device=0)
    b1, skip = timeshard.layout_with_headroom(shards[1], info)
    e1.bind_device(b1)
    e1.halo_import(buf.data_ptr(), info)
    e1.run()
    r0, r1 = timeshard.shard_bounds(N, world, 1)
    before = e1.download()
    assert not np.array_equal(np.isnan(before["norm"][:, skip:]), np.isnan(whole["norm"][:, r0:r1]))   # the halo starts after row 300
    e1.poison_apply(state0[0], [r0 - skip])
    got = e1.download()
    assert np.array_equal(np.isnan(got["norm"][:, skip:]), np.isnan(whole["norm"][:, r0:r1]))
    fin = np.isfinite(whole["norm"][:, r0:r1])
    assert np.allclose(got["norm"][:, skip:][fin], whole["norm"][:, r0:r1][fin], rtol=0, atol=1e-7)
    e0.close()
    e1.close()
