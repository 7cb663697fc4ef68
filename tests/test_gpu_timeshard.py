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

    monkeypatch.setenv("MRP_FG_SEGMENT_ROWS", "16384")
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
