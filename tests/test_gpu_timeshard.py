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
