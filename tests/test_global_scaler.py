"""The Python stage's StandardScaler over the z-scored columns (regime_classifier/python/classifier.py:35-37,
data_checks.py:78-80) as combinable moments: host rule vs sklearn (CPU), two ranks combining with one all-reduce (gloo, CPU),
and the CUDA moments / scaling kernels vs sklearn on the GPU's own z-scores (GPU)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp
from sklearn.preprocessing import StandardScaler

import mrp_b200


def _moments(x):   # [19, n] -> [19, 3], NaN cells missing
    ok = ~np.isnan(x)
    z = np.where(ok, x, 0.0)
    return np.stack([ok.sum(1).astype(np.float64), z.sum(1), (z * z).sum(1)], 1)


def _matrix(n, seed):
    rng = np.random.default_rng(seed)
    x = rng.standard_normal((19, n)) * rng.uniform(0.2, 3.0, (19, 1)) + rng.uniform(-1, 1, (19, 1))
    x[4, rng.integers(0, n, n // 7)] = np.nan      # a column with missing cells (aggressor_bias before the first trade)
    x[9] = 0.25                                     # a constant column: sklearn leaves it unscaled
    return x


def test_mean_scale_rule_matches_sklearn():
    x = _matrix(5000, 1)
    ms = mrp_b200.mean_scale(_moments(x))
    sk = StandardScaler().fit(x.T)
    assert np.allclose(ms[:, 0], sk.mean_, rtol=1e-12, atol=1e-15)
    assert np.allclose(ms[:, 1], sk.scale_, rtol=1e-10)
    assert ms[9, 1] == 1.0 and sk.scale_[9] == 1.0


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, results):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        parts = [_matrix(3000, 10), _matrix(4500, 11)]           # two "dates", one per rank
        mom = torch.from_numpy(_moments(parts[rank]))
        dist.all_reduce(mom)                                     # the single collective: 57 doubles
        ms = mrp_b200.mean_scale(mom.numpy())
        sk = StandardScaler().fit(np.concatenate(parts, 1).T)
        assert np.allclose(ms[:, 0], sk.mean_, rtol=1e-12, atol=1e-15) and np.allclose(ms[:, 1], sk.scale_, rtol=1e-10)
        mine = (parts[rank] - ms[:, :1]) / ms[:, 1:]
        whole = sk.transform(np.concatenate(parts, 1).T).T
        lo = 0 if rank == 0 else 3000
        assert np.allclose(mine, whole[:, lo:lo + parts[rank].shape[1]], rtol=1e-9, atol=1e-12, equal_nan=True)
        results[rank] = "ok"
    finally:
        dist.destroy_process_group()


def test_two_ranks_combine_moments_with_one_allreduce():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    results = ctx.Manager().dict()
    procs = [ctx.Process(target=_worker, args=(r, world, port, results)) for r in range(world)]
    for pr in procs:
        pr.start()
    for pr in procs:
        pr.join(300)
    assert all(pr.exitcode == 0 for pr in procs), [pr.exitcode for pr in procs]
    assert dict(results) == {0: "ok", 1: "ok"}


@pytest.mark.gpu
def test_gpu_moments_and_scaling_match_sklearn():
    from mrp_b200 import synth

    from helpers import make_host_batch

    specs = [synth.SynthSpec(n_rows=n, seed=70 + i, warm_trades=0 if i == 0 else 1500) for i, n in enumerate((9000, 4001))]
    hv = mrp_b200.HostView(make_host_batch(specs))
    eng = mrp_b200.Engine(mrp_b200.default_params(), device=0)
    eng.upload(hv)
    eng.run()
    z = eng.download()["norm"]
    mom = eng.norm_moments()
    assert np.array_equal(mom[:, 0], (~np.isnan(z)).sum(1))
    ref = _moments(z)
    assert np.allclose(mom, ref, rtol=1e-12, atol=1e-9)
    assert np.array_equal(mom, eng.norm_moments())              # deterministic: same partial sums, same order
    scaled = eng.standardize(mom)
    eng.close()
    sk = StandardScaler().fit_transform(z.T).T
    assert np.array_equal(np.isnan(scaled), np.isnan(sk))
    assert np.allclose(scaled, sk, rtol=1e-9, atol=1e-10, equal_nan=True)
