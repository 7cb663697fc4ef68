"""vol-of-vol extension (BASELINE.json configs[3]).  PARITY UNPINNED: the reference has no such feature, so the chain is
definition (include/mrp_fg.h) -> oracle restatement (direct window loops, checked here against numpy) -> CUDA (checked
against the oracle on the GPU)."""
import numpy as np
import pytest

import mrp_b200
from mrp_b200 import synth
from oracle import binding as ob

from helpers import make_host_batch, nan_pattern_equal, norm_check

HORIZONS = (600, 3000, 30000)


def test_oracle_restatement_matches_numpy():
    rng = np.random.default_rng(3)
    x = np.abs(rng.standard_normal(3000)) * 1e-4 + 2e-4
    ro = [0, 1100, 3000]
    raw, norm = ob.vol_of_vol(x, ro, (64, 320), norm_window=640)
    for k, h in enumerate((64, 320)):
        for t in (0, 1, 63, 64, 500, 1099, 1100, 1101, 1163, 2999):
            a = 0 if t < 1100 else 1100
            w = x[max(a, t - h + 1):t + 1]
            assert raw[k, t] == pytest.approx(w.std() if len(w) > 1 else 0.0, rel=1e-11, abs=1e-18)
            v = raw[k, max(a, t - 640 + 1):t + 1]
            z = (raw[k, t] - v.mean()) / (v.std() if v.var() > 0 else 1.0)
            assert norm[k, t] == pytest.approx(z, rel=1e-7, abs=1e-9)
    # first row of every stream: one value, variance 0 -> std 0, z 0
    assert raw[:, [0, 1100]].tolist() == [[0.0, 0.0], [0.0, 0.0]] and norm[:, [0, 1100]].tolist() == [[0.0, 0.0], [0.0, 0.0]]


def _gpu_case(specs, horizons, params=None, mode="upload"):
    from helpers import run_gpu  # noqa: F401  (keeps the import next to its only GPU use)

    p = params or mrp_b200.default_params()
    hv = mrp_b200.HostView(make_host_batch(specs))
    eng = mrp_b200.Engine(p, device=0)
    eng.upload(hv)
    eng.run()
    got = eng.download()
    raw, norm = eng.vol_of_vol(horizons)
    eng.close()
    ro = hv.arrays["row_offsets"]
    eraw, enorm = ob.vol_of_vol(got["raw"][mrp_b200.FEATURES.index("ewm_volatility")], ro, horizons, p.norm_window)
    assert nan_pattern_equal(raw, eraw) and nan_pattern_equal(norm, enorm)
    scale = np.maximum(np.nanmax(np.abs(eraw), axis=1, keepdims=True), 1e-300)
    err = np.abs(raw - eraw) / np.maximum(np.abs(eraw), 1e-6 * scale)
    assert np.nanmax(err) < 1e-9, np.nanmax(err)                      # the fp64 tolerance of the 19 reference columns
    wa, wb, _, frac = norm_check(norm, eraw, enorm, ro, p.norm_window, 1e-9)   # conditioning-aware, as for their z-scores
    assert frac > 0.5 and (wa <= 1.0).all() and (wb <= 1.0).all(), (wa, wb, frac)
    return raw, norm


@pytest.mark.gpu
def test_gpu_three_horizons_one_stream():
    _gpu_case([synth.SynthSpec(n_rows=33500, seed=61)], HORIZONS)


@pytest.mark.gpu
def test_gpu_ragged_streams_and_short_windows():
    specs = [synth.SynthSpec(n_rows=n, seed=62 + i, variable_cursors=bool(i & 1)) for i, n in enumerate((1, 2, 700, 2501, 1030))]
    _gpu_case(specs, (64, 320, 640), params=mrp_b200.default_params(norm_window=640))


@pytest.mark.gpu
def test_gpu_rejects_windows_without_a_common_tile():
    hv = mrp_b200.HostView(make_host_batch([synth.SynthSpec(n_rows=500, seed=63)]))
    eng = mrp_b200.Engine(mrp_b200.default_params(), device=0)
    eng.upload(hv)
    eng.run()
    with pytest.raises(mrp_b200.MrpError):
        eng.vol_of_vol((601, 3000, 30000))
    eng.close()
