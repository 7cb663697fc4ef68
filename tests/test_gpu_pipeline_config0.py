"""GPU: BASELINE.json configs[0] end to end — the reference's OWN features_to_csv pipeline (EventParser -> OrderEngine ->
OrderBookManager -> FeatureEngine -> DualFeaturePipeline::run -> CsvWriter) on a seeded synthetic MBO day, run twice:

  oracle/_ref/libmrp_refpipe.so   the unmodified reference (the checker),
  oracle/_ref/libmrp_gpupipe.so   the same sources with the INTEGRATION.md §3 patch applied by sed to temporary copies
                                  (oracle/Makefile): GpuFeatureProcessor (compiled against the reference's real headers)
                                  -> C ABI -> CUDA, the reference's CsvWriter still the sink.

The four CSV files must agree: identical row sets, timestamps and bit-exact columns byte for byte; the other raw columns
within 1e-9 relative; the z-scores with the conditioning-aware check of the parity tests.  (databento-cpp is not vendored
in the reference, so oracle/mbo_synth_reader.cpp stands in for DbnMboReader at link time — SURVEY.md §8c.)"""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

import mrp_b200
from oracle import binding as ob

from helpers import nan_pattern_equal, norm_check, parity_record, raw_errors

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REFPIPE = os.path.join(ROOT, "oracle", "_ref", "libmrp_refpipe.so")
GPUPIPE = os.path.join(ROOT, "oracle", "_ref", "libmrp_gpupipe.so")


def _run(lib, root, events, seed):
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "run_pipe.py"), lib, str(root), "--events", str(events), "--seed", str(seed)],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-2000:]
    info = json.loads(r.stdout.strip().splitlines()[-1])
    assert info["rc"] == 0, r.stderr[-2000:]
    return info


def _read(path):
    import pandas as pd

    d = pd.read_csv(path)
    assert list(d.columns) == ["timestamp_ns", "instrument"] + list(mrp_b200.FEATURES)
    return d["timestamp_ns"].to_numpy(), d["instrument"].to_numpy(), d[list(mrp_b200.FEATURES)].to_numpy(dtype=np.float64).T.copy()


@pytest.mark.skipif(not (os.path.exists(REFPIPE) and os.path.exists(GPUPIPE)), reason="oracle/_ref pipelines were not built (no /root/reference at build time)")
def test_reference_pipeline_with_and_without_the_patch(tmp_path):
    events, seed = 600_000, 24301
    ref = _run(REFPIPE, tmp_path / "ref", events, seed)
    gpu = _run(GPUPIPE, tmp_path / "gpu", events, seed)
    report = {}
    for name in ("base_SPY", "future_ES"):
        ts_r, ins_r, raw_r = _read(os.path.join(ref["dir"], name + "_raw.csv"))
        ts_g, ins_g, raw_g = _read(os.path.join(gpu["dir"], name + "_raw.csv"))
        _, _, norm_r = _read(os.path.join(ref["dir"], name + "_norm.csv"))
        _, _, norm_g = _read(os.path.join(gpu["dir"], name + "_norm.csv"))
        n = raw_r.shape[1]
        assert n > 1500 and raw_g.shape[1] == n
        assert np.array_equal(ts_r, ts_g) and np.array_equal(ins_r, ins_g)
        # bit-exact columns: the text is identical
        for col in ("midprice", "market_depth", "signed_volume_pressure", "order_arrival_rate", "depth_imbalance", "price_gap"):
            c = mrp_b200.FEATURES.index(col)
            assert np.array_equal(raw_r[c], raw_g[c], equal_nan=True), (name, col)
        assert nan_pattern_equal(raw_g, raw_r) and nan_pattern_equal(norm_g, norm_r)
        re = raw_errors(raw_g, raw_r)
        assert (re <= 1.0).all(), (name, {mrp_b200.FEATURES[i]: re[i] for i in np.nonzero(re > 1.0)[0]})
        wa, wb, drift, frac = norm_check(norm_g, raw_r, norm_r, [0, n], 30000, 1e-9)
        assert (wa <= 1.0).all() and (wb <= 1.0).all(), (name, wa, wb)
        exact = [mrp_b200.FEATURES[c] for c in range(19) if np.array_equal(raw_r[c], raw_g[c], equal_nan=True)]
        report[name] = {"rows": n, "raw_bit_exact_columns": exact, "raw_worst_error_over_tolerance": float(re.max()),
                        "z_worst_ratio_vs_exact": float(wa.max()), "z_worst_ratio_vs_reference": float(wb.max())}
    parity_record("config0_pipeline", events_per_instrument=events, seconds_reference=ref["seconds"], seconds_patched=gpu["seconds"], files=report)
    print("config0:", json.dumps(report)[:600], "wall: reference %.2f s, patched %.2f s" % (ref["seconds"], gpu["seconds"]))
