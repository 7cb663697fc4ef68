"""The C++ CsvWriter of the host adaptor against golden CSV files written by the reference's OWN CsvWriter
(features_to_csv.cpp compiled where it lies, tests/golden/make_golden.py): byte-for-byte, both the per-row sink path
(DataReciever::ingest_feature_set) and the threaded column path; plus, where oracle/_ref is present, a live
comparison on random and special values."""
import os

import numpy as np
import pytest

import mrp_b200
from mrp_b200 import hostlib
from oracle import binding as ob

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "csv")


def _read(d, name):
    with open(os.path.join(d, name), "rb") as f:
        return f.read()


@pytest.mark.parametrize("bulk", [False, True], ids=["per_row_sink", "bulk_columns"])
def test_csv_matches_reference_golden(tmp_path, bulk):
    z = np.load(os.path.join(GOLD, "values.npz"))
    d = hostlib.write_csv(str(tmp_path), "base_SYN", 500_000_000, "20250505", "SYN", z["grid"], z["raw"], z["norm"],
                          row_ts=z["row_ts"], bulk=bulk)
    assert os.path.relpath(d, str(tmp_path)) == str(z["rel_dir"])      # directory naming, features_to_csv.cpp:18-24
    for f in ("base_SYN_raw.csv", "base_SYN_norm.csv"):
        assert _read(d, f) == _read(GOLD, f), f


def test_header_and_column_order():
    import mrp_b200

    head = _read(GOLD, "base_SYN_raw.csv").split(b"\n", 1)[0].decode().split(",")
    assert head[:2] == ["timestamp_ns", "instrument"]
    assert tuple(head[2:]) == mrp_b200.FEATURES       # CSV order == FeatureSet order == mrp_fg_feature order


def test_live_against_reference_writer(tmp_path):
    if not ob.ref_csv_available():
        pytest.skip("oracle/_ref/libmrp_refcsv.so not built (needs /root/reference at build time)")
    rng = np.random.default_rng(9)
    n = 5000
    raw = rng.standard_normal((19, n)) * 10.0 ** rng.integers(-300, 300, (19, n))
    norm = rng.standard_normal((19, n))
    specials = [0.0, -0.0, np.inf, -np.inf, np.nan, np.copysign(np.nan, -1), 5e-324, 2.2250738585072014e-308,
                1.7976931348623157e308, 1.0, -1.0, 0.1, 1e15, 9.999999999999999e22, 123456789012345678.0]
    raw[:, :len(specials)] = np.array(specials)[None, :]
    grid = (np.arange(n, dtype=np.uint64) * np.uint64(500_000_000) + np.uint64(1_746_453_700_000_000_000))
    a = ob.ref_write_csv(str(tmp_path / "ref"), "future_ES", 250_000_000, "20250506", "ES", grid, grid, raw, norm)
    for i, bulk in enumerate((False, True)):
        b = hostlib.write_csv(str(tmp_path / f"ours{i}"), "future_ES", 250_000_000, "20250506", "ES", grid, raw, norm, bulk=bulk)
        assert os.path.relpath(a, str(tmp_path / "ref")) == os.path.relpath(b, str(tmp_path / f"ours{i}"))
        for f in ("future_ES_raw.csv", "future_ES_norm.csv"):
            assert _read(a, f) == _read(b, f), (f, bulk)


def test_adaptor_refuses_to_run_without_a_gpu(tmp_path):
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    import mrp_b200
    from mrp_b200 import synth

    hv = mrp_b200.HostView(synth.to_numpy(synth.make_batch([synth.SynthSpec(n_rows=16, seed=1)])))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        hostlib.features_to_csv(hv, str(tmp_path), "base_SYN", "20250505", "SYN", 0, 500_000_000)


def test_binary_sink_is_mappable_and_equals_the_csv(tmp_path):
    """microregime::BinarySink: NPY files in the reference's directory, np.load(mmap_mode="r")-able, holding
    the doubles themselves (the CSV's 16 significant digits lose the last bit or two of most of them)."""
    import pandas as pd

    from mrp_b200 import hostlib

    rng = np.random.default_rng(3)
    n = 1234
    raw = rng.standard_normal((19, n)) * 10.0 ** rng.integers(-12, 6, (19, 1))
    norm = rng.standard_normal((19, n))
    raw[3, 5] = np.nan
    norm[7, 9] = np.inf
    grid = (1_746_453_700_000_000_000 + np.arange(n) * 500_000_000).astype(np.uint64)
    d1 = hostlib.write_binary(str(tmp_path), "base_SYN", 500_000_000, "20250505", "SYN", grid, raw, norm, bulk=True)
    ts, r1, z1, meta = hostlib.load_binary(d1, "base_SYN")
    assert isinstance(r1, np.memmap) and r1.shape == (19, n) and meta["columns"] == list(mrp_b200.FEATURES) and meta["rows"] == n
    assert np.array_equal(ts, grid) and np.array_equal(r1, raw, equal_nan=True) and np.array_equal(z1, norm, equal_nan=True)
    # the per-row DataReciever path writes the same files
    d2 = hostlib.write_binary(str(tmp_path / "rows"), "base_SYN", 500_000_000, "20250505", "SYN", grid, raw, norm, bulk=False)
    _, r2, z2, _ = hostlib.load_binary(d2, "base_SYN", mmap=False)
    assert np.array_equal(r2, raw, equal_nan=True) and np.array_equal(z2, norm, equal_nan=True)
    # and the CSV of the same rows parses to the same matrix (what classifier.py would have read)
    d3 = hostlib.write_csv(str(tmp_path / "csv"), "base_SYN", 500_000_000, "20250505", "SYN", grid, raw, norm, bulk=True)
    df = pd.read_csv(os.path.join(d3, "base_SYN_norm.csv"))
    assert np.allclose(df[list(mrp_b200.FEATURES)].to_numpy().T, np.asarray(z1), rtol=2e-15, atol=0, equal_nan=True)
