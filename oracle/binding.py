"""ctypes bindings of the CHECKERS: oracle/libmrp_oracle.so (plain-C restatement) and, when
present, oracle/_ref/libmrp_ref*.so (the unmodified reference compiled from /root/reference).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs, never by the product package.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from typing import Dict, Optional

import numpy as np

from mrp_b200.cabi import NUM_COUNTS, NUM_FEATURES, HostView, Params, SoaView, default_params

HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE_SO = os.path.join(HERE, "libmrp_oracle.so")
REF_SO = os.path.join(HERE, "_ref", "libmrp_ref.so")
REF_L50_SO = os.path.join(HERE, "_ref", "libmrp_ref_L50.so")


def build(quiet: bool = True) -> None:
    """(Re)build the checkers with oracle/Makefile (reference part only if /root/reference exists)."""
    env = dict(os.environ)
    env.pop("CC", None)   # the image's CC/CXX wrappers lack libgomp: the Makefile names the PATH compilers itself
    env.pop("CXX", None)
    subprocess.run(["make", "-C", HERE, "all"], check=True, env=env,
                   stdout=subprocess.DEVNULL if quiet else None)


_oracle = None
_refs: Dict[str, C.CDLL] = {}


def oracle_lib() -> C.CDLL:
    global _oracle
    if _oracle is None:
        if not os.path.exists(ORACLE_SO):
            build()
        lib = C.CDLL(ORACLE_SO)
        lib.mrp_oracle_run.argtypes = [C.POINTER(Params), C.POINTER(SoaView)] + [C.c_void_p] * 5 + [C.c_int]
        _oracle = lib
    return _oracle


def ref_available(levels: int = 10) -> bool:
    return os.path.exists(REF_SO if levels == 10 else REF_L50_SO) and levels in (10, 50)


def ref_lib(levels: int = 10) -> C.CDLL:
    path = REF_SO if levels == 10 else REF_L50_SO
    if path not in _refs:
        lib = C.CDLL(path)
        lib.mrp_ref_run.argtypes = [C.POINTER(SoaView), C.c_void_p, C.c_int] + [C.c_void_p] * 6 + [C.c_int]
        assert lib.mrp_ref_depth_levels() == levels
        _refs[path] = lib
    return _refs[path]


def _ptr(a: Optional[np.ndarray]):
    return C.c_void_p(a.ctypes.data) if a is not None else C.c_void_p()


def run_oracle(hv: HostView, params: Optional[Params] = None, nthreads: int = 0,
               want_dirs: bool = True, want_counts: bool = True) -> Dict[str, np.ndarray]:
    """The C restatement on a host batch."""
    p = params or default_params()
    n, L = hv.n_rows, p.depth_levels
    out = {
        "raw": np.empty((NUM_FEATURES, n), np.float64), "norm": np.empty((NUM_FEATURES, n), np.float64),
        "bid_dir": np.empty((L, n), np.int8) if want_dirs else None,
        "ask_dir": np.empty((L, n), np.int8) if want_dirs else None,
        "counts": np.zeros((NUM_COUNTS, n), np.int32) if want_counts else None,
    }
    rc = oracle_lib().mrp_oracle_run(C.byref(p), C.byref(hv.view), _ptr(out["raw"]), _ptr(out["norm"]),
                                     _ptr(out["bid_dir"]), _ptr(out["ask_dir"]), _ptr(out["counts"]),
                                     nthreads or (os.cpu_count() or 1))
    if rc != 0:
        raise RuntimeError(f"mrp_oracle_run failed ({rc})")
    return out


def vol_of_vol(ewm: np.ndarray, row_offsets, horizons, norm_window: int = 30000):
    """vol-of-vol extension, PARITY UNPINNED: the restated definition (direct window loops in extended precision).
    Returns (raw, norm), each [n_horizons][n_rows]."""
    ewm = np.ascontiguousarray(ewm, dtype=np.float64)
    ro = np.ascontiguousarray(row_offsets, dtype=np.int64)
    h = np.ascontiguousarray(horizons, dtype=np.int32)
    n = ewm.shape[0]
    raw = np.empty((len(h), n), np.float64)
    norm = np.empty((len(h), n), np.float64)
    lib = oracle_lib()
    lib.mrp_oracle_vol_of_vol.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
    rc = lib.mrp_oracle_vol_of_vol(_ptr(ewm), _ptr(ro), len(ro) - 1, _ptr(h), len(h), norm_window, _ptr(raw), _ptr(norm))
    if rc != 0:
        raise RuntimeError(f"mrp_oracle_vol_of_vol failed ({rc})")
    return raw, norm


def run_reference(hv: HostView, trade_size: Optional[np.ndarray] = None, use_book: bool = False,
                  nthreads: int = 1) -> Dict[str, np.ndarray]:
    """The compiled, unmodified reference (L = 10 or 50 as hv.levels says)."""
    n, L = hv.n_rows, hv.levels
    lib = ref_lib(L)
    ts = np.ascontiguousarray(trade_size, dtype=np.float64) if trade_size is not None else None
    out = {
        "raw": np.empty((NUM_FEATURES, n), np.float64), "norm": np.empty((NUM_FEATURES, n), np.float64),
        "bid_dir": np.empty((L, n), np.int8), "ask_dir": np.empty((L, n), np.int8),
        "snap_buy": np.empty(n, np.float64), "snap_sell": np.empty(n, np.float64),
    }
    rc = lib.mrp_ref_run(C.byref(hv.view), _ptr(ts), int(use_book), _ptr(out["raw"]), _ptr(out["norm"]),
                         _ptr(out["bid_dir"]), _ptr(out["ask_dir"]), _ptr(out["snap_buy"]), _ptr(out["snap_sell"]),
                         nthreads)
    if rc != 0:
        raise RuntimeError(f"mrp_ref_run failed ({rc})")
    return out


# ---- the reference's own CsvWriter (oracle/_ref/libmrp_refcsv.so, built from features_to_csv.cpp where it lies)
REF_CSV_SO = os.path.join(HERE, "_ref", "libmrp_refcsv.so")


def ref_csv_available() -> bool:
    return os.path.exists(REF_CSV_SO)


def ref_write_csv(project_root: str, base_filename: str, interval_ns: int, date: str, instrument: str,
                  grid_ts: np.ndarray, row_ts: np.ndarray, raw: np.ndarray, norm: np.ndarray) -> str:
    """Format rows with the UNMODIFIED reference CsvWriter; returns the directory it wrote into."""
    lib = C.CDLL(REF_CSV_SO)
    lib.mrp_ref_write_csv.argtypes = [C.c_char_p, C.c_uint64, C.c_char_p, C.c_char_p, C.c_int64] + [C.c_void_p] * 4
    grid_ts = np.ascontiguousarray(grid_ts, dtype=np.uint64)
    row_ts = np.ascontiguousarray(row_ts, dtype=np.uint64)
    raw = np.ascontiguousarray(raw, dtype=np.float64)
    norm = np.ascontiguousarray(norm, dtype=np.float64)
    n = raw.shape[1]
    old = os.environ.get("MRP_REF_PROJECT_ROOT")
    os.environ["MRP_REF_PROJECT_ROOT"] = project_root
    try:
        rc = lib.mrp_ref_write_csv(base_filename.encode(), interval_ns, date.encode(), instrument.encode(), n,
                                   _ptr(grid_ts), _ptr(row_ts), _ptr(raw), _ptr(norm))
    finally:
        if old is None:
            os.environ.pop("MRP_REF_PROJECT_ROOT", None)
        else:
            os.environ["MRP_REF_PROJECT_ROOT"] = old
    if rc != 0:
        raise RuntimeError(f"mrp_ref_write_csv failed ({rc})")
    secs = "%f" % (interval_ns / 1e9)
    return os.path.join(project_root, "data", f"output_Snapshot{secs}_Window30000_Events1500", date)
