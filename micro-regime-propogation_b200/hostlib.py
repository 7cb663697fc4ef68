"""ctypes view of host/libmrp_host.so — the C++20 adaptor (GpuFeatureProcessor, CsvWriter) driven from Python by
the tests and the validation harness.  The library itself is the product's host side; this file adds nothing to it."""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

import numpy as np

from .cabi import HostView, SoaView

HERE = os.path.dirname(os.path.abspath(__file__))
HOST_LIB_PATH = os.path.join(HERE, "host", "libmrp_host.so")

_lib = None


def load_host() -> C.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(HOST_LIB_PATH):
            raise RuntimeError(f"{HOST_LIB_PATH} not found: build it with `make -C micro-regime-propogation_b200/host`")
        lib = C.CDLL(HOST_LIB_PATH)
        lib.mrp_host_last_error.restype = C.c_char_p
        lib.mrp_host_write_csv.argtypes = [C.c_char_p, C.c_char_p, C.c_uint64, C.c_char_p, C.c_char_p, C.c_int64] + [C.c_void_p] * 4 + [C.c_int]
        lib.mrp_host_write_binary.argtypes = lib.mrp_host_write_csv.argtypes
        lib.mrp_host_features_to_csv.argtypes = [C.POINTER(SoaView), C.c_char_p, C.c_char_p, C.c_char_p, C.c_char_p, C.c_uint64,
                                                 C.c_uint64, C.c_int, C.c_int]
        lib.mrp_host_rows_by_reference_calls.argtypes = [C.POINTER(SoaView), C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_int]
        _lib = lib
    return _lib


def csv_directory(project_root: str, interval_ns: int, date: str, window: int = 30000, events: int = 1500) -> str:
    return os.path.join(project_root, "data", "output_Snapshot%f_Window%d_Events%d" % (interval_ns / 1e9, window, events), date)


def write_csv(project_root: str, base_filename: str, interval_ns: int, date: str, instrument: str, grid_ts: np.ndarray,
              raw: np.ndarray, norm: np.ndarray, row_ts: Optional[np.ndarray] = None, bulk: bool = False) -> str:
    lib = load_host()
    grid_ts = np.ascontiguousarray(grid_ts, dtype=np.uint64)
    raw = np.ascontiguousarray(raw, dtype=np.float64)
    norm = np.ascontiguousarray(norm, dtype=np.float64)
    rts = np.ascontiguousarray(row_ts, dtype=np.uint64) if row_ts is not None else None
    rc = lib.mrp_host_write_csv(project_root.encode(), base_filename.encode(), interval_ns, date.encode(), instrument.encode(),
                                raw.shape[1], grid_ts.ctypes.data, rts.ctypes.data if rts is not None else None,
                                raw.ctypes.data, norm.ctypes.data, int(bulk))
    if rc != 0:
        raise RuntimeError(lib.mrp_host_last_error().decode())
    return csv_directory(project_root, interval_ns, date)


def write_binary(project_root: str, base_filename: str, interval_ns: int, date: str, instrument: str, grid_ts: np.ndarray,
                 raw: np.ndarray, norm: np.ndarray, row_ts: Optional[np.ndarray] = None, bulk: bool = True) -> str:
    """The rows through microregime::BinarySink: NPY files in the reference's directory (np.load(..., mmap_mode="r"))."""
    lib = load_host()
    grid_ts = np.ascontiguousarray(grid_ts, dtype=np.uint64)
    raw = np.ascontiguousarray(raw, dtype=np.float64)
    norm = np.ascontiguousarray(norm, dtype=np.float64)
    rts = np.ascontiguousarray(row_ts, dtype=np.uint64) if row_ts is not None else None
    rc = lib.mrp_host_write_binary(project_root.encode(), base_filename.encode(), interval_ns, date.encode(), instrument.encode(),
                                   raw.shape[1], grid_ts.ctypes.data, rts.ctypes.data if rts is not None else None,
                                   raw.ctypes.data, norm.ctypes.data, int(bulk))
    if rc != 0:
        raise RuntimeError(lib.mrp_host_last_error().decode())
    return csv_directory(project_root, interval_ns, date)


def load_binary(directory: str, base_filename: str, mmap: bool = True):
    """What the Python stage does instead of pandas.read_csv(<base>_norm.csv): (timestamps [n], raw [19, n], norm [19, n], column names),
    memory-mapped.  `norm.T` is the (n, 19) matrix classifier.py:15-37 builds from the CSV."""
    import json

    mode = "r" if mmap else None
    with open(os.path.join(directory, base_filename + "_columns.json")) as f:
        meta = json.load(f)
    return (np.load(os.path.join(directory, base_filename + "_timestamp_ns.npy"), mmap_mode=mode),
            np.load(os.path.join(directory, base_filename + "_raw.npy"), mmap_mode=mode),
            np.load(os.path.join(directory, base_filename + "_norm.npy"), mmap_mode=mode), meta)


def features_to_csv(hv: HostView, project_root: str, base_filename: str, date: str, instrument: str, first_grid_ns: int,
                    interval_ns: int, device: int = 0, bulk: bool = False) -> str:
    """Stage one stream through microregime::GpuFeatureProcessor (ticks, trades and snapshots in event order), run the
    GPU pass and write the reference-layout CSV files.  Raises when no GPU is present (no CPU fallback)."""
    lib = load_host()
    rc = lib.mrp_host_features_to_csv(C.byref(hv.view), project_root.encode(), base_filename.encode(), date.encode(),
                                      instrument.encode(), first_grid_ns, interval_ns, device, int(bulk))
    if rc != 0:
        raise RuntimeError(lib.mrp_host_last_error().decode())
    return csv_directory(project_root, interval_ns, date)


def rows_by_reference_calls(hv: HostView, first_row: int = 0, params=None, device: int = 0):
    """Replay one stream through the adaptor's reference-signature calls — GetRawFeatureSet(snapshot) then
    GetProcessedFeatureSet(raw), one row per call (feature_processor.hpp:14-15) — for the rows >= first_row (earlier rows
    are only staged).  Returns (raw, norm), each [19][n - first_row]."""
    lib = load_host()
    m = hv.n_rows - first_row
    raw = np.empty((19, m), dtype=np.float64)
    norm = np.empty((19, m), dtype=np.float64)
    rc = lib.mrp_host_rows_by_reference_calls(C.byref(hv.view), C.byref(params) if params is not None else None, first_row,
                                              raw.ctypes.data, norm.ctypes.data, device)
    if rc != 0:
        raise RuntimeError(lib.mrp_host_last_error().decode())
    return raw, norm
