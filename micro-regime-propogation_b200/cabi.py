"""ctypes view of include/mrp_fg.h — a thin Python host layer over the C ABI.

The product's host side is C++20 (host/), because the reference is C++; this module exists so
that the parity tests, the benchmark driver and the validation harness can call the SAME C-ABI
shared library (`csrc/libmrp_fg.so`) from Python.  It contains no feature arithmetic and no
fallback: if the library is missing, `load()` raises.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Dict, Optional, Sequence

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "csrc", "libmrp_fg.so")

NUM_FEATURES = 19
FEATURES = (
    "midprice", "log_spread", "log_return", "ewm_volatility", "realized_variance",
    "directional_volatility", "spread_volatility", "ofi", "signed_volume_pressure",
    "order_arrival_rate", "depth_imbalance", "market_depth", "lob_slope", "price_gap",
    "tick_direction_entropy", "reversal_rate", "aggressor_bias", "shannon_entropy", "liquidity_stress",
)
COUNTS = (
    "ret_valid", "ret_up", "tick_up", "tick_down", "tick_total", "trade_pos", "trade_neg",
    "trade_reversals", "trade_total", "trade_dir_sum", "norm_n",
)
NUM_COUNTS = len(COUNTS)
MEM_DEVICE, MEM_HOST = 0, 1
NORM_COMPENSATED, NORM_FAITHFUL = 0, 1


class Params(C.Structure):
    _fields_ = [
        ("depth_levels", C.c_int32), ("mid_window", C.c_int32), ("dir_window", C.c_int32),
        ("spread_divisor", C.c_int32), ("ewm_span", C.c_int32), ("norm_window", C.c_int32),
        ("return_lag", C.c_int32), ("stress_levels", C.c_int32),
        ("ofi_depth_decay", C.c_double), ("ofi_alpha", C.c_double), ("ofi_min_total_vol", C.c_double),
        ("stress_min_size", C.c_double), ("stress_distance_decay", C.c_double), ("stress_alpha", C.c_double),
        ("norm_mode", C.c_int32), ("emit_dirs", C.c_int32), ("emit_counts", C.c_int32), ("reserved", C.c_int32),
    ]


class HaloInfo(C.Structure):
    _fields_ = [("n_rows", C.c_int64), ("n_ticks", C.c_int64), ("n_trades", C.c_int64), ("bytes", C.c_int64)]


class SoaView(C.Structure):
    _fields_ = [
        ("n_streams", C.c_int32), ("memory", C.c_int32),
        ("row_offsets", C.c_void_p), ("tick_offsets", C.c_void_p), ("trade_offsets", C.c_void_p),
        ("ts_ns", C.c_void_p), ("bid_px", C.c_void_p), ("ask_px", C.c_void_p),
        ("bid_sz", C.c_void_p), ("ask_sz", C.c_void_p), ("buy_vol", C.c_void_p), ("sell_vol", C.c_void_p),
        ("adds", C.c_void_p), ("tick_cursor", C.c_void_p), ("trade_cursor", C.c_void_p),
        ("mid", C.c_void_p), ("spread", C.c_void_p), ("trade_dir", C.c_void_p),
    ]


_ARRAY_FIELDS = ("ts_ns", "bid_px", "ask_px", "bid_sz", "ask_sz", "buy_vol", "sell_vol", "adds",
                 "tick_cursor", "trade_cursor", "mid", "spread", "trade_dir")
_OPTIONAL_FIELDS = ("buy_vol", "sell_vol", "adds")   # mrp_fg_build_state can derive them
_HOST_DTYPES = {
    "ts_ns": np.uint64, "bid_px": np.float64, "ask_px": np.float64, "bid_sz": np.int32, "ask_sz": np.int32,
    "buy_vol": np.float64, "sell_vol": np.float64, "adds": np.int32, "tick_cursor": np.uint32,
    "trade_cursor": np.uint32, "mid": np.float64, "spread": np.float64, "trade_dir": np.int8,
}


def default_params(**over) -> Params:
    """Reference constants (common_constants.hpp:8-11, feature_processor.cpp:64,114-116,274-277)."""
    p = Params(10, 1800, 1500, 1500, 1500, 30000, 20, 5, 0.5, 0.2, 1e-6, 5.0, 10.0, 0.1, 0, 0, 0, 0)
    for k, v in over.items():
        setattr(p, k, v)
    return p


class HostView:
    """Keeps numpy arrays alive and exposes them as a mrp_fg_soa_view (memory = HOST)."""

    def __init__(self, batch: Dict[str, object]):
        self.arrays = {}
        for k in _ARRAY_FIELDS:
            if k in _OPTIONAL_FIELDS and batch.get(k) is None:
                continue   # derived on the device (Engine.build_state)
            a = np.ascontiguousarray(np.asarray(batch[k]), dtype=_HOST_DTYPES[k])
            self.arrays[k] = a
        for k in ("row_offsets", "tick_offsets", "trade_offsets"):
            self.arrays[k] = np.ascontiguousarray(np.asarray(batch[k], dtype=np.int64))
        self.n_streams = len(self.arrays["row_offsets"]) - 1
        self.n_rows = int(self.arrays["row_offsets"][-1])
        self.n_ticks = int(self.arrays["tick_offsets"][-1])
        self.n_trades = int(self.arrays["trade_offsets"][-1])
        self.levels = self.arrays["bid_px"].size // max(self.n_rows, 1)
        v = SoaView()
        v.n_streams, v.memory = self.n_streams, MEM_HOST
        for k, a in self.arrays.items():
            setattr(v, k, a.ctypes.data)
        self.view = v


def device_view(batch: Dict[str, object]) -> "tuple[SoaView, list]":
    """mrp_fg_soa_view over CUDA torch tensors (memory = DEVICE).  Returns (view, keepalive)."""
    keep = []
    v = SoaView()
    v.memory = MEM_DEVICE
    for k in ("row_offsets", "tick_offsets", "trade_offsets"):
        a = np.ascontiguousarray(np.asarray(batch[k], dtype=np.int64))
        keep.append(a)
        setattr(v, k, a.ctypes.data)
    v.n_streams = len(keep[0]) - 1
    for k in _ARRAY_FIELDS:
        t = batch.get(k) if k in _OPTIONAL_FIELDS else batch[k]
        if t is None:
            continue   # derived on the device (Engine.build_state)
        assert t.is_cuda and t.is_contiguous(), k
        keep.append(t)
        setattr(v, k, t.data_ptr())
    return v, keep


_lib = None


def load(path: Optional[str] = None) -> C.CDLL:
    """Load libmrp_fg.so.  There is no fallback: a missing library is an error."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    p = path or os.environ.get("MRP_FG_LIB") or LIB_PATH   # MRP_FG_LIB: an experiment build of the same ABI
    if not os.path.exists(p):
        raise RuntimeError(
            f"{p} not found: the CUDA extension is not built (run `python -c 'import __graft_entry__ as g; g.build()'`). "
            "This package has no CPU fallback.")
    lib = C.CDLL(p)
    vp, i32, i64 = C.c_void_p, C.c_int, C.c_int64
    PP = C.POINTER(Params)
    PV = C.POINTER(SoaView)
    lib.mrp_fg_default_params.argtypes = [PP]
    lib.mrp_fg_create.argtypes = [PP, i32, C.POINTER(vp)]
    lib.mrp_fg_destroy.argtypes = [vp]
    lib.mrp_fg_bind_inputs.argtypes = [vp, PV]
    lib.mrp_fg_upload.argtypes = [vp, PV, vp]
    lib.mrp_fg_run.argtypes = [vp, vp]
    lib.mrp_fg_process_host.argtypes = [vp, PV, vp, vp, i64]
    lib.mrp_fg_outputs.argtypes = [vp] + [C.POINTER(vp)] * 5
    lib.mrp_fg_download.argtypes = [vp, vp, vp, vp, vp, vp, vp]
    lib.mrp_fg_download_rows.argtypes = [vp, C.c_int64, C.c_int64, vp, vp, vp]
    lib.mrp_fg_vol_of_vol.argtypes = [vp, vp, C.c_int32, C.POINTER(vp), C.POINTER(vp), vp]
    lib.mrp_fg_download_vol_of_vol.argtypes = [vp, vp, vp, vp]
    lib.mrp_fg_moments.argtypes = [vp, vp, vp]
    lib.mrp_fg_norm_moments.argtypes = [vp, vp, vp]
    lib.mrp_fg_moments_rows.argtypes = [vp, i64, i64, vp, vp]
    lib.mrp_fg_norm_moments_rows.argtypes = [vp, i64, i64, vp, vp]
    lib.mrp_fg_halo_rows.argtypes = [vp]
    lib.mrp_fg_poison_state.argtypes = [vp, vp, vp, vp]
    lib.mrp_fg_poison_apply.argtypes = [vp, vp, vp, vp]
    lib.mrp_fg_build_state.argtypes = [vp, vp, vp, i32, vp]
    lib.mrp_fg_halo_query.argtypes = [vp, C.POINTER(HaloInfo), vp]
    lib.mrp_fg_halo_export.argtypes = [vp, vp, i64, C.POINTER(vp), C.POINTER(HaloInfo), vp]
    lib.mrp_fg_halo_import.argtypes = [vp, vp, C.POINTER(HaloInfo), vp]
    lib.mrp_fg_halo_rows_for.argtypes = [PP]
    lib.mrp_fg_mean_scale.argtypes = [vp, vp]
    lib.mrp_fg_standardize.argtypes = [vp, vp, vp, C.POINTER(vp), vp]
    lib.mrp_fg_sync.argtypes = [vp, vp]
    lib.mrp_fg_num_kernels.argtypes = []
    lib.mrp_fg_kernel_name.argtypes = [i32]
    lib.mrp_fg_kernel_name.restype = C.c_char_p
    lib.mrp_fg_enable_timing.argtypes = [vp, i32]
    lib.mrp_fg_kernel_times_ms.argtypes = [vp, vp]
    lib.mrp_fg_launch_count.argtypes = [vp]
    lib.mrp_fg_launch_count.restype = i64
    lib.mrp_fg_last_error.argtypes = []
    lib.mrp_fg_last_error.restype = C.c_char_p
    lib.mrp_fg_abi_version.argtypes = []
    if path is None:
        _lib = lib
    return lib


class MrpError(RuntimeError):
    pass


class Engine:
    """One mrp_fg_ctx.  Mirrors the C ABI one to one."""

    def __init__(self, params: Optional[Params] = None, device: int = 0):
        self.lib = load()
        self.params = params or default_params()
        self.ctx = C.c_void_p()
        self._keep = None
        self.n_rows = 0
        self._check(self.lib.mrp_fg_create(C.byref(self.params), device, C.byref(self.ctx)))

    def _check(self, rc: int):
        if rc != 0:
            raise MrpError(f"mrp_fg error {rc}: {self.lib.mrp_fg_last_error().decode()}")

    def close(self):
        if self.ctx:
            self.lib.mrp_fg_destroy(self.ctx)
            self.ctx = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- inputs
    def bind_device(self, batch: Dict[str, object]):
        view, keep = device_view(batch)
        self._keep = keep
        self.n_rows = int(np.asarray(batch["row_offsets"])[-1])
        self._check(self.lib.mrp_fg_bind_inputs(self.ctx, C.byref(view)))

    def upload(self, hv: HostView, stream: int = 0):
        self._keep = hv
        self.n_rows = hv.n_rows
        self._check(self.lib.mrp_fg_upload(self.ctx, C.byref(hv.view), C.c_void_p(stream)))

    # -- host in, host out, pipelined over PCIe (mrp_fg_process_host)
    def process_host(self, hv: HostView, raw: Optional[np.ndarray] = None, norm: Optional[np.ndarray] = None, chunk_rows: int = 0,
                     raw_ptr: int = 0, norm_ptr: int = 0):
        """raw / norm: numpy [19, n] outputs (allocated when omitted), or raw_ptr / norm_ptr: addresses of pinned buffers."""
        n = hv.n_rows
        if not raw_ptr and not norm_ptr:
            raw = raw if raw is not None else np.empty((NUM_FEATURES, n), np.float64)
            norm = norm if norm is not None else np.empty((NUM_FEATURES, n), np.float64)
            raw_ptr, norm_ptr = raw.ctypes.data, norm.ctypes.data
        # with raw_ptr / norm_ptr, ONE of the two may be 0: that matrix is not copied back
        self._check(self.lib.mrp_fg_process_host(self.ctx, C.byref(hv.view), C.c_void_p(raw_ptr or None), C.c_void_p(norm_ptr or None), chunk_rows))
        return raw, norm

    # -- rolling state from per-event arrays (mrp_fg_build_state)
    def build_state(self, trade_size=None, add_cursor=None, stream: int = 0):
        """trade_size: int32 [n_trades], add_cursor: uint32 [n_rows]; numpy arrays (host) or CUDA torch tensors (device), or None."""
        def ptr(a, dt):
            if a is None:
                return 0, None
            if isinstance(a, np.ndarray):
                a = np.ascontiguousarray(a, dtype=dt)
                return a.ctypes.data, a
            assert a.is_cuda and a.is_contiguous()
            return a.data_ptr(), a
        ps, ks = ptr(trade_size, np.int32)
        pa, ka = ptr(add_cursor, np.uint32)
        host = any(isinstance(x, np.ndarray) for x in (ks, ka))
        dev = any(x is not None and not isinstance(x, np.ndarray) for x in (ks, ka))
        assert not (host and dev), "trade_size and add_cursor must live in the same memory"
        self._state_keep = (ks, ka)
        self._check(self.lib.mrp_fg_build_state(self.ctx, C.c_void_p(ps), C.c_void_p(pa), MEM_HOST if host else MEM_DEVICE, C.c_void_p(stream)))
        if host:
            self.sync(stream)   # the staging copies read the numpy arrays

    # -- compute
    def run(self, stream: int = 0):
        self._check(self.lib.mrp_fg_run(self.ctx, C.c_void_p(stream)))

    def sync(self, stream: int = 0):
        self._check(self.lib.mrp_fg_sync(self.ctx, C.c_void_p(stream)))

    # -- outputs
    def output_ptrs(self):
        ptrs = [C.c_void_p() for _ in range(5)]
        self._check(self.lib.mrp_fg_outputs(self.ctx, *[C.byref(p) for p in ptrs]))
        return [p.value for p in ptrs]

    def download(self, stream: int = 0, raw=True, norm=True):
        """Returns dict of numpy arrays: raw/norm [19, n], bid_dir/ask_dir [L, n], counts [NUM_COUNTS, n]."""
        n, L = self.n_rows, self.params.depth_levels
        out = {}
        out["raw"] = np.empty((NUM_FEATURES, n), np.float64) if raw else None
        out["norm"] = np.empty((NUM_FEATURES, n), np.float64) if norm else None
        out["bid_dir"] = np.empty((L, n), np.int8) if self.params.emit_dirs else None
        out["ask_dir"] = np.empty((L, n), np.int8) if self.params.emit_dirs else None
        out["counts"] = np.empty((NUM_COUNTS, n), np.int32) if self.params.emit_counts else None
        p = lambda a: C.c_void_p(a.ctypes.data) if a is not None else C.c_void_p()
        self._check(self.lib.mrp_fg_download(self.ctx, p(out["raw"]), p(out["norm"]), p(out["bid_dir"]),
                                             p(out["ask_dir"]), p(out["counts"]), C.c_void_p(stream)))
        self.sync(stream)
        return out

    def download_rows(self, row_begin: int, row_end: int, stream: int = 0):
        """raw / norm of the rows [row_begin, row_end) only: ([19, m], [19, m])."""
        m = row_end - row_begin
        raw = np.empty((NUM_FEATURES, m), np.float64)
        norm = np.empty((NUM_FEATURES, m), np.float64)
        self._check(self.lib.mrp_fg_download_rows(self.ctx, C.c_int64(row_begin), C.c_int64(row_end), C.c_void_p(raw.ctypes.data),
                                                  C.c_void_p(norm.ctypes.data), C.c_void_p(stream)))
        self.sync(stream)
        return raw, norm

    def vol_of_vol(self, horizons, stream: int = 0, download: bool = True):
        """The vol-of-vol extension (no reference counterpart; see include/mrp_fg.h).  Returns (raw, norm) [n_horizons, n] on
        the host, or the two device pointers when download is False."""
        h = np.ascontiguousarray(horizons, dtype=np.int32)
        praw, pnorm = C.c_void_p(), C.c_void_p()
        self._check(self.lib.mrp_fg_vol_of_vol(self.ctx, C.c_void_p(h.ctypes.data), len(h), C.byref(praw), C.byref(pnorm), C.c_void_p(stream)))
        if not download:
            return praw.value, pnorm.value
        raw = np.empty((len(h), self.n_rows), np.float64)
        norm = np.empty((len(h), self.n_rows), np.float64)
        self._check(self.lib.mrp_fg_download_vol_of_vol(self.ctx, C.c_void_p(raw.ctypes.data), C.c_void_p(norm.ctypes.data), C.c_void_p(stream)))
        self.sync(stream)
        return raw, norm

    def moments(self, stream: int = 0, rows: Optional[Sequence[int]] = None) -> np.ndarray:
        """[19, 3] = {non-NaN count, sum x, sum x^2} of the raw columns over all bound rows or over rows = (begin, end)."""
        m = np.empty((NUM_FEATURES, 3), np.float64)
        a, b = rows if rows is not None else (0, self.n_rows)
        self._check(self.lib.mrp_fg_moments_rows(self.ctx, a, b, C.c_void_p(m.ctypes.data), C.c_void_p(stream)))
        return m

    # -- one stream split by time: the tail of this shard as one packed device buffer / a neighbour's tail into the headroom
    def halo_query(self, stream: int = 0) -> "HaloInfo":
        """What the packed tail of the bound stream holds (mrp_fg_halo_query); .bytes sizes the message buffer."""
        info = HaloInfo()
        self._check(self.lib.mrp_fg_halo_query(self.ctx, C.byref(info), C.c_void_p(stream)))
        return info

    def halo_export(self, dst_ptr: int = 0, capacity: int = 0, stream: int = 0):
        """Pack the tail of the bound stream into dst_ptr (device; 0 = a context-owned buffer).  Returns (device pointer, HaloInfo)."""
        ptr, info = C.c_void_p(), HaloInfo()
        self._check(self.lib.mrp_fg_halo_export(self.ctx, C.c_void_p(dst_ptr), capacity, C.byref(ptr), C.byref(info), C.c_void_p(stream)))
        return ptr.value, info

    def halo_import(self, packed_ptr: int, info: "HaloInfo", stream: int = 0):
        """Write a packed halo into the front (headroom) of the bound stream (mrp_fg_halo_import)."""
        self._check(self.lib.mrp_fg_halo_import(self.ctx, C.c_void_p(packed_ptr), C.byref(info), C.c_void_p(stream)))

    # -- the normaliser's "poisoned for ever" state across time shards
    def poison_state(self, row_base=None, stream: int = 0) -> np.ndarray:
        """[n_streams, 19, 3] first NaN / +inf / -inf row per column in absolute stream rows (int64 max = none)."""
        n_streams = len(np.atleast_1d(row_base)) if row_base is not None else 1
        st = np.empty((n_streams, NUM_FEATURES, 3), np.int64)
        rb = np.ascontiguousarray(np.atleast_1d(row_base), dtype=np.int64) if row_base is not None else None
        self._check(self.lib.mrp_fg_poison_state(self.ctx, C.c_void_p(rb.ctypes.data) if rb is not None else C.c_void_p(),
                                                 C.c_void_p(st.ctypes.data), C.c_void_p(stream)))
        return st

    def poison_apply(self, state_before: np.ndarray, row_base=None, stream: int = 0):
        sb = np.ascontiguousarray(state_before, dtype=np.int64)
        rb = np.ascontiguousarray(np.atleast_1d(row_base), dtype=np.int64) if row_base is not None else None
        self._check(self.lib.mrp_fg_poison_apply(self.ctx, C.c_void_p(sb.ctypes.data), C.c_void_p(rb.ctypes.data) if rb is not None else C.c_void_p(),
                                                 C.c_void_p(stream)))

    def halo_rows(self) -> int:
        """Rows of history a time shard / host chunk needs in front of its own rows (mrp_fg_halo_rows)."""
        return int(self.lib.mrp_fg_halo_rows(self.ctx))

    # -- the Python stage's StandardScaler over the z-scored columns (classifier.py:35-37), combinable across GPUs
    def norm_moments(self, stream: int = 0, rows: Optional[Sequence[int]] = None) -> np.ndarray:
        """[19, 3] = {non-NaN count, sum z, sum z^2} per z-scored column (over rows = (begin, end) when given: a time shard's
        own rows); ADD these across contexts / ranks."""
        m = np.empty((NUM_FEATURES, 3), np.float64)
        a, b = rows if rows is not None else (0, self.n_rows)
        self._check(self.lib.mrp_fg_norm_moments_rows(self.ctx, a, b, C.c_void_p(m.ctypes.data), C.c_void_p(stream)))
        return m

    def standardize(self, moments: np.ndarray, stream: int = 0, download: bool = True):
        """(z - mean) / scale with mean / scale from (combined) moments: [19, n] on the host, or the device pointer."""
        m = np.ascontiguousarray(moments, dtype=np.float64)
        out = np.empty((NUM_FEATURES, self.n_rows), np.float64) if download else None
        dev = C.c_void_p()
        self._check(self.lib.mrp_fg_standardize(self.ctx, C.c_void_p(m.ctypes.data), C.c_void_p(out.ctypes.data) if download else C.c_void_p(),
                                                C.byref(dev), C.c_void_p(stream)))
        self.sync(stream)
        return out if download else dev.value

    # -- introspection
    def enable_timing(self, on: bool = True):
        self._check(self.lib.mrp_fg_enable_timing(self.ctx, int(on)))

    def kernel_times_ms(self) -> Dict[str, float]:
        k = self.lib.mrp_fg_num_kernels()
        buf = (C.c_float * k)()
        self._check(self.lib.mrp_fg_kernel_times_ms(self.ctx, C.cast(buf, C.c_void_p)))
        return {self.lib.mrp_fg_kernel_name(i).decode(): float(buf[i]) for i in range(k)}

    def launch_count(self) -> int:
        return int(self.lib.mrp_fg_launch_count(self.ctx))


def plan_shape(row_offsets, params: Optional[Params] = None, sm_count: int = 0):
    """mrp_fg_plan_shape: (rows per segment, rows per normaliser tile) for a batch with these row offsets (host arithmetic, no GPU)."""
    p = params or default_params()
    ro = np.ascontiguousarray(row_offsets, dtype=np.int64)
    seg, tile = C.c_int64(0), C.c_int32(0)
    lib = load()
    lib.mrp_fg_plan_shape.argtypes = [C.c_void_p, C.c_int32, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p]
    rc = lib.mrp_fg_plan_shape(C.byref(p), len(ro) - 1, ro.ctypes.data, sm_count, C.byref(seg), C.byref(tile))
    if rc:
        raise MrpError(lib.mrp_fg_last_error().decode())
    return int(seg.value), int(tile.value)


def plan_host_chunks(rows: int, chunk_rows: int = 0, first_stream: bool = True, last_stream: bool = True):
    """mrp_fg_plan_host_chunks: the chunk lengths mrp_fg_process_host cuts one stream into (host arithmetic, no GPU)."""
    lib = load()
    lib.mrp_fg_plan_host_chunks.argtypes = [C.c_int64, C.c_int64, C.c_int, C.c_int, C.c_void_p, C.c_int]
    n = lib.mrp_fg_plan_host_chunks(rows, chunk_rows, int(first_stream), int(last_stream), None, 0)
    if n < 0:
        raise MrpError("bad argument")
    out = np.zeros(max(n, 1), np.int64)
    lib.mrp_fg_plan_host_chunks(rows, chunk_rows, int(first_stream), int(last_stream), out.ctypes.data, n)
    return out[:n].tolist()


def halo_rows(params: Optional[Params] = None) -> int:
    """mrp_fg_halo_rows_for: rows of history a time shard / host chunk needs in front of its own rows (host arithmetic, no GPU)."""
    p = params or default_params()
    h = int(load().mrp_fg_halo_rows_for(C.byref(p)))
    if h < 0:
        raise MrpError("invalid EMA parameters")
    return h


def mean_scale(moments: np.ndarray) -> np.ndarray:
    """[19, 2] = {mean, scale} from [19, 3] combined moments (mrp_fg_mean_scale: host arithmetic, no GPU needed)."""
    m = np.ascontiguousarray(moments, dtype=np.float64)
    out = np.empty((NUM_FEATURES, 2), np.float64)
    rc = load().mrp_fg_mean_scale(C.c_void_p(m.ctypes.data), C.c_void_p(out.ctypes.data))
    if rc != 0:
        raise MrpError(f"mrp_fg_mean_scale failed ({rc})")
    return out
