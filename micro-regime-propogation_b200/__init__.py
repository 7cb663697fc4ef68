"""micro-regime-propogation_b200 — B200-native feature_generation engine (hot path only).

Layout
  csrc/    hand-written sm_100a CUDA kernels + the C ABI (include/mrp_fg.h) -> csrc/libmrp_fg.so
  host/    C++20 host adaptor mirroring the reference's FeatureProcessor / FeatureSet / DataReciever / CsvWriter
  cabi.py  ctypes view of the C ABI (tests, bench, validation)
  synth.py synthetic snapshot streams (test / bench input)
  build.py nvcc / g++ recipes used by __graft_entry__.build()

The directory name carries a hyphen (it mirrors the reference repository's name), so it is imported
through the root-level shim module `mrp_b200` (see /mrp_b200.py): `import mrp_b200`.
"""
from .cabi import (COUNTS, FEATURES, HaloInfo, NUM_COUNTS, NUM_FEATURES, Engine, HostView, MrpError, Params, SoaView,
                   default_params, device_view, halo_rows, load, mean_scale, plan_host_chunks, plan_shape, LIB_PATH, NORM_COMPENSATED,
                   NORM_FAITHFUL)

__all__ = ["COUNTS", "FEATURES", "HaloInfo", "NUM_COUNTS", "NUM_FEATURES", "Engine", "HostView", "MrpError", "Params",
           "SoaView", "default_params", "device_view", "halo_rows", "load", "mean_scale", "plan_host_chunks", "plan_shape", "LIB_PATH", "NORM_COMPENSATED", "NORM_FAITHFUL"]
