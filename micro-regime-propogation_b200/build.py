"""Build recipes: the sm_100a CUDA extension (csrc/libmrp_fg.so) and the C++20 host adaptor objects.

Everything is built IN-TREE so that the binaries travel with the repository snapshot to the GPU box.
"""
from __future__ import annotations

import os
import shutil
import subprocess
from typing import List

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
HOST = os.path.join(HERE, "host")
LIB = os.path.join(CSRC, "libmrp_fg.so")
CU_SOURCES = ["kernels_window.cu", "kernels_row.cu", "kernels_norm.cu", "kernels_volvol.cu", "kernels_state.cu", "api.cu"]

NVCC_FLAGS = [
    "-std=c++17", "-O3",
    "-gencode", "arch=compute_100a,code=sm_100a",   # Blackwell B200 only: no PTX for other architectures
    "-lineinfo",
    "-fmad=false",   # a*b+c is rounded as the reference's SSE2 build rounds it; fma() is explicit where wanted
    "-Xcompiler", "-fPIC", "-shared",
]


def _nvcc() -> str:
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: cannot build the CUDA extension (no CPU fallback exists)")


def _stale(target: str, deps: List[str]) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build_cuda(force: bool = False, verbose: bool = False) -> str:
    srcs = [os.path.join(CSRC, s) for s in CU_SOURCES]
    deps = srcs + [os.path.join(CSRC, "mrp_device.cuh"), os.path.join(HERE, "..", "include", "mrp_fg.h"), __file__]
    extra = os.environ.get("MRP_FG_NVCC_EXTRA", "").split()   # experiment switches (-DNAME=value), empty in normal builds
    # the flags a library was built with sit next to it: a library left behind by an experiment build is rebuilt
    stamp = LIB + ".flags"
    want = " ".join(NVCC_FLAGS + extra)
    have = open(stamp).read() if os.path.exists(stamp) else None
    if force or have != want or _stale(LIB, deps):
        cmd = [_nvcc()] + NVCC_FLAGS + extra + (["-Xptxas", "-v"] if verbose else []) + ["-o", LIB] + srcs
        # nvcc's host compiler must be the PATH g++ (the image's CC/CXX wrappers are not usable by nvcc)
        env = dict(os.environ)
        env.pop("CC", None)
        env.pop("CXX", None)
        subprocess.run(cmd, check=True, cwd=CSRC, env=env)
        with open(stamp, "w") as f:
            f.write(want)
    return LIB


def build_host(force: bool = False) -> str:
    """C++20 host adaptor + validation tool (links against libmrp_fg.so)."""
    mk = os.path.join(HOST, "Makefile")
    if os.path.exists(mk):
        env = dict(os.environ)
        env.pop("CC", None)
        env.pop("CXX", None)
        subprocess.run(["make", "-C", HOST] + (["-B"] if force else []), check=True, stdout=subprocess.DEVNULL, env=env)
    return HOST


def build_all(force: bool = False, verbose: bool = False) -> None:
    build_cuda(force, verbose)
    build_host(force)
