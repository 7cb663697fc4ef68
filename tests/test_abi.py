"""The C-ABI shared library: loads, exports every symbol include/mrp_fg.h declares, reports errors without
exceptions, and refuses to run without a GPU (there is no CPU fallback).  No compute calls here."""
import ctypes as C
import os
import re

import pytest

import mrp_b200

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    names = set()
    inc = os.path.join(ROOT, "include")
    for fn in sorted(os.listdir(inc)):
        if fn.endswith(".h"):
            src = open(os.path.join(inc, fn)).read()
            src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
            names |= set(re.findall(r"\b(mrp_[a-z0-9_]+)\s*\(", src))
    return sorted(names)


def test_header_symbols_are_exported():
    lib = C.CDLL(mrp_b200.LIB_PATH)
    names = _declared_symbols()
    assert len(names) >= 17
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing


def test_abi_version_and_default_params_are_the_reference_constants():
    lib = mrp_b200.load()
    assert lib.mrp_fg_abi_version() >= 1
    p = mrp_b200.Params()
    assert lib.mrp_fg_default_params(C.byref(p)) == 0
    # common_constants.hpp:8-11, feature_engine.cpp:146, feature_processor.cpp:37,64,114-116,274-277
    assert (p.depth_levels, p.mid_window, p.dir_window, p.spread_divisor, p.ewm_span, p.norm_window) == (10, 1800, 1500, 1500, 1500, 30000)
    assert (p.return_lag, p.stress_levels) == (20, 5)
    assert (p.ofi_depth_decay, p.ofi_alpha, p.ofi_min_total_vol) == (0.5, 0.2, 1e-6)
    assert (p.stress_min_size, p.stress_distance_decay, p.stress_alpha) == (5.0, 10.0, 0.1)
    q = mrp_b200.default_params()
    for f, _ in mrp_b200.Params._fields_:
        assert getattr(p, f) == getattr(q, f), f
    assert lib.mrp_fg_default_params(None) == 1 and b"NULL" in lib.mrp_fg_last_error()


def test_struct_layout_matches_header():
    # 8 int32 + 6 double + 4 int32 ; 2 int32 + 16 pointers
    assert C.sizeof(mrp_b200.Params) == 8 * 4 + 6 * 8 + 4 * 4
    assert C.sizeof(mrp_b200.SoaView) == 8 + 16 * 8
    assert len(mrp_b200.FEATURES) == 19 and len(set(mrp_b200.FEATURES)) == 19


def test_bad_params_are_rejected_before_touching_the_device():
    lib = mrp_b200.load()
    ctx = C.c_void_p()
    for field, val in (("depth_levels", 1), ("stress_levels", 11), ("norm_window", 0), ("return_lag", 4000), ("norm_mode", 7)):
        p = mrp_b200.default_params(**{field: val})
        assert lib.mrp_fg_create(C.byref(p), 0, C.byref(ctx)) == 1, field
        assert not ctx.value
        assert lib.mrp_fg_last_error()


def test_no_gpu_means_error_not_fallback():
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    lib = mrp_b200.load()
    ctx = C.c_void_p()
    p = mrp_b200.default_params()
    rc = lib.mrp_fg_create(C.byref(p), 0, C.byref(ctx))
    assert rc == 2 and not ctx.value
    assert b"no CPU fallback" in lib.mrp_fg_last_error()
    with pytest.raises(mrp_b200.MrpError):
        mrp_b200.Engine(p)


def test_null_context_calls_fail_cleanly():
    lib = mrp_b200.load()
    assert lib.mrp_fg_run(None, None) == 1
    assert lib.mrp_fg_destroy(None) == 0
    assert lib.mrp_fg_launch_count(None) == 0
    assert lib.mrp_fg_num_kernels() >= 1 and lib.mrp_fg_kernel_name(0)


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "micro-regime-propogation_b200")
    for dirpath, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".cpp", ".hpp", ".h")) or fn == "Makefile":
                src = open(os.path.join(dirpath, fn)).read()
                assert "oracle" not in src.lower().replace("# oracle-free", ""), os.path.join(dirpath, fn)


def test_graft_entry_build_passes():
    """The driver's "does it build" check: __graft_entry__.build() (nvcc cross-compile, host adaptor, checkers, ABI version
    of the built library == the header's)."""
    import subprocess
    import sys

    r = subprocess.run([sys.executable, "-c", "import __graft_entry__ as g; g.build(); print('ok')"], cwd=ROOT, capture_output=True, text=True,
                       timeout=900)
    assert r.returncode == 0 and r.stdout.strip().endswith("ok"), r.stderr[-2000:]


def test_cmake_sm_100a_target_builds(tmp_path):
    """The CMake target the north star names: configure + build both libraries, sm_100a code only, every ABI symbol exported."""
    import shutil
    import subprocess

    cmake, ninja, nvcc = shutil.which("cmake"), shutil.which("ninja"), shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not (cmake and ninja and os.path.exists(nvcc)):
        pytest.skip("cmake / ninja / nvcc not available")
    env = dict(os.environ)
    env.pop("CC", None)
    env.pop("CXX", None)
    gxx = shutil.which("g++")
    cfg = [cmake, "-S", ROOT, "-B", str(tmp_path), "-G", "Ninja", f"-DCMAKE_CUDA_COMPILER={nvcc}", f"-DCMAKE_CXX_COMPILER={gxx}",
           f"-DCMAKE_CUDA_HOST_COMPILER={gxx}"]
    if os.path.isdir("/root/reference/feature_generation/include"):
        cfg.append("-DMRP_REFERENCE_ROOT=/root/reference")     # the adaptor against the reference's own headers
    subprocess.run(cfg, check=True, env=env, capture_output=True)
    subprocess.run([cmake, "--build", str(tmp_path)], check=True, env=env, capture_output=True)
    lib = C.CDLL(os.path.join(str(tmp_path), "libmrp_fg.so"))
    for name in _declared_symbols():
        assert hasattr(lib, name), name
    dump = subprocess.run([os.path.join(os.path.dirname(nvcc), "cuobjdump"), "--list-elf", os.path.join(str(tmp_path), "libmrp_fg.so")],
                          capture_output=True, text=True).stdout
    archs = {l.split(".")[-2] for l in dump.splitlines() if ".cubin" in l}
    assert archs == {"sm_100a"}, archs


def test_header_is_plain_c_and_a_c_caller_links(tmp_path):
    """include/mrp_fg.h compiles as C99 and a C program linked against libmrp_fg.so runs the host-only entry points (and gets a clean
    refusal from mrp_fg_create when there is no GPU)."""
    import shutil
    import subprocess

    gcc = shutil.which("gcc")
    if gcc is None:
        pytest.skip("gcc not installed")
    exe = tmp_path / "abi_smoke"
    libdir = os.path.dirname(mrp_b200.LIB_PATH)
    env = dict(os.environ)
    env.pop("CC", None)
    r = subprocess.run([gcc, "-std=c99", "-Wall", "-Werror", "-pedantic", "-I", os.path.join(ROOT, "include"),
                        os.path.join(ROOT, "tests", "c", "abi_smoke.c"), "-o", str(exe), "-L", libdir, "-lmrp_fg", f"-Wl,-rpath,{libdir}"],
                       capture_output=True, text=True, env=env)
    assert r.returncode == 0, r.stderr[-3000:]
    r = subprocess.run([str(exe)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "abi_smoke ok" in r.stdout, (r.returncode, r.stdout, r.stderr[-2000:])
