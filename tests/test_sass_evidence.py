"""The shipped library is what DESIGN.md says it is (checked on the CPU with cuobjdump): built for sm_100a only, the three hot kernels
stage their inputs with bulk asynchronous copies (SASS UBLKCP) tracked by mbarriers (SYNCS), nothing on the path is a tensor-core
contraction, and the kernel names the profiles quote exist."""
import os
import re
import shutil
import subprocess

import pytest

import mrp_b200

CUOBJDUMP = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
pytestmark = pytest.mark.skipif(not os.path.exists(CUOBJDUMP), reason="cuobjdump not installed")


def _sass_by_function():
    out = subprocess.run([CUOBJDUMP, "-sass", mrp_b200.LIB_PATH], capture_output=True, text=True, timeout=600).stdout
    fns, cur = {}, None
    for ln in out.splitlines():
        m = re.search(r"Function : (\S+)", ln)
        if m:
            cur = m.group(1)
            fns[cur] = []
        elif cur and re.match(r"\s+/\*[0-9a-f]{4,5}\*/", ln):
            fns[cur].append(ln)
    return out, fns


def test_library_is_sm_100a_with_bulk_copy_staged_kernels():
    out, fns = _sass_by_function()
    archs = set(re.findall(r"arch = (sm_\w+)", out))
    assert archs == {"sm_100a"}, archs

    def kernel(sub):
        hits = [k for k in fns if sub in k]
        assert hits, sub
        return max(hits, key=lambda k: len(fns[k]))

    hot = {"k_windows": kernel("k_windows"), "k_book": kernel("k_bookILi5ELi2E"), "k_normalize_runs": kernel("k_normalize_runs")}
    for name, fn in hot.items():
        text = "\n".join(fns[fn])
        assert "UBLKCP" in text, f"{name}: no bulk asynchronous copy in the SASS"
        assert "SYNCS" in text, f"{name}: no mbarrier instruction in the SASS"
        assert not re.search(r"\b(HMMA|IMMA|DMMA|UTCHMMA|UTCIMMA|QGMMA)\b", text), f"{name}: tensor-core instruction on a path without a contraction"
    # the support kernels the docs name
    for sub in ("k_norm_poison", "k_col_moments", "k_standardize", "k_state_rows", "k_vol_prefix", "k_roll_apply", "k_rebase"):
        assert any(sub in k for k in fns), sub
