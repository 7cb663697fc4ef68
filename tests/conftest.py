import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session", autouse=True)
def _built():
    """Make sure the C-ABI library and the checkers exist (nvcc cross-compiles without a GPU)."""
    import mrp_b200
    from mrp_b200 import build as b
    from oracle import binding as ob

    if not os.path.exists(mrp_b200.LIB_PATH):
        b.build_cuda()
    if not os.path.exists(ob.ORACLE_SO):
        ob.build()
    from mrp_b200 import hostlib

    if not os.path.exists(hostlib.HOST_LIB_PATH):
        b.build_host()
    yield


@pytest.fixture
def small_segments(monkeypatch):
    """Row segments of 257 rows: every halo / restart / partial-chunk path of the segment kernels runs many times."""
    monkeypatch.setenv("MRP_FG_SEGMENT_ROWS", "257")
    yield
