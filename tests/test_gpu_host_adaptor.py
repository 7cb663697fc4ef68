"""GPU: the C++20 adaptor end to end (GpuFeatureProcessor staging -> C ABI -> CUDA -> CsvWriter) against the reference's
own CSV files, through the validation harness."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("extra", [[], ["--cold", "--bulk"]], ids=["warm_per_row_sink", "cold_bulk_csv"])
def test_validation_harness_passes(tmp_path, extra):
    r = subprocess.run([sys.executable, os.path.join(ROOT, "validation", "validate.py"), "--rows", "6000", "--out", str(tmp_path)] + extra,
                       capture_output=True, text=True)
    print(r.stdout[-3000:], r.stderr[-2000:])
    assert r.returncode == 0
