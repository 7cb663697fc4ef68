"""GPU parity for the shapes the bulk-copy staging of the book kernel has to get right: row counts and level strides that
are odd (level runs start at any 8 / 4-byte boundary), every level-group width (depth % 5, % 4, % 2, odd), streams that
begin at odd rows, and caller-owned device arrays whose base pointers are not 16-byte aligned."""
import pytest

import mrp_b200
from mrp_b200 import synth

from helpers import make_host_batch
from test_gpu_parity import _check

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("levels", [10, 5, 4, 6, 3, 1 + 6])
@pytest.mark.parametrize("mode", ["upload", "device_offset"])
def test_odd_rows_every_level_group(levels, mode):
    p = mrp_b200.default_params(emit_dirs=1, emit_counts=1, depth_levels=levels, stress_levels=min(5, levels))
    _check(make_host_batch([synth.SynthSpec(n_rows=2501, seed=70 + levels, levels=levels)]), params=p, mode=mode,
           min_frac=0.2 if levels < 5 else 0.95)


@pytest.mark.parametrize("mode", ["device", "device_offset"])
def test_streams_starting_at_odd_rows(mode):
    # stream starts at rows 0, 1, 4, 261, 774, 2775: chunks of later streams begin at every alignment
    specs = [synth.SynthSpec(n_rows=n, seed=80 + i, variable_cursors=bool(i & 1)) for i, n in enumerate((1, 3, 257, 513, 2001, 1030))]
    _check(make_host_batch(specs), mode=mode)


def test_segment_tail_shorter_than_a_chunk(small_segments):
    # segment ends cut chunks at arbitrary rows: the last bulk copy of a segment is short and ends unaligned
    _check(make_host_batch([synth.SynthSpec(n_rows=3001, seed=90, variable_cursors=True)]), mode="device_offset")
