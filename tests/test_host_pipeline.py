"""Host-side logic of the chunked host pipeline (scipy_dae_b200/host_pipeline.py); no GPU needed."""
import numpy as np
import pytest

from scipy_dae_b200.host_pipeline import _chunk_bounds


@pytest.mark.parametrize("N", [0, 1, 7, 64, 1000, 1 << 20])
@pytest.mark.parametrize("chunks", [1, 2, 4, 5, 8, 16, 33])
@pytest.mark.parametrize("taper", [False, True])
def test_chunk_bounds_partition_the_instance_axis(N, chunks, taper):
    b = _chunk_bounds(N, chunks, taper)
    assert b.dtype == np.int64 and len(b) == chunks + 1
    assert b[0] == 0 and b[-1] == N
    assert (np.diff(b) >= 0).all()            # empty chunks are allowed (skipped by the pipeline)


def test_tapered_ends_are_short():
    N = 1 << 20
    b = _chunk_bounds(N, 16, True)
    sizes = np.diff(b)
    interior = sizes[2:-2]
    assert interior.max() - interior.min() <= 1
    assert abs(sizes[0] - interior[0] / 4) <= 2 and abs(sizes[-1] - interior[0] / 4) <= 2
    assert abs(sizes[1] - 3 * interior[0] / 4) <= 2 and abs(sizes[-2] - 3 * interior[0] / 4) <= 2
    # fewer than five chunks: nothing to taper
    assert (np.diff(_chunk_bounds(N, 4, True)) == N // 4).all()
