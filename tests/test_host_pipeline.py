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


def test_radau_tables_have_the_normalisation_the_kernel_relies_on():
    """The thread-per-system Radau kernel assumes T[s-1] = (1, 0, 1, 0, ..., 1) and a dense first
    row of TI (csrc/radau.cuh; the launcher rejects other tables).  Both follow from the
    reference's eigenvector scaling (radau.py:72-84); pinned here for every stage count."""
    from scipy_dae_b200.constants import radau_tables
    for s in (1, 3, 5, 7):
        tb = radau_tables(s)
        T, TI = np.asarray(tb.T).reshape(s, s), np.asarray(tb.TI).reshape(s, s)
        assert np.array_equal(T[-1], np.array([1.0 if l % 2 == 0 else 0.0 for l in range(s)]))
        assert (TI[0] != 0.0).all()
