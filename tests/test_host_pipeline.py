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


@pytest.mark.gpu
@pytest.mark.parametrize("problem,method,N,sigma,tol", [
    ("robertson", "Radau", 5000, 0.1, 1e-6),      # radau_kernel (headline)
    ("robertson", "BDF", 1500, 0.1, 1e-6),        # bdf_kernel
    ("akzo_nobel", "BDF", 300, 0.05, 1e-8),       # unrolled thread regime
    ("transamp", "Radau", 70, 0.02, 1e-8),        # radau_group_kernel
    ("transamp", "BDF", 70, 0.02, 1e-8),          # bdf_group_kernel
])
def test_zero_copy_launch_is_bit_identical_to_the_device_path(problem, method, N, sigma, tol):
    """io_layout = 1 (include/b200dae.h): ONE launch reads the pinned (N, n) / (N, p) host arrays and
    writes the pinned result buffers in place.  Same kernels, same arithmetic: every output must equal
    the device-resident solve bit for bit."""
    import torch
    from scipy_dae_b200 import (solve_dae_batched, solve_dae_batched_host, alloc_host_result, perturbed_params,
                                initial_conditions, get_problem)
    pr = get_problem(problem)
    params = perturbed_params(problem, N, sigma, seed=11)
    y0, yp0 = initial_conditions(problem, params)
    hp, hy, hyp = (torch.from_numpy(np.ascontiguousarray(a)).pin_memory() for a in (params, y0, yp0))
    out = alloc_host_result(N, pr.n)
    r = solve_dae_batched_host(problem, pr.t_span, hy, hyp, hp, method=method, rtol=tol, atol=tol, out=out, zero_copy=True)
    ref = solve_dae_batched(problem, pr.t_span, y0, yp0, params, method=method, rtol=tol, atol=tol)
    torch.cuda.synchronize()
    ref = ref.numpy()
    assert r.extra["zero_copy"] and (ref.status == 0).all()
    assert np.array_equal(r.status.numpy(), ref.status) and np.array_equal(r.t_final.numpy(), ref.t_final)
    assert np.array_equal(r.y_final.numpy(), ref.y_final) and np.array_equal(r.yp_final.numpy(), ref.yp_final)
    for key in ("nfev", "njev", "nlu", "nsolve", "n_accepted", "n_err_rejected", "n_newton_fail", "n_newton_iter"):
        assert np.array_equal(getattr(r, key).numpy(), getattr(ref, key)), key


@pytest.mark.gpu
def test_zero_copy_reports_non_finite_initial_values_like_the_reference():
    import torch
    from scipy_dae_b200 import solve_dae_batched_host, alloc_host_result, perturbed_params, initial_conditions
    params = perturbed_params("robertson", 64, 0.1, seed=0)
    y0, yp0 = initial_conditions("robertson", params)
    y0[17, 1] = np.nan
    hp, hy, hyp = (torch.from_numpy(np.ascontiguousarray(a)).pin_memory() for a in (params, y0, yp0))
    with pytest.raises(ValueError, match="must be finite"):          # base.py:33-37
        solve_dae_batched_host("robertson", (0.0, 1e3), hy, hyp, hp, rtol=1e-6, atol=1e-6,
                               out=alloc_host_result(64, 3), zero_copy=True)


@pytest.mark.gpu
def test_sharded_zero_copy_gather_matches_the_plain_solve():
    """distributed.solve_dae_sharded(gather="host", zero_copy=True) at world size 1: the launch writes the
    shard straight into the shared, page-locked result arrays (cudaHostRegister'ed /dev/shm mapping)."""
    import os
    import torch
    from scipy_dae_b200 import solve_dae_batched, perturbed_params, initial_conditions
    from scipy_dae_b200.distributed import SharedHostResult, solve_dae_sharded
    N = 3000
    params = perturbed_params("robertson", N, 0.1, seed=4)
    y0, yp0 = initial_conditions("robertson", params)
    hp, hy, hyp = (torch.from_numpy(np.ascontiguousarray(a)).pin_memory() for a in (params, y0, yp0))
    shared = SharedHostResult(f"b200dae_test_{os.getpid()}", N, 3)
    try:
        res = solve_dae_sharded("robertson", (0.0, 1e7), hy, hyp, hp, method="Radau", rtol=1e-6, atol=1e-6,
                                gather="host", host_result=shared, zero_copy=True)
        ref = solve_dae_batched("robertson", (0.0, 1e7), y0, yp0, params, method="Radau", rtol=1e-6, atol=1e-6)
        torch.cuda.synchronize()
        ref = ref.numpy()
        assert np.array_equal(res["status"].numpy(), ref.status)
        assert np.array_equal(res["y_final"].numpy(), ref.y_final)
        assert np.array_equal(res["n_accepted"].numpy(), ref.n_accepted)
    finally:
        shared.close()
