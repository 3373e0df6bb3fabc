"""t_eval output layouts (b200dae_outputs: eval_stride_row / _comp / _inst).

The reference returns `sol.y[n, m]` per instance (dae.py:496-515, 545-547); the batched result is the
`[N, n, m]` view of either device layout.  Both layouts must hold bit-identical states for every
kernel family, and the C-ABI must reject negative strides."""
import ctypes

import numpy as np
import pytest


@pytest.mark.gpu
@pytest.mark.parametrize("problem,method,N,sigma,tol,opts", [
    ("robertson", "Radau", 300, 0.1, 1e-6, {}),             # radau_kernel (lean twin with t_eval output)
    ("robertson", "BDF", 300, 0.1, 1e-6, {}),               # bdf_kernel
    ("akzo_nobel", "Radau", 65, 0.05, 1e-8, {}),            # unrolled thread regime
    ("transamp", "Radau", 33, 0.02, 1e-8, {}),              # radau_group_kernel
    ("transamp", "BDF", 33, 0.02, 1e-8, {}),                # bdf_group_kernel
    ("robertson", "Radau", 40, 0.1, 1e-6, {"stages": 5}),   # full-option kernel
])
def test_layouts_hold_identical_states(problem, method, N, sigma, tol, opts):
    import torch
    from scipy_dae_b200 import solve_dae_batched, perturbed_params, initial_conditions, get_problem
    pr = get_problem(problem)
    params = perturbed_params(problem, N, sigma, seed=5)
    y0, yp0 = initial_conditions(problem, params)
    t0, t1 = pr.t_span
    t_eval = t0 + (t1 - t0) * np.geomspace(1e-6, 1.0, 37)
    res = {}
    for layout in ("instance", "soa"):
        r = solve_dae_batched(problem, pr.t_span, y0, yp0, params, method=method, rtol=tol, atol=tol,
                              t_eval=t_eval, eval_layout=layout, **opts)
        torch.cuda.synchronize()
        assert tuple(r.y.shape) == (N, pr.n, len(t_eval))
        res[layout] = r.numpy()
    a, b = res["instance"], res["soa"]
    assert (a.status == 0).all()
    assert np.array_equal(a.n_eval_valid, b.n_eval_valid) and (a.n_eval_valid == len(t_eval)).all()
    assert np.array_equal(a.y, b.y) and np.array_equal(a.yp, b.yp)
    assert np.array_equal(a.y_final, b.y_final) and np.array_equal(a.n_accepted, b.n_accepted)
    assert np.isfinite(a.y).all()


def test_negative_strides_are_rejected():
    from scipy_dae_b200 import _lib
    from scipy_dae_b200.constants import radau_tables_c
    L = _lib.load()
    pid, n, npar = _lib.problem_lookup("robertson")
    out = _lib.Outputs()
    out.y_final = out.t_final = out.status = 8          # never dereferenced: the call fails on validation
    out.t_eval, out.m, out.y_eval = 8, 1, 8
    out.eval_stride_row = -1
    ro = _lib.RadauOpts(rtol=1e-6, max_step=float("inf"), first_step=0.0, continuous_error_weight=0.0,
                        jac_recompute_rate=1e-3, deadband_lo=1.0, deadband_hi=1.2, newton_max_iter=0,
                        newton_iter_embedded=1, max_steps=0, analytic_jac=0)
    tb = radau_tables_c(3)
    atol = (ctypes.c_double * n)(*([1e-6] * n))
    rc = L.b200dae_solve_radau(pid, 1, 1, 8, 8, 8, 0.0, 1.0, atol, ctypes.byref(ro), ctypes.byref(tb),
                               ctypes.byref(out), None)
    assert rc == -1 and b"strides" in L.b200dae_last_error()


def test_unknown_io_layout_is_rejected():
    """b200dae_radau_opts.io_layout / b200dae_bdf_opts.io_layout: 0 ([n][ld]) or 1 (instance-major)."""
    from scipy_dae_b200 import _lib
    from scipy_dae_b200.constants import radau_tables_c
    L = _lib.load()
    pid, n, npar = _lib.problem_lookup("robertson")
    out = _lib.Outputs()
    out.y_final = out.t_final = out.status = 8          # never dereferenced: the call fails on validation
    atol = (ctypes.c_double * n)(*([1e-6] * n))
    ro = _lib.RadauOpts(rtol=1e-6, max_step=float("inf"), first_step=0.0, continuous_error_weight=0.0,
                        jac_recompute_rate=1e-3, deadband_lo=1.0, deadband_hi=1.2, newton_max_iter=0,
                        newton_iter_embedded=1, max_steps=0, analytic_jac=0, io_layout=2)
    tb = radau_tables_c(3)
    rc = L.b200dae_solve_radau(pid, 1, 1, 8, 8, 8, 0.0, 1.0, atol, ctypes.byref(ro), ctypes.byref(tb), ctypes.byref(out), None)
    assert rc == -1 and b"io_layout" in L.b200dae_last_error()
    bo = _lib.BdfOpts(rtol=1e-6, max_step=float("inf"), first_step=0.0, max_order=5, ndf_strategy=0, max_steps=0,
                      analytic_jac=0, io_layout=7)
    rc = L.b200dae_solve_bdf(pid, 1, 1, 8, 8, 8, 0.0, 1.0, atol, ctypes.byref(bo), ctypes.byref(out), None)
    assert rc == -1 and b"io_layout" in L.b200dae_last_error()
