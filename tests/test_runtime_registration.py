"""Residuals registered at run time (b200dae_register_problem, scipy_dae_b200.register_problem).

The reference takes any Python callable (base.py:44-45); here a residual is CUDA source compiled
with NVRTC against the library's own kernel templates.  The compile step needs no device, so the
CPU tests cover registration, the NVRTC error path and the variant compile; the GPU tests check
that a run-time registered copy of a built-in functor gives BIT-IDENTICAL results to the built-in
(thread regime: Robertson, group regime: Akzo Nobel), and that a brand-new problem (van der Pol as
an implicit ODE) agrees with the unmodified reference integrating its NumPy twin."""
import os
import re

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
PROBLEMS_CUH = os.path.join(HERE, "..", "scipy_dae_b200", "csrc", "problems.cuh")


def builtin_source(struct_name: str) -> str:
    """The source text of a built-in functor, renamed to `Problem` (so the registered copy is the
    same code, whatever the built-in looks like)."""
    text = open(PROBLEMS_CUH).read()
    m = re.search(r"^struct %s \{.*?^\};" % struct_name, text, flags=re.S | re.M)
    assert m, struct_name
    return m.group(0).replace("struct %s {" % struct_name, "struct Problem {", 1)


VDP_SRC = """
struct Problem {
    static constexpr int N = 2, NP = 1;
    static constexpr bool kHasJac = false;
    __device__ static void eval(double, const double* y, const double* yp, const double* p, double* out) {
        out[0] = yp[0] - y[1];
        out[1] = yp[1] - (p[0] * ((1.0 - y[0] * y[0]) * y[1]) - y[0]);
    }
};
"""


def vdp_residual(t, y, yp, mu):
    return np.array([yp[0] - y[1], yp[1] - (mu * ((1.0 - y[0] * y[0]) * y[1]) - y[0])])


@pytest.fixture(scope="module", autouse=True)
def rt_cache(tmp_path_factory):
    old = os.environ.get("B200DAE_RT_CACHE")
    if old is None:
        os.environ["B200DAE_RT_CACHE"] = str(tmp_path_factory.mktemp("rt_cache"))
    yield
    if old is None:
        os.environ.pop("B200DAE_RT_CACHE", None)


def test_register_lookup_and_compile_without_a_device():
    from scipy_dae_b200 import _lib, register_problem, precompile, get_problem
    before = _lib.load().b200dae_num_problems()
    prob = register_problem("rt_robertson", builtin_source("Robertson"), n=3, n_params=3)
    assert get_problem("rt_robertson") is prob
    pid, n, npar = _lib.problem_lookup("rt_robertson")
    assert (n, npar) == (3, 3) and pid >= before - 1
    assert _lib.load().b200dae_problem_name(pid) == b"rt_robertson"
    assert _lib.load().b200dae_problem_num_events(pid) == _lib.load().b200dae_problem_num_events(
        _lib.problem_lookup("robertson")[0])
    # registering the same source again is idempotent, a different one under the same name is refused
    assert _lib.load().b200dae_register_problem(b"rt_robertson", builtin_source("Robertson").encode(), 3, 3) == pid
    with pytest.raises(ValueError, match="already registered"):
        register_problem("rt_robertson", VDP_SRC, n=3, n_params=3)
    # the headline kernel variant and the BDF kernel compile for sm_100a here, without a GPU
    precompile("rt_robertson", "Radau", stages=3)
    precompile("rt_robertson", "BDF")
    precompile("robertson", "Radau")          # built-in: nothing to do


def test_nvrtc_errors_reach_the_caller():
    from scipy_dae_b200 import _lib, register_problem
    with pytest.raises(_lib.B200DaeError, match="no_such_symbol"):
        register_problem("rt_broken", VDP_SRC.replace("y[1];", "y[1] + no_such_symbol;", 1), n=2, n_params=1)
    with pytest.raises(_lib.B200DaeError, match="N / NP differ"):
        register_problem("rt_wrong_size", VDP_SRC, n=3, n_params=1)
    with pytest.raises(ValueError, match="built-in"):
        register_problem("robertson", VDP_SRC, n=2, n_params=1)
    with pytest.raises(ValueError, match="identifier"):
        register_problem("bad name", VDP_SRC, n=2, n_params=1)


def _same(a, b):
    return a.shape == b.shape and np.array_equal(a, b, equal_nan=True)


@pytest.mark.gpu
@pytest.mark.parametrize("builtin,struct,N,sigma,tol", [("robertson", "Robertson", 2048, 0.1, 1e-6),
                                                        ("akzo_nobel", "AkzoNobel", 256, 0.05, 1e-8)])
def test_registered_copy_is_bit_identical_to_the_builtin(builtin, struct, N, sigma, tol):
    import torch
    from scipy_dae_b200 import (solve_dae_batched, register_problem, perturbed_params, initial_conditions,
                                get_problem, device_residual)
    pr = get_problem(builtin)
    name = "rt_copy_" + builtin
    register_problem(name, builtin_source(struct), n=pr.n, n_params=pr.n_params)
    params = perturbed_params(builtin, N, sigma, seed=3)
    y0, yp0 = initial_conditions(builtin, params)
    fa = device_residual(builtin, pr.t_span[0], y0, yp0, params)
    fb = device_residual(name, pr.t_span[0], y0, yp0, params)
    assert _same(fa.cpu().numpy(), fb.cpu().numpy())
    t_eval = np.geomspace(1e-3, 1.0, 9) * pr.t_span[1]
    for method, kw in (("Radau", {}), ("BDF", {}), ("Radau", {"t_eval": t_eval}), ("BDF", {"t_eval": t_eval})):
        a = solve_dae_batched(builtin, pr.t_span, y0, yp0, params, method=method, rtol=tol, atol=tol, **kw)
        b = solve_dae_batched(name, pr.t_span, y0, yp0, params, method=method, rtol=tol, atol=tol, **kw)
        torch.cuda.synchronize()
        a, b = a.numpy(), b.numpy()
        assert (a.status == 0).all(), (method, kw)
        for key in ("status", "t_final", "y_final", "yp_final", "nfev", "njev", "nlu", "n_accepted",
                    "n_err_rejected", "n_newton_fail", "n_newton_iter"):
            assert _same(np.asarray(getattr(a, key)), np.asarray(getattr(b, key))), (method, key)
        if kw:
            assert _same(a.y, b.y) and _same(a.yp, b.yp)


@pytest.mark.gpu
def test_new_problem_against_the_reference():
    import torch
    from oracle import reference as REF
    from scipy_dae_b200 import solve_dae_batched, register_problem
    if not REF.available():
        pytest.skip("oracle/_ref is missing: run oracle/make_ref.sh")
    prob = register_problem("rt_vdp", VDP_SRC, n=2, n_params=1, residual=vdp_residual, default_params=(5.0,),
                            t_span=(0.0, 12.0))
    mus = np.array([[0.5], [5.0], [50.0], [200.0]])
    y0 = np.tile([2.0, 0.0], (len(mus), 1))
    yp0 = np.tile([0.0, -2.0], (len(mus), 1))
    rtol, atol = 1e-7, 1e-9
    t_eval = np.linspace(0.5, 12.0, 24)
    for method in ("Radau", "BDF"):
        r = solve_dae_batched("rt_vdp", prob.t_span, y0, yp0, mus, method=method, rtol=rtol, atol=atol, t_eval=t_eval)
        torch.cuda.synchronize()
        r = r.numpy()
        assert (r.status == 0).all()
        for i, mu in enumerate(mus):
            ref = REF.run_instance(prob, mu, y0[i], yp0[i], prob.t_span, method, rtol, atol, t_eval, {})
            assert ref["status"] == 0
            # the north_star band: max(10 rtol |y|, atol) against the reference's own states
            band = np.maximum(10 * rtol * np.abs(ref["y_eval"]), atol)
            ratio = float((np.abs(r.y[i] - ref["y_eval"]) / band).max())
            assert ratio <= 1.0, (method, mu, ratio)
            assert abs(int(r.n_accepted[i]) - ref["n_accepted"]) <= 0.05 * ref["n_accepted"] + 3, (method, mu)
