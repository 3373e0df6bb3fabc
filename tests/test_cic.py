"""consistent_initial_conditions: the CPU oracle against goldens minted
from the unmodified reference (CPU), and the CUDA kernel against the same goldens and against
the oracle on larger seeded ensembles (GPU).  Bar: the returned point is the reference's point
to rounding (the Newton / chord iteration and both pivoted QRs are restated step by step); where
the reference raises, the batch reports the matching status code."""
import glob
import json
import os

import numpy as np
import pytest

from conftest import GOLDEN_DIR

CIC_DIR = os.path.join(GOLDEN_DIR, "cic")
# numerically singular set-ups in which the reference's own failure mode flips between
# "Convergence failed" and "Too many fixed components" from instance to instance
FAILURE_MODE_UNSTABLE = {"robertson_fixed_yp"}
# The amplifier's dF/dy' has pairs of columns with IDENTICAL norms ((-C1, C1) and (C1, -C1), ...).
# With finite-difference Jacobians the pivoted QR breaks those ties on rounding noise, and the
# device functor (FMA contraction, reciprocal constants) rounds differently from NumPy, so the
# kernel may pick the other column of a pair: a different but equally consistent basic solution.
# For that case the point is compared where it coincides and consistency is required everywhere;
# with analytic Jacobians (exact ties -> first index, as LAPACK) the points agree.
TIE_PRONE_FD = {"transamp_free"}


def cic_cases():
    return sorted(os.path.splitext(os.path.basename(f))[0] for f in glob.glob(os.path.join(CIC_DIR, "*.npz")))


def load(case):
    d = np.load(os.path.join(CIC_DIR, case + ".npz"))
    return {k: d[k] for k in d.files if k != "meta"}, json.loads(str(d["meta"]))


def check(case, out, g, rel=1e-9, tie_prone=False):
    st = np.asarray(out["status"])
    if case in FAILURE_MODE_UNSTABLE:
        assert ((st != 0) == (g["status"] != 0)).all()
    else:
        assert np.array_equal(st, g["status"]), (st, g["status"])
    ok = g["status"] == 0
    if ok.any():
        for k in ("y0", "yp0"):
            a, b = np.asarray(out[k])[ok], g[k][ok]
            scale = np.maximum(np.abs(b).max(axis=1, keepdims=True), 1e-300)
            same = (np.abs(a - b) <= rel * scale).all(axis=1)
            if tie_prone:
                assert same.any(), (case, k)
            else:
                assert same.all(), (case, k, np.abs(a - b).max())


@pytest.mark.parametrize("case", cic_cases())
def test_oracle_matches_reference_cic(case, oracle):
    g, meta = load(case)
    out = oracle.consistent_initial_conditions(meta["problem"], 0.0, g["y0_in"], g["yp0_in"], g["params"],
                                               jac=meta["jac"], fixed_y0=meta["fixed_y0"],
                                               fixed_yp0=meta["fixed_yp0"], **meta["options"])
    check(case, out, g)


@pytest.mark.gpu
@pytest.mark.parametrize("case", cic_cases())
def test_cuda_matches_reference_cic(case):
    from scipy_dae_b200 import consistent_initial_conditions_batched
    g, meta = load(case)
    r = consistent_initial_conditions_batched(meta["problem"], 0.0, g["y0_in"], g["yp0_in"], g["params"],
                                              jac=meta["jac"], fixed_y0=meta["fixed_y0"],
                                              fixed_yp0=meta["fixed_yp0"], **meta["options"])
    out = dict(y0=r.y0.cpu().numpy(), yp0=r.yp0.cpu().numpy(), status=r.status.cpu().numpy())
    # FMA contraction and fused norms on the device: same point to 1e-7 of the vector's scale
    # (the iteration stops at |F| < 0.5 * (atol + rtol |F0|), not at machine precision)
    check(case, out, g, rel=1e-7, tie_prone=case in TIE_PRONE_FD)
    ok = g["status"] == 0
    if ok.any():
        assert np.abs(r.f.cpu().numpy()[ok]).max() <= 10 * max(np.abs(g["f"][ok]).max(), 1e-12)


@pytest.mark.gpu
@pytest.mark.parametrize("problem,N,sigma", [("robertson", 4096, 0.1), ("akzo_nobel", 1024, 0.05),
                                             ("transamp", 512, 0.02), ("andrews", 256, 0.05)])
def test_cuda_cic_matches_oracle_on_ensembles(problem, N, sigma, oracle):
    """Larger ensembles than the goldens: device vs oracle, then the device result is consistent
    enough to start the integrator from (F(t0, y0, yp0) ~ 0, solve succeeds)."""
    from scipy_dae_b200 import (consistent_initial_conditions_batched, perturbed_params, initial_conditions,
                                device_residual, solve_dae_batched, get_problem)
    rng = np.random.default_rng(9)
    P = perturbed_params(problem, N, sigma, seed=21)
    Y0, YP0 = initial_conditions(problem, P)
    Y0 = Y0 * (1 + 1e-3 * rng.standard_normal(Y0.shape))
    YP0 = YP0 + 1e-3 * rng.standard_normal(YP0.shape) * np.maximum(np.abs(YP0), 1e-2)
    jac = "analytic" if problem == "transamp" else None      # see TIE_PRONE_FD
    r = consistent_initial_conditions_batched(problem, 0.0, Y0, YP0, P, jac=jac)
    ref = oracle.consistent_initial_conditions(problem, 0.0, Y0, YP0, P, jac=jac)
    st = r.status.cpu().numpy()
    assert np.array_equal(st, ref["status"]) and (st == 0).all()
    for k, a in (("y0", r.y0), ("yp0", r.yp0)):
        a, b = a.cpu().numpy(), ref[k]
        scale = np.maximum(np.abs(b).max(axis=1, keepdims=True), 1e-300)
        assert (np.abs(a - b) <= 1e-7 * scale).all(), (k, np.abs(a - b).max())
    F = device_residual(problem, 0.0, r.y0.cpu().numpy(), r.yp0.cpu().numpy(), P).cpu().numpy()
    assert np.abs(F).max() < 1e-6
    pr = get_problem(problem)
    s = solve_dae_batched(problem, pr.t_span, r.y0[:64], r.yp0[:64], P[:64], rtol=1e-4, atol=1e-6)
    assert bool((s.status == 0).all())


@pytest.mark.gpu
def test_cic_argument_errors():
    from scipy_dae_b200 import consistent_initial_conditions_batched as cic
    y0, yp0 = np.array([[1.0, 0, 0]]), np.array([[-0.04, 0.04, 0]])
    with pytest.raises(ValueError, match="Too many components fixed"):
        cic("robertson", 0.0, y0, yp0, fixed_y0=[0, 1, 2], fixed_yp0=[0])
    with pytest.raises(ValueError, match="Relative tolerance"):
        cic("robertson", 0.0, y0, yp0, rtol=-1.0)
    with pytest.raises(ValueError, match="Absolute tolerance"):
        cic("robertson", 0.0, y0, yp0, atol=0.0)
    with pytest.raises(ValueError, match="Too many fixed components"):
        cic("robertson", 0.0, y0, yp0, fixed_y0=[0, 1, 2], raise_on_failure=True)
