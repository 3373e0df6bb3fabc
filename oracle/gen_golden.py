#!/usr/bin/env python3
"""Mint golden vectors by running the UNMODIFIED reference (``/root/reference``).

TEST INFRASTRUCTURE — runs only in the build container (the reference tree does not
exist on the GPU box).  It imports ``scipy_dae.integrate.solve_dae`` from
``/root/reference`` and integrates the problems of ``scipy_dae_b200/problems.py``
instance by instance, recording for each instance what the parity contract needs
(SURVEY.md §8c, App. G):

    status, t_final, y_final, yp_final, nfev, njev, nlu,
    n_accepted, n_newton_calls, n_newton_fail, n_newton_iter, n_err_rejected
    and, when a t_eval grid is part of the case, y(t_eval), yp(t_eval).

The per-instance runner (with the accepted / rejected counts derived by rebinding the
module-level Newton solvers) lives in ``oracle/reference.py``, which the live-reference tests on
the GPU box share.

Output: ``tests/golden/<case>.npz`` (small, committed).  Usage::

    python oracle/gen_golden.py            # all cases
    python oracle/gen_golden.py robertson  # cases whose name contains the substring
    python oracle/gen_golden.py -transamp1 # all cases but those
"""
import json
import os
import sys
import time
import warnings

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)

import scipy  # noqa: E402

from oracle.reference import import_reference, run_instance, ref_source  # noqa: E402
from scipy_dae_b200.problems import (  # noqa: E402
    get_problem, perturbed_params, initial_conditions)

import_reference()            # the unmodified reference, from oracle/_ref (checksummed) or /root/reference
GOLDEN_DIR = os.path.join(REPO, "tests", "golden")


TWIN_EPS = 1e-15
TWINS = (1, -1, 2, -2)          # parameters * (1 + k * TWIN_EPS): the reference against itself
TWIN_KEYS = ("status", "t_final", "y_final", "nfev", "njev", "nlu", "n_accepted", "n_newton_fail", "n_err_rejected",
             "n_eval_valid", "y_eval", "t_events", "n_events_found")


def _run_all(problem, params, y0, yp0, t_span, method, rtol, atol, t_eval, options, record_steps):
    return [run_instance(problem, params[i], y0[i], yp0[i], t_span, method, rtol, atol,
                         t_eval, options, record_steps=record_steps and i == 0)
            for i in range(params.shape[0])]


def _job(args):
    return _run_all(get_problem(args[0]), *args[1:])


def run_case(name, problem, params, method, rtol, atol, t_span=None, t_eval=None,
             options=None, record_steps=False, y0=None, yp0=None, twins=TWINS, pool=None):
    """One golden file: the case itself plus its TWINS, the same case with every parameter multiplied
    by (1 + k * 1e-15) and the consistent initial values recomputed from them.  The twins measure how
    far the UNMODIFIED reference moves against itself under a perturbation at rounding level; the
    parity tests (tests/parity.py) allow the CUDA path and the oracle the north_star band or twice that
    self-distance, whichever is larger, so every widened band is a measured quantity of this file."""
    problem = get_problem(problem)
    options = dict(options or {})
    params = np.atleast_2d(np.asarray(params, dtype=float))
    N = params.shape[0]
    given_ic = y0 is not None
    if not given_ic:
        y0, yp0 = initial_conditions(problem, params)
    t_span = tuple(t_span or problem.t_span)
    t0 = time.perf_counter()
    jobs = [(problem.name, params, y0, yp0, t_span, method, rtol, atol, t_eval, options, record_steps)]
    for k in twins:
        pk = params * (1.0 + k * TWIN_EPS)
        yk, ypk = (y0, yp0) if given_ic else initial_conditions(problem, pk)
        jobs.append((problem.name, pk, yk, ypk, t_span, method, rtol, atol, t_eval, options, False))
    runs = pool.map(_job, jobs) if pool is not None else [_job(j) for j in jobs]
    res = runs[0]
    wall = time.perf_counter() - t0
    arrays = dict(params=params, y0=y0, yp0=yp0)
    for key in ("status", "t_final", "nfev", "njev", "nlu", "n_accepted", "n_newton_calls",
                "n_newton_fail", "n_newton_iter", "n_err_rejected"):
        arrays[key] = np.array([r[key] for r in res])
    arrays["y_final"] = np.stack([r["y_final"] for r in res])
    arrays["yp_final"] = np.stack([r["yp_final"] for r in res])
    if t_eval is not None:
        arrays["t_eval"] = np.asarray(t_eval, dtype=float)
        arrays["n_eval_valid"] = np.array([r["n_eval_valid"] for r in res])
        arrays["y_eval"] = np.stack([r["y_eval"] for r in res])    # [N, n, m]
        arrays["yp_eval"] = np.stack([r["yp_eval"] for r in res])
    if record_steps:
        arrays["t_steps"] = res[0]["t_steps"]
        arrays["y_steps"] = res[0]["y_steps"]
    if "t_events" in res[0]:
        for key in ("t_events", "y_events", "yp_events", "n_events_found"):
            arrays[key] = np.stack([r[key] for r in res])
    if twins:
        arrays["twin_factor"] = np.array([1.0 + k * TWIN_EPS for k in twins])
        for key in TWIN_KEYS:
            if key in res[0]:
                arrays["twin_" + key] = np.stack([np.stack([np.asarray(r[key]) for r in run]) for run in runs[1:]])
    meta = dict(case=name, problem=problem.name, method=method, rtol=rtol,
                atol=atol if np.isscalar(atol) else list(atol), t_span=list(t_span),
                options=options, N=N, reference="JonasBreuling/scipy_dae v0.1.0 (" + ref_source() + ")",
                scipy=scipy.__version__, numpy=np.__version__, wall_s=round(wall, 3),
                twins=[f"params * (1 + {k}e-15)" for k in twins],
                generator="oracle/gen_golden.py")
    arrays["meta"] = np.array(json.dumps(meta))
    os.makedirs(GOLDEN_DIR, exist_ok=True)
    np.savez_compressed(os.path.join(GOLDEN_DIR, name + ".npz"), **arrays)
    print(f"{name}: N={N} wall={wall:.2f}s status={arrays['status'].tolist()[:8]} "
          f"acc={arrays['n_accepted'].tolist()[:8]} nfev={arrays['nfev'].tolist()[:4]}", flush=True)


def cases():
    rob_nom = np.array([[0.04, 1e4, 3e7]])
    te_log = np.logspace(-6, 7, 60)
    # --- C1: single nominal Robertson (README.md:37-66), with the full step trace
    yield dict(name="robertson_radau3_c1", problem="robertson", params=rob_nom, method="Radau",
               rtol=1e-6, atol=1e-6, t_eval=te_log, record_steps=True)
    yield dict(name="robertson_bdf_c1", problem="robertson", params=rob_nom, method="BDF",
               rtol=1e-6, atol=1e-6, t_eval=te_log, record_steps=True)
    # --- C2: first instances of the bench ensemble (sigma 0.1, seed 0) and a stress set
    yield dict(name="robertson_radau3_ens01", problem="robertson",
               params=perturbed_params("robertson", 48, 0.1, 0), method="Radau",
               rtol=1e-6, atol=1e-6)
    yield dict(name="robertson_radau3_ens05", problem="robertson",
               params=perturbed_params("robertson", 32, 0.5, 1), method="Radau",
               rtol=1e-6, atol=1e-6, t_eval=np.logspace(-5, 7, 13))
    yield dict(name="robertson_bdf_ens01", problem="robertson",
               params=perturbed_params("robertson", 32, 0.1, 0), method="BDF",
               rtol=1e-6, atol=1e-6)
    # other tolerances / stage counts / options (row N4 of SURVEY §8f)
    yield dict(name="robertson_radau5_ens01", problem="robertson",
               params=perturbed_params("robertson", 8, 0.1, 0), method="Radau",
               rtol=1e-6, atol=1e-6, options=dict(stages=5))
    yield dict(name="robertson_radau7_ens01", problem="robertson",
               params=perturbed_params("robertson", 4, 0.1, 0), method="Radau",
               rtol=1e-6, atol=1e-6, options=dict(stages=7))
    yield dict(name="robertson_radau3_tight", problem="robertson",
               params=perturbed_params("robertson", 8, 0.1, 0), method="Radau",
               rtol=1e-9, atol=1e-12)
    yield dict(name="robertson_radau3_1e8_vecatol", problem="robertson",
               params=perturbed_params("robertson", 8, 0.1, 3), method="Radau",
               rtol=1e-5, atol=[1e-7, 1e-10, 1e-7], t_span=(0.0, 1e8))
    yield dict(name="robertson_radau3_emb0", problem="robertson",
               params=perturbed_params("robertson", 4, 0.1, 0), method="Radau",
               rtol=1e-6, atol=1e-6, options=dict(newton_iter_embedded=0))
    yield dict(name="robertson_radau3_emb2", problem="robertson",
               params=perturbed_params("robertson", 4, 0.1, 0), method="Radau",
               rtol=1e-6, atol=1e-6, options=dict(newton_iter_embedded=2))
    yield dict(name="robertson_radau3_cew05", problem="robertson",
               params=perturbed_params("robertson", 4, 0.1, 0), method="Radau",
               rtol=1e-6, atol=1e-6, options=dict(continuous_error_weight=0.5))
    yield dict(name="robertson_radau3_maxstep", problem="robertson",
               params=perturbed_params("robertson", 4, 0.1, 0), method="Radau",
               rtol=1e-6, atol=1e-6, t_span=(0.0, 1e3),
               options=dict(max_step=20.0, first_step=1e-4))
    yield dict(name="robertson_bdf_variants", problem="robertson",
               params=perturbed_params("robertson", 4, 0.1, 0), method="BDF",
               rtol=1e-6, atol=1e-6, options=dict(NDF_strategy="efficiency", max_order=4))
    # --- C3: Akzo Nobel, BDF default, rtol=atol=1e-8
    yield dict(name="akzo_bdf_ens", problem="akzo_nobel",
               params=perturbed_params("akzo_nobel", 16, 0.05, 0), method="BDF",
               rtol=1e-8, atol=1e-8, t_eval=np.linspace(0, 180, 10))
    yield dict(name="akzo_radau3_ens", problem="akzo_nobel",
               params=perturbed_params("akzo_nobel", 8, 0.05, 0), method="Radau",
               rtol=1e-8, atol=1e-8)
    # --- C5: two-transistor amplifier, rtol=1e-10, atol=1e-6, Radau s=3 and BDF side by side
    yield dict(name="transamp_radau3_ens", problem="transamp",
               params=perturbed_params("transamp", 4, 0.02, 1), method="Radau",
               rtol=1e-10, atol=1e-6, t_eval=np.linspace(0, 0.2, 9))
    yield dict(name="transamp_bdf_ens", problem="transamp",
               params=perturbed_params("transamp", 4, 0.02, 1), method="BDF",
               rtol=1e-10, atol=1e-6)
    yield dict(name="transamp_radau3_loose", problem="transamp",
               params=perturbed_params("transamp", 4, 0.02, 1), method="Radau",
               rtol=1e-6, atol=1e-6)
    # --- C4: Andrews' squeezer (n = 26, stabilised index-1 form), Radau s=5 and s=3
    yield dict(name="andrews_radau5_ens", problem="andrews",
               params=perturbed_params("andrews", 6, 0.05, 2), method="Radau",
               rtol=1e-4, atol=1e-4, options=dict(stages=5), t_eval=np.linspace(0, 0.03, 7))
    yield dict(name="andrews_radau3_ens", problem="andrews",
               params=perturbed_params("andrews", 4, 0.05, 2), method="Radau",
               rtol=1e-6, atol=1e-6, options=dict(stages=3))
    # --- option matrix and stage counts on the group-per-system kernels (n > 4)
    yield dict(name="akzo_radau3_emb0", problem="akzo_nobel", params=perturbed_params("akzo_nobel", 4, 0.05, 3),
               method="Radau", rtol=1e-6, atol=1e-8, options=dict(newton_iter_embedded=0))
    yield dict(name="akzo_radau3_emb2", problem="akzo_nobel", params=perturbed_params("akzo_nobel", 4, 0.05, 3),
               method="Radau", rtol=1e-6, atol=1e-8, options=dict(newton_iter_embedded=2))
    yield dict(name="akzo_radau3_cew05", problem="akzo_nobel", params=perturbed_params("akzo_nobel", 4, 0.05, 3),
               method="Radau", rtol=1e-6, atol=1e-8, options=dict(continuous_error_weight=0.5),
               t_eval=np.linspace(0, 180, 7))
    yield dict(name="akzo_radau5_ens", problem="akzo_nobel", params=perturbed_params("akzo_nobel", 4, 0.05, 3),
               method="Radau", rtol=1e-8, atol=1e-8, options=dict(stages=5))
    yield dict(name="akzo_bdf_variants", problem="akzo_nobel", params=perturbed_params("akzo_nobel", 4, 0.05, 3),
               method="BDF", rtol=1e-6, atol=1e-8, options=dict(NDF_strategy="efficiency", max_order=3, max_step=5.0))
    yield dict(name="andrews_bdf_ens", problem="andrews", params=perturbed_params("andrews", 3, 0.05, 2),
               method="BDF", rtol=1e-5, atol=1e-5, t_eval=np.linspace(0, 0.03, 4))
    # --- row N3 of SURVEY §8f: events (dae.py:30-109, 426-490) as threshold events
    yield dict(name="robertson_radau3_events", problem="robertson",
               params=perturbed_params("robertson", 8, 0.1, 0), method="Radau", rtol=1e-6, atol=1e-6,
               t_eval=np.logspace(-5, 7, 13),
               options=dict(events=[dict(component=0, value=0.9), dict(component=1, value=3e-5, direction=-1),
                                    dict(component=2, value=0.5, terminal=True)]))
    yield dict(name="robertson_bdf_events", problem="robertson",
               params=perturbed_params("robertson", 8, 0.1, 0), method="BDF", rtol=1e-6, atol=1e-6,
               options=dict(events=[dict(component=0, value=0.9), dict(component=1, value=3e-5, direction=-1),
                                    dict(component=2, value=0.5, terminal=True)]))
    # an event on y' (the Radau dense output reproduces y' at both step ends; with the reference's
    # BDF interpolant the same event makes brentq raise "f(a) and f(b) must have different signs")
    yield dict(name="robertson_radau3_events_yp", problem="robertson",
               params=perturbed_params("robertson", 4, 0.1, 0), method="Radau", rtol=1e-6, atol=1e-6,
               t_span=(0.0, 1e3),
               options=dict(events=[dict(component=1, of="yp", value=0.0), dict(component=0, of="yp", value=-0.01, direction=1)]))
    yield dict(name="transamp_radau3_events", problem="transamp",
               params=perturbed_params("transamp", 4, 0.02, 1), method="Radau", rtol=1e-8, atol=1e-8,
               options=dict(events=[dict(component=7, value=-1.0, direction=1),
                                    dict(component=7, value=-1.0, direction=-1, terminal=5),
                                    dict(component=4, value=3.0)]))
    yield dict(name="akzo_bdf_events", problem="akzo_nobel",
               params=perturbed_params("akzo_nobel", 8, 0.05, 0), method="BDF", rtol=1e-8, atol=1e-8,
               options=dict(events=[dict(component=0, value=0.2, terminal=True), dict(component=2, value=0.1)]))
    # event functions that are not thresholds on one component (the problem functor's own g(t, y, y'; p),
    # given to the reference as the NumPy twins Problem.events[k]): y - y' and y - a t / 2 on the event
    # test problem of tests/test_events.py, y1 - y3 and a y' combination on Robertson / Akzo Nobel
    yield dict(name="cosine_radau3_functor_events", problem="cosine", params=np.array([[1.0], [1.3]]),
               method="Radau", rtol=1e-10, atol=1e-10,
               options=dict(events=[dict(functor=0), dict(functor=1, direction=-1), dict(functor=0, direction=1, terminal=2)]))
    yield dict(name="cosine_bdf_functor_events", problem="cosine", params=np.array([[1.0], [1.3]]),
               method="BDF", rtol=1e-10, atol=1e-10,
               options=dict(events=[dict(functor=0), dict(functor=1, direction=-1)]))
    yield dict(name="robertson_radau3_functor_events", problem="robertson",
               params=perturbed_params("robertson", 8, 0.1, 0), method="Radau", rtol=1e-6, atol=1e-6,
               options=dict(events=[dict(functor=1), dict(functor=0, terminal=True)]))
    yield dict(name="akzo_radau3_functor_events", problem="akzo_nobel",
               params=perturbed_params("akzo_nobel", 4, 0.05, 0), method="Radau", rtol=1e-8, atol=1e-8,
               options=dict(events=[dict(functor=0), dict(component=1, value=1e-3, direction=-1)]))
    yield dict(name="akzo_bdf_functor_events", problem="akzo_nobel",
               params=perturbed_params("akzo_nobel", 4, 0.05, 0), method="BDF", rtol=1e-8, atol=1e-8,
               options=dict(events=[dict(functor=0, terminal=True)]))
    # --- row N2 of SURVEY §8f: analytic Jacobian callables (base.py:299-331)
    yield dict(name="robertson_radau3_jac", problem="robertson",
               params=perturbed_params("robertson", 16, 0.1, 0), method="Radau",
               rtol=1e-6, atol=1e-6, options=dict(jac="analytic"), t_eval=np.logspace(-5, 7, 13))
    yield dict(name="robertson_bdf_jac", problem="robertson",
               params=perturbed_params("robertson", 8, 0.1, 0), method="BDF",
               rtol=1e-6, atol=1e-6, options=dict(jac="analytic"))
    yield dict(name="akzo_bdf_jac", problem="akzo_nobel",
               params=perturbed_params("akzo_nobel", 8, 0.05, 0), method="BDF",
               rtol=1e-8, atol=1e-8, options=dict(jac="analytic"))
    yield dict(name="akzo_radau3_jac", problem="akzo_nobel",
               params=perturbed_params("akzo_nobel", 8, 0.05, 0), method="Radau",
               rtol=1e-8, atol=1e-8, options=dict(jac="analytic"))
    yield dict(name="transamp_radau3_jac", problem="transamp",
               params=perturbed_params("transamp", 4, 0.02, 1), method="Radau",
               rtol=1e-10, atol=1e-6, options=dict(jac="analytic"))
    # --- failure paths (dae.py:451-453: solve_dae breaks BEFORE recording the failed step)
    # deterministic: y' = p y^2 blows up at t* = 1/p inside t_span; both methods end with
    # TOO_SMALL_STEP (base.py:158, radau.py:515-516, bdf.py:297-298) just before t*
    yield dict(name="blowup_radau3_fail", problem="blowup", params=np.array([[1.0], [0.8], [1.25], [0.6]]),
               method="Radau", rtol=1e-6, atol=1e-8, t_eval=np.linspace(0.0, 2.0, 41))
    yield dict(name="blowup_bdf_fail", problem="blowup", params=np.array([[1.0], [0.8], [1.25], [0.6]]),
               method="BDF", rtol=1e-6, atol=1e-8, t_eval=np.linspace(0.0, 2.0, 41))
    # the reference's own n = 5 circuit (examples/daes/stiff_transistor_amplifier.py) fails at
    # rtol = atol = 1e-6 with both methods (BASELINE.md §2 caveat); WHERE it fails is chaotic in the
    # reference itself (the twins of this file end between t = 8e-4 and 0.1)
    yield dict(name="transamp1_radau3_fail", problem="transamp1", params=perturbed_params("transamp1", 6, 0.02, 1),
               method="Radau", rtol=1e-6, atol=1e-6, t_eval=np.linspace(0.0, 0.2, 41))
    yield dict(name="transamp1_bdf_fail", problem="transamp1", params=perturbed_params("transamp1", 4, 0.02, 1),
               method="BDF", rtol=1e-6, atol=1e-6, t_eval=np.linspace(0.0, 0.2, 41))
    # ... and succeeds with the example's own settings (atol = 1e-3, rtol = 1e-6, stages = 5)
    yield dict(name="transamp1_radau5_example", problem="transamp1", params=perturbed_params("transamp1", 4, 0.02, 1),
               method="Radau", rtol=1e-6, atol=1e-3, options=dict(stages=5), t_eval=np.linspace(0.0, 0.2, 21))
    # non-finite iteration matrix: scipy's lu_factor(check_finite=True) raises ValueError (base.py:234)
    # from DaeSolver.__init__ / _step_impl; recorded as status -2
    p_nan = perturbed_params("robertson", 3, 0.1, 0); p_nan[1, 1] = np.nan
    yield dict(name="robertson_radau3_nonfinite", problem="robertson", params=p_nan, method="Radau",
               rtol=1e-6, atol=1e-6, twins=())
    yield dict(name="robertson_bdf_nonfinite", problem="robertson", params=p_nan, method="BDF",
               rtol=1e-6, atol=1e-6, twins=())
    # --- tighter tolerances where the step sequences of the n > 4 problems are NOT chaotic: the
    # unwidened north_star band has to hold per instance (measured: twin distances << 1)
    yield dict(name="andrews_radau5_tight", problem="andrews", params=perturbed_params("andrews", 4, 0.05, 2),
               method="Radau", rtol=1e-8, atol=1e-8, options=dict(stages=5), t_eval=np.linspace(0, 0.03, 7))
    yield dict(name="transamp_bdf_tight", problem="transamp", params=perturbed_params("transamp", 4, 0.02, 1),
               method="BDF", rtol=1e-12, atol=1e-9)
    # backward integration (direction = -1) of a non-stiff stretch
    yield dict(name="robertson_radau3_backward", problem="robertson",
               params=rob_nom, method="Radau", rtol=1e-6, atol=1e-8, t_span=(1e-3, 0.0),
               y0=np.array([[9.99960003e-01, 3.64862817e-05, 3.51085270e-06]]),
               yp0=None, t_eval=np.linspace(1e-3, 0.0, 6))


def main():
    import multiprocessing as mp
    sel = [a for a in sys.argv[1:] if not a.startswith("-")]
    skip = [a[1:] for a in sys.argv[1:] if a.startswith("-")]      # -substring: leave those cases alone
    with mp.get_context("fork").Pool(min(5, os.cpu_count() or 1)) as pool:
        for c in cases():
            if (sel and not any(s in c["name"] for s in sel)) or any(s in c["name"] for s in skip):
                continue
            if c["name"] == "robertson_radau3_backward":
                # consistent yp0 for the given y0 (differential rows explicit, algebraic row's yp3 = -(yp1+yp2))
                y = c["y0"][0]
                k1, k2, k3 = c["params"][0]
                yp = np.array([-k1 * y[0] + k2 * y[1] * y[2],
                               k1 * y[0] - k2 * y[1] * y[2] - k3 * y[1] ** 2, 0.0])
                yp[2] = -(yp[0] + yp[1])
                y[2] = 1.0 - y[0] - y[1]
                c["yp0"] = yp[None, :]
            run_case(pool=pool, **c)


if __name__ == "__main__":
    main()
