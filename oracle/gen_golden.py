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

Accepted / rejected counts are not exposed by the reference; they are derived
without editing it by rebinding the module-level ``solve_collocation_system`` /
``solve_bdf_system`` (looked up as globals inside ``_step_impl``: radau.py:540,
bdf.py:323).

Output: ``tests/golden/<case>.npz`` (small, committed).  Usage::

    python oracle/gen_golden.py            # all cases
    python oracle/gen_golden.py robertson  # cases whose name contains the substring
"""
import json
import os
import sys
import time
import warnings

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
sys.path.insert(0, "/root/reference")

import scipy  # noqa: E402
import scipy_dae.integrate._dae.radau as R  # noqa: E402
import scipy_dae.integrate._dae.bdf as B  # noqa: E402
from scipy_dae.integrate import solve_dae  # noqa: E402

from scipy_dae_b200.problems import (  # noqa: E402
    get_problem, perturbed_params, initial_conditions)

GOLDEN_DIR = os.path.join(REPO, "tests", "golden")

_cnt = dict(calls=0, nonconv=0, iters=0)


def _wrap(orig):
    def w(*a, **k):
        out = orig(*a, **k)  # out[0] = converged, out[1] = n_iter
        _cnt["calls"] += 1
        _cnt["nonconv"] += (not out[0])
        _cnt["iters"] += out[1]
        return out
    return w


R.solve_collocation_system = _wrap(R.solve_collocation_system)
B.solve_bdf_system = _wrap(B.solve_bdf_system)


def run_instance(problem, params, y0, yp0, t_span, method, rtol, atol, t_eval, options,
                 record_steps=False):
    for k in _cnt:
        _cnt[k] = 0
    options = dict(options)
    ev_spec = options.pop("events", None)
    ev_cap = options.pop("event_cap", 16)
    if ev_spec:                               # threshold events as the reference's event callables
        def make(e):
            j, c, of_yp = int(e["component"]), float(e["value"]), e.get("of", "y") == "yp"
            fn = (lambda t, y, yp, *a: yp[j] - c) if of_yp else (lambda t, y, yp, *a: y[j] - c)
            fn.direction = float(e.get("direction", 0))
            fn.terminal = e.get("terminal", False)
            return fn
        options["events"] = [make(e) for e in ev_spec]
    if options.get("jac") == "analytic":     # the problem's analytic Jacobians as the `jac` callable
        options["jac"] = problem.jacobian
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        # pass 1: t_eval=None => sol.t are the step ends => accepted count, final state
        sol = solve_dae(problem.residual, t_span, y0, yp0, method=method, rtol=rtol,
                        atol=atol, args=tuple(params), **options)
    accepted = len(sol.t) - 1
    out = dict(
        status=sol.status, t_final=sol.t[-1], y_final=sol.y[:, -1], yp_final=sol.yp[:, -1],
        nfev=sol.nfev, njev=sol.njev, nlu=sol.nlu, n_accepted=accepted,
        n_newton_calls=_cnt["calls"], n_newton_fail=_cnt["nonconv"],
        n_newton_iter=_cnt["iters"],
        n_err_rejected=_cnt["calls"] - _cnt["nonconv"] - accepted,
    )
    if record_steps:
        out["t_steps"] = sol.t
        out["y_steps"] = sol.y
    if ev_spec:
        ne = len(ev_spec)
        te = np.full((ne, ev_cap), np.nan); yev = np.full((ne, ev_cap, problem.n), np.nan)
        ypev = np.full((ne, ev_cap, problem.n), np.nan); nf = np.zeros(ne, dtype=np.int64)
        for k in range(ne):
            nf[k] = len(sol.t_events[k])
            for i in range(min(nf[k], ev_cap)):
                te[k, i] = sol.t_events[k][i]; yev[k, i] = sol.y_events[k][i]; ypev[k, i] = sol.yp_events[k][i]
        out.update(t_events=te, y_events=yev, yp_events=ypev, n_events_found=nf)
    if t_eval is not None:
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            sol2 = solve_dae(problem.residual, t_span, y0, yp0, method=method, rtol=rtol,
                             atol=atol, args=tuple(params), t_eval=t_eval, **options)
        m = len(t_eval)
        ye = np.full((problem.n, m), np.nan)
        ype = np.full((problem.n, m), np.nan)
        nv = 0 if isinstance(sol2.t, list) else len(sol2.t)
        if nv:
            ye[:, :nv] = sol2.y
            ype[:, :nv] = sol2.yp
        out["n_eval_valid"] = nv
        out["y_eval"] = ye
        out["yp_eval"] = ype
    return out


def run_case(name, problem, params, method, rtol, atol, t_span=None, t_eval=None,
             options=None, record_steps=False, y0=None, yp0=None):
    problem = get_problem(problem)
    options = dict(options or {})
    params = np.atleast_2d(np.asarray(params, dtype=float))
    N = params.shape[0]
    if y0 is None:
        y0, yp0 = initial_conditions(problem, params)
    t_span = tuple(t_span or problem.t_span)
    t0 = time.perf_counter()
    res = [run_instance(problem, params[i], y0[i], yp0[i], t_span, method, rtol, atol,
                        t_eval, options, record_steps=record_steps and i == 0)
           for i in range(N)]
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
    meta = dict(case=name, problem=problem.name, method=method, rtol=rtol,
                atol=atol if np.isscalar(atol) else list(atol), t_span=list(t_span),
                options=options, N=N, reference="JonasBreuling/scipy_dae v0.1.0 (/root/reference)",
                scipy=scipy.__version__, numpy=np.__version__, wall_s=round(wall, 3),
                generator="oracle/gen_golden.py")
    arrays["meta"] = np.array(json.dumps(meta))
    os.makedirs(GOLDEN_DIR, exist_ok=True)
    np.savez_compressed(os.path.join(GOLDEN_DIR, name + ".npz"), **arrays)
    print(f"{name}: N={N} wall={wall:.2f}s status={arrays['status'].tolist()[:8]} "
          f"acc={arrays['n_accepted'].tolist()[:8]} nfev={arrays['nfev'].tolist()[:4]}")


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
    # backward integration (direction = -1) of a non-stiff stretch
    yield dict(name="robertson_radau3_backward", problem="robertson",
               params=rob_nom, method="Radau", rtol=1e-6, atol=1e-8, t_span=(1e-3, 0.0),
               y0=np.array([[9.99960003e-01, 3.64862817e-05, 3.51085270e-06]]),
               yp0=None, t_eval=np.linspace(1e-3, 0.0, 6))


def main():
    sel = sys.argv[1:]
    for c in cases():
        if sel and not any(s in c["name"] for s in sel):
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
        run_case(**c)


if __name__ == "__main__":
    main()
