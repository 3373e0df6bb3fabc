"""Live runs of the UNMODIFIED Python reference (JonasBreuling/scipy_dae) — TEST INFRASTRUCTURE.

``oracle/make_ref.sh`` copies the reference's eight hot-path files byte for byte into
``oracle/_ref/`` (git-ignored, travels to the GPU box) together with their SHA-256 sums;
``import_reference()`` verifies the sums and imports ``scipy_dae`` from there (or from
``/root/reference`` in the build container).  On top of it:

* ``run_instance`` / ``solve_ensemble``: the reference's ``solve_dae(fun, t_span, y0, yp0, method=,
  rtol=, atol=, t_eval=, args=params_i, **options)`` (dae.py:115-547, ``args`` path dae.py:365-383)
  instance by instance over a ``multiprocessing.Pool``, returning the same fields as
  ``oracle.oracle.solve_ensemble`` and the product's ``BatchedDaeResult`` so that parity tests compare
  them one to one.  Accepted / rejected counts are not exposed by the reference; they are derived
  without editing it by rebinding the module-level ``solve_collocation_system`` /
  ``solve_bdf_system`` (looked up as globals inside ``_step_impl``: radau.py:540, bdf.py:323).
* ``python -m oracle.reference --bench ...``: the CPU baseline of BASELINE.md §3 (the untouched
  ``solve_dae`` over the bench ensemble, ``Pool(os.cpu_count())``, BLAS threads pinned to 1), run as
  its own process so that the thread-count environment is in place before NumPy loads and no
  CUDA context is forked.  Prints one JSON line.

Only ``tests/``, ``oracle/gen_golden*.py``, ``__graft_entry__`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs use this module.  ``scipy_dae_b200`` never imports it.
"""
from __future__ import annotations

import hashlib
import json
import os
import sys
import time
import warnings

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)
REF_DIR = os.path.join(HERE, "_ref")
REF_TREE = "/root/reference"

_ref_modules = None


def verify_ref(ref_dir: str = REF_DIR) -> int:
    """Check oracle/_ref against its SHA256SUMS; returns the number of files verified."""
    sums = os.path.join(ref_dir, "SHA256SUMS")
    n = 0
    with open(sums) as fh:
        for line in fh:
            digest, name = line.split()
            with open(os.path.join(ref_dir, name), "rb") as f:
                if hashlib.sha256(f.read()).hexdigest() != digest:
                    raise RuntimeError(f"oracle/_ref/{name} differs from the reference it was copied from")
            n += 1
    return n


def available() -> bool:
    return os.path.exists(os.path.join(REF_DIR, "SHA256SUMS")) or os.path.isdir(os.path.join(REF_TREE, "scipy_dae"))


def ref_source() -> str:
    """Where the reference is imported from: oracle/_ref (verified copy) or /root/reference."""
    if os.path.exists(os.path.join(REF_DIR, "SHA256SUMS")):
        return REF_DIR
    if os.path.isdir(os.path.join(REF_TREE, "scipy_dae")):
        return REF_TREE
    raise ImportError("the reference is not available: run oracle/make_ref.sh in the build container "
                      "(it creates oracle/_ref/, which travels to the GPU box)")


def import_reference():
    """Returns (solve_dae, radau_module, bdf_module, base_module) of the unmodified reference, with the Newton
    solvers wrapped by counters (see module docstring)."""
    global _ref_modules
    if _ref_modules is not None:
        return _ref_modules
    src = ref_source()
    if src == REF_DIR:
        verify_ref()
    if src not in sys.path:
        sys.path.insert(0, src)
    import scipy_dae.integrate._dae.radau as R
    import scipy_dae.integrate._dae.bdf as B
    import scipy_dae.integrate._dae.base as BASE
    from scipy_dae.integrate import solve_dae
    if not os.path.abspath(R.__file__).startswith(os.path.abspath(src)):
        raise ImportError(f"scipy_dae was imported from {R.__file__}, not from {src}")
    if not getattr(R.solve_collocation_system, "_counted", False):
        R.solve_collocation_system = _wrap(R.solve_collocation_system)
        B.solve_bdf_system = _wrap(B.solve_bdf_system)
    _ref_modules = (solve_dae, R, B, BASE)
    return _ref_modules


_cnt = dict(calls=0, nonconv=0, iters=0)


def _wrap(orig):
    def w(*a, **k):
        out = orig(*a, **k)  # out[0] = converged, out[1] = n_iter
        _cnt["calls"] += 1
        _cnt["nonconv"] += (not out[0])
        _cnt["iters"] += out[1]
        return out
    w._counted = True
    return w


def _threshold_event(e):
    j, c, of_yp = int(e["component"]), float(e["value"]), e.get("of", "y") == "yp"
    fn = (lambda t, y, yp, *a: yp[j] - c) if of_yp else (lambda t, y, yp, *a: y[j] - c)
    fn.direction = float(e.get("direction", 0))
    fn.terminal = e.get("terminal", False)
    return fn


def run_instance(problem, params, y0, yp0, t_span, method, rtol, atol, t_eval, options,
                 record_steps=False):
    """One solve_dae call of the reference (two with t_eval: the first without it gives the step
    ends and with them the accepted count).  `problem` is a scipy_dae_b200.problems.Problem (its
    NumPy residual is the reference's `fun`), `params` its `args`."""
    import numpy as np
    solve_dae = import_reference()[0]
    for k in _cnt:
        _cnt[k] = 0
    options = dict(options)
    ev_spec = options.pop("events", None)
    ev_cap = options.pop("event_cap", 16)
    if ev_spec:                               # events as the reference's event callables
        options["events"] = [e if callable(e) else
                             (problem.event_callable(e["functor"], e) if "functor" in e else _threshold_event(e))
                             for e in ev_spec]
    if options.get("jac") == "analytic":     # the problem's analytic Jacobians as the `jac` callable
        options["jac"] = problem.jacobian
    try:
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            # pass 1: t_eval=None => sol.t are the step ends => accepted count, final state
            sol = solve_dae(problem.residual, t_span, y0, yp0, method=method, rtol=rtol,
                            atol=atol, args=tuple(params), **options)
    except ValueError as exc:
        # scipy's lu_factor / lu_solve(check_finite=True) on a non-finite iteration matrix or
        # right-hand side (base.py:234,238): the batched path reports it as status -2
        if "infs or NaNs" not in str(exc):
            raise
        nan = np.full(problem.n, np.nan)
        out = dict(status=-2, t_final=float(t_span[0]), y_final=np.array(y0, dtype=float), yp_final=np.array(yp0, dtype=float),
                   nfev=0, njev=0, nlu=0, n_accepted=0, n_newton_calls=0, n_newton_fail=0, n_newton_iter=0,
                   n_err_rejected=0, message=str(exc))
        if t_eval is not None:
            m = len(t_eval)
            out.update(n_eval_valid=0, y_eval=np.full((problem.n, m), np.nan), yp_eval=np.full((problem.n, m), np.nan))
        return out
    accepted = len(sol.t) - 1
    out = dict(
        status=sol.status, t_final=sol.t[-1], y_final=sol.y[:, -1], yp_final=sol.yp[:, -1],
        nfev=sol.nfev, njev=sol.njev, nlu=sol.nlu, n_accepted=accepted,
        n_newton_calls=_cnt["calls"], n_newton_fail=_cnt["nonconv"],
        n_newton_iter=_cnt["iters"],
        n_err_rejected=_cnt["calls"] - _cnt["nonconv"] - accepted,
    )
    if sol.status == -1:
        # the failing step's attempts end with the TOO_SMALL_STEP return, not with an accepted step
        # or a counted rejection (radau.py:515-516, bdf.py:297-298): keep the derived count >= 0
        out["n_err_rejected"] = max(out["n_err_rejected"], 0)
    if record_steps:
        out["t_steps"] = sol.t
        out["y_steps"] = sol.y
        out["yp_steps"] = sol.yp
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


# ---- ensembles over a process pool ------------------------------------------------------------
_job = {}


def _pool_init(job):
    for k in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS"):
        os.environ.setdefault(k, "1")
    if REPO not in sys.path:
        sys.path.insert(0, REPO)
    from scipy_dae_b200.problems import get_problem
    _job.update(job)
    _job["problem"] = get_problem(job["problem"])
    import_reference()


def _pool_task(i):
    j = _job
    return run_instance(j["problem"], j["params"][i], j["y0"][i], j["yp0"][i], j["t_span"], j["method"],
                        j["rtol"], j["atol"], j["t_eval"], j["options"])


def _pool_task_plain(i):
    """The bare call BASELINE.md §3 times: one solve_dae per instance, nothing else."""
    j = _job
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        sol = import_reference()[0](j["problem"].residual, j["t_span"], j["y0"][i], j["yp0"][i], method=j["method"],
                                    rtol=j["rtol"], atol=j["atol"], args=tuple(j["params"][i]), **j["options"])
    return sol.status, sol.y[:, -1]


def _make_pool(job, processes):
    import multiprocessing as mp
    ctx = mp.get_context("fork" if "torch" not in sys.modules else "spawn")
    return ctx.Pool(processes, initializer=_pool_init, initargs=(job,))


def solve_ensemble(problem, params, y0, yp0, t_span, method="Radau", rtol=1e-3, atol=1e-6, t_eval=None,
                   processes=None, **options):
    """N instances through the live reference.  Host AoS arrays params[N, p], y0[N, n], yp0[N, n];
    returns dict(status, t_final, y_final[N, n], yp_final, nfev, njev, nlu, n_accepted, n_err_rejected,
    n_newton_fail, n_newton_iter [, n_eval_valid, y_eval[N, n, m], yp_eval])."""
    import numpy as np
    params = np.atleast_2d(np.asarray(params, dtype=float))
    N = params.shape[0]
    name = problem if isinstance(problem, str) else problem.name
    job = dict(problem=name, params=params, y0=np.atleast_2d(y0), yp0=np.atleast_2d(yp0), t_span=tuple(t_span),
               method=method, rtol=rtol, atol=atol, t_eval=t_eval, options=options)
    processes = processes or os.cpu_count() or 1
    if processes == 1 or N == 1:
        _pool_init(job)
        res = [_pool_task(i) for i in range(N)]
    else:
        with _make_pool(job, min(processes, N)) as pool:
            res = pool.map(_pool_task, range(N), chunksize=max(1, N // (8 * processes)))
    out = {}
    for key in ("status", "t_final", "nfev", "njev", "nlu", "n_accepted", "n_newton_calls", "n_newton_fail",
                "n_newton_iter", "n_err_rejected"):
        out[key] = np.array([r[key] for r in res])
    out["y_final"] = np.stack([r["y_final"] for r in res])
    out["yp_final"] = np.stack([r["yp_final"] for r in res])
    if t_eval is not None:
        out["n_eval_valid"] = np.array([r["n_eval_valid"] for r in res])
        out["y_eval"] = np.stack([r["y_eval"] for r in res])       # [N, n, m]
        out["yp_eval"] = np.stack([r["yp_eval"] for r in res])
    if "t_events" in res[0]:
        for key in ("t_events", "y_events", "yp_events", "n_events_found"):
            out[key] = np.stack([r[key] for r in res])
    return out


# ---- the CPU baseline of BASELINE.md §3 --------------------------------------------------------
def bench(problem="robertson", method="Radau", rtol=1e-6, atol=1e-6, t_span=(0.0, 1e7), sigma=0.1, seed=0,
          sample=16384, budget_s=30.0, steps=1, warmup=0, processes=None, options=None):
    """solve_dae over the first `sample` instances of the seeded ensemble (time-bounded: the sample
    shrinks to what `budget_s` seconds per step admit at the rate a probe measures), on a pool of
    `processes` (default os.cpu_count()) single-threaded workers; pool start-up is not timed."""
    import numpy as np
    if REPO not in sys.path:
        sys.path.insert(0, REPO)
    from scipy_dae_b200.problems import perturbed_params, initial_conditions
    P = processes or os.cpu_count() or 1
    params = perturbed_params(problem, sample, sigma, seed)
    y0, yp0 = initial_conditions(problem, params)
    job = dict(problem=problem, params=params, y0=y0, yp0=yp0, t_span=tuple(t_span), method=method, rtol=rtol,
               atol=atol, t_eval=None, options=dict(options or {}))
    with _make_pool(job, P) as pool:
        probe = min(sample, 16 * P)
        pool.map(_pool_task_plain, range(min(sample, P)), chunksize=1)      # workers up, imports done
        t = time.perf_counter()
        pool.map(_pool_task_plain, range(probe), chunksize=1)
        rate = probe / (time.perf_counter() - t)
        n = int(min(sample, max(probe, budget_s * rate)))
        chunk = max(1, n // (8 * P))
        for _ in range(warmup):
            pool.map(_pool_task_plain, range(n), chunksize=chunk)
        times, ok = [], True
        for _ in range(steps):
            t = time.perf_counter()
            r = pool.map(_pool_task_plain, range(n), chunksize=chunk)
            times.append(time.perf_counter() - t)
            ok = ok and all(s == 0 for s, _ in r)
    total = float(sum(times))
    return dict(value=n * steps / total, unit="solves/s", cores=P, kind="reference", instances=n,
                steps=steps, warmup=warmup, seconds_per_step=total / steps, all_finished=bool(ok),
                source=os.path.relpath(ref_source(), REPO) if ref_source().startswith(REPO) else ref_source(),
                sample=f"first {n} instances of the seed-{seed} {problem} ensemble per step, unmodified "
                       f"scipy_dae solve_dae(method={method!r}, rtol={rtol:g}, atol={atol:g}, args=params_i), "
                       f"multiprocessing.Pool({P}) with BLAS/OMP threads = 1, {total / steps:.2f} s per step")


def main(argv=None):
    import argparse
    ap = argparse.ArgumentParser()
    ap.add_argument("--bench", action="store_true")
    ap.add_argument("--problem", default="robertson")
    ap.add_argument("--method", default="Radau")
    ap.add_argument("--rtol", type=float, default=1e-6)
    ap.add_argument("--atol", type=float, default=1e-6)
    ap.add_argument("--t-span", type=float, nargs=2, default=(0.0, 1e7))
    ap.add_argument("--sigma", type=float, default=0.1)
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--sample", type=int, default=16384)
    ap.add_argument("--budget-s", type=float, default=30.0)
    ap.add_argument("--steps", type=int, default=1)
    ap.add_argument("--warmup", type=int, default=0)
    ap.add_argument("--processes", type=int, default=0)
    ap.add_argument("--options", default="{}")
    a = ap.parse_args(argv)
    if a.bench:
        print(json.dumps(bench(a.problem, a.method, a.rtol, a.atol, tuple(a.t_span), a.sigma, a.seed, a.sample,
                               a.budget_s, a.steps, a.warmup, a.processes or None, json.loads(a.options))))


if __name__ == "__main__":
    for _k in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS"):
        os.environ.setdefault(_k, "1")
    main()
