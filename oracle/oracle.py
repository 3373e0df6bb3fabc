"""ctypes binding of the CPU oracle (oracle/libdae_oracle.so) — TEST INFRASTRUCTURE.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this module.  The product package
``scipy_dae_b200`` never does.

``solve_ensemble`` integrates N independent instances with the plain-C restatement of the
reference's RadauDAE / BDFDAE step loops (dae_oracle.c) and returns the same fields as the
product's batched entry point, so the parity tests can compare them one to one.
"""
from __future__ import annotations

import ctypes
import os
import subprocess
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)
if REPO not in sys.path:
    sys.path.insert(0, REPO)

from scipy_dae_b200.constants import CRadauTables, radau_tables_c, NDF_STRATEGIES  # noqa: E402

LIB_PATH = os.path.join(HERE, "libdae_oracle.so")

COUNTER_NAMES = ("nfev", "njev", "nlu", "nsolve", "n_accepted", "n_err_rejected",
                 "n_newton_fail", "n_newton_iter")


class RadauOpts(ctypes.Structure):
    _fields_ = [("rtol", ctypes.c_double), ("max_step", ctypes.c_double),
                ("first_step", ctypes.c_double), ("continuous_error_weight", ctypes.c_double),
                ("jac_recompute_rate", ctypes.c_double),
                ("deadband_lo", ctypes.c_double), ("deadband_hi", ctypes.c_double),
                ("newton_max_iter", ctypes.c_int), ("newton_iter_embedded", ctypes.c_int),
                ("max_steps", ctypes.c_int64)]


class BdfOpts(ctypes.Structure):
    _fields_ = [("rtol", ctypes.c_double), ("max_step", ctypes.c_double),
                ("first_step", ctypes.c_double), ("max_order", ctypes.c_int),
                ("ndf_strategy", ctypes.c_int), ("max_steps", ctypes.c_int64)]


def build(force: bool = False) -> str:
    srcs = [os.path.join(HERE, f) for f in ("dae_oracle.c", "cic_oracle.c", "problems.c", "dae_oracle.h")]
    if (not force and os.path.exists(LIB_PATH)
            and os.path.getmtime(LIB_PATH) >= max(os.path.getmtime(s) for s in srcs)):
        return LIB_PATH
    subprocess.run(["make", "-C", HERE, "-B", "libdae_oracle.so"], check=True,
                   stdout=subprocess.DEVNULL)
    return LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = ctypes.CDLL(LIB_PATH)
        L.orc_problem_lookup.restype = ctypes.c_int
        L.orc_problem_lookup.argtypes = [ctypes.c_char_p, ctypes.POINTER(ctypes.c_int),
                                         ctypes.POINTER(ctypes.c_int)]
        L.orc_solve_ensemble.restype = ctypes.c_int
        L.orc_problem_eval.restype = ctypes.c_int
        L.orc_select_initial_step.restype = ctypes.c_double
        _lib = L
    return _lib


def _p(a, ctype=ctypes.c_double):
    return None if a is None else a.ctypes.data_as(ctypes.POINTER(ctype))


def problem_id(name: str):
    n, npar = ctypes.c_int(), ctypes.c_int()
    pid = lib().orc_problem_lookup(name.encode(), ctypes.byref(n), ctypes.byref(npar))
    if pid < 0:
        raise ValueError(f"oracle does not know problem {name!r}")
    return pid, n.value, npar.value


def residual(name, t, y, yp, params):
    pid, n, _ = problem_id(name)
    y = np.ascontiguousarray(y, dtype=float)
    yp = np.ascontiguousarray(yp, dtype=float)
    p = np.ascontiguousarray(params, dtype=float)
    out = np.empty(n)
    lib().orc_problem_eval(ctypes.c_int(pid), ctypes.c_double(t), _p(y), _p(yp), _p(p), _p(out))
    return out


def jacobian(name, t, y, yp, params):
    pid, n, _ = problem_id(name)
    y = np.ascontiguousarray(y, dtype=float)
    yp = np.ascontiguousarray(yp, dtype=float)
    p = np.ascontiguousarray(params, dtype=float)
    Jy, Jyp = np.empty((n, n)), np.empty((n, n))
    rc = lib().orc_problem_jac_eval(ctypes.c_int(pid), ctypes.c_double(t), _p(y), _p(yp), _p(p), _p(Jy), _p(Jyp))
    if rc != 0:
        raise ValueError(f"oracle has no analytic Jacobian for {name!r}")
    return Jy, Jyp


MAX_EVENTS = 4


class Events(ctypes.Structure):
    _fields_ = [("n_events", ctypes.c_int32), ("of_yp", ctypes.c_int32 * MAX_EVENTS),
                ("component", ctypes.c_int32 * MAX_EVENTS), ("value", ctypes.c_double * MAX_EVENTS),
                ("direction", ctypes.c_int32 * MAX_EVENTS), ("max_events", ctypes.c_int32 * MAX_EVENTS)]


def events_struct(events, cls=Events):
    """events: list of dicts(component, value, of='y'|'yp', direction=0, terminal=False|True|int) or
    dicts(functor=index, value=0, direction, terminal) for the problem's own event functions."""
    ev = cls()
    ev.n_events = len(events)
    for k, e in enumerate(events):
        if "functor" in e:
            ev.of_yp[k], ev.component[k] = 2, int(e["functor"])
        else:
            ev.of_yp[k] = 1 if e.get("of", "y") == "yp" else 0
            ev.component[k] = int(e["component"])
        ev.value[k] = float(e.get("value", 0.0))
        ev.direction[k] = int(np.sign(e.get("direction", 0)))
        term = e.get("terminal", False)
        ev.max_events[k] = 0 if term in (False, None) else (1 if term is True else int(term))
    return ev


def solve_ensemble(problem, params, y0, yp0, t_span, method="Radau", rtol=1e-3, atol=1e-6,
                   t_eval=None, nthreads=1, max_steps=0, events=None, event_cap=16, **options):
    """N instances through the C oracle.  Arrays are AoS on the host: params[N,p], y0[N,n]."""
    pid, n, npar = problem_id(problem)
    params = np.ascontiguousarray(np.atleast_2d(params), dtype=float)
    y0 = np.ascontiguousarray(np.atleast_2d(y0), dtype=float)
    yp0 = np.ascontiguousarray(np.atleast_2d(yp0), dtype=float)
    N = params.shape[0]
    assert params.shape == (N, npar) and y0.shape == (N, n) and yp0.shape == (N, n)
    atol_v = np.ascontiguousarray(np.broadcast_to(np.asarray(atol, dtype=float), (n,)))
    t0, t_bound = map(float, t_span)

    jac = options.pop("jac", None)
    if jac not in (None, "analytic"):
        raise ValueError("oracle: jac must be None (finite differences) or 'analytic'")
    ropts = bopts = tb = None
    if method == "Radau":
        s = int(options.pop("stages", 3))
        tb = radau_tables_c(s)
        db = options.pop("controller_deadband", (1.0, 1.2))
        nmi = options.pop("newton_max_iter", None)
        ropts = RadauOpts(rtol=rtol, max_step=float(options.pop("max_step", np.inf)),
                          first_step=float(options.pop("first_step", None) or 0.0),
                          continuous_error_weight=float(options.pop("continuous_error_weight", 0.0)),
                          jac_recompute_rate=float(options.pop("jac_recompute_rate", 1e-3)),
                          deadband_lo=db[0], deadband_hi=db[1],
                          newton_max_iter=int(nmi or 0),
                          newton_iter_embedded=int(options.pop("newton_iter_embedded", 1)),
                          max_steps=max_steps)
        mcode = 0
    elif method == "BDF":
        bopts = BdfOpts(rtol=rtol, max_step=float(options.pop("max_step", np.inf)),
                        first_step=float(options.pop("first_step", None) or 0.0),
                        max_order=int(options.pop("max_order", 5)),
                        ndf_strategy=NDF_STRATEGIES[options.pop("NDF_strategy", "stability")],
                        max_steps=max_steps)
        mcode = 1
    else:
        raise ValueError(method)
    if options:
        raise TypeError(f"unknown options {sorted(options)}")

    m = 0
    te = ye = ype = None
    if t_eval is not None:
        te = np.ascontiguousarray(t_eval, dtype=float)
        m = te.shape[0]
        ye = np.full((N, m, n), np.nan)
        ype = np.full((N, m, n), np.nan)
    nev = np.zeros(N, dtype=np.int64)
    y_final = np.empty((N, n)); yp_final = np.empty((N, n)); t_final = np.empty(N)
    status = np.zeros(N, dtype=np.int32)
    counters = np.zeros((N, 8), dtype=np.int64)

    lib().orc_set_analytic_jac(ctypes.c_int(1 if jac == "analytic" else 0))
    ev = None
    if events:
        ev = events_struct(events)
        ne = ev.n_events
        t_events = np.full((N, ne, event_cap), np.nan)
        y_events = np.full((N, ne, event_cap, n), np.nan)
        yp_events = np.full((N, ne, event_cap, n), np.nan)
        n_found = np.zeros((N, ne), dtype=np.int32)
        lib().orc_set_events(ctypes.byref(ev), ctypes.c_int64(event_cap), _p(t_events), _p(y_events),
                             _p(yp_events), _p(n_found, ctypes.c_int32))
    rc = lib().orc_solve_ensemble(
        ctypes.c_int(pid), ctypes.c_int(mcode), ctypes.c_int64(N),
        _p(params), _p(y0), _p(yp0), ctypes.c_double(t0), ctypes.c_double(t_bound), _p(atol_v),
        ctypes.byref(ropts) if ropts is not None else None,
        ctypes.byref(tb) if tb is not None else None,
        ctypes.byref(bopts) if bopts is not None else None,
        _p(te), ctypes.c_int64(m), _p(ye), _p(ype), _p(nev, ctypes.c_int64),
        _p(y_final), _p(yp_final), _p(t_final), _p(status, ctypes.c_int32),
        _p(counters, ctypes.c_int64), ctypes.c_int(nthreads))
    lib().orc_set_analytic_jac(ctypes.c_int(0))
    lib().orc_set_events(None, ctypes.c_int64(0), None, None, None, None)
    if rc != 0:
        raise RuntimeError(f"orc_solve_ensemble returned {rc}")
    out = dict(status=status, t_final=t_final, y_final=y_final, yp_final=yp_final,
               n_eval_valid=nev)
    for i, name in enumerate(COUNTER_NAMES):
        out[name] = counters[:, i].copy()
    if t_eval is not None:
        out["y_eval"] = ye      # [N, m, n]
        out["yp_eval"] = ype
    if ev is not None:
        out.update(t_events=t_events, y_events=y_events, yp_events=yp_events, n_events_found=n_found)
    return out


CIC_MESSAGES = {0: "converged", -1: "Convergence failed.", -2: "Too many fixed components",
                -3: "Index greater than one."}


def consistent_initial_conditions(problem, t0, y0, yp0, params, jac=None, fixed_y0=None, fixed_yp0=None,
                                  rtol=1e-8, atol=1e-8, newton_maxiter=10, chord_iter=3, safety=0.5):
    """N instances of the reference's consistent_initial_conditions (common.py:186-258) through the
    C oracle.  Returns dict(y0[N,n], yp0[N,n], f[N,n], status[N], n_iter[N])."""
    pid, n, npar = problem_id(problem)
    params = np.ascontiguousarray(np.atleast_2d(params), dtype=float)
    y0 = np.array(np.atleast_2d(y0), dtype=float, order="C")
    yp0 = np.array(np.atleast_2d(yp0), dtype=float, order="C")
    N = y0.shape[0]
    assert params.shape == (N, npar) and y0.shape == (N, n) and yp0.shape == (N, n)
    fy = np.zeros(n, dtype=np.int32); fyp = np.zeros(n, dtype=np.int32)
    if fixed_y0 is not None:
        fy[np.asarray(fixed_y0, dtype=int)] = 1
    if fixed_yp0 is not None:
        fyp[np.asarray(fixed_yp0, dtype=int)] = 1
    atol_v = np.ascontiguousarray(np.broadcast_to(np.asarray(atol, dtype=float), (n,)))
    f = np.empty((N, n)); status = np.zeros(N, dtype=np.int32); n_iter = np.zeros(N, dtype=np.int32)
    rc = lib().orc_consistent_ic_ensemble(
        ctypes.c_int(pid), ctypes.c_int(1 if jac == "analytic" else 0), ctypes.c_int64(N), _p(params),
        ctypes.c_double(float(t0)), _p(y0), _p(yp0), _p(fy, ctypes.c_int32), _p(fyp, ctypes.c_int32),
        ctypes.c_double(float(rtol)), _p(atol_v), ctypes.c_int(newton_maxiter), ctypes.c_int(chord_iter),
        ctypes.c_double(float(safety)), _p(f), _p(status, ctypes.c_int32), _p(n_iter, ctypes.c_int32))
    if rc != 0:
        raise RuntimeError(f"orc_consistent_ic_ensemble returned {rc}")
    return dict(y0=y0, yp0=yp0, f=f, status=status, n_iter=n_iter)


def solve_trace(problem, params, y0, yp0, t_span, method="Radau", rtol=1e-3, atol=1e-6, cap=1 << 14, **options):
    """ONE instance with its accepted-step trace (what solve_dae returns as t / y when t_eval is
    None, dae.py:492-495): dict(t_steps[k], y_steps[n, k], yp_steps[n, k], status, nfev, njev, nlu,
    n_accepted).  yp at the step ends comes from a second pass with t_eval = step ends."""
    pid, n, npar = problem_id(problem)
    L = lib()
    fn = ctypes.c_void_p(ctypes.cast(L.orc_problem_residual, ctypes.c_void_p).value)
    L.orc_problem_residual.restype = ctypes.c_void_p
    fnp = L.orc_problem_residual(ctypes.c_int(pid))
    p = np.ascontiguousarray(params, dtype=float).reshape(npar)
    y0 = np.ascontiguousarray(y0, dtype=float).reshape(n)
    yp0 = np.ascontiguousarray(yp0, dtype=float).reshape(n)
    atol_v = np.ascontiguousarray(np.broadcast_to(np.asarray(atol, dtype=float), (n,)))
    t0, t1 = map(float, t_span)
    t_steps = np.empty(cap); y_steps = np.empty((cap, n)); nsteps = ctypes.c_int64(0)
    y_final = np.empty(n); yp_final = np.empty(n); t_final = ctypes.c_double(0.0)
    counters = np.zeros(8, dtype=np.int64); nev = ctypes.c_int64(0)
    if method == "Radau":
        s = int(options.pop("stages", 3))
        tb = radau_tables_c(s)
        db = options.pop("controller_deadband", (1.0, 1.2))
        ro = RadauOpts(rtol=rtol, max_step=float(options.pop("max_step", np.inf)),
                       first_step=float(options.pop("first_step", None) or 0.0),
                       continuous_error_weight=float(options.pop("continuous_error_weight", 0.0)),
                       jac_recompute_rate=float(options.pop("jac_recompute_rate", 1e-3)),
                       deadband_lo=db[0], deadband_hi=db[1], newton_max_iter=int(options.pop("newton_max_iter", None) or 0),
                       newton_iter_embedded=int(options.pop("newton_iter_embedded", 1)), max_steps=0)
        L.orc_solve_radau.restype = ctypes.c_int
        st = L.orc_solve_radau(ctypes.c_void_p(fnp), ctypes.c_int(n), _p(p), ctypes.c_double(t0), ctypes.c_double(t1),
                               _p(y0), _p(yp0), _p(atol_v), ctypes.byref(ro), ctypes.byref(tb),
                               None, ctypes.c_int64(0), None, None, ctypes.byref(nev),
                               _p(y_final), _p(yp_final), ctypes.byref(t_final), _p(counters, ctypes.c_int64),
                               _p(t_steps), _p(y_steps), ctypes.c_int64(cap), ctypes.byref(nsteps))
        opts2 = dict(stages=s)
    else:
        bo = BdfOpts(rtol=rtol, max_step=float(options.pop("max_step", np.inf)),
                     first_step=float(options.pop("first_step", None) or 0.0), max_order=int(options.pop("max_order", 5)),
                     ndf_strategy=NDF_STRATEGIES[options.pop("NDF_strategy", "stability")], max_steps=0)
        L.orc_solve_bdf.restype = ctypes.c_int
        st = L.orc_solve_bdf(ctypes.c_void_p(fnp), ctypes.c_int(n), _p(p), ctypes.c_double(t0), ctypes.c_double(t1),
                             _p(y0), _p(yp0), _p(atol_v), ctypes.byref(bo),
                             None, ctypes.c_int64(0), None, None, ctypes.byref(nev),
                             _p(y_final), _p(yp_final), ctypes.byref(t_final), _p(counters, ctypes.c_int64),
                             _p(t_steps), _p(y_steps), ctypes.c_int64(cap), ctypes.byref(nsteps))
        opts2 = {}
    k = int(nsteps.value)
    ts = t_steps[:k].copy()
    yps = np.empty((n, k)); yps[:, 0] = yp0
    if k > 1:
        r2 = solve_ensemble(problem, p[None, :], y0[None, :], yp0[None, :], t_span, method=method, rtol=rtol,
                            atol=atol, t_eval=ts[1:], **opts2)
        nv = int(r2["n_eval_valid"][0])
        yps[:, 1:1 + nv] = r2["yp_eval"][0, :nv].T
    return dict(t_steps=ts, y_steps=y_steps[:k].T.copy(), yp_steps=yps, status=int(st), nfev=int(counters[0]),
                njev=int(counters[1]), nlu=int(counters[2]), n_accepted=int(counters[4]))
