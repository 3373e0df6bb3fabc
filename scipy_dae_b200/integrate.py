"""Single-system surface: ``solve_dae(fun, t_span, y0, yp0, method, t_eval, args, **options)``
with the reference's signature and result fields (dae.py:115-547), executed on the B200.

``fun`` must be a residual registered in the CUDA library (a name or a
``scipy_dae_b200.problems.Problem``): the step loop runs entirely on the device, so a Python
callable cannot be evaluated there.  ``args`` are the problem's parameters (the reference's
``fun(t, y, yp, *args)`` convention, dae.py:365-383).
"""
from __future__ import annotations

import numpy as np

from .batched import MESSAGES, METHODS, solve_dae_batched
from .dense import DaeSolution  # noqa: F401  (re-exported: the type of DaeResult.sol)
from .problems import Problem, get_problem


class DaeResult(dict):
    """Same fields as the reference's DaeResult (an OptimizeResult, dae.py:26-27,545-547)."""
    __getattr__ = dict.get
    __setattr__ = dict.__setitem__

    def __repr__(self):
        return "\n".join(f"{k:>10}: {v!r}" for k, v in self.items())


def solve_dae(fun, t_span, y0=None, yp0=None, method="Radau", t_eval=None, dense_output=False,
              events=None, vectorized=False, args=None, var_index=None, device=None,
              max_steps_recorded: int = 1 << 12, **options) -> DaeResult:
    """Integrate one registered system; returns ``t, y[n, m], yp[n, m], nfev, njev, nlu, status,
    message, success`` exactly shaped like the reference's result.

    Without ``t_eval`` the accepted step ends (including ``t0``) are returned (dae.py:492-495);
    with ``t_eval`` the dense output at those points (dae.py:496-515).  A failed integration is
    not an exception: ``status = -1`` and the points reached so far are returned (dae.py:451-453).
    """
    if var_index is not None:
        raise NotImplementedError("This feature is not implemented yet.")          # dae.py:356-357
    if isinstance(method, str) and method not in METHODS:
        raise ValueError(f"`method` must be one of {METHODS} or DaeSolver class.")  # dae.py:359-361
    if not isinstance(fun, (str, Problem)):
        raise TypeError("scipy_dae_b200.solve_dae integrates residuals that are registered as CUDA "
                        "device functors: pass a problem name or Problem (see scipy_dae_b200.problems); "
                        "a Python callable cannot run inside the device step loop.")
    if events is not None and not isinstance(events, (list, tuple)):
        events = [events]                                                           # dae.py:427 prepare_events
    prob = get_problem(fun)
    if args is not None:
        try:
            params = np.asarray([*args], dtype=float)
        except TypeError as exp:
            raise TypeError("Supplied 'args' cannot be unpacked. Please supply `args` as a tuple "
                            f"(e.g. `args=({args},)`)") from exp
    else:
        params = np.asarray(prob.default_params, dtype=float)
    for name, v in (("y0", y0), ("yp0", yp0)):
        if v is not None and np.asarray(v).ndim != 1:
            raise ValueError(f"`{name}` must be 1-dimensional.")                    # base.py:31-34
    if y0 is not None and yp0 is not None and np.shape(y0) != np.shape(yp0):
        raise ValueError("`y0` and `yp0` must be of same shape.")                   # base.py:36-37
    y0b = None if y0 is None else np.asarray(y0, dtype=float)[None, :]
    yp0b = None if yp0 is None else np.asarray(yp0, dtype=float)[None, :]

    cap = 0 if (t_eval is not None and not dense_output) else int(max_steps_recorded)
    while True:
        r = solve_dae_batched(prob, t_span, y0b, yp0b, params[None, :], method=method, t_eval=t_eval,
                              device=device, record_steps=cap, record_dense=bool(dense_output and cap),
                              events=events, **options).numpy()
        if cap == 0 or r.n_steps[0] < cap or r.n_accepted[0] + 1 <= cap:
            break
        cap *= 4       # the step trace overflowed: rerun with a larger trace

    status = int(r.status[0])
    if t_eval is None:
        k = int(r.n_steps[0])
        t, y, yp = r.t_steps[0, :k], r.y_steps[0, :, :k], r.yp_steps[0, :, :k]
    else:
        k = int(r.n_eval_valid[0])
        t, y, yp = r.t[:k], r.y[0, :, :k], r.yp[0, :, :k]
        if k == 0:
            t, y, yp = [], [], []                                                   # dae.py:523-535
    t_ev = y_ev = yp_ev = None
    if events:                       # lists per event, like dae.py:537-540
        nf = np.minimum(r.n_events_found[0], r.t_events.shape[2])
        t_ev = [r.t_events[0, k, :nf[k]] for k in range(len(events))]
        y_ev = [r.y_events[0, k, :nf[k]] for k in range(len(events))]
        yp_ev = [r.yp_events[0, k, :nf[k]] for k in range(len(events))]
    sol = None
    if dense_output:
        # DaeSolution over the accepted steps (dae.py:520-526 with common.py:9-126), built from the
        # per-step coefficients the same launch recorded: no re-integration
        sol = r.solution(0) if int(r.n_steps[0]) > 1 else None
    return DaeResult(t=t, y=y, yp=yp, sol=sol, t_events=t_ev, y_events=y_ev, yp_events=yp_ev,
                     nfev=int(r.nfev[0]), njev=int(r.njev[0]), nlu=int(r.nlu[0]),
                     status=-1 if status < 0 else status,
                     message=MESSAGES.get(status, ""), success=status >= 0,
                     n_accepted=int(r.n_accepted[0]), n_err_rejected=int(r.n_err_rejected[0]),
                     n_newton_fail=int(r.n_newton_fail[0]))
