"""`DaeSolver` adapters: the B200 path behind the reference's own plug-in API.

The reference's extension point is ``solve_dae(..., method=<DaeSolver subclass>)``
(scipy_dae/integrate/_dae/dae.py:359-361, 405-408); its driver consumes exactly ``step()``,
``status``, ``t_old``, ``t``, ``y``, ``yp``, ``direction``, ``dense_output()``, ``nfev``, ``njev``,
``nlu`` (dae.py:446-518, 545-547).  ``make_dae_solver_adapters`` builds two subclasses of the
reference's ``DaeSolver`` (passed in, so this module never imports the reference) that satisfy
that protocol for ONE registered system:

    from scipy_dae.integrate import solve_dae
    from scipy_dae.integrate._dae.base import DaeSolver, DAEDenseOutput
    from scipy_dae_b200.adapter import make_dae_solver_adapters
    B200Radau, B200BDF = make_dae_solver_adapters(DaeSolver, DAEDenseOutput)
    sol = solve_dae(get_problem("robertson"), (0, 1e7), y0, yp0, method=B200Radau,
                    rtol=1e-6, atol=1e-6, params=(0.04, 1e4, 3e7), t_eval=...)

The device integrates the whole interval in one launch, so the adapter runs it on the first
``step()`` with the accepted steps AND their dense-output coefficients recorded
(``record_dense``) and then hands them to the driver one by one; a step's ``dense_output()`` is
that step's interpolant built from the recorded coefficients (scipy_dae_b200/dense.py): ONE launch
per ``solve_dae`` call, whatever ``t_eval`` / ``dense_output`` ask for.
``fun`` must be a registered ``Problem`` (it is callable like the reference's ``fun``) or the
keyword ``problem=`` must name one; parameters go in ``params=`` (with ``args=`` the reference
wraps ``fun`` in a lambda, dae.py:380-381, and the functor could not be identified).
"""
from __future__ import annotations

import numpy as np

from .dense import step_interpolants
from .problems import Problem, get_problem

MESSAGES = {-1: "Required step size is less than spacing between numbers.",
            -2: "array must not contain infs or NaNs",
            -3: "max_steps accepted steps were taken before reaching t_bound."}


def _device_solve(problem, t_span, y0, yp0, params, method, t_eval=None, record_steps=0, **options):
    from .batched import solve_dae_batched
    return solve_dae_batched(problem, t_span, y0[None, :], yp0[None, :], params[None, :], method=method,
                             t_eval=t_eval, record_steps=record_steps, record_dense=record_steps > 0,
                             **options).numpy()


def make_dae_solver_adapters(DaeSolver, DAEDenseOutput, solve=_device_solve):
    """Returns ``(B200RadauDAE, B200BDFDAE)``, subclasses of the given ``DaeSolver``.

    ``solve(problem, t_span, y0, yp0, params, method, t_eval=, record_steps=, **options)`` must
    return an object with the fields of ``BatchedDaeResult.numpy()`` for one instance; the default
    runs the CUDA path (tests on a machine without a GPU inject the CPU oracle here)."""

    class _TraceDenseOutput(DAEDenseOutput):
        """A step's interpolant from the device's coefficient trace (no launch)."""
        def __init__(self, t_old, t, interpolant):
            super().__init__(t_old, t)
            self._ip = interpolant

        def _call_impl(self, t):
            return self._ip(t)

    class _ReplayDenseOutput(DAEDenseOutput):
        """Fallback for solve hooks that return no coefficient trace (the CPU-oracle hook of the
        tests): evaluates by a deterministic replay with t_eval."""
        def __init__(self, t_old, t, owner):
            super().__init__(t_old, t)
            self._owner = owner

        def _call_impl(self, t):
            tt = np.atleast_1d(np.asarray(t, dtype=float))
            y, yp = self._owner._evaluate(tt)
            return (y[:, 0], yp[:, 0]) if np.ndim(t) == 0 else (y, yp)

    class _B200Solver(DaeSolver):
        _method = "Radau"
        TOO_SMALL_STEP = "Required step size is less than spacing between numbers."

        def __init__(self, fun, t0, y0, yp0, t_bound, vectorized=False, problem=None, params=None,
                     rtol=1e-3, atol=1e-6, **options):
            # DaeSolver.__init__ is not run: it would form an FD Jacobian with the Python callable
            # (base.py:297); the device does that itself.
            prob = problem if problem is not None else fun
            if not isinstance(prob, (str, Problem)):
                raise TypeError("pass the registered Problem as `fun` or name it with `problem=`; with `args=` "
                                "the reference wraps `fun` in a lambda, use `params=` instead")
            self._prob = get_problem(prob)
            self._params = np.asarray(self._prob.default_params if params is None else params, dtype=float)
            self._opts = dict(options, rtol=rtol, atol=atol)
            self.fun = self._prob
            self.t_old = None
            self.t = float(t0)
            self._t0 = float(t0)
            self.y = np.array(y0, dtype=float)
            self.yp = np.array(yp0, dtype=float)
            if self.y.ndim != 1 or self.y.shape != self.yp.shape:
                raise ValueError("`y0` and `yp0` must be 1-dimensional and of same shape.")
            self._y0, self._yp0 = self.y.copy(), self.yp.copy()
            self.t_bound = float(t_bound)
            self.n = self.y.size
            self.direction = np.sign(t_bound - t0) if t_bound != t0 else 1
            self.status = "running"
            self.nfev = self.njev = self.nlu = 0
            self._trace = None
            self._k = 0

        def _run(self, **kw):
            return solve(self._prob, (self._t0, self.t_bound), self._y0, self._yp0, self._params, self._method,
                         **kw, **self._opts)

        def _ensure_trace(self):
            if self._trace is not None:
                return
            cap = 1 << 12
            while True:
                r = self._run(record_steps=cap)
                if r.n_steps[0] < cap or r.n_accepted[0] + 1 <= cap:
                    break
                cap *= 4
            k = int(r.n_steps[0])
            self._trace = (r.t_steps[0, :k], r.y_steps[0, :, :k], r.yp_steps[0, :, :k], int(r.status[0]),
                           (int(r.nfev[0]), int(r.njev[0]), int(r.nlu[0])))
            self._interpolants = None
            if getattr(r, "dense_coef", None) is not None:
                self._interpolants = step_interpolants(
                    self._method, int(self._opts.get("stages", 3)), r.t_steps[0, :k], r.y_steps[0, :, :k],
                    r.h_steps[0, :k], r.t_anchor_steps[0, :k], r.order_steps[0, :k], r.dense_coef[0, :k])

        def _step_impl(self):
            self._ensure_trace()
            ts, ys, yps, status, counters = self._trace
            if self._k + 1 >= ts.shape[0]:                 # the device run stopped here (dae.py:451-453)
                self.nfev, self.njev, self.nlu = counters
                return False, MESSAGES.get(status, self.TOO_SMALL_STEP)
            self._k += 1
            self.t = float(ts[self._k])
            self.y = ys[:, self._k].copy()
            self.yp = yps[:, self._k].copy()
            if self._k + 1 == ts.shape[0]:                 # counters of the whole run become visible at the end
                self.nfev, self.njev, self.nlu = counters
            return True, None

        def _evaluate(self, times):
            """(y, yp) of the recorded piecewise interpolant at `times` (any order, within [t0, t])."""
            lo, hi = min(self._t0, self.t), max(self._t0, self.t)
            if np.any(times < lo) or np.any(times > hi):
                raise ValueError("dense output of the device path does not extrapolate")
            uniq, inv = np.unique(times, return_inverse=True)
            order = uniq if self.direction > 0 else uniq[::-1]
            r = self._run(t_eval=order)
            y, yp = r.y[0], r.yp[0]
            if self.direction < 0:
                y, yp = y[:, ::-1], yp[:, ::-1]
            return y[:, inv], yp[:, inv]

        def _dense_output_impl(self):
            if self._interpolants is not None:
                return _TraceDenseOutput(self.t_old, self.t, self._interpolants[self._k - 1])
            return _ReplayDenseOutput(self.t_old, self.t, self)

    class B200RadauDAE(_B200Solver):
        _method = "Radau"

    class B200BDFDAE(_B200Solver):
        _method = "BDF"

    return B200RadauDAE, B200BDFDAE
