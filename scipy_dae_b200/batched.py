"""Batched entry point: integrate N independent systems F(t, y, y'; p_i) = 0 at once on a B200.

``solve_dae_batched`` keeps the keyword surface of the reference's
``solve_dae(fun, t_span, y0, yp0, method, t_eval, **options)`` (dae.py:115-547) and of its two
solver classes (RadauDAE radau.py:393-399, BDFDAE bdf.py:224-227), with

* ``fun`` replaced by a problem registered in the CUDA library as a device functor
  (``scipy_dae_b200.problems``),
* a leading instance axis on ``y0``, ``yp0`` and the new per-instance ``params``,
* results returned with a leading instance axis (``BatchedDaeResult``).

Host code only validates arguments, lays the data out structure-of-arrays on the device and
calls the C-ABI (``include/b200dae.h``); the whole step loop runs in one kernel launch.
There is no CPU fallback.
"""
from __future__ import annotations

import ctypes
import warnings
from dataclasses import dataclass, field
from typing import Optional

import numpy as np
import torch

from . import _lib
from .constants import NDF_STRATEGIES, radau_tables_c
from .problems import get_problem

METHODS = ("Radau", "BDF")
MESSAGES = {0: "The solver successfully reached the end of the integration interval.",
            1: "A termination event occurred.",
            -1: "Required step size is less than spacing between numbers.",
            -2: "array must not contain infs or NaNs",
            -3: "max_steps accepted steps were taken before reaching t_bound."}

_RADAU_KEYS = ("stages", "max_step", "continuous_error_weight", "first_step", "newton_max_iter",
               "jac_recompute_rate", "newton_iter_embedded", "controller_deadband")
_BDF_KEYS = ("max_step", "first_step", "max_order", "NDF_strategy")
# accepted for signature compatibility: dense per-instance LUs only, no Python callables.
# `jac` is None (finite differences, the reference's jac=None) or "analytic" / True: the analytic
# Jacobians registered with the device functor, i.e. what the reference does with a `jac`
# callable (base.py:299-331).
_IGNORED_NONE_KEYS = ("jac_sparsity", "vectorized")


@dataclass(frozen=True)
class ThresholdEvent:
    """A device event  g(t, y, y') = x[component] - value,  x = y (``of="y"``) or y' (``of="yp"``),
    with the attributes of a ``solve_dae`` event function (dae.py:226-247): ``direction`` (> 0:
    only - to +, < 0: only + to -, 0: both) and ``terminal`` (False, True = stop at the first
    occurrence, or an int k = stop at the k-th)."""
    component: int
    value: float = 0.0
    of: str = "y"
    direction: float = 0.0
    terminal: object = False


@dataclass(frozen=True)
class FunctorEvent:
    """Event function number ``index`` of the problem's device functor, ``g(t, y, y'; p) - value``
    (``Problem.events[index]`` is its NumPy twin, the callable the reference would be given), with
    the ``direction`` / ``terminal`` attributes of a ``solve_dae`` event function."""
    index: int
    value: float = 0.0
    direction: float = 0.0
    terminal: object = False


def _events_struct(events, n, n_functor_events=0):
    ev = _lib.Events()
    if len(events) > _lib.MAX_EVENTS:
        raise ValueError(f"at most {_lib.MAX_EVENTS} device events are supported")
    ev.spec.n_events = len(events)
    for k, e in enumerate(events):
        if isinstance(e, dict):
            e = (FunctorEvent(index=e["functor"], **{a: b for a, b in e.items() if a != "functor"})
                 if "functor" in e else ThresholdEvent(**e))
        if isinstance(e, FunctorEvent):
            if not 0 <= int(e.index) < n_functor_events:
                raise ValueError(f"the problem's functor has {n_functor_events} event functions; got index {e.index}")
            ev.spec.of_yp[k] = 2
            ev.spec.component[k] = int(e.index)
        elif isinstance(e, ThresholdEvent):
            if e.of not in ("y", "yp"):
                raise ValueError("ThresholdEvent.of must be 'y' or 'yp'")
            if not 0 <= int(e.component) < n:
                raise ValueError("ThresholdEvent.component out of range")
            ev.spec.of_yp[k] = 1 if e.of == "yp" else 0
            ev.spec.component[k] = int(e.component)
        else:
            raise TypeError("device events are ThresholdEvent / FunctorEvent instances (or dicts with the same "
                            "fields): a Python event callable cannot run inside the device step loop")
        ev.spec.value[k] = float(e.value)
        ev.spec.direction[k] = int(np.sign(e.direction))
        term = e.terminal
        ev.spec.max_events[k] = 0 if term in (False, None) else (1 if term is True else int(term))
    return ev


@dataclass
class BatchedDaeResult:
    """Fields of the reference's DaeResult (dae.py:545-547) with a leading instance axis.

    ``y``/``yp`` hold the ``t_eval`` states as ``[N, n, m]`` (NaN where an instance failed before
    reaching the point, cf. ``n_eval_valid``); without ``t_eval`` they are ``None`` and the final
    state is in ``y_final``/``yp_final`` ``[N, n]``.  With ``record_steps=K`` the accepted step
    ends are in ``t_steps [N, K]`` / ``y_steps [N, n, K]`` (first ``n_steps[i]`` valid).
    """
    t_final: object
    y_final: object
    yp_final: object
    status: object
    nfev: object
    njev: object
    nlu: object
    nsolve: object
    n_accepted: object
    n_err_rejected: object
    n_newton_fail: object
    n_newton_iter: object
    t: object = None
    y: object = None
    yp: object = None
    n_eval_valid: object = None
    t_steps: object = None
    y_steps: object = None
    yp_steps: object = None
    n_steps: object = None
    h_steps: object = None         # [N, K]  h of step k's interpolant (record_dense)
    t_anchor_steps: object = None  # [N, K]  step end the interpolant is anchored at
    order_steps: object = None     # [N, K]  BDF order of the interpolant (Radau: stages)
    dense_coef: object = None      # [N, K, rows]  Radau ZP[n, s] / BDF D[7, n] flattened (include/b200dae.h)
    t_events: object = None        # [N, n_events, cap] (NaN beyond n_events_found)
    y_events: object = None        # [N, n_events, cap, n]
    yp_events: object = None
    n_events_found: object = None  # [N, n_events]
    method: str = ""
    problem: str = ""
    extra: dict = field(default_factory=dict)

    @property
    def success(self):
        return self.status >= 0

    @property
    def N(self):
        return int(self.status.shape[0])

    def message(self, i: int) -> str:
        return MESSAGES.get(int(self.status[i]), "")

    def solution(self, i: int = 0):
        """The continuous solution of instance ``i`` (``solve_dae(dense_output=True).sol``,
        common.py:9-126) from the recorded per-step interpolants; needs ``record_dense=True``."""
        from .dense import solution_from_trace
        return solution_from_trace(self, i)

    def numpy(self) -> "BatchedDaeResult":
        """Copy every tensor to host NumPy arrays (one device->host transfer per field)."""
        out = {}
        for k, v in self.__dict__.items():
            out[k] = v.detach().cpu().numpy() if isinstance(v, torch.Tensor) else v
        return BatchedDaeResult(**out)


def _validate_tol(rtol, atol, n):
    """scipy validate_tol (scipy/integrate/_ivp/common.py:44-60) + Radau's scalar rtol."""
    rtol = float(rtol)
    eps = float(np.finfo(float).eps)
    if rtol < 100 * eps:
        warnings.warn("At least one element of `rtol` is too small. "
                      f"Setting `rtol = np.maximum(rtol, {100 * eps})`.", stacklevel=3)
        rtol = 100 * eps
    atol = np.asarray(atol, dtype=float)
    if atol.ndim > 0 and atol.shape != (n,):
        raise ValueError("`atol` has wrong shape.")
    if np.any(atol < 0):
        raise ValueError("`atol` must be positive.")
    return rtol, np.ascontiguousarray(np.broadcast_to(atol, (n,)), dtype=float)


def _validate_t_eval(t_eval, t0, t_bound):
    """dae.py:384-403 (the device wants t_eval sorted in the direction of integration)."""
    t_eval = np.asarray(t_eval, dtype=float)
    if t_eval.ndim != 1:
        raise ValueError("`t_eval` must be 1-dimensional.")
    if np.any(t_eval < min(t0, t_bound)) or np.any(t_eval > max(t0, t_bound)):
        raise ValueError("Values in `t_eval` are not within `t_span`.")
    d = np.diff(t_eval)
    if t_bound > t0 and np.any(d <= 0) or t_bound < t0 and np.any(d >= 0):
        raise ValueError("Values in `t_eval` are not properly sorted.")
    return np.array(t_eval, dtype=float, copy=True)     # owns its data (reversed views have negative strides)


def _as_device_soa(x, N, width, name, device, check_finite=True):
    """[N, width] (or [width], broadcast) host/device array -> contiguous float64 [width, N] on device."""
    if isinstance(x, torch.Tensor):
        t = x.to(dtype=torch.float64)
    else:
        a = np.asarray(x)
        if np.iscomplexobj(a):
            raise ValueError("`y0` or `yp0` is complex, but the chosen solver does not support "
                             "integration in a complex domain.")
        t = torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64))
    if t.ndim == 1:
        t = t.unsqueeze(0).expand(N, -1)
    if t.ndim != 2 or t.shape[1] != width or t.shape[0] != N:
        raise ValueError(f"`{name}` must have shape ({N}, {width}) or ({width},), got {tuple(t.shape)}.")
    t = t.to(device, non_blocking=True)
    soa = t.t().contiguous()
    if check_finite and not bool(torch.isfinite(soa).all()):
        raise ValueError(f"All components of the initial state `{name}` must be finite.")
    return soa


def _warn_extraneous(extraneous):
    if extraneous:
        warnings.warn("The following arguments have no effect for a chosen solver: "
                      f"{', '.join(f'`{x}`' for x in extraneous)}.", stacklevel=3)


def _host_device_ptr(L, t):
    """Device address of a page-locked host tensor (cudaHostGetDevicePointer through the C-ABI)."""
    dp = ctypes.c_void_p()
    _lib.check(L.b200dae_host_device_pointer(ctypes.c_void_p(t.data_ptr()), ctypes.byref(dp)))
    return dp.value


def solve_dae_batched(problem, t_span, y0=None, yp0=None, params=None, method="Radau",
                      t_eval=None, rtol=1e-3, atol=1e-6, *, device=None, stream=None,
                      record_steps: int = 0, record_dense: bool = False, max_steps: int = 0,
                      check_finite: bool = True, eval_layout: str = "soa",
                      events=None, event_cap: int = 16, **options) -> BatchedDaeResult:
    """Integrate ``N`` instances of a registered problem from ``t_span[0]`` to ``t_span[1]``.

    Parameters mirror ``solve_dae`` (dae.py:115-353); differences:

    problem : str or Problem   name of a device functor compiled into libb200dae.so
    y0, yp0 : [N, n] or [n]    initial state; ``None`` -> the problem's consistent initial values
    params  : [N, p] or [p]    per-instance parameters (the reference's ``args``)
    device  : torch device     default: current CUDA device
    record_steps : int         > 0: also return up to that many accepted steps per instance
    record_dense : bool        with record_steps: also return every recorded step's dense-output
                               coefficients (the reference's per-step RadauDenseOutput /
                               BdfDenseOutput objects); ``result.solution(i)`` is then instance i's
                               ``DaeSolution`` without any re-integration
    max_steps : int            > 0: stop every instance after that many accepted steps (status -3)
    eval_layout : str          device layout of the ``t_eval`` states behind ``result.y`` / ``result.yp``
                               (both are ``[N, n, m]`` views either way): ``"soa"`` = ``[m][n][N]`` (default:
                               the lanes of a warp write a row together, measured 34.5 ms against 55.0 ms
                               for 2^20 Robertson instances at m = 1000), ``"instance"`` = ``[N][m][n]``
                               (every instance's ``(m, n)`` block contiguous, e.g. for per-instance D2H)
    events : list of ThresholdEvent / FunctorEvent   located on the dense output with Brent's method as in the
                               reference (dae.py:30-109, 466-490); a terminal event ends the instance
                               with status 1; up to ``event_cap`` occurrences per event are stored

    Returns a ``BatchedDaeResult`` of device tensors (call ``.numpy()`` for host arrays).
    """
    if eval_layout not in ("instance", "soa"):
        raise ValueError("eval_layout must be 'instance' or 'soa'")
    host_io = options.pop("_host_io", None)      # host_pipeline.py: zero-copy launch on pinned host arrays
    L = _lib.load()
    prob = get_problem(problem)
    pid, n, npar = _lib.problem_lookup(prob.name)
    if method not in METHODS:
        raise ValueError(f"`method` must be one of {METHODS}.")
    if not torch.cuda.is_available():
        raise _lib.B200DaeError("scipy_dae_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
    device = torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}")
    t0, t_bound = map(float, t_span)

    # ---- per-instance inputs ---------------------------------------------------------------
    if params is None:
        params = np.asarray(prob.default_params, dtype=float)
    p_shape = tuple(params.shape)
    N = None
    for arr in (params, y0, yp0):
        if arr is not None and len(arr.shape if hasattr(arr, "shape") else np.shape(arr)) == 2:
            N = int(arr.shape[0] if hasattr(arr, "shape") else np.shape(arr)[0])
            break
    if N is None:
        N = 1
    if (y0 is None) != (yp0 is None):
        raise ValueError("`y0` and `yp0` must be of same shape.")
    if y0 is None:
        from .problems import initial_conditions
        p_host = params.detach().cpu().numpy() if isinstance(params, torch.Tensor) else np.asarray(params, dtype=float)
        p_host = np.broadcast_to(np.atleast_2d(p_host), (N, npar)) if len(p_shape) == 1 else p_host
        y0, yp0 = initial_conditions(prob, p_host)
    else:
        s0 = tuple(y0.shape) if hasattr(y0, "shape") else np.shape(y0)
        s1 = tuple(yp0.shape) if hasattr(yp0, "shape") else np.shape(yp0)
        if s0 != s1:
            raise ValueError("`y0` and `yp0` must be of same shape.")
    rtol, atol_v = _validate_tol(rtol, atol, n)

    # every piece of device work of this call (copies, relayouts, output fills, the launch, the
    # status post-processing) is queued on ONE stream: `stream` if given, after it has waited for
    # the caller's current stream, else the current stream
    import contextlib
    if stream is not None:
        if stream != torch.cuda.current_stream(device):
            stream.wait_stream(torch.cuda.current_stream(device))
        stream_ctx = torch.cuda.stream(stream)
    else:
        stream_ctx = contextlib.nullcontext()
    with torch.cuda.device(device), stream_ctx:
        ld = N
        f64 = dict(dtype=torch.float64, device=device)
        i32 = dict(dtype=torch.int32, device=device)
        out = _lib.Outputs()
        if host_io is None:
            y0_d = _as_device_soa(y0, N, n, "y0", device, check_finite)
            yp0_d = _as_device_soa(yp0, N, n, "yp0", device, check_finite)
            p_d = _as_device_soa(params, N, npar, "params", device, False) if npar > 0 else None
            y0_ptr, yp0_ptr = y0_d.data_ptr(), yp0_d.data_ptr()
            p_ptr = p_d.data_ptr() if p_d is not None else None
            # ---- outputs -------------------------------------------------------------------
            y_final = torch.empty((n, ld), **f64)
            yp_final = torch.empty((n, ld), **f64)
            t_final = torch.empty((ld,), **f64)
            status = torch.empty((ld,), **i32)
            counters = torch.empty((_lib.NUM_COUNTERS, ld), **i32)
            out.y_final, out.yp_final, out.t_final = y_final.data_ptr(), yp_final.data_ptr(), t_final.data_ptr()
            out.status, out.counters = status.data_ptr(), counters.data_ptr()
        else:
            # in place on page-locked HOST arrays of the reference's own shapes (io_layout = 1): the
            # launch reads (N, n) / (N, p) inputs and writes (N, n) / (N, 8) results over PCIe
            if t_eval is not None or record_steps or events:
                raise NotImplementedError("t_eval / record_steps / events are not available in place on host arrays")
            hb = host_io["out"]
            y0_d, yp0_d, p_d = host_io["y0"], host_io["yp0"], host_io["params"]
            y0_ptr, yp0_ptr = _host_device_ptr(L, y0_d), _host_device_ptr(L, yp0_d)
            p_ptr = _host_device_ptr(L, p_d) if npar > 0 else None
            y_final, yp_final, t_final, status, counters = hb.y_final, hb.yp_final, hb.t_final, hb.status, hb.counters
            out.y_final, out.yp_final = _host_device_ptr(L, y_final), _host_device_ptr(L, yp_final)
            out.t_final, out.status = _host_device_ptr(L, t_final), _host_device_ptr(L, status)
            out.counters = _host_device_ptr(L, counters)
            if host_io.get("flag") is not None:
                out.bad_input = _host_device_ptr(L, host_io["flag"])
        te_d = y_eval = yp_eval = n_eval_valid = None
        if t_eval is not None:
            te = _validate_t_eval(t_eval, t0, t_bound)
            m = te.shape[0]
            te_d = torch.from_numpy(te).to(device)
            # "soa" = [m][n][ld] (default); "instance" = [ld][m][n] through the strides of b200dae_outputs
            if eval_layout == "instance":
                y_eval = torch.full((ld, m, n), float("nan"), **f64)
                yp_eval = torch.full((ld, m, n), float("nan"), **f64)
                out.eval_stride_row, out.eval_stride_comp, out.eval_stride_inst = n, 1, m * n
            else:
                y_eval = torch.full((m, n, ld), float("nan"), **f64)
                yp_eval = torch.full((m, n, ld), float("nan"), **f64)
            n_eval_valid = torch.empty((ld,), **i32)
            out.t_eval, out.m = te_d.data_ptr(), m
            out.y_eval, out.yp_eval = y_eval.data_ptr(), yp_eval.data_ptr()
            out.n_eval_valid = n_eval_valid.data_ptr()
        t_steps = y_steps = yp_steps = n_steps = None
        if record_steps and record_steps > 0:
            K = int(record_steps)
            t_steps = torch.full((K, ld), float("nan"), **f64)
            y_steps = torch.full((K, n, ld), float("nan"), **f64)
            yp_steps = torch.full((K, n, ld), float("nan"), **f64)
            n_steps = torch.empty((ld,), **i32)
            out.steps_cap = K
            out.t_steps, out.y_steps, out.yp_steps = t_steps.data_ptr(), y_steps.data_ptr(), yp_steps.data_ptr()
            out.n_steps = n_steps.data_ptr()
        h_steps = order_steps = dense_coef = None
        if record_dense:
            if t_steps is None:
                raise ValueError("`record_dense` needs `record_steps` > 0.")
            rows = n * int(options.get("stages", 3)) if method == "Radau" else 7 * n
            h_steps = torch.full((K, 2, ld), float("nan"), **f64)
            order_steps = torch.zeros((K, ld), **i32)
            dense_coef = torch.full((K, rows, ld), float("nan"), **f64)
            out.h_steps, out.order_steps = h_steps.data_ptr(), order_steps.data_ptr()
            out.dense_coef, out.coef_rows = dense_coef.data_ptr(), rows

        ev_struct = t_events = y_events = yp_events = n_found = None
        if events:
            ev_struct = _events_struct(list(events), n, L.b200dae_problem_num_events(pid))
            ne, cap = ev_struct.spec.n_events, int(event_cap)
            t_events = torch.full((ne, cap, ld), float("nan"), **f64)
            y_events = torch.full((ne, cap, n, ld), float("nan"), **f64)
            yp_events = torch.full((ne, cap, n, ld), float("nan"), **f64)
            n_found = torch.zeros((ne, ld), **i32)
            ev_struct.cap = cap
            ev_struct.t_events, ev_struct.y_events = t_events.data_ptr(), y_events.data_ptr()
            ev_struct.yp_events, ev_struct.n_found = yp_events.data_ptr(), n_found.data_ptr()
            out.events = ctypes.addressof(ev_struct)

        # ---- options -----------------------------------------------------------------------
        opts = dict(options)
        jac = opts.pop("jac", None)
        if jac is True or jac == "analytic":
            analytic_jac = 1
        elif jac is None or jac is False:
            analytic_jac = 0
        else:
            raise NotImplementedError(
                "`jac` must be None (finite differences) or 'analytic' (the Jacobians registered with "
                "the device functor): a Python callable cannot run inside the device step loop.")
        for k in _IGNORED_NONE_KEYS:
            v = opts.pop(k, None)
            if v not in (None, False):
                raise NotImplementedError(
                    f"`{k}` is not supported on the device path (dense per-instance Jacobians).")
        max_step = float(opts.pop("max_step", np.inf))
        if max_step <= 0:
            raise ValueError("`max_step` must be positive.")
        first_step = opts.pop("first_step", None)
        if first_step is not None:
            first_step = float(first_step)
            if first_step <= 0:
                raise ValueError("`first_step` must be positive.")
            if first_step > abs(t_bound - t0):
                raise ValueError("`first_step` exceeds bounds.")
        cu_stream = ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)
        atol_p = atol_v.ctypes.data_as(ctypes.POINTER(ctypes.c_double))
        io_layout = 0 if host_io is None else 1
        if method == "Radau":
            stages = int(opts.pop("stages", 3))
            if stages % 2 != 1:
                raise ValueError("`stages` must be odd.")
            db = opts.pop("controller_deadband", (1.0, 1.2))
            nmi = opts.pop("newton_max_iter", None)
            if nmi is not None and not 1 <= int(nmi) <= 32:
                # the reference asserts newton_max_iter >= 1 (radau.py:438-441); the device evaluates
                # rate ** (newton_max_iter - k) with a 5-bit square-and-multiply, hence the upper bound
                raise ValueError("`newton_max_iter` must be in 1 .. 32 on the device path.")
            ro = _lib.RadauOpts(
                rtol=rtol, max_step=max_step, first_step=first_step or 0.0,
                continuous_error_weight=float(opts.pop("continuous_error_weight", 0.0)),
                jac_recompute_rate=float(opts.pop("jac_recompute_rate", 1e-3)),
                deadband_lo=float(db[0]), deadband_hi=float(db[1]),
                newton_max_iter=int(nmi) if nmi is not None else 0,
                newton_iter_embedded=int(opts.pop("newton_iter_embedded", 1)),
                max_steps=int(max_steps), analytic_jac=analytic_jac, io_layout=io_layout)
            _warn_extraneous(opts)
            tb = radau_tables_c(stages)
            rc = L.b200dae_solve_radau(pid, N, ld, y0_ptr, yp0_ptr, p_ptr,
                                       t0, t_bound, atol_p, ctypes.byref(ro), ctypes.byref(tb),
                                       ctypes.byref(out), cu_stream)
        else:
            ndf = opts.pop("NDF_strategy", "stability")
            bo = _lib.BdfOpts(rtol=rtol, max_step=max_step, first_step=first_step or 0.0,
                              max_order=int(opts.pop("max_order", 5)),
                              ndf_strategy=NDF_STRATEGIES.get(ndf, 2), max_steps=int(max_steps),
                              analytic_jac=analytic_jac, io_layout=io_layout)
            if bo.max_order == 6:
                warnings.warn("Choosing `max_order = 6` is not recomended due to its poor "
                              "stability properties.", stacklevel=2)
            _warn_extraneous(opts)
            rc = L.b200dae_solve_bdf(pid, N, ld, y0_ptr, yp0_ptr, p_ptr,
                                     t0, t_bound, atol_p, ctypes.byref(bo), ctypes.byref(out),
                                     cu_stream)
        _lib.check(rc)
        if ev_struct is not None:  # device code 2 -> solve_dae's status 1 ("A termination event occurred.")
            status = torch.where(status == _lib.ST_EVENT, torch.ones_like(status), status)
    if host_io is not None:
        cn = counters
        return BatchedDaeResult(
            t_final=t_final, y_final=y_final, yp_final=yp_final, status=status,
            nfev=cn[:, 0], njev=cn[:, 1], nlu=cn[:, 2], nsolve=cn[:, 3], n_accepted=cn[:, 4],
            n_err_rejected=cn[:, 5], n_newton_fail=cn[:, 6], n_newton_iter=cn[:, 7],
            method=method, problem=prob.name, extra={"_keepalive": (y0_d, yp0_d, p_d)})
    res = BatchedDaeResult(
        t_final=t_final, y_final=y_final.t(), yp_final=yp_final.t(), status=status,
        nfev=counters[0], njev=counters[1], nlu=counters[2], nsolve=counters[3],
        n_accepted=counters[4], n_err_rejected=counters[5], n_newton_fail=counters[6],
        n_newton_iter=counters[7], method=method, problem=prob.name)
    if t_eval is not None:
        res.t = te_d
        if eval_layout == "instance":
            res.y = y_eval.permute(0, 2, 1)      # [N, n, m]
            res.yp = yp_eval.permute(0, 2, 1)
        else:
            res.y = y_eval.permute(2, 1, 0)      # [N, n, m]
            res.yp = yp_eval.permute(2, 1, 0)
        res.n_eval_valid = n_eval_valid
    if t_steps is not None:
        res.t_steps = t_steps.t()            # [N, K]
        res.y_steps = y_steps.permute(2, 1, 0)
        res.yp_steps = yp_steps.permute(2, 1, 0)
        res.n_steps = n_steps
    if dense_coef is not None:
        res.h_steps = h_steps[:, 0].t()              # [N, K]
        res.t_anchor_steps = h_steps[:, 1].t()
        res.order_steps = order_steps.t()
        res.dense_coef = dense_coef.permute(2, 0, 1)  # [N, K, rows]
        res.extra["stages"] = int(options.get("stages", 3)) if method == "Radau" else 0
    if ev_struct is not None:
        res.t_events = t_events.permute(2, 0, 1)            # [N, ne, cap]
        res.y_events = y_events.permute(3, 0, 1, 2)         # [N, ne, cap, n]
        res.yp_events = yp_events.permute(3, 0, 1, 2)
        res.n_events_found = n_found.t()
    # keep the inputs alive until the caller synchronises
    res.extra["_keepalive"] = (y0_d, yp0_d, p_d)
    return res


def device_residual(problem, t, y, yp, params, device=None):
    """Evaluate the registered device functor for N instances: returns [N, n] (torch, device)."""
    L = _lib.load()
    prob = get_problem(problem)
    pid, n, npar = _lib.problem_lookup(prob.name)
    device = torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}")
    y = np.atleast_2d(np.asarray(y, dtype=float))
    N = y.shape[0]
    with torch.cuda.device(device):
        y_d = _as_device_soa(y, N, n, "y", device, False)
        yp_d = _as_device_soa(yp, N, n, "yp", device, False)
        p_d = _as_device_soa(params, N, npar, "params", device, False)
        out = torch.empty((n, N), dtype=torch.float64, device=device)
        rc = L.b200dae_residual(pid, N, N, float(t), y_d.data_ptr(), yp_d.data_ptr(), p_d.data_ptr(),
                                out.data_ptr(), ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream))
        _lib.check(rc)
    return out.t()
