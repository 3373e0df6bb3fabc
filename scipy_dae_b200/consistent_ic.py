"""Batched ``consistent_initial_conditions``: the reference's pre-processing step
(scipy_dae/integrate/_dae/common.py:186-258) for N independent instances in one kernel launch.

Same keyword surface as the reference (``jac, fixed_y0, fixed_yp0, rtol, atol, newton_maxiter,
chord_iter, safety``); ``fun`` is a residual registered as a device functor and ``args`` become
the per-instance ``params``.  Where the reference raises for a single system, the batch reports a
per-instance ``status`` (and ``raise_on_failure=True`` turns the first failure into the
reference's exception).
"""
from __future__ import annotations

import ctypes
from dataclasses import dataclass

import numpy as np
import torch

from . import _lib
from .batched import _as_device_soa
from .problems import get_problem

CIC_MESSAGES = {0: "converged", -1: "Convergence failed.",
                -2: "Too many fixed components, cannot solve the problem.",
                -3: "Index greater than one."}


@dataclass
class ConsistentICResult:
    y0: torch.Tensor        # [N, n]
    yp0: torch.Tensor       # [N, n]
    f: torch.Tensor         # [N, n] residual at the returned point
    status: torch.Tensor    # [N] int32: 0 ok, -1 / -2 / -3 see CIC_MESSAGES
    n_iter: torch.Tensor    # [N] int32 chord iterations taken

    @property
    def success(self):
        return self.status == 0


def consistent_initial_conditions_batched(problem, t0, y0, yp0, params=None, jac=None, fixed_y0=None,
                                          fixed_yp0=None, rtol=1e-8, atol=1e-8, newton_maxiter=10,
                                          chord_iter=3, safety=0.5, *, device=None, stream=None,
                                          raise_on_failure: bool = False) -> ConsistentICResult:
    """Consistent ``(y0, yp0)`` with ``F(t0, y0, yp0; p_i) = 0`` for every instance ``i``.

    y0, yp0 : [N, n] (or [n])   initial guesses; components listed in ``fixed_y0`` / ``fixed_yp0``
                                (index lists, shared by all instances) are kept
    jac     : None -> forward differences as the reference's default; "analytic" -> the
              Jacobians registered with the device functor
    """
    L = _lib.load()
    prob = get_problem(problem)
    pid, n, npar = _lib.problem_lookup(prob.name)
    if not torch.cuda.is_available():
        raise _lib.B200DaeError("scipy_dae_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
    device = torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}")
    if not (isinstance(rtol, float) and rtol > 0):
        raise ValueError("Relative tolerance must be a positive scalar.")           # common.py:227-228
    atol_v = np.ascontiguousarray(np.broadcast_to(np.asarray(atol, dtype=float), (n,)))
    if np.any(atol_v <= 0):
        raise ValueError("Absolute tolerance must be positive.")                    # common.py:234-235
    if not 0 < safety <= 1:
        raise AssertionError("safety factor has to be in (0, 1]")                   # common.py:237
    if jac not in (None, False, True, "analytic"):
        raise NotImplementedError("`jac` must be None (finite differences) or 'analytic'")
    masks = []
    for fixed in (fixed_y0, fixed_yp0):
        m = np.zeros(n, dtype=np.int32)
        if fixed is not None:
            m[np.asarray(fixed, dtype=int)] = 1
        masks.append(m)
    if (n - masks[0].sum()) + (n - masks[1].sum()) < n:
        raise ValueError("Too many components fixed, cannot solve the problem.")    # common.py:224-225
    if params is None:
        params = np.asarray(prob.default_params, dtype=float)
    N = 1
    for arr in (params, y0, yp0):
        shp = tuple(arr.shape) if hasattr(arr, "shape") else np.shape(arr)
        if len(shp) == 2:
            N = int(shp[0])
            break
    import contextlib
    if stream is not None:          # all device work of this call on ONE stream (see solve_dae_batched)
        if stream != torch.cuda.current_stream(device):
            stream.wait_stream(torch.cuda.current_stream(device))
        stream_ctx = torch.cuda.stream(stream)
    else:
        stream_ctx = contextlib.nullcontext()
    with torch.cuda.device(device), stream_ctx:
        y_d = _as_device_soa(y0, N, n, "y0", device, False).clone()
        yp_d = _as_device_soa(yp0, N, n, "yp0", device, False).clone()
        p_d = _as_device_soa(params, N, npar, "params", device, False) if npar > 0 else None
        f_d = torch.empty((n, N), dtype=torch.float64, device=device)
        status = torch.empty((N,), dtype=torch.int32, device=device)
        n_iter = torch.empty((N,), dtype=torch.int32, device=device)
        opts = _lib.CicOpts(rtol=float(rtol), safety=float(safety), newton_maxiter=int(newton_maxiter),
                            chord_iter=int(chord_iter), analytic_jac=1 if jac in (True, "analytic") else 0)
        cu_stream = ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)
        i32p = ctypes.POINTER(ctypes.c_int32)
        rc = L.b200dae_consistent_ic(pid, N, N, float(t0), y_d.data_ptr(), yp_d.data_ptr(),
                                     p_d.data_ptr() if p_d is not None else None,
                                     atol_v.ctypes.data_as(ctypes.POINTER(ctypes.c_double)),
                                     masks[0].ctypes.data_as(i32p), masks[1].ctypes.data_as(i32p),
                                     ctypes.byref(opts), f_d.data_ptr(), status.data_ptr(),
                                     n_iter.data_ptr(), cu_stream)
        _lib.check(rc)
    res = ConsistentICResult(y0=y_d.t(), yp0=yp_d.t(), f=f_d.t(), status=status, n_iter=n_iter)
    if raise_on_failure:
        st = status.cpu().numpy()
        bad = np.nonzero(st != 0)[0]
        if bad.size:
            code = int(st[bad[0]])
            exc = RuntimeError if code == -1 else ValueError
            raise exc(f"instance {int(bad[0])}: {CIC_MESSAGES[code]}")
    return res
