"""Continuous solutions from the device's dense-output trace.

The reference keeps one interpolant object per accepted step (``RadauDenseOutput``
radau.py:672-699, ``BdfDenseOutput`` bdf.py:422-548) and ``DaeSolution`` (common.py:9-126) selects
the right one for a requested time.  The kernels record the data of exactly those objects per step
(``record_dense``: ``t_old``, ``h``, anchor, order, ``ZP`` / ``D[:order+1]``, include/b200dae.h), so
``sol(t)`` and a step's ``dense_output()`` are evaluated here on the host from the trace of ONE
launch, with the formulas of the reference's ``_call_impl`` methods.
"""
from __future__ import annotations

from math import factorial

import numpy as np


class StepInterpolant:
    """Local interpolant over one step, callable like the reference's ``DAEDenseOutput``
    (base.py:419-454): ``t`` scalar -> ``(y[n], yp[n])``, ``t[m]`` -> ``(y[n, m], yp[n, m])``."""

    def __init__(self, t_old, t):
        self.t_old, self.t = float(t_old), float(t)
        self.t_min, self.t_max = min(self.t, self.t_old), max(self.t, self.t_old)

    def __call__(self, t):
        t = np.asarray(t, dtype=float)
        if t.ndim > 1:
            raise ValueError("`t` must be a float or a 1-D array.")
        y, yp = self._evaluate(np.atleast_1d(t))
        return (y[:, 0], yp[:, 0]) if t.ndim == 0 else (y, yp)


class RadauStepInterpolant(StepInterpolant):
    """y(t) = y_old + ZP @ x^(1..s), y'(t) = ZP @ (k / h) x^(k-1), x = (t - t_old) / h (radau.py:681-699)."""

    def __init__(self, t_old, t, h, y_old, ZP):
        super().__init__(t_old, t)
        self.h, self.y_old, self.ZP = float(h), np.asarray(y_old, dtype=float), np.asarray(ZP, dtype=float)

    def _evaluate(self, t):
        x = (t - self.t_old) / self.h
        c = np.arange(1, self.ZP.shape[1] + 1)[:, None]
        return self.ZP @ (x ** c) + self.y_old[:, None], self.ZP @ ((c / self.h) * x ** (c - 1))


class BdfStepInterpolant(StepInterpolant):
    """Newton form on the backward differences D[:order + 1] anchored at the step end, and its
    derivative (bdf.py:525-548)."""

    def __init__(self, t_old, t, t_anchor, h, order, D):
        super().__init__(t_old, t)
        self.t_anchor, self.h, self.order, self.D = float(t_anchor), float(h), int(order), np.asarray(D, dtype=float)

    def _evaluate(self, t):
        y = np.repeat(self.D[0][:, None], t.size, axis=1)
        yp = np.zeros_like(y)
        for j in range(1, self.order + 1):
            fac, dfac = np.ones_like(t), np.zeros_like(t)
            for m in range(j):
                fac = fac * (t - (self.t_anchor - m * self.h))
                dprod = np.ones_like(t)
                for i in range(j):
                    if i != m:
                        dprod = dprod * (t - (self.t_anchor - i * self.h))
                dfac = dfac + dprod
            denom = factorial(j) * self.h ** j
            y = y + self.D[j][:, None] * fac / denom
            yp = yp + self.D[j][:, None] * dfac / denom
        return y, yp


class DaeSolution:
    """Piecewise continuous solution over the accepted steps, with the segment selection of the
    reference's ``DaeSolution`` (common.py:42-126): a breakpoint belongs to the segment with the
    lower index; times outside ``[t_min, t_max]`` use the first / last segment (extrapolation,
    "not forbidden, but the accuracy is not guaranteed")."""

    def __init__(self, ts, interpolants):
        ts = np.asarray(ts, dtype=float)
        d = np.diff(ts)
        if not ((ts.size == 2 and ts[0] == ts[-1]) or np.all(d > 0) or np.all(d < 0)):
            raise ValueError("`ts` must be strictly increasing or decreasing.")
        if ts.shape != (len(interpolants) + 1,):
            raise ValueError("Numbers of time stamps and interpolants don't match.")
        self.ts, self.interpolants, self.n_segments = ts, list(interpolants), len(interpolants)
        self.ascending = bool(ts[-1] >= ts[0])
        self.t_min, self.t_max = (ts[0], ts[-1]) if self.ascending else (ts[-1], ts[0])
        self._sorted = ts if self.ascending else ts[::-1]
        self._side = "left" if self.ascending else "right"

    def _segments(self, t):
        seg = np.clip(np.searchsorted(self._sorted, t, side=self._side) - 1, 0, self.n_segments - 1)
        return seg if self.ascending else self.n_segments - 1 - seg

    def __call__(self, t):
        t = np.asarray(t, dtype=float)
        if t.ndim > 1:
            raise ValueError("`t` must be a float or a 1-D array.")
        ts = np.atleast_1d(t)
        seg = self._segments(ts)
        n = self.interpolants[0](ts[:1])[0].shape[0]
        y, yp = np.empty((n, ts.size)), np.empty((n, ts.size))
        for s in np.unique(seg):
            m = seg == s
            y[:, m], yp[:, m] = self.interpolants[s](ts[m])
        return (y[:, 0], yp[:, 0]) if t.ndim == 0 else (y, yp)


def step_interpolants(method, stages, t_steps, y_steps, h_steps, t_anchor, order_steps, dense_coef):
    """Interpolant objects of the recorded steps of ONE instance: arrays t_steps[k], y_steps[n, k],
    h_steps[k], t_anchor[k], order_steps[k], dense_coef[k, rows] (row 0 = t0 carries no interpolant)."""
    n = y_steps.shape[0]
    out = []
    for k in range(1, len(t_steps)):
        if method == "Radau":
            out.append(RadauStepInterpolant(t_steps[k - 1], t_steps[k], h_steps[k], y_steps[:, k - 1],
                                            dense_coef[k].reshape(n, stages)))
        else:
            o = int(order_steps[k])
            out.append(BdfStepInterpolant(t_steps[k - 1], t_steps[k], t_anchor[k], h_steps[k], o,
                                          dense_coef[k].reshape(-1, n)[:o + 1]))
    return out


def solution_from_trace(res, i: int = 0) -> DaeSolution:
    """``DaeSolution`` of instance ``i`` of a ``BatchedDaeResult`` obtained with ``record_dense=True``."""
    if res.dense_coef is None:
        raise ValueError("the result carries no dense-output trace: call solve_dae_batched(..., record_steps=K, record_dense=True)")
    to_np = lambda a: a.detach().cpu().numpy() if hasattr(a, "detach") else np.asarray(a)      # noqa: E731
    k = int(to_np(res.n_steps)[i])
    ts = to_np(res.t_steps[i])[:k]
    ips = step_interpolants(res.method, int(res.extra.get("stages", 3)), ts, to_np(res.y_steps[i])[:, :k],
                            to_np(res.h_steps[i])[:k], to_np(res.t_anchor_steps[i])[:k],
                            to_np(res.order_steps[i])[:k], to_np(res.dense_coef[i])[:k])
    if k == 1:          # no step was taken: the reference builds DaeSolution only after a step
        raise ValueError("no accepted step was recorded")
    return DaeSolution(ts, ips)
