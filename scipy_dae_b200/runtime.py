"""Residuals registered at run time.

The reference takes any Python callable ``fun(t, y, yp, *args)`` (base.py:44-45). Here the step
loop runs on the device, so a residual is CUDA source text: ``register_problem`` hands it to
``b200dae_register_problem`` (include/b200dae.h), which compiles it with NVRTC against the library's
own kernel templates — the same instantiations the built-in functors get from nvcc — one kernel
variant at a time on first use (a few seconds each, cubins cached on disk).

    src = '''
    struct Problem {
        static constexpr int N = 2, NP = 1;
        static constexpr bool kHasJac = false;
        __device__ static void eval(double t, const double* y, const double* yp, const double* p, double* out) {
            out[0] = yp[0] - y[1];
            out[1] = yp[1] - p[0] * (1.0 - y[0] * y[0]) * y[1] + y[0];
        }
    };'''
    register_problem("vdp", src, n=2, n_params=1, default_params=(5.0,),
                     initial=lambda p: (np.array([2.0, 0.0]), np.array([0.0, -2.0])), t_span=(0.0, 10.0))
    solve_dae_batched("vdp", ...)
"""
from __future__ import annotations

import ctypes

import numpy as np

from . import _lib
from .problems import PROBLEMS, Problem

METHOD_IDS = {"Radau": 0, "BDF": 1}


def _no_twin(name):
    def residual(t, y, yp, *params):
        raise NotImplementedError(f"problem {name!r} was registered without a NumPy twin of its residual")
    return residual


def register_problem(name: str, cuda_src: str, n: int, n_params: int, *, residual=None, jacobian=None,
                     default_params=(), initial=None, initial_batch=None, t_span=(0.0, 1.0), solution=None,
                     events=(), flops: int = 0, doc: str = "", param_names=()) -> Problem:
    """Compile-check ``cuda_src`` (NVRTC, no device needed) and make ``name`` usable with every
    entry point of the package. ``residual`` / ``jacobian`` / ``events`` are optional NumPy twins
    (used by host-side checks and to drive the reference in tests), never by the solve."""
    L = _lib.load()
    pid = L.b200dae_register_problem(name.encode(), cuda_src.encode(), int(n), int(n_params))
    if pid < 0:
        _lib.check(pid)
    if initial is None:
        def initial(params):            # noqa: E306
            raise ValueError(f"problem {name!r} has no `initial`; pass y0 and yp0 explicitly")
    prob = Problem(name=name, n=int(n), n_params=int(n_params), residual=residual or _no_twin(name),
                   default_params=tuple(default_params), initial=initial, t_span=tuple(t_span), doc=doc,
                   param_names=tuple(param_names), initial_batch=initial_batch, jacobian=jacobian, solution=solution,
                   events=tuple(events), flops=int(flops))
    PROBLEMS[name] = prob
    return prob


def precompile(problem, method: str = "Radau", stages: int = 3, t_eval: bool = False, jac=None) -> None:
    """Compile the kernel a solve with these settings would launch (no device needed); a no-op for
    built-in problems."""
    name = problem.name if isinstance(problem, Problem) else problem
    pid, _, _ = _lib.problem_lookup(name)
    _lib.check(_lib.load().b200dae_precompile(pid, METHOD_IDS[method], int(stages), int(bool(t_eval)),
                                              int(jac == "analytic")))
