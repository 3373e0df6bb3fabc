"""ctypes binding of libb200dae.so (include/b200dae.h).

There is NO fallback: if the CUDA library is missing or does not export the expected symbols
the import of the product fails loudly (`B200DaeLibraryError`)."""
from __future__ import annotations

import ctypes
import os

from .constants import CRadauTables

HERE = os.path.dirname(os.path.abspath(__file__))
# B200DAE_LIB selects an alternative build of the same library (kernel A/B experiments)
LIB_PATH = os.environ.get("B200DAE_LIB") or os.path.join(HERE, "libb200dae.so")

NUM_COUNTERS = 8
COUNTER_NAMES = ("nfev", "njev", "nlu", "nsolve", "n_accepted", "n_err_rejected",
                 "n_newton_fail", "n_newton_iter")
ST_FINISHED, ST_TOO_SMALL_STEP, ST_NONFINITE, ST_MAX_STEPS, ST_RUNNING, ST_EVENT = 0, -1, -2, -3, 1, 2

EXPORTS = ("b200dae_version", "b200dae_last_error", "b200dae_num_problems",
           "b200dae_problem_lookup", "b200dae_problem_name", "b200dae_residual",
           "b200dae_solve_radau", "b200dae_solve_bdf", "b200dae_launch_count",
           "b200dae_fp64_peak_kernel", "b200dae_selftest_fastmath", "b200dae_consistent_ic",
           "b200dae_problem_num_events", "b200dae_register_problem", "b200dae_precompile", "b200dae_host_device_pointer")


class B200DaeLibraryError(ImportError):
    pass


class B200DaeError(RuntimeError):
    pass


class RadauOpts(ctypes.Structure):
    _fields_ = [("rtol", ctypes.c_double), ("max_step", ctypes.c_double),
                ("first_step", ctypes.c_double), ("continuous_error_weight", ctypes.c_double),
                ("jac_recompute_rate", ctypes.c_double),
                ("deadband_lo", ctypes.c_double), ("deadband_hi", ctypes.c_double),
                ("newton_max_iter", ctypes.c_int32), ("newton_iter_embedded", ctypes.c_int32),
                ("max_steps", ctypes.c_int64), ("analytic_jac", ctypes.c_int32),
                ("io_layout", ctypes.c_int32)]


class BdfOpts(ctypes.Structure):
    _fields_ = [("rtol", ctypes.c_double), ("max_step", ctypes.c_double),
                ("first_step", ctypes.c_double), ("max_order", ctypes.c_int32),
                ("ndf_strategy", ctypes.c_int32), ("max_steps", ctypes.c_int64),
                ("analytic_jac", ctypes.c_int32), ("io_layout", ctypes.c_int32)]


class CicOpts(ctypes.Structure):
    _fields_ = [("rtol", ctypes.c_double), ("safety", ctypes.c_double),
                ("newton_maxiter", ctypes.c_int32), ("chord_iter", ctypes.c_int32),
                ("analytic_jac", ctypes.c_int32), ("reserved", ctypes.c_int32)]


class Outputs(ctypes.Structure):
    _fields_ = [("y_final", ctypes.c_void_p), ("yp_final", ctypes.c_void_p),
                ("t_final", ctypes.c_void_p), ("status", ctypes.c_void_p),
                ("counters", ctypes.c_void_p), ("t_eval", ctypes.c_void_p),
                ("m", ctypes.c_int64), ("y_eval", ctypes.c_void_p), ("yp_eval", ctypes.c_void_p),
                ("n_eval_valid", ctypes.c_void_p), ("steps_cap", ctypes.c_int64),
                ("t_steps", ctypes.c_void_p), ("y_steps", ctypes.c_void_p),
                ("yp_steps", ctypes.c_void_p), ("n_steps", ctypes.c_void_p),
                ("h_steps", ctypes.c_void_p), ("order_steps", ctypes.c_void_p),
                ("dense_coef", ctypes.c_void_p), ("coef_rows", ctypes.c_int64),
                ("events", ctypes.c_void_p),
                ("eval_stride_row", ctypes.c_int64), ("eval_stride_comp", ctypes.c_int64),
                ("eval_stride_inst", ctypes.c_int64), ("bad_input", ctypes.c_void_p)]


MAX_EVENTS = 4


class EventSpec(ctypes.Structure):
    _fields_ = [("n_events", ctypes.c_int32), ("of_yp", ctypes.c_int32 * MAX_EVENTS),
                ("component", ctypes.c_int32 * MAX_EVENTS), ("value", ctypes.c_double * MAX_EVENTS),
                ("direction", ctypes.c_int32 * MAX_EVENTS), ("max_events", ctypes.c_int32 * MAX_EVENTS)]


class Events(ctypes.Structure):
    _fields_ = [("spec", EventSpec), ("cap", ctypes.c_int64), ("t_events", ctypes.c_void_p),
                ("y_events", ctypes.c_void_p), ("yp_events", ctypes.c_void_p),
                ("n_found", ctypes.c_void_p)]


_lib = None


def load():
    """Load libb200dae.so; raises B200DaeLibraryError when it is absent or incomplete."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise B200DaeLibraryError(
            f"{LIB_PATH} is missing: build it with `python -m scipy_dae_b200.build` "
            "(nvcc, sm_100a). scipy_dae_b200 has no CPU fallback.")
    try:
        L = ctypes.CDLL(LIB_PATH)
    except OSError as exc:
        raise B200DaeLibraryError(f"cannot load {LIB_PATH}: {exc}") from exc
    missing = [s for s in EXPORTS if not hasattr(L, s)]
    if missing:
        raise B200DaeLibraryError(f"{LIB_PATH} does not export {missing}")
    vp, i64, i32p = ctypes.c_void_p, ctypes.c_int64, ctypes.POINTER(ctypes.c_int32)
    dbl, dblp = ctypes.c_double, ctypes.POINTER(ctypes.c_double)
    L.b200dae_version.restype = ctypes.c_int
    L.b200dae_last_error.restype = ctypes.c_char_p
    L.b200dae_num_problems.restype = ctypes.c_int
    L.b200dae_problem_lookup.restype = ctypes.c_int
    L.b200dae_problem_lookup.argtypes = [ctypes.c_char_p, i32p, i32p]
    L.b200dae_problem_name.restype = ctypes.c_char_p
    L.b200dae_problem_name.argtypes = [ctypes.c_int]
    L.b200dae_problem_num_events.restype = ctypes.c_int
    L.b200dae_problem_num_events.argtypes = [ctypes.c_int]
    L.b200dae_launch_count.restype = i64
    L.b200dae_register_problem.restype = ctypes.c_int
    L.b200dae_register_problem.argtypes = [ctypes.c_char_p, ctypes.c_char_p, ctypes.c_int32, ctypes.c_int32]
    L.b200dae_host_device_pointer.restype = ctypes.c_int
    L.b200dae_host_device_pointer.argtypes = [ctypes.c_void_p, ctypes.POINTER(ctypes.c_void_p)]
    L.b200dae_precompile.restype = ctypes.c_int
    L.b200dae_precompile.argtypes = [ctypes.c_int] + [ctypes.c_int32] * 4
    L.b200dae_residual.restype = ctypes.c_int
    L.b200dae_residual.argtypes = [ctypes.c_int, i64, i64, dbl, vp, vp, vp, vp, vp]
    L.b200dae_solve_radau.restype = ctypes.c_int
    L.b200dae_solve_radau.argtypes = [ctypes.c_int, i64, i64, vp, vp, vp, dbl, dbl, dblp,
                                      ctypes.POINTER(RadauOpts), ctypes.POINTER(CRadauTables),
                                      ctypes.POINTER(Outputs), vp]
    L.b200dae_solve_bdf.restype = ctypes.c_int
    L.b200dae_solve_bdf.argtypes = [ctypes.c_int, i64, i64, vp, vp, vp, dbl, dbl, dblp,
                                    ctypes.POINTER(BdfOpts), ctypes.POINTER(Outputs), vp]
    L.b200dae_consistent_ic.restype = ctypes.c_int
    L.b200dae_consistent_ic.argtypes = [ctypes.c_int, i64, i64, dbl, vp, vp, vp, dblp, i32p, i32p,
                                        ctypes.POINTER(CicOpts), vp, vp, vp, vp]
    L.b200dae_fp64_peak_kernel.restype = ctypes.c_int
    L.b200dae_fp64_peak_kernel.argtypes = [ctypes.c_int, ctypes.c_int, i64, vp, vp]
    L.b200dae_selftest_fastmath.restype = ctypes.c_int
    L.b200dae_selftest_fastmath.argtypes = [i64, vp, vp, vp, vp]
    if L.b200dae_version() != 106:
        raise B200DaeLibraryError(f"{LIB_PATH}: ABI version {L.b200dae_version()} != 106")
    _lib = L
    return L


def check(rc: int):
    if rc != 0:
        msg = load().b200dae_last_error().decode(errors="replace")
        if rc in (-1, -2):
            raise ValueError(msg)
        if rc == -4:
            raise B200DaeError(f"run-time compilation failed:\n{msg}")
        raise B200DaeError(f"libb200dae error {rc}: {msg}")


def problem_lookup(name: str):
    n, npar = ctypes.c_int32(), ctypes.c_int32()
    pid = load().b200dae_problem_lookup(name.encode(), ctypes.byref(n), ctypes.byref(npar))
    if pid < 0:
        raise ValueError(f"problem {name!r} is not compiled into libb200dae.so")
    return pid, n.value, npar.value


def launch_count() -> int:
    return int(load().b200dae_launch_count())
