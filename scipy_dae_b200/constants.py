"""Method tables for the device integrators, computed once on the host in float64.

Radau IIA (any odd number of stages ``s``; 3/5/7 are exercised): Butcher matrix ``A``,
nodes ``c``, the real block-diagonalising transformation ``T``/``TI`` of ``A``
(``TI @ A @ T = blockdiag(gamma, [[alpha, beta], [-beta, alpha]], ...)``), the
eigenvalues ``mu_real = gamma`` and ``mu_complex = alpha - i beta`` that enter the
iteration matrices ``Jyp + h*mu*Jy``, the interpolation matrix ``P`` of the
collocation polynomial, and the weights of the explicit / implicit embedded error
estimators (Hairer-Wanner IV.8; de Swart-Soederlind 1997).

These are the same mathematical objects the reference derives in
``radau_constants`` (radau.py:51-134) and they are obtained from the same
NumPy/SciPy primitives (polynomial roots, ``eig``, ``cdf2rdf``, Vandermonde
inverses) on purpose: SURVEY.md App. E notes the tables carry ~1e-14 eigen-solver
noise (e.g. ``c[-1] = 1.0000000000000082`` for s=5) and that step sequences are
sensitive to it, so textbook literals would drift from the reference.
``tests/test_constants.py`` pins them against values frozen from the reference.

BDF/NDF (bdf.py:242-252): ``kappa`` per strategy, ``gamma_k = sum_{j<=k} 1/j``,
``alpha = (1 - kappa) * gamma``, ``error_const = kappa*gamma + 1/(k+1)``.
"""
from __future__ import annotations

import ctypes
from dataclasses import dataclass
from functools import lru_cache

import numpy as np
from numpy.polynomial import Polynomial
from scipy.linalg import cdf2rdf, eig

EPS = float(np.finfo(float).eps)
MAX_STAGES = 7
MAX_ORDER = 6

# de Swart / Soederlind damping ratio b_hat1 / gamma of the implicit embedded method
# (radau.py:9: DAMPING_RATIO_ERROR_ESTIMATE)
DAMPING_RATIO_ERROR_ESTIMATE = 0.05


@dataclass(frozen=True)
class RadauTables:
    s: int
    A: np.ndarray            # (s, s)
    c: np.ndarray            # (s,)
    T: np.ndarray            # (s, s)
    TI: np.ndarray           # (s, s)
    P: np.ndarray            # (s, s) interpolation: ZP = Z.T @ P
    P2: np.ndarray           # (s, s) inverse Vandermonde in c (collocation error uses row 0)
    b_hat1_implicit: float
    v_implicit: np.ndarray   # (s,)
    v_explicit: np.ndarray   # (s,)
    b_hat_implicit: np.ndarray
    b_hat_explicit: np.ndarray
    mu_real: float
    mu_complex: np.ndarray   # (s//2,) complex
    order: int


@lru_cache(maxsize=None)
def radau_tables(s: int) -> RadauTables:
    if s < 1 or s % 2 != 1 or s > MAX_STAGES:
        raise ValueError(f"Radau IIA needs an odd number of stages in [1, {MAX_STAGES}], got {s}")
    # nodes: zeros of d^{s-1}/dx^{s-1} [ x^{s-1} (x-1)^s ]   (Hairer-Wanner IV.5 (5.2))
    c = (Polynomial([0, 1]) ** (s - 1) * Polynomial([-1, 1]) ** s).deriv(s - 1).roots()
    # collocation conditions  A V = diag(c) V diag(1/k)
    V = np.vander(c, increasing=True)
    A = np.diag(c) @ V @ np.diag(1.0 / np.arange(1, s + 1)) @ np.linalg.inv(V)
    b = A[-1, :]

    lam, Q = eig(A)
    order_desc = np.argsort(lam)[::-1]
    lam, Q = lam[order_desc], Q[:, order_desc]
    for i in range(s):
        Q[:, i] /= Q[-1, i]
    _, T = cdf2rdf(lam, Q)
    TI = np.linalg.inv(T)

    is_real = lam.imag == 0
    gamma = lam[is_real].real[0]
    pair = lam[~is_real][::2]
    mu_complex = pair.real - 1j * pair.imag

    c_hat = np.concatenate(([0.0], c))
    Vh = np.vander(c_hat, increasing=True).T
    moments = 1.0 / np.arange(1, s + 1)

    rhs = moments.copy()
    rhs[0] -= gamma
    b_hat_explicit = np.linalg.solve(Vh[:-1, 1:], rhs)
    v_explicit = b_hat_explicit - b

    b_hat1 = DAMPING_RATIO_ERROR_ESTIMATE * gamma
    rhs = moments.copy()
    rhs[0] -= b_hat1
    rhs -= gamma
    b_hat_implicit = np.linalg.solve(Vh[:-1, 1:], rhs)
    v_implicit = b - b_hat_implicit

    P = np.linalg.inv(Vh[1:, 1:])
    P2 = np.linalg.inv(np.vander(c, increasing=True))

    return RadauTables(s=s, A=A, c=c, T=T, TI=TI, P=P, P2=P2, b_hat1_implicit=float(b_hat1),
                       v_implicit=v_implicit, v_explicit=v_explicit,
                       b_hat_implicit=b_hat_implicit, b_hat_explicit=b_hat_explicit,
                       mu_real=float(gamma), mu_complex=mu_complex, order=2 * s - 1)


def radau_newton_tol(rtol: float, s: int) -> float:
    """radau.py:422,432-433."""
    expmns = (s + 1) / (2 * s)
    return max(10 * EPS / rtol, min(0.03, rtol ** (1 / expmns - 1)))


def radau_newton_max_iter(s: int) -> int:
    """radau.py:440-441."""
    return 7 + int((s - 3) * 2.5)


@dataclass(frozen=True)
class BdfTables:
    max_order: int
    kappa: np.ndarray
    gamma: np.ndarray
    alpha: np.ndarray
    error_const: np.ndarray


NDF_STRATEGIES = {"stability": 0, "efficiency": 1, None: 2, "none": 2, "bdf": 2}


def bdf_tables(max_order: int = 5, NDF_strategy="stability") -> BdfTables:
    if not 1 <= max_order <= MAX_ORDER:
        raise ValueError("Ensure that 1 <= max_order <= 6.")
    if NDF_strategy == "stability":
        kappa = np.array([0, -37 / 200, -1 / 9, 0.0834, 0.0665, 0.0551, 0.0464])
    elif NDF_strategy == "efficiency":
        kappa = np.array([0, -37 / 200, -1 / 9, -0.0823, -0.0415, 0, 0])
    else:
        kappa = np.array([0, -37 / 200, -1 / 9, 0, 0, 0, 0])
    kappa = kappa[:max_order + 1]
    gamma = np.hstack((0, np.cumsum(1 / np.arange(1, max_order + 1))))
    alpha = (1 - kappa) * gamma
    error_const = kappa * gamma + 1 / np.arange(1, max_order + 2)
    return BdfTables(max_order, kappa, gamma, alpha, error_const)


def bdf_newton_tol(rtol: float) -> float:
    """bdf.py:234."""
    return max(10 * EPS / rtol, min(0.03, rtol ** 0.5))


# --------------------------------------------------------------------------- #
# ctypes mirrors of the structs in include/b200dae.h (shared with oracle/dae_oracle.h)
# --------------------------------------------------------------------------- #
class CRadauTables(ctypes.Structure):
    _fields_ = [
        ("s", ctypes.c_int),
        ("order", ctypes.c_int),
        ("A", ctypes.c_double * (MAX_STAGES * MAX_STAGES)),
        ("C", ctypes.c_double * MAX_STAGES),
        ("T", ctypes.c_double * (MAX_STAGES * MAX_STAGES)),
        ("TI", ctypes.c_double * (MAX_STAGES * MAX_STAGES)),
        ("P", ctypes.c_double * (MAX_STAGES * MAX_STAGES)),
        ("P2row0", ctypes.c_double * MAX_STAGES),
        ("v_implicit", ctypes.c_double * MAX_STAGES),
        ("v_explicit", ctypes.c_double * MAX_STAGES),
        ("b_hat_implicit", ctypes.c_double * MAX_STAGES),
        ("mu_real", ctypes.c_double),
        ("mu_re", ctypes.c_double * (MAX_STAGES // 2)),
        ("mu_im", ctypes.c_double * (MAX_STAGES // 2)),
        ("b_hat1_implicit", ctypes.c_double),
    ]


def radau_tables_c(s: int) -> CRadauTables:
    """Pack ``radau_tables(s)`` into the C struct; matrices are row-major with stride ``s``."""
    tb = radau_tables(s)
    ct = CRadauTables()
    ct.s = s
    ct.order = tb.order
    for name, arr in (("A", tb.A), ("T", tb.T), ("TI", tb.TI), ("P", tb.P)):
        flat = np.ascontiguousarray(arr, dtype=float).ravel()
        dst = getattr(ct, name)
        for i, v in enumerate(flat):
            dst[i] = v
    for name, arr in (("C", tb.c), ("P2row0", tb.P2[0]), ("v_implicit", tb.v_implicit),
                      ("v_explicit", tb.v_explicit), ("b_hat_implicit", tb.b_hat_implicit)):
        dst = getattr(ct, name)
        for i, v in enumerate(np.asarray(arr, dtype=float)):
            dst[i] = v
    ct.mu_real = tb.mu_real
    for i, mu in enumerate(tb.mu_complex):
        ct.mu_re[i] = mu.real
        ct.mu_im[i] = mu.imag
    ct.b_hat1_implicit = tb.b_hat1_implicit
    return ct
