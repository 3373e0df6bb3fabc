"""Problem zoo: host-side (NumPy) statements of the residuals F(t, y, y', p) = 0
that are compiled into the CUDA library as device functors.

Every entry here has a twin with the same name and the same arithmetic in
``scipy_dae_b200/csrc/problems.cuh`` (device functor) and in ``oracle/problems.c``
(CPU oracle).  The NumPy versions are used

* by ``oracle/gen_golden.py`` to drive the *unmodified* reference ``solve_dae``
  (they are plain ``fun(t, y, yp, *params)`` callables, the reference's residual
  convention, base.py:44-45 / dae.py:365-383),
* by host-side consistency checks (``F(t0, y0, yp0) == 0``) and tests.

They are never on the product hot path: the hot path evaluates the device functor.

Problem definitions follow SURVEY.md App. D:
  robertson  : README.md:37-47, tests/test_integration_robertson.py:73-87 (parameterised k1,k2,k3)
  akzo_nobel : examples/daes/akzo_nobel.py:18-50 (parameterised rate constants)
  transamp   : two-transistor amplifier, IVP test set "transamp" (n=8), SURVEY.md App. D.5
  andrews    : examples/daes/andrews.py:20-225 (n=26 stabilised index-1 form)
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Callable, Sequence

import numpy as np


@dataclass(frozen=True)
class Problem:
    """A residual known to the CUDA library under ``name``.

    ``residual(t, y, yp, *params)`` is the NumPy twin of the device functor.
    ``initial(params)`` returns consistent ``(y0, yp0)`` for one parameter vector.
    """
    name: str
    n: int
    n_params: int
    residual: Callable
    default_params: tuple
    initial: Callable
    t_span: tuple
    doc: str = ""
    param_names: tuple = field(default_factory=tuple)
    initial_batch: Callable = None   # optional vectorised initial(params[N, p]) -> (y0[N, n], yp0[N, n])
    jacobian: Callable = None        # optional analytic jac(t, y, yp, *params) -> (Jy[n, n], Jyp[n, n])
    solution: Callable = None        # optional exact solution(t, *params) -> (y[n, ...], yp[n, ...])
    t0_default: float = None         # start of t_span (for problems whose ICs depend on it)
    events: tuple = ()               # NumPy twins g(t, y, yp, *params) of the functor's event functions
    flops: int = 0                   # flops of one residual as written (the functor's kFlops; roofline model)

    def event_callable(self, index, spec=None):
        """Event function `index` as the callable the reference's `events=` takes (dae.py:226-247),
        with `value` subtracted and the `direction` / `terminal` attributes of `spec` set."""
        spec = spec or {}
        g, c = self.events[int(index)], float(spec.get("value", 0.0))
        # (for n = 1 the reference's dense output hands a 0-d array to the event function: np.squeeze,
        # radau.py:696-697)
        fn = lambda t, y, yp, *a: g(t, np.atleast_1d(y), np.atleast_1d(yp), *a) - c          # noqa: E731
        fn.direction = float(spec.get("direction", 0))
        fn.terminal = spec.get("terminal", False)
        return fn

    def __call__(self, t, y, yp, *params):
        if not params:
            params = self.default_params
        return self.residual(t, y, yp, *params)


# --------------------------------------------------------------------------- #
# Robertson (n = 3, index 1)
# --------------------------------------------------------------------------- #
def robertson_residual(t, y, yp, k1, k2, k3):
    y1, y2, y3 = y
    y1p, y2p, y3p = yp
    return np.array([
        y1p - (-k1 * y1 + k2 * y2 * y3),
        y2p - (k1 * y1 - k2 * y2 * y3 - k3 * y2 * y2),
        y1 + y2 + y3 - 1,
    ])


def robertson_jacobian(t, y, yp, k1, k2, k3):
    """(dF/dy, dF/dy') of robertson_residual (reference: examples/daes/robertson.py:38-47)."""
    Jy = np.array([[k1, -k2 * y[2], -k2 * y[1]],
                   [-k1, k2 * y[2] + 2.0 * k3 * y[1], k2 * y[1]],
                   [1.0, 1.0, 1.0]])
    Jyp = np.diag([1.0, 1.0, 0.0])
    return Jy, Jyp


ROBERTSON_EVENTS = (lambda t, y, yp, k1, k2, k3: y[0] - y[2],
                    lambda t, y, yp, k1, k2, k3: yp[0] + yp[1] + k1 * y[0] * 0.5)
AKZO_EVENTS = (lambda t, y, yp, *p: y[0] - y[2],)
COSINE_EVENTS = (lambda t, y, yp, a: y[0] - yp[0], lambda t, y, yp, a: y[0] - a * t * 0.5)


def robertson_initial(params):
    k1 = params[0]
    return np.array([1.0, 0.0, 0.0]), np.array([-k1, k1, 0.0])


def robertson_initial_batch(params):
    N = params.shape[0]
    y0 = np.zeros((N, 3))
    y0[:, 0] = 1.0
    yp0 = np.zeros((N, 3))
    yp0[:, 0] = -params[:, 0]
    yp0[:, 1] = params[:, 0]
    return y0, yp0


# --------------------------------------------------------------------------- #
# Akzo Nobel chemical reaction (n = 6, index 1)
# --------------------------------------------------------------------------- #
AKZO_PO2 = 0.9
AKZO_HEN = 737.0


def akzo_rhs(y, k1, k2, k3, k4, kbig, kla, ks):
    sq = np.sqrt(y[1])
    y0sq = y[0] * y[0]
    r1 = k1 * (y0sq * y0sq) * sq
    r2 = k2 * y[2] * y[3]
    r3 = k2 / kbig * y[0] * y[4]
    r4 = k3 * y[0] * (y[3] * y[3])
    r5 = k4 * (y[5] * y[5]) * sq
    fin = kla * (AKZO_PO2 / AKZO_HEN - y[1])
    return np.array([
        -2.0 * r1 + r2 - r3 - r4,
        -0.5 * r1 - r4 - 0.5 * r5 + fin,
        r1 - r2 + r3,
        -r2 + r3 - 2.0 * r4,
        r2 - r3 + r5,
        ks * y[0] * y[3] - y[5],
    ])


def akzo_residual(t, y, yp, k1, k2, k3, k4, kbig, kla, ks):
    with np.errstate(invalid="ignore"):
        r = akzo_rhs(y, k1, k2, k3, k4, kbig, kla, ks)
    out = np.empty(6)
    out[:5] = yp[:5] - r[:5]
    out[5] = -r[5]
    return out


def akzo_jacobian(t, y, yp, k1, k2, k3, k4, kbig, kla, ks):
    """(dF/dy, dF/dy') of akzo_residual (reference: examples/daes/akzo_nobel.py:52-103)."""
    with np.errstate(invalid="ignore", divide="ignore"):
        sq = np.sqrt(y[1])
        hsq = 0.5 / sq
    y0c = y[0] * y[0] * y[0]
    r1_0, r1_1 = 4.0 * k1 * y0c * sq, k1 * (y0c * y[0]) * hsq
    r2_2, r2_3 = k2 * y[3], k2 * y[2]
    kk = k2 / kbig
    r3_0, r3_4 = kk * y[4], kk * y[0]
    r4_0, r4_3 = k3 * (y[3] * y[3]), 2.0 * k3 * y[0] * y[3]
    r5_1, r5_5 = k4 * (y[5] * y[5]) * hsq, 2.0 * k4 * y[5] * sq
    df = np.zeros((6, 6))
    df[0, 0] = -2.0 * r1_0 - r3_0 - r4_0; df[0, 1] = -2.0 * r1_1; df[0, 2] = r2_2
    df[0, 3] = r2_3 - r4_3; df[0, 4] = -r3_4
    df[1, 0] = -0.5 * r1_0 - r4_0; df[1, 1] = -0.5 * r1_1 - 0.5 * r5_1 - kla
    df[1, 3] = -r4_3; df[1, 5] = -0.5 * r5_5
    df[2, 0] = r1_0 + r3_0; df[2, 1] = r1_1; df[2, 2] = -r2_2; df[2, 3] = -r2_3; df[2, 4] = r3_4
    df[3, 0] = r3_0 - 2.0 * r4_0; df[3, 2] = -r2_2; df[3, 3] = -r2_3 - 2.0 * r4_3; df[3, 4] = r3_4
    df[4, 0] = -r3_0; df[4, 1] = r5_1; df[4, 2] = r2_2; df[4, 3] = r2_3; df[4, 4] = -r3_4; df[4, 5] = r5_5
    df[5, 0] = ks * y[3]; df[5, 3] = ks * y[0]; df[5, 5] = -1.0
    Jy = -df
    Jyp = np.diag([1.0, 1.0, 1.0, 1.0, 1.0, 0.0])
    return Jy, Jyp


def akzo_initial(params):
    ks = params[6]
    y0 = np.array([0.444, 0.00123, 0.0, 0.007, 0.0, ks * 0.444 * 0.007])
    yp0 = akzo_rhs(y0, *params)
    yp0[5] = 0.0 * yp0[5]  # algebraic row: F does not depend on yp[5]
    return y0, yp0


# --------------------------------------------------------------------------- #
# Two-transistor amplifier "transamp" (n = 8, index 1)
# --------------------------------------------------------------------------- #
TA_UF = 0.026
TA_ALPHA = 0.99
TA_R = 9000.0


TA_RUF = 1.0 / TA_UF
TA_RR = 1.0 / TA_R
TA_R2R = 2.0 / TA_R


def transamp_residual(t, y, yp, Ub, R0, beta):
    """Divisions by the constants UF, R and by the parameter R0 are written as multiplications by
    their reciprocals, operation by operation as in the device functor and the C twin: the three
    are bit-identical up to exp / sin (the problem is restated here)."""
    C1, C2, C3, C4, C5 = 1e-6, 2e-6, 3e-6, 4e-6, 5e-6
    rUF, rR, r2R = TA_RUF, TA_RR, TA_R2R
    rR0 = 1.0 / R0
    Ue = 0.1 * np.sin((200.0 * np.pi) * t)
    g23 = beta * (np.exp((y[1] - y[2]) * rUF) - 1.0)
    g56 = beta * (np.exp((y[4] - y[5]) * rUF) - 1.0)
    UbR = Ub * rR
    f = np.array([
        (y[0] - Ue) * rR0,
        -UbR + y[1] * r2R - (TA_ALPHA - 1.0) * g23,
        -g23 + y[2] * rR,
        -UbR + y[3] * rR + TA_ALPHA * g23,
        -UbR + y[4] * r2R - (TA_ALPHA - 1.0) * g56,
        -g56 + y[5] * rR,
        -UbR + y[6] * rR + TA_ALPHA * g56,
        y[7] * rR,
    ])
    My = np.array([
        -C1 * yp[0] + C1 * yp[1],
        C1 * yp[0] - C1 * yp[1],
        -C2 * yp[2],
        -C3 * yp[3] + C3 * yp[4],
        C3 * yp[3] - C3 * yp[4],
        -C4 * yp[5],
        -C5 * yp[6] + C5 * yp[7],
        C5 * yp[6] - C5 * yp[7],
    ])
    return My - f


def transamp_jacobian(t, y, yp, Ub, R0, beta):
    """(dF/dy, dF/dy') of transamp_residual: F = M yp - f(t, y), so Jyp = M and Jy = -df/dy."""
    C1, C2, C3, C4, C5 = 1e-6, 2e-6, 3e-6, 4e-6, 5e-6
    R, a = TA_R, TA_ALPHA
    e23 = beta * np.exp((y[1] - y[2]) / TA_UF) / TA_UF
    e56 = beta * np.exp((y[4] - y[5]) / TA_UF) / TA_UF
    df = np.zeros((8, 8))
    df[0, 0] = 1.0 / R0
    df[1, 1] = 2.0 / R - (a - 1.0) * e23; df[1, 2] = (a - 1.0) * e23
    df[2, 1] = -e23; df[2, 2] = e23 + 1.0 / R
    df[3, 1] = a * e23; df[3, 2] = -a * e23; df[3, 3] = 1.0 / R
    df[4, 4] = 2.0 / R - (a - 1.0) * e56; df[4, 5] = (a - 1.0) * e56
    df[5, 4] = -e56; df[5, 5] = e56 + 1.0 / R
    df[6, 4] = a * e56; df[6, 5] = -a * e56; df[6, 6] = 1.0 / R
    df[7, 7] = 1.0 / R
    M = np.zeros((8, 8))
    M[0, 0] = -C1; M[0, 1] = C1; M[1, 0] = C1; M[1, 1] = -C1
    M[2, 2] = -C2
    M[3, 3] = -C3; M[3, 4] = C3; M[4, 3] = C3; M[4, 4] = -C3
    M[5, 5] = -C4
    M[6, 6] = -C5; M[6, 7] = C5; M[7, 6] = C5; M[7, 7] = -C5
    return -df, M


def transamp_initial(params):
    Ub = params[0]
    R = TA_R
    C2, C4 = 2e-6, 4e-6
    y0 = np.array([0.0, Ub / 2.0, Ub / 2.0, Ub, Ub / 2.0, Ub / 2.0, Ub, 0.0])
    yp0 = np.array([
        51.338775, 51.338775, -Ub / (2.0 * C2 * R),
        -24.9757667, -24.9757667, -Ub / (2.0 * C4 * R),
        -10.00564453, -10.00564453,
    ])
    return y0, yp0


# --------------------------------------------------------------------------- #
# Andrews' squeezer, stabilised index-1 form (n = 26); p = (mom, c0)
# --------------------------------------------------------------------------- #
_AND = dict(
    m1=0.04325, m2=0.00365, m3=0.02373, m4=0.00706, m5=0.07050, m6=0.00706, m7=0.05498,
    I1=2.194e-6, I2=4.410e-7, I3=5.255e-6, I4=5.667e-7, I5=1.169e-5, I6=5.667e-7, I7=1.912e-5,
    xa=-0.06934, ya=-0.00227, xb=-0.03635, yb=0.03273, xc=0.014, yc=0.072,
    d=0.028, da=0.0115, e=0.02, ea=0.01421, zf=0.02, fa=0.01421,
    rr=0.007, ra=0.00092, ss=0.035, sa=0.01874, sb=0.01043, sc=0.018,
    sd=0.02, zt=0.04, ta=0.02308, tb=0.00916, u=0.04, ua=0.01228, ub=0.00449, l0=0.07785)

ANDREWS_Q0 = np.array([
    -0.0617138900142764496358948458001, 0.0, 0.455279819163070380255912382449,
    0.222668390165885884674473185609, 0.487364979543842550225598953530,
    -0.222668390165885884674473185609, 1.23054744454982119249735015568])


def _andrews_terms(q, v, mom, c0):
    """Mass matrix M(q), forces h(q, v), constraints g(q) and their Jacobian G(q) (6 x 7)."""
    c = _AND
    sin, cos = np.sin, np.cos
    sbe, cbe = sin(q[0]), cos(q[0])
    sth, cth = sin(q[1]), cos(q[1])
    sga, cga = sin(q[2]), cos(q[2])
    sph, cph = sin(q[3]), cos(q[3])
    sde, cde = sin(q[4]), cos(q[4])
    som, com = sin(q[5]), cos(q[5])
    sep, cep = sin(q[6]), cos(q[6])
    sbt, cbt = sin(q[0] + q[1]), cos(q[0] + q[1])
    spd, cpd = sin(q[3] + q[4]), cos(q[3] + q[4])
    soe, coe = sin(q[5] + q[6]), cos(q[5] + q[6])
    eea, zfa = c["e"] - c["ea"], c["zf"] - c["fa"]

    M = np.zeros((7, 7))
    M[0, 0] = c["m1"] * c["ra"] ** 2 + c["m2"] * (c["rr"] ** 2 - 2 * c["da"] * c["rr"] * cth + c["da"] ** 2) + c["I1"] + c["I2"]
    M[0, 1] = M[1, 0] = c["m2"] * (c["da"] ** 2 - c["da"] * c["rr"] * cth) + c["I2"]
    M[1, 1] = c["m2"] * c["da"] ** 2 + c["I2"]
    M[2, 2] = c["m3"] * (c["sa"] ** 2 + c["sb"] ** 2) + c["I3"]
    M[3, 3] = c["m4"] * eea ** 2 + c["I4"]
    M[3, 4] = M[4, 3] = c["m4"] * (eea ** 2 + c["zt"] * eea * sph) + c["I4"]
    M[4, 4] = c["m4"] * (c["zt"] ** 2 + 2 * c["zt"] * eea * sph + eea ** 2) + c["m5"] * (c["ta"] ** 2 + c["tb"] ** 2) + c["I4"] + c["I5"]
    M[5, 5] = c["m6"] * zfa ** 2 + c["I6"]
    M[5, 6] = M[6, 5] = c["m6"] * (zfa ** 2 - c["u"] * zfa * som) + c["I6"]
    M[6, 6] = c["m6"] * (zfa ** 2 - 2 * c["u"] * zfa * som + c["u"] ** 2) + c["m7"] * (c["ua"] ** 2 + c["ub"] ** 2) + c["I6"] + c["I7"]

    xd = c["sd"] * cga + c["sc"] * sga + c["xb"]
    yd = c["sd"] * sga - c["sc"] * cga + c["yb"]
    L = np.sqrt((xd - c["xc"]) ** 2 + (yd - c["yc"]) ** 2)
    Fs = -c0 * (L - c["l0"]) / L
    Fx, Fy = Fs * (xd - c["xc"]), Fs * (yd - c["yc"])
    h = np.array([
        mom - c["m2"] * c["da"] * c["rr"] * v[1] * (v[1] + 2 * v[0]) * sth,
        c["m2"] * c["da"] * c["rr"] * v[0] ** 2 * sth,
        Fx * (c["sc"] * cga - c["sd"] * sga) + Fy * (c["sd"] * cga + c["sc"] * sga),
        c["m4"] * c["zt"] * eea * v[4] ** 2 * cph,
        -c["m4"] * c["zt"] * eea * v[3] * (v[3] + 2 * v[4]) * cph,
        -c["m6"] * c["u"] * zfa * v[6] ** 2 * com,
        c["m6"] * c["u"] * zfa * v[5] * (v[5] + 2 * v[6]) * com])

    gc = c["rr"] * cbe - c["d"] * cbt
    gs = c["rr"] * sbe - c["d"] * sbt
    g = np.array([
        gc - c["ss"] * sga - c["xb"],
        gs + c["ss"] * cga - c["yb"],
        gc - c["e"] * spd - c["zt"] * cde - c["xa"],
        gs + c["e"] * cpd - c["zt"] * sde - c["ya"],
        gc - c["zf"] * coe - c["u"] * sep - c["xa"],
        gs - c["zf"] * soe + c["u"] * cep - c["ya"]])
    a0, a1 = -c["rr"] * sbe + c["d"] * sbt, c["d"] * sbt
    b0, b1 = c["rr"] * cbe - c["d"] * cbt, -c["d"] * cbt
    G = np.zeros((6, 7))
    G[0::2, 0], G[0::2, 1] = a0, a1
    G[1::2, 0], G[1::2, 1] = b0, b1
    G[0, 2], G[1, 2] = -c["ss"] * cga, -c["ss"] * sga
    G[2, 3], G[2, 4] = -c["e"] * cpd, -c["e"] * cpd + c["zt"] * sde
    G[3, 3], G[3, 4] = -c["e"] * spd, -c["e"] * spd - c["zt"] * cde
    G[4, 5], G[4, 6] = c["zf"] * soe, c["zf"] * soe - c["u"] * cep
    G[5, 5], G[5, 6] = -c["zf"] * coe, -c["zf"] * coe - c["u"] * sep
    return M, h, g, G


# rows k of G = g_q with a structural non-zero in column i
_AND_NZ = {0: range(6), 1: range(6), 2: (0, 1), 3: (2, 3), 4: (2, 3), 5: (4, 5), 6: (4, 5)}


def andrews_residual(t, y, yp, mom, c0):
    q, v = y[:7], y[7:14]
    qd, vd, la, mu = yp[:7], yp[7:14], yp[14:20], yp[20:26]
    M, h, g, G = _andrews_terms(q, v, mom, c0)
    # products written out term by term in row order (the structural non-zeros of M and G), the
    # order of the C and device twins; a BLAS matvec would sum in an implementation-defined order
    gtm = [sum((G[k, i] * mu[k] for k in range(6) if k in _AND_NZ[i]), 0.0) for i in range(7)]
    gtl = [sum((G[k, i] * la[k] for k in range(6) if k in _AND_NZ[i]), 0.0) for i in range(7)]
    Mvd = [M[0, 0] * vd[0] + M[0, 1] * vd[1], M[0, 1] * vd[0] + M[1, 1] * vd[1], M[2, 2] * vd[2],
           M[3, 3] * vd[3] + M[3, 4] * vd[4], M[3, 4] * vd[3] + M[4, 4] * vd[4],
           M[5, 5] * vd[5] + M[5, 6] * vd[6], M[5, 6] * vd[5] + M[6, 6] * vd[6]]
    Gv = [sum((G[k, i] * v[i] for i in range(7) if k in _AND_NZ[i]), 0.0) for k in range(6)]
    return np.array([qd[i] - v[i] + gtm[i] for i in range(7)] + [Mvd[i] - h[i] - gtl[i] for i in range(7)]
                    + Gv + list(g))


def andrews_initial(params):
    """q0 of the IVP test set at rest (u0 = 0); the accelerations and multipliers follow from
    [[M, -G^T], [G, 0]] [u'; lambda] = [h; 0]  (g_qq(u, u) = 0 because u0 = 0), so they stay
    consistent for every (mom, c0) draw."""
    mom, c0 = params
    q0 = ANDREWS_Q0
    M, h, _, G = _andrews_terms(q0, np.zeros(7), mom, c0)
    K = np.block([[M, -G.T], [G, np.zeros((6, 6))]])
    sol = np.linalg.solve(K, np.concatenate([h, np.zeros(6)]))
    y0 = np.concatenate([q0, np.zeros(7 + 6 + 6)])
    yp0 = np.concatenate([np.zeros(7), sol[:7], sol[7:], np.zeros(6)])
    return y0, yp0


# --------------------------------------------------------------------------- #
# registry
# --------------------------------------------------------------------------- #
# --------------------------------------------------------------------------- #
# The reference's own known-answer test problems
# --------------------------------------------------------------------------- #
def rational_residual(t, y, yp):
    """tests/test_integration_rational.py:12-13 with scipy's test_ivp.fun_rational."""
    return np.array([yp[0] - y[1] / t,
                     yp[1] - y[1] * (y[0] + 2 * y[1] - 1) / (t * (y[0] - 1))])


def rational_jacobian(t, y, yp):
    Jy = -np.array([[0, 1 / t],
                    [-2 * y[1] ** 2 / (t * (y[0] - 1) ** 2), (y[0] + 4 * y[1] - 1) / (t * (y[0] - 1))]])
    return Jy, np.eye(2)


def rational_solution(t):
    t = np.asarray(t, dtype=float)
    return (np.array([t / (t + 10), 10 * t / (t + 10) ** 2]),
            np.array([10 / (t + 10) ** 2, 10 * (10 - t) / (t + 10) ** 3]))


def linear_residual(t, y, yp):
    """tests/test_integration_linear.py:9-10 with scipy's test_ivp.fun_linear."""
    return np.array([yp[0] - (-y[0] - 5 * y[1]), yp[1] - (y[0] + y[1])])


def linear_jacobian(t, y, yp):
    return np.array([[1.0, 5.0], [-1.0, -1.0]]), np.eye(2)


def linear_solution(t):
    t = np.asarray(t, dtype=float)
    return (np.array([-5 * np.sin(2 * t), 2 * np.cos(2 * t) + np.sin(2 * t)]),
            np.array([-10 * np.cos(2 * t), -4 * np.sin(2 * t) + 2 * np.cos(2 * t)]))


def cosine_residual(t, y, yp, a):
    """tests/test_events.py:8-9 (a = 1)."""
    return np.array([yp[0] - a * np.cos(t)])


def cosine_solution(t, a=1.0):
    t = np.asarray(t, dtype=float)
    return np.array([a * np.sin(t)]), np.array([a * np.cos(t)])


def weissinger_residual(t, y, yp, a):
    """benchmarks/weissinger/weissinger.py:11-17 for z = a*y."""
    u, up = y[0] / a, yp[0] / a
    return np.array([t * (u * u) * (up * up * up) - (u * u * u) * (up * up) + t * (t * t + 1) * up - (t * t) * u])


def weissinger_solution(t, a=1.0):
    t = np.asarray(t, dtype=float)
    r = np.sqrt(t ** 2 + 0.5)
    return np.array([a * r]), np.array([a * t / r])


def brenan_residual(t, y, yp, A):
    """benchmarks/brenan/brenan.py:11-19 with amplitude A."""
    return np.array([yp[0] - t * yp[1] + y[0] - (1 + t) * y[1], y[1] - A * np.sin(t)])


def brenan_solution(t, A=1.0, C=1.0):
    t = np.asarray(t, dtype=float)
    return (np.array([C * np.exp(-t) + A * t * np.sin(t), A * np.sin(t)]),
            np.array([-C * np.exp(-t) + A * (np.sin(t) + t * np.cos(t)), A * np.cos(t)]))


# --------------------------------------------------------------------------- #
# One-transistor amplifier of the reference's examples/daes/stiff_transistor_amplifier.py (n = 5)
# --------------------------------------------------------------------------- #
def transamp1_residual(t, y, yp, Ub, R0, beta):
    """examples/daes/stiff_transistor_amplifier.py:20-50 with (Ub, R0, beta) as parameters."""
    U1, U2, U3, U4, U5 = y
    Up1, Up2, Up3, Up4, Up5 = yp
    R = TA_R
    C1, C2, C3 = 1e-6, 2e-6, 3e-6
    Ue = 0.4 * np.sin((200.0 * np.pi) * t)
    f23 = beta * (np.exp((U2 - U3) / TA_UF) - 1.0)
    return np.array([
        (Ue - U1) / R0 + C1 * (Up2 - Up1),
        (Ub - U2) / R - U2 / R + C1 * (Up1 - Up2) - (1.0 - TA_ALPHA) * f23,
        f23 - U3 / R - C2 * Up3,
        (Ub - U4) / R + C3 * (Up5 - Up4) - TA_ALPHA * f23,
        -U5 / R + C3 * (Up4 - Up5),
    ])


def transamp1_initial(params):
    """The example's start values (stiff_transistor_amplifier.py:33-37,64-65): they satisfy
    F(0, y0, yp0) = 0 for every (Ub, R0, beta) because f(0) = 0 and Ue(0) = 0."""
    Ub = params[0]
    y0 = np.array([0.0, Ub / 2.0, Ub / 2.0, Ub, 0.0])
    yp0 = np.array([0.0, 0.0, -(Ub / 2.0) / TA_R / 2e-6, 0.0, 0.0])
    return y0, yp0


# --------------------------------------------------------------------------- #
# y' = p y^2: finite-time blow-up at t* = 1 / (p y0), the deterministic failure-path problem
# --------------------------------------------------------------------------- #
def blowup_residual(t, y, yp, p):
    return np.array([yp[0] - p * y[0] * y[0]])


def blowup_solution(t, p=1.0):
    t = np.asarray(t, dtype=float)
    return np.array([1.0 / (1.0 - p * t)]), np.array([p / (1.0 - p * t) ** 2])


def _ic_from_solution(sol, t0):
    def initial(params):
        y, yp = sol(t0, *params)
        return np.asarray(y, dtype=float).reshape(-1), np.asarray(yp, dtype=float).reshape(-1)
    return initial


PROBLEMS: dict[str, Problem] = {}


def _register(p: Problem):
    PROBLEMS[p.name] = p
    return p


ROBERTSON = _register(Problem(
    name="robertson", flops=15, n=3, n_params=3, residual=robertson_residual,
    default_params=(0.04, 1e4, 3e7), initial=robertson_initial, t_span=(0.0, 1e7),
    initial_batch=robertson_initial_batch, jacobian=robertson_jacobian,
    param_names=("k1", "k2", "k3"), events=ROBERTSON_EVENTS,
    doc="Robertson chemical kinetics as an index-1 DAE"))

AKZO_NOBEL = _register(Problem(
    name="akzo_nobel", flops=52, n=6, n_params=7, residual=akzo_residual,
    default_params=(18.7, 0.58, 0.09, 0.42, 34.4, 3.3, 115.83),
    initial=akzo_initial, t_span=(0.0, 180.0), jacobian=akzo_jacobian,
    param_names=("k1", "k2", "k3", "k4", "kbig", "kla", "ks"), events=AKZO_EVENTS,
    doc="Akzo Nobel chemical reaction, index 1"))

TRANSAMP = _register(Problem(
    name="transamp", flops=75, n=8, n_params=3, residual=transamp_residual,
    default_params=(6.0, 1000.0, 1e-6), initial=transamp_initial, t_span=(0.0, 0.2),
    jacobian=transamp_jacobian,
    param_names=("Ub", "R0", "beta"),
    doc="Two-transistor amplifier (IVP test set 'transamp'), index 1"))

ANDREWS = _register(Problem(
    name="andrews", flops=520, n=26, n_params=2, residual=andrews_residual,
    default_params=(0.033, 4530.0), initial=andrews_initial, t_span=(0.0, 0.03),
    param_names=("mom", "c0"),
    doc="Andrews' squeezer mechanism, stabilised index-1 formulation (n = 26)"))

RATIONAL = _register(Problem(
    name="rational", n=2, n_params=0, residual=rational_residual, default_params=(),
    initial=_ic_from_solution(rational_solution, 5.0), t_span=(5.0, 9.0), jacobian=rational_jacobian,
    solution=rational_solution, t0_default=5.0,
    doc="2-D rational ODE of the reference's test_integration_rational.py"))

LINEAR = _register(Problem(
    name="linear", n=2, n_params=0, residual=linear_residual, default_params=(),
    initial=_ic_from_solution(linear_solution, 0.0), t_span=(0.0, 2.0), jacobian=linear_jacobian,
    solution=linear_solution, t0_default=0.0,
    doc="linear 2-D ODE of the reference's test_integration_linear.py"))

COSINE = _register(Problem(
    name="cosine", n=1, n_params=1, residual=cosine_residual, default_params=(1.0,),
    initial=_ic_from_solution(cosine_solution, -0.1), t_span=(-0.1, -0.1 + 2 * np.pi),
    solution=cosine_solution, t0_default=-0.1, param_names=("a",), events=COSINE_EVENTS,
    doc="y' = a cos t, the problem of the reference's test_events.py"))

WEISSINGER = _register(Problem(
    name="weissinger", n=1, n_params=1, residual=weissinger_residual, default_params=(1.0,),
    initial=_ic_from_solution(weissinger_solution, np.sqrt(0.5)), t_span=(np.sqrt(0.5), 10.0),
    solution=weissinger_solution, t0_default=np.sqrt(0.5), param_names=("a",),
    doc="Weissinger's fully implicit differential equation (scaled by a)"))

BRENAN = _register(Problem(
    name="brenan", n=2, n_params=1, residual=brenan_residual, default_params=(1.0,),
    initial=_ic_from_solution(brenan_solution, 0.0), t_span=(0.0, 10.0),
    solution=brenan_solution, t0_default=0.0, param_names=("A",),
    doc="Brenan's index-1 DAE with a time-dependent leading matrix (amplitude A)"))


TRANSAMP1 = _register(Problem(
    name="transamp1", flops=40, n=5, n_params=3, residual=transamp1_residual,
    default_params=(6.0, 1000.0, 1e-6), initial=transamp1_initial, t_span=(0.0, 0.2),
    param_names=("Ub", "R0", "beta"),
    doc="One-transistor amplifier of the reference's examples/daes/stiff_transistor_amplifier.py (n = 5)"))

BLOWUP = _register(Problem(
    name="blowup", flops=3, n=1, n_params=1, residual=blowup_residual, default_params=(1.0,),
    initial=_ic_from_solution(blowup_solution, 0.0), t_span=(0.0, 2.0),
    solution=blowup_solution, t0_default=0.0, param_names=("p",),
    doc="y' = p y^2, y(0) = 1: blows up at t = 1 / p (failure-path test problem)"))


def get_problem(name_or_problem) -> Problem:
    if isinstance(name_or_problem, Problem):
        return name_or_problem
    try:
        return PROBLEMS[name_or_problem]
    except KeyError:
        raise ValueError(
            f"unknown problem {name_or_problem!r}; registered: {sorted(PROBLEMS)}") from None


def perturbed_params(problem, N: int, sigma: float, seed: int = 0) -> np.ndarray:
    """Ensemble parameters ``p_i = p_nominal * exp(sigma * u_i)``, ``u ~ U(-1, 1)``.

    Returns an ``(N, n_params)`` array (AoS on the host; the library wants SoA
    ``[n_params][N]`` and the batched entry point transposes).  SURVEY.md §8(d) C2:
    ``numpy.random.default_rng(seed)``, sigma = 0.1 headline, 0.5 divergence stress.
    """
    problem = get_problem(problem)
    rng = np.random.default_rng(seed)
    u = rng.uniform(-1.0, 1.0, size=(N, problem.n_params))
    return np.asarray(problem.default_params, dtype=float).reshape(1, problem.n_params) * np.exp(sigma * u)


def initial_conditions(problem, params: np.ndarray):
    """Consistent ``(y0[N,n], yp0[N,n])`` for an ``(N, n_params)`` parameter array."""
    problem = get_problem(problem)
    params = np.asarray(params, dtype=float)
    params = params.reshape(-1, problem.n_params) if params.ndim != 2 else params
    if problem.initial_batch is not None:
        return problem.initial_batch(params)
    y0 = np.empty((params.shape[0], problem.n))
    yp0 = np.empty_like(y0)
    for i, p in enumerate(params):
        y0[i], yp0[i] = problem.initial(p)
    return y0, yp0
