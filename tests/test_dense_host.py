"""Host-side interpolants (scipy_dae_b200/dense.py) against the reference's own RadauDenseOutput /
BdfDenseOutput / DaeSolution objects (radau.py:672-699, bdf.py:422-548, common.py:9-126) built from
the same data.  CPU; the reference comes from oracle/_ref or /root/reference."""
import numpy as np
import pytest

from oracle import reference as REF
from scipy_dae_b200.dense import BdfStepInterpolant, DaeSolution, RadauStepInterpolant, step_interpolants

pytestmark = pytest.mark.skipif(not REF.available(), reason="oracle/_ref is missing: run oracle/make_ref.sh")


def _segments(rng, ascending):
    ts = np.cumsum(rng.uniform(0.1, 1.0, 6))
    return ts if ascending else -ts


@pytest.mark.parametrize("ascending", [True, False])
def test_radau_interpolants_and_solution(ascending):
    _, R, _, _ = REF.import_reference()
    import scipy_dae.integrate._dae.common as C
    rng = np.random.default_rng(3)
    n, s = 4, 3
    ts = _segments(rng, ascending)
    ys = rng.standard_normal((n, ts.size))
    ZPs = rng.standard_normal((ts.size, n, s))
    mine, ref = [], []
    for k in range(1, ts.size):
        mine.append(RadauStepInterpolant(ts[k - 1], ts[k], ts[k] - ts[k - 1], ys[:, k - 1], ZPs[k]))
        ref.append(R.RadauDenseOutput(ts[k - 1], ts[k], ys[:, k - 1], ZPs[k], ZPs[k]))
    tq = np.concatenate([rng.uniform(ts.min(), ts.max(), 40), ts, [ts.min() - 0.3, ts.max() + 0.2]])
    for a, b in zip(mine, ref):
        ya, ypa = a(tq[:5]); yb, ypb = b(tq[:5])
        assert np.allclose(ya, yb, rtol=1e-14, atol=0) and np.allclose(ypa, ypb, rtol=1e-14, atol=0)
        assert np.allclose(a(tq[0])[0], b(np.asarray(tq[0]))[0], rtol=1e-14, atol=0)
    A, B = DaeSolution(ts, mine), C.DaeSolution(ts, ref)
    ya, ypa = A(tq); yb, ypb = B(tq)
    assert np.allclose(ya, yb, rtol=1e-13, atol=0) and np.allclose(ypa, ypb, rtol=1e-13, atol=0)   # same segments, incl. breakpoints
    assert np.allclose(A(ts[2])[0], B(ts[2])[0], rtol=1e-13, atol=0)
    assert (A.t_min, A.t_max) == (B.t_min, B.t_max)


@pytest.mark.parametrize("ascending", [True, False])
def test_bdf_interpolants_and_solution(ascending):
    _, _, Bm, _ = REF.import_reference()
    import scipy_dae.integrate._dae.common as C
    rng = np.random.default_rng(4)
    n = 3
    ts = _segments(rng, ascending)
    mine, ref = [], []
    K = ts.size
    t_anchor, hs, orders = np.zeros(K), np.zeros(K), np.zeros(K, dtype=int)
    coef = np.full((K, 7 * n), np.nan)
    ysteps = rng.standard_normal((n, K))
    for k in range(1, K):
        order = int(rng.integers(1, 6))
        h = (ts[k] - ts[k - 1]) * rng.uniform(0.5, 2.0)            # the rescaled h of the NEXT step
        D = rng.standard_normal((order + 1, n))
        t_anchor[k], hs[k], orders[k] = ts[k], h, order
        coef[k, :(order + 1) * n] = D.reshape(-1)
        ref.append(Bm.BdfDenseOutput(ts[k - 1], ts[k], h, order, D))
    mine = step_interpolants("BDF", 0, ts, ysteps, hs, t_anchor, orders, coef)
    tq = np.concatenate([rng.uniform(ts.min(), ts.max(), 40), ts])
    ya, ypa = DaeSolution(ts, mine)(tq); yb, ypb = C.DaeSolution(ts, ref)(tq)
    assert np.allclose(ya, yb, rtol=1e-12, atol=1e-13) and np.allclose(ypa, ypb, rtol=1e-12, atol=1e-12)
    one = BdfStepInterpolant(ts[0], ts[1], t_anchor[1], hs[1], orders[1], coef[1].reshape(-1, n)[:orders[1] + 1])
    assert np.allclose(one(tq[3])[1], ref[0](np.asarray(tq[3]))[1], rtol=1e-12, atol=1e-13)
