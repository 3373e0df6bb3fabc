"""The reference's own option product on the Robertson DAE (tests/test_integration_robertson.py:70-126:
BDF x NDF_strategy {stability, efficiency, None} x max_order 1..6; Radau x stages {3, 5, 7} x
continuous_error_weight {0, 0.5, 1} x newton_iter_embedded {0, 1, 2}; t in [0, 1e8], rtol = atol = 1e-6)
on the CUDA path, with the reference's own work-counter ceilings (:104-126) and accepted-step counts
within 5 % of the CPU oracle for every option tuple.  CPU twin: the oracle under the same ceilings."""
from itertools import product

import numpy as np
import pytest

BDF_MATRIX = list(product(["stability", "efficiency", None], [1, 2, 3, 4, 5, 6]))
RADAU_MATRIX = list(product([3, 5, 7], [0.0, 0.5, 1.0], [0, 1, 2]))
Y0, YP0, P = np.array([[1.0, 0.0, 0.0]]), np.array([[-0.04, 0.04, 0.0]]), np.array([[0.04, 1e4, 3e7]])
TSPAN = (0.0, 1e8)


def _ceilings(method, res, max_order=None):
    nfev, njev, nlu = int(res["nfev"][0]), int(res["njev"][0]), int(res["nlu"][0])
    if method == "BDF":
        if max_order == 1:
            assert nfev < 5000 and nlu < 1100, (nfev, nlu)
        else:
            assert nfev < 1600 and nlu < 210, (nfev, nlu)
        assert njev < 32, njev
    else:
        assert nfev < 2150 and njev < 90 and nlu < 300, (nfev, njev, nlu)


def _bdf_opts(ndf, max_order):
    return dict(max_order=max_order, NDF_strategy=ndf)


def _radau_opts(s, cew, emb):
    return dict(stages=s, continuous_error_weight=cew, newton_iter_embedded=emb)


@pytest.mark.parametrize("ndf,max_order", BDF_MATRIX)
def test_oracle_bdf_option_product(ndf, max_order, oracle):
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        r = oracle.solve_ensemble("robertson", P, Y0, YP0, TSPAN, method="BDF", rtol=1e-6, atol=1e-6, **_bdf_opts(ndf, max_order))
    assert r["status"][0] == 0
    _ceilings("BDF", r, max_order)


@pytest.mark.parametrize("s,cew,emb", RADAU_MATRIX)
def test_oracle_radau_option_product(s, cew, emb, oracle):
    r = oracle.solve_ensemble("robertson", P, Y0, YP0, TSPAN, method="Radau", rtol=1e-6, atol=1e-6, **_radau_opts(s, cew, emb))
    assert r["status"][0] == 0
    _ceilings("Radau", r)


def _cuda(method, opts):
    import warnings
    from scipy_dae_b200 import solve_dae_batched
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        r = solve_dae_batched("robertson", TSPAN, Y0, YP0, P, method=method, rtol=1e-6, atol=1e-6, **opts).numpy()
    return dict(status=r.status, nfev=r.nfev, njev=r.njev, nlu=r.nlu, n_accepted=r.n_accepted, y_final=r.y_final)


@pytest.mark.gpu
@pytest.mark.parametrize("ndf,max_order", BDF_MATRIX)
def test_cuda_bdf_option_product(ndf, max_order, oracle):
    import warnings
    opts = _bdf_opts(ndf, max_order)
    r = _cuda("BDF", opts)
    assert r["status"][0] == 0
    _ceilings("BDF", r, max_order)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        ref = oracle.solve_ensemble("robertson", P, Y0, YP0, TSPAN, method="BDF", rtol=1e-6, atol=1e-6, **opts)
    assert abs(int(r["n_accepted"][0]) - int(ref["n_accepted"][0])) <= 0.05 * ref["n_accepted"][0] + 2
    # (BDF on this problem moves by up to ~1.4 bands under a 1e-15 parameter perturbation of the
    # reference itself, tests/golden/robertson_bdf_*: twin_* arrays)
    assert (np.abs(r["y_final"] - ref["y_final"]) <= 4 * np.maximum(1e-5 * np.abs(ref["y_final"]), 1e-6)).all()


@pytest.mark.gpu
@pytest.mark.parametrize("s,cew,emb", RADAU_MATRIX)
def test_cuda_radau_option_product(s, cew, emb, oracle):
    opts = _radau_opts(s, cew, emb)
    r = _cuda("Radau", opts)
    assert r["status"][0] == 0
    _ceilings("Radau", r)
    ref = oracle.solve_ensemble("robertson", P, Y0, YP0, TSPAN, method="Radau", rtol=1e-6, atol=1e-6, **opts)
    assert abs(int(r["n_accepted"][0]) - int(ref["n_accepted"][0])) <= 0.05 * ref["n_accepted"][0] + 2
    assert (np.abs(r["y_final"] - ref["y_final"]) <= np.maximum(1e-5 * np.abs(ref["y_final"]), 1e-6)).all()
