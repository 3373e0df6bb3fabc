"""Pins the CPU oracle (oracle/dae_oracle.c) against golden vectors minted from the UNMODIFIED
Python reference (oracle/gen_golden.py).  CPU only."""
import numpy as np
import pytest

from conftest import golden_cases, load_golden
from parity import check_events, compare


@pytest.mark.parametrize("case", golden_cases())
def test_oracle_matches_reference_golden(case, oracle):
    g, meta = load_golden(case)
    te = g.get("t_eval")
    res = oracle.solve_ensemble(meta["problem"], g["params"], g["y0"], g["yp0"], meta["t_span"],
                                method=meta["method"], rtol=meta["rtol"], atol=meta["atol"],
                                t_eval=te, nthreads=4, **meta["options"])
    n = g["y0"].shape[1]
    atol_vec = np.broadcast_to(np.asarray(meta["atol"], dtype=float), (n,))
    fails = compare(res, g, meta, case, atol_vec, te is not None, eval_axes_ref="nm", eval_axes_res="mn", t_eval=te)
    assert not fails, "; ".join(fails)
    if "events" in meta["options"]:
        check_events(res["t_events"], res["y_events"], res["n_events_found"], res["t_final"], g, meta)


def test_exact_counters_on_headline_config(oracle):
    """Config 1 (Robertson, Radau s=3, rtol=atol=1e-6): the oracle reproduces the reference's
    work counters exactly (80 steps, nfev 911, njev 71, nlu 170)."""
    g, meta = load_golden("robertson_radau3_c1")
    res = oracle.solve_ensemble("robertson", g["params"], g["y0"], g["yp0"], meta["t_span"],
                                method="Radau", rtol=1e-6, atol=1e-6)
    for k in ("nfev", "njev", "nlu", "n_accepted", "n_newton_fail", "n_newton_iter", "n_err_rejected"):
        assert res[k][0] == g[k][0], k
    assert (res["nfev"][0], res["njev"][0], res["nlu"][0], res["n_accepted"][0]) == (911, 71, 170, 80)
    assert np.allclose(res["y_final"], g["y_final"], rtol=1e-7, atol=0)


def test_oracle_residual_twins():
    from oracle import oracle as O
    from scipy_dae_b200.problems import PROBLEMS
    rng = np.random.default_rng(5)
    for name, pr in PROBLEMS.items():
        p = np.array(pr.default_params) * np.exp(0.05 * rng.uniform(-1, 1, pr.n_params))
        y0, yp0 = pr.initial(p)
        y = np.abs(y0 + 0.01 * rng.standard_normal(pr.n) * np.maximum(np.abs(y0), 0.1)) if name == "akzo_nobel" \
            else y0 + 0.01 * rng.standard_normal(pr.n) * np.maximum(np.abs(y0), 0.1)
        yp = yp0 + 0.01 * rng.standard_normal(pr.n)
        a = pr.residual(0.3, y, yp, *p)
        b = O.residual(name, 0.3, y, yp, p)
        assert np.allclose(a, b, rtol=1e-12, atol=1e-14 * max(1.0, np.abs(a).max())), name
        assert np.abs(pr.residual(pr.t_span[0], y0, yp0, *p)).max() < 1e-9, name
