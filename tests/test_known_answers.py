"""The reference's own known-answer tests restated on this repository's surfaces:
tests/test_integration_rational.py, test_integration_linear.py, test_events.py of the reference,
plus Weissinger's fully implicit equation and Brenan's index-1 DAE from its benchmarks.  The
exact solutions are the oracle here, so these tests need neither the reference nor goldens.

CPU: the C oracle against the exact solutions.  GPU: the CUDA path through the single-system
surface `scipy_dae_b200.integrate.solve_dae` (same calls as the reference's tests) and through
parameter-perturbed ensembles."""
import numpy as np
import pytest

from scipy_dae_b200.problems import get_problem


def compute_error(y, y_true, rtol, atol):           # scipy test_ivp.compute_error
    e = (y - y_true) / (atol + rtol * np.abs(y_true))
    return np.linalg.norm(e, axis=0) / np.sqrt(e.shape[0])


EVENT_CASES = [   # tests/test_events.py:28-84: (events, t_events, y_events, yp_events); event = y or y' crossing 0
    ([dict(component=0, of="y")], [[0, np.pi]], [[0, 0]], [[1, -1]]),
    ([dict(component=0, of="yp")], [[np.pi / 2, 3 * np.pi / 2]], [[1, -1]], [[0, 0]]),
    ([dict(component=0, of="y"), dict(component=0, of="yp")],
     [[0, np.pi], [np.pi / 2, 3 * np.pi / 2]], [[0, 0], [1, -1]], [[1, -1], [0, 0]]),
    ([dict(component=0, of="y", terminal=True)], [[0]], [[0]], [[1]]),
    ([dict(component=0, of="yp", terminal=True)], [[np.pi / 2]], [[1]], [[0]]),
]


# ------------------------------------------------------------------------------------------------
# CPU: oracle vs exact solutions
@pytest.mark.parametrize("method", ["BDF", "Radau"])
def test_oracle_functor_events_known_answers(method, oracle):
    """Event functions of the problem functor (not thresholds on one component): y - y' and y - t / 2."""
    pr = get_problem("cosine")
    y0, yp0 = pr.initial((1.0,))
    r = oracle.solve_ensemble("cosine", np.array([[1.0]]), y0[None, :], yp0[None, :], pr.t_span, method=method,
                              rtol=1e-12, atol=1e-12, events=[dict(functor=0), dict(functor=1, direction=-1, terminal=True)])
    assert r["status"][0] == 1 and r["n_events_found"][0].tolist() == [1, 1]
    assert abs(r["t_events"][0, 0, 0] - np.pi / 4) < 1e-7 and abs(r["t_events"][0, 1, 0] - 1.8954942670339809) < 1e-7



# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("method", ["Radau", "BDF"])
@pytest.mark.parametrize("name,t_span", [("rational", (5.0, 9.0)), ("rational", (5.0, 1.0)), ("linear", (0.0, 2.0)),
                                         ("weissinger", None), ("brenan", None), ("cosine", None)])
def test_oracle_known_answers(name, t_span, method, oracle):
    pr = get_problem(name)
    t_span = t_span or pr.t_span
    p = np.array(pr.default_params, dtype=float)
    y0, yp0 = pr.initial(p)
    # the reference's own bars for its two test problems (error < 6 at rtol 1e-3); for the
    # benchmark problems the global error over the longer intervals is 1 .. 27 tolerance units at
    # rtol 1e-6 (measured with the oracle), bar 60
    rtol, atol, bar = (1e-3, 1e-6, 6) if name in ("rational", "linear") else (1e-6, 1e-9, 60)
    te = np.linspace(*t_span, 12)[1:]
    r = oracle.solve_ensemble(name, p[None, :], y0[None, :], yp0[None, :], t_span, method=method, rtol=rtol,
                              atol=atol, t_eval=te)
    assert r["status"][0] == 0
    y_true, _ = pr.solution(te, *p)
    assert np.all(compute_error(r["y_eval"][0].T, y_true, rtol, atol) < bar)


@pytest.mark.parametrize("method", ["Radau", "BDF"])
@pytest.mark.parametrize("case", range(len(EVENT_CASES)))
def test_oracle_events_known_answers(case, method, oracle):
    events, t_ev, y_ev, yp_ev = EVENT_CASES[case]
    pr = get_problem("cosine")
    y0, yp0 = pr.initial((1.0,))
    r = oracle.solve_ensemble("cosine", np.array([[1.0]]), y0[None, :], yp0[None, :], pr.t_span, method=method,
                              rtol=1e-12, atol=1e-12, events=[dict(value=0.0, **e) for e in events], event_cap=4)
    for k in range(len(events)):
        m = len(t_ev[k])
        assert r["n_events_found"][0, k] == m
        assert np.allclose(r["t_events"][0, k, :m], t_ev[k], rtol=1e-7, atol=1e-7)
        assert np.allclose(r["y_events"][0, k, :m, 0], y_ev[k], rtol=1e-7, atol=1e-7)
        assert np.allclose(r["yp_events"][0, k, :m, 0], yp_ev[k], rtol=1e-7, atol=1e-7)


# ------------------------------------------------------------------------------------------------
# GPU: the CUDA path, written like the reference's tests
# ------------------------------------------------------------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("first_step", [None, 0.1])
@pytest.mark.parametrize("method", ["BDF", "Radau"])
@pytest.mark.parametrize("t_span", [[5, 9], [5, 1]])
@pytest.mark.parametrize("jac", [None, "analytic"])
def test_integration_rational(first_step, method, t_span, jac):
    """tests/test_integration_rational.py:47-95 (vectorized / sparse variants have no device analogue)."""
    from scipy_dae_b200.integrate import solve_dae
    rtol, atol = 1e-3, 1e-6
    pr = get_problem("rational")
    y0 = [1 / 3, 2 / 9]
    yp0 = [y0[1] / 5, y0[1] * (y0[0] + 2 * y0[1] - 1) / (5 * (y0[0] - 1))]
    res = solve_dae("rational", t_span, y0, yp0, rtol=rtol, atol=atol, method=method, dense_output=True,
                    jac=jac, first_step=first_step)
    assert res.t[0] == t_span[0] and res.t_events is None and res.y_events is None
    assert res.success and res.status == 0
    if method == "BDF":
        assert 0 < res.nfev < 31 and 0 < res.njev < 3 and 0 < res.nlu < 7
    else:
        assert 0 < res.nfev < 46 and 0 < res.njev < 3 and 0 < res.nlu < 13
    assert np.all(compute_error(res.y, pr.solution(res.t)[0], rtol, atol) < 6)
    tc = np.linspace(*t_span)
    assert np.all(compute_error(res.sol(tc)[0], pr.solution(tc)[0], rtol, atol) < 6)
    tc = (t_span[0] + t_span[-1]) / 2
    assert np.all(compute_error(res.sol(tc)[0][:, None], pr.solution(tc)[0][:, None], rtol, atol) < 5)
    assert np.allclose(res.sol(res.t)[0], res.y, rtol=1e-13, atol=1e-15)


@pytest.mark.gpu
@pytest.mark.parametrize("method", ["BDF", "Radau"])
@pytest.mark.parametrize("jac", [None, "analytic"])
def test_integration_linear(method, jac):
    """tests/test_integration_linear.py:30-48."""
    from scipy_dae_b200.integrate import solve_dae
    rtol, atol = 1e-3, 1e-6
    pr = get_problem("linear")
    res = solve_dae("linear", [0, 2], [0, 2], [-10, 2], rtol=rtol, atol=atol, method=method, dense_output=True, jac=jac)
    assert res.t[0] == 0 and res.t_events is None and res.success and res.status == 0
    assert np.all(compute_error(res.y, pr.solution(res.t)[0], rtol, atol) < 5)


@pytest.mark.gpu
@pytest.mark.parametrize("method", ["BDF", "Radau"])
@pytest.mark.parametrize("case", range(len(EVENT_CASES)))
def test_events(case, method):
    """tests/test_events.py:91-105: y' = cos t, rtol = atol = 1e-12, events y = 0 / y' = 0."""
    from scipy_dae_b200 import ThresholdEvent
    from scipy_dae_b200.integrate import solve_dae
    events, t_ev, y_ev, yp_ev = EVENT_CASES[case]
    pr = get_problem("cosine")
    y0, yp0 = pr.initial((1.0,))
    sol = solve_dae("cosine", pr.t_span, y0, yp0, method=method, rtol=1e-12, atol=1e-12,
                    events=[ThresholdEvent(value=0.0, **e) for e in events])
    for k in range(len(events)):
        assert np.allclose(sol.t_events[k], t_ev[k], rtol=1e-7, atol=1e-7)
        assert np.allclose(np.ravel(sol.y_events[k]), y_ev[k], rtol=1e-7, atol=1e-7)
        assert np.allclose(np.ravel(sol.yp_events[k]), yp_ev[k], rtol=1e-7, atol=1e-7)


@pytest.mark.gpu
@pytest.mark.parametrize("method", ["BDF", "Radau"])
def test_functor_events(method):
    """tests/test_events.py restated with event functions that are NOT thresholds on one component:
    the device functor's own g(t, y, y'; p) (FunctorEvent): y - y' = 0 at t = pi/4, 5pi/4 and, terminal,
    y - t/2 = 0 (sin t = t / 2 at t = 1.895494267...)."""
    from scipy_dae_b200 import FunctorEvent
    from scipy_dae_b200.integrate import solve_dae
    pr = get_problem("cosine")
    y0, yp0 = pr.initial((1.0,))
    sol = solve_dae("cosine", pr.t_span, y0, yp0, method=method, rtol=1e-12, atol=1e-12, events=[FunctorEvent(0)])
    assert np.allclose(sol.t_events[0], [np.pi / 4, 5 * np.pi / 4], rtol=1e-7, atol=1e-7)
    assert np.allclose(np.ravel(sol.y_events[0]), np.sin(sol.t_events[0]), rtol=1e-7, atol=1e-7)
    assert np.allclose(np.ravel(sol.y_events[0]), np.ravel(sol.yp_events[0]), rtol=1e-7, atol=1e-7)
    sol = solve_dae("cosine", pr.t_span, y0, yp0, method=method, rtol=1e-12, atol=1e-12,
                    events=[FunctorEvent(0), FunctorEvent(1, direction=-1, terminal=True)])
    t_star = 1.8954942670339809
    assert sol.status == 1 and np.allclose(sol.t_events[1], [t_star], rtol=1e-7, atol=1e-7) and sol.t[-1] == sol.t_events[1][0]
    assert np.allclose(sol.t_events[0], [np.pi / 4], rtol=1e-7, atol=1e-7)          # 5 pi / 4 lies beyond the stop
    with pytest.raises(ValueError, match="event functions"):
        solve_dae("cosine", pr.t_span, y0, yp0, method=method, rtol=1e-6, atol=1e-6, events=[FunctorEvent(5)])


@pytest.mark.gpu
@pytest.mark.parametrize("method", ["Radau", "BDF"])
@pytest.mark.parametrize("name,sigma", [("weissinger", 0.3), ("brenan", 0.3), ("cosine", 0.3)])
def test_ensembles_against_exact_solutions(name, sigma, method):
    """Parameter-perturbed ensembles whose exact solutions are known: the fully implicit
    Weissinger equation (dF/dy' depends on y and y') and Brenan's DAE (dF/dy' depends on t)."""
    from scipy_dae_b200 import solve_dae_batched, perturbed_params, initial_conditions
    pr = get_problem(name)
    N = 4096
    P = perturbed_params(name, N, sigma, seed=4)
    y0, yp0 = initial_conditions(name, P)
    rtol, atol = 1e-6, 1e-9
    te = np.linspace(*pr.t_span, 9)[1:]
    r = solve_dae_batched(name, pr.t_span, y0, yp0, P, method=method, rtol=rtol, atol=atol, t_eval=te).numpy()
    # Brenan's iteration matrix turns singular for isolated (t, h) pairs: the reference method itself
    # gives up there ("step size too small": 4-6 of these 4096 instances in the CPU oracle, 1-3 on
    # the device, near t = 4.914 and t = 7.979), so a failure rate is checked, not zero failures
    ok = r.status == 0
    assert ok.mean() > 0.995 and (r.status[~ok] == -1).all()
    for i in np.nonzero(ok)[0][[0, 1, ok.sum() // 2, -1]]:
        y_true, yp_true = pr.solution(te, *P[i])
        assert np.all(compute_error(r.y[i], y_true, rtol, atol) < 60)       # global error in tolerance units
    idx = np.nonzero(ok)[0][::64]
    y_true_all = np.stack([pr.solution(pr.t_span[1], *P[i])[0] for i in idx]).reshape(-1, pr.n)
    e = np.abs(r.y_final[idx] - y_true_all) / (atol + rtol * np.abs(y_true_all))
    assert e.max() < 120
