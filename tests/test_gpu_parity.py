"""Parity of the CUDA path (through the C-ABI, libb200dae.so) with the reference: against the
golden vectors minted from the unmodified Python reference and against the CPU oracle on the
same seeded inputs.  Bands: parity.py (north_star: states within max(10*rtol*|y|, atol),
accepted / rejected step counts within 5 %)."""
import numpy as np
import pytest
import torch

from conftest import golden_cases, load_golden
from parity import check_events, compare, counts_within_band, worst_state_violation

pytestmark = pytest.mark.gpu


def _to_dict(r):
    r = r.numpy()
    d = dict(status=r.status, y_final=r.y_final, yp_final=r.yp_final, t_final=r.t_final,
             n_accepted=r.n_accepted, n_err_rejected=r.n_err_rejected,
             n_newton_fail=r.n_newton_fail, n_newton_iter=r.n_newton_iter,
             nfev=r.nfev, njev=r.njev, nlu=r.nlu, n_eval_valid=r.n_eval_valid)
    if r.y is not None:
        d["y_eval"] = r.y      # [N, n, m]
        d["yp_eval"] = r.yp
    return d


@pytest.mark.parametrize("case", golden_cases())
def test_cuda_matches_reference_golden(case):
    from scipy_dae_b200 import solve_dae_batched
    g, meta = load_golden(case)
    te = g.get("t_eval")
    r = solve_dae_batched(meta["problem"], meta["t_span"], g["y0"], g["yp0"], g["params"],
                          method=meta["method"], rtol=meta["rtol"], atol=meta["atol"], t_eval=te,
                          **meta["options"])
    torch.cuda.synchronize()
    res = _to_dict(r)
    if "events" in meta["options"]:
        rn = r.numpy()
        check_events(rn.t_events, rn.y_events, rn.n_events_found, rn.t_final, g, meta)
    n = g["y0"].shape[1]
    atol_vec = np.broadcast_to(np.asarray(meta["atol"], dtype=float), (n,))
    fails = compare(res, g, meta, case, atol_vec, te is not None, eval_axes_ref="nm", eval_axes_res="nm", t_eval=te)
    assert not fails, "; ".join(fails)
    if "events" not in meta["options"]:
        done = g["status"] == 0
        assert np.array_equal(res["t_final"][done], g["t_final"][done])
    # work counters are not part of the north_star contract (only accepted / rejected steps are);
    # they are chaotic at tight tolerances (the oracle itself differs from the golden by 9 % in nlu
    # on robertson_radau3_tight), so they only get a coarse sanity band
    fin = (g["status"] == 0) & (res["status"] == 0)         # (where an instance gives up is chaotic, parity.py)
    for k in ("nfev", "njev", "nlu"):
        assert counts_within_band(res[k][fin], g[k][fin], band=0.25, slack=4).all(), (k, res[k], g[k])


def test_headline_counters_match_reference():
    """Config 1: the reference's step sequence (80 accepted, 5 Newton failures, nfev 911, njev 71,
    nlu 170, 277 Newton iterations).  The control flow is the reference's, but the
    FD-Jacobian noise depends on the last bits of every residual, so a recompilation (other FMA
    contractions) moves the counters by a few evaluations: builds of this source have given nfev
    911, 914 and 908.  The bar is 1 % on the work counters, the exact number of
    accepted steps and Newton failures, and the drift bands on the accepted times below."""
    from scipy_dae_b200 import solve_dae_batched
    g, meta = load_golden("robertson_radau3_c1")
    lean = solve_dae_batched("robertson", meta["t_span"], g["y0"], g["yp0"], g["params"],
                             method="Radau", rtol=1e-6, atol=1e-6).numpy()
    r = solve_dae_batched("robertson", meta["t_span"], g["y0"], g["yp0"], g["params"],
                          method="Radau", rtol=1e-6, atol=1e-6, record_steps=128).numpy()
    for res in (lean, r):            # the untraced kernel bench.py times, and the traced variant
        assert (res.n_accepted[0], res.n_newton_fail[0], res.n_err_rejected[0]) == (80, 5, 0)
        for got, ref in zip((res.nfev[0], res.njev[0], res.nlu[0], res.n_newton_iter[0]), (911, 71, 170, 277)):
            assert abs(int(got) - ref) <= max(2, 0.01 * ref), (got, ref)
    k = r.n_steps[0]
    assert k == len(g["t_steps"])
    # same control flow; the step SIZES drift apart through rounding: every FD-Jacobian refresh
    # amplifies a 1e-16 difference in the state by eps*|F|/h_fd, the discrete num_jac decisions
    # flip on it, and the controller compounds it.  Measured drift of the accepted times for three
    # builds of this source (scripts/step_drift.py on B200): <= 2e-12 for steps 1-7, <= 2e-6 up
    # to step 16, then jumps to 1e-5 .. 3e-2 at build-dependent places (the last steps are 1e6 long).
    assert np.allclose(r.t_steps[0, :8], g["t_steps"][:8], rtol=1e-7, atol=0)
    assert np.allclose(r.t_steps[0, :16], g["t_steps"][:16], rtol=1e-5, atol=0)
    assert np.allclose(r.t_steps[0, :k], g["t_steps"], rtol=5e-2, atol=0)
    assert np.allclose(r.y_steps[0, :, :16], g["y_steps"][:, :16], rtol=1e-4, atol=1e-9)
    assert r.t_steps[0, k - 1] == g["t_steps"][-1]


@pytest.mark.parametrize("problem,method,N,sigma,rtol,atol,opts", [
    ("robertson", "Radau", 2048, 0.1, 1e-6, 1e-6, {}),
    ("robertson", "Radau", 512, 0.5, 1e-6, 1e-6, {}),
    ("robertson", "Radau", 256, 0.1, 1e-6, 1e-6, {"stages": 5}),
    ("robertson", "BDF", 512, 0.1, 1e-6, 1e-6, {}),
    ("akzo_nobel", "BDF", 256, 0.05, 1e-8, 1e-8, {}),
    ("akzo_nobel", "Radau", 128, 0.05, 1e-8, 1e-8, {}),
    ("transamp", "Radau", 32, 0.02, 1e-10, 1e-6, {}),
    ("transamp", "BDF", 32, 0.02, 1e-10, 1e-6, {}),
    ("andrews", "Radau", 16, 0.05, 1e-4, 1e-4, {"stages": 5}),
    ("andrews", "Radau", 16, 0.05, 1e-6, 1e-6, {"stages": 3}),
])
def test_cuda_matches_oracle_on_seeded_ensembles(problem, method, N, sigma, rtol, atol, opts, oracle):
    """Larger seeded ensembles than the goldens hold: CUDA vs the CPU oracle, instance by instance."""
    from scipy_dae_b200 import solve_dae_batched, perturbed_params, initial_conditions, get_problem
    pr = get_problem(problem)
    params = perturbed_params(problem, N, sigma, seed=7)
    y0, yp0 = initial_conditions(problem, params)
    r = solve_dae_batched(problem, pr.t_span, y0, yp0, params, method=method, rtol=rtol, atol=atol, **opts)
    res = _to_dict(r)
    ref = oracle.solve_ensemble(problem, params, y0, yp0, pr.t_span, method=method, rtol=rtol,
                                atol=atol, nthreads=8, **opts)
    assert np.array_equal(res["status"], ref["status"])
    ok = ref["status"] == 0
    assert ok.mean() > 0.9
    ratio = np.abs(res["y_final"] - ref["y_final"]) / np.maximum(10 * rtol * np.abs(ref["y_final"]), atol)
    ratio = ratio[ok].max(axis=1)
    # chaotic cases (parity.py STATE_SLACK): BDF at loose rtol and the Newton-failure-heavy circuit
    slack = 4.0 if (method == "BDF" or problem == "transamp") else (10.0 if problem == "andrews" else 1.0)
    assert (ratio <= slack).mean() >= 0.99, f"{(ratio > slack).sum()} of {ok.sum()} outside band, worst {ratio.max():.3g}"
    assert np.median(ratio) < 0.5
    if problem == "andrews":
        assert (ratio <= 1.0).mean() >= 0.8
    # Andrews: the reference's own step count moves by up to 22 % under a 1e-15 parameter
    # perturbation (parity.py STEP_BAND_CASE); per-instance band widened, ensemble sums below still 2-5 %
    step_band = 0.30 if problem == "andrews" else 0.05
    acc_ok = counts_within_band(res["n_accepted"], ref["n_accepted"], band=step_band)
    assert acc_ok.mean() >= 0.99, (res["n_accepted"][~acc_ok], ref["n_accepted"][~acc_ok])
    # ensemble-level work agrees to 3 % (builds of this source have differed by up to 2.5 % in nlu
    # on Akzo Nobel: every LU set follows a discrete controller decision)
    for k in ("n_accepted", "nfev", "nlu", "n_newton_iter"):
        a, b = res[k][ok].sum(), ref[k][ok].sum()
        assert abs(a - b) <= (0.06 if problem == "andrews" else 0.03) * b, (k, a, b)


def test_device_functors_match_numpy_twins():
    from scipy_dae_b200 import device_residual, PROBLEMS
    rng = np.random.default_rng(11)
    for name, pr in PROBLEMS.items():
        N = 64
        p = np.array(pr.default_params)[None, :] * np.exp(0.05 * rng.uniform(-1, 1, (N, pr.n_params)))
        y = np.empty((N, pr.n)); yp = np.empty((N, pr.n))
        for i in range(N):
            y0, yp0 = pr.initial(p[i])
            y[i] = y0 + 0.01 * rng.standard_normal(pr.n) * np.maximum(np.abs(y0), 0.1)
            yp[i] = yp0 + 0.01 * rng.standard_normal(pr.n)
        if name == "akzo_nobel":
            y = np.abs(y)
        out = device_residual(name, 0.3, y, yp, p).cpu().numpy()
        ref = np.stack([pr.residual(0.3, y[i], yp[i], *p[i]) for i in range(N)])
        assert np.allclose(out, ref, rtol=1e-11, atol=1e-13 * max(1.0, np.abs(ref).max())), name


def test_edge_cases():
    from scipy_dae_b200 import solve_dae_batched
    # t0 == t_bound: finishes immediately (base.py:376-381)
    r = solve_dae_batched("robertson", (0.0, 0.0), rtol=1e-6, atol=1e-6).numpy()
    assert r.status[0] == 0 and r.n_accepted[0] == 0 and np.allclose(r.y_final[0], [1, 0, 0])
    # ragged batch sizes around the block size
    for N in (1, 31, 129):
        p = np.tile([0.04, 1e4, 3e7], (N, 1))
        r = solve_dae_batched("robertson", (0.0, 1e2), params=p, rtol=1e-6, atol=1e-6).numpy()
        assert (r.status == 0).all() and np.ptp(r.y_final, axis=0).max() == 0.0
    # max_steps cap -> status -3, t_final < t_bound
    r = solve_dae_batched("robertson", (0.0, 1e7), rtol=1e-6, atol=1e-6, max_steps=10).numpy()
    assert r.status[0] == -3 and r.n_accepted[0] == 10 and r.t_final[0] < 1e7
    # a NaN parameter poisons only its own instance
    p = np.tile([0.04, 1e4, 3e7], (64, 1)); p[5, 1] = np.nan
    y0 = np.tile([1.0, 0, 0], (64, 1)); yp0 = np.tile([-0.04, 0.04, 0], (64, 1))
    r = solve_dae_batched("robertson", (0.0, 1e3), y0, yp0, p, rtol=1e-6, atol=1e-6).numpy()
    assert r.status[5] < 0 and (np.delete(r.status, 5) == 0).all()
    # argument errors mirror the reference's messages
    with pytest.raises(ValueError, match="not properly sorted"):
        solve_dae_batched("robertson", (0.0, 1.0), t_eval=[0.5, 0.2], rtol=1e-6, atol=1e-6)
    with pytest.raises(ValueError, match="not within"):
        solve_dae_batched("robertson", (0.0, 1.0), t_eval=[0.5, 2.0], rtol=1e-6, atol=1e-6)
    with pytest.raises(ValueError, match="must be finite"):
        solve_dae_batched("robertson", (0.0, 1.0), [[np.inf, 0, 0]], [[0.0, 0, 0]], rtol=1e-6, atol=1e-6)


def test_full_size_properties(oracle):
    """The bench ensemble itself (config 2: 2^20 Robertson instances, seed 0): size-independent
    properties (mass conservation of the algebraic constraint, all finished, placement
    independence) and the SURVEY.md §8d parity sample (the first 4 096 instances and 4 096 random
    ones against the oracle)."""
    from scipy_dae_b200 import solve_dae_batched, perturbed_params, initial_conditions
    N = 1 << 20
    params = perturbed_params("robertson", N, 0.1, seed=0)
    y0, yp0 = initial_conditions("robertson", params)
    r = solve_dae_batched("robertson", (0.0, 1e7), y0, yp0, params, rtol=1e-6, atol=1e-6)
    torch.cuda.synchronize()
    assert bool((r.status == 0).all())
    idx = np.concatenate([np.arange(4096), np.random.default_rng(123).choice(N, 4096, replace=False)])
    ref = oracle.solve_ensemble("robertson", params[idx], y0[idx], yp0[idx], (0.0, 1e7), method="Radau",
                                rtol=1e-6, atol=1e-6, nthreads=8)
    yg = r.y_final[torch.from_numpy(idx).to(r.y_final.device)].cpu().numpy()
    ratio = np.abs(yg - ref["y_final"]) / np.maximum(1e-5 * np.abs(ref["y_final"]), 1e-6)
    assert (ref["status"] == 0).all() and ratio.max() <= 1.0, ratio.max()
    acc_g = r.n_accepted[torch.from_numpy(idx).to(r.n_accepted.device)].cpu().numpy()
    assert (np.abs(acc_g - ref["n_accepted"]) <= 0.05 * ref["n_accepted"] + 2).all()
    y = r.y_final
    assert float((y.sum(dim=1) - 1.0).abs().max()) < 1e-5          # y1 + y2 + y3 = 1 (algebraic row)
    assert float(y.min()) > -1e-6
    acc = r.n_accepted
    assert 70 <= int(acc.min()) and int(acc.max()) <= 95            # 78..81 at sigma = 0.1
    # identical parameters -> bit-identical results wherever they sit in the batch
    params2 = params.copy(); params2[N // 2 + 3] = params[17]
    y0b, yp0b = initial_conditions("robertson", params2)
    r2 = solve_dae_batched("robertson", (0.0, 1e7), y0b, yp0b, params2, rtol=1e-6, atol=1e-6)
    assert torch.equal(r2.y_final[N // 2 + 3], r.y_final[17])


def test_host_pipeline_matches_device_call():
    """solve_dae_batched_host (chunked, multi-stream, pinned buffers) returns bit-identical
    per-instance results to one solve_dae_batched call, for ragged chunkings too."""
    from scipy_dae_b200 import (solve_dae_batched, solve_dae_batched_host, alloc_host_result,
                                perturbed_params, initial_conditions)
    N = 5000
    params = perturbed_params("robertson", N, 0.1, seed=5)
    y0, yp0 = initial_conditions("robertson", params)
    ref = solve_dae_batched("robertson", (0.0, 1e7), y0, yp0, params, rtol=1e-6, atol=1e-6).numpy()
    out = alloc_host_result(N, 3)
    for chunks in (1, 3, 7):
        r = solve_dae_batched_host("robertson", (0.0, 1e7), y0, yp0, params, rtol=1e-6, atol=1e-6,
                                   chunks=chunks, out=out)
        assert np.array_equal(r.y_final.numpy(), ref.y_final)
        assert np.array_equal(r.yp_final.numpy(), ref.yp_final)
        assert np.array_equal(r.status.numpy(), ref.status)
        assert np.array_equal(r.nfev.numpy(), ref.nfev) and np.array_equal(r.n_accepted.numpy(), ref.n_accepted)
    bad = y0.copy(); bad[17, 1] = np.inf
    with pytest.raises(ValueError, match="must be finite"):
        solve_dae_batched_host("robertson", (0.0, 1.0), bad, yp0, params, rtol=1e-6, atol=1e-6, chunks=4)
    # CUDA-graph replay of the same pipeline: pinned inputs, new data in the same buffers
    hp, hy, hyp = (torch.from_numpy(a).pin_memory() for a in (params, y0, yp0))
    for rep in range(3):
        if rep == 2:                       # the graph reads the buffers at replay time
            hp.copy_(torch.from_numpy(perturbed_params("robertson", N, 0.1, seed=6)))
            ref = solve_dae_batched("robertson", (0.0, 1e7), hy, hyp, hp, rtol=1e-6, atol=1e-6).numpy()
        r = solve_dae_batched_host("robertson", (0.0, 1e7), hy, hyp, hp, rtol=1e-6, atol=1e-6,
                                   chunks=4, out=out, use_graph=True)
        assert r.extra["graph"] and np.array_equal(r.y_final.numpy(), ref.y_final)
        assert np.array_equal(r.n_accepted.numpy(), ref.n_accepted)


def test_group_and_thread_regimes_agree():
    """n > 4 systems run on the cooperative group-per-system kernels by default; the
    one-thread-per-system kernels (B200DAE_REGIME=thread, read once per process) implement the
    same arithmetic, so a subprocess run must agree to rounding."""
    import subprocess, sys, os, json
    code = (
        "import json,sys; sys.path.insert(0, %r);"
        "from scipy_dae_b200 import solve_dae_batched, perturbed_params, initial_conditions, get_problem;"
        "pr=get_problem('transamp'); p=perturbed_params('transamp',16,0.02,seed=2); y0,yp0=initial_conditions('transamp',p);"
        "r=solve_dae_batched('transamp',pr.t_span,y0,yp0,p,method='Radau',rtol=1e-10,atol=1e-6).numpy();"
        "print(json.dumps(dict(y=r.y_final.tolist(), acc=r.n_accepted.tolist(), st=r.status.tolist())))"
    ) % os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    outs = []
    for regime in ("group", "thread"):
        env = dict(os.environ, B200DAE_REGIME=regime)
        o = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, check=True)
        outs.append(json.loads(o.stdout.strip().splitlines()[-1]))
    a, b = outs
    assert a["st"] == b["st"] == [0] * 16
    ya, yb = np.array(a["y"]), np.array(b["y"])
    assert (np.abs(ya - yb) <= 4.0 * np.maximum(10 * 1e-10 * np.abs(yb), 1e-6)).all()
    assert (np.abs(np.array(a["acc"]) - np.array(b["acc"])) <= 0.05 * np.array(b["acc"]) + 2).all()


def test_single_system_surface_events_and_dense_output():
    """scipy_dae_b200.integrate.solve_dae: the reference's single-system signature with events
    (ThresholdEvent) and dense_output=True (sol(t) from the dense-output trace of the same launch)."""
    from scipy_dae_b200 import ThresholdEvent, _lib
    from scipy_dae_b200.integrate import solve_dae
    ev = [ThresholdEvent(component=0, value=0.9), ThresholdEvent(component=2, value=0.5, terminal=True)]
    n0 = _lib.launch_count()
    r = solve_dae("robertson", (0.0, 1e7), method="Radau", rtol=1e-6, atol=1e-6, events=ev, dense_output=True)
    assert r.status == 1 and r.message == "A termination event occurred."
    assert len(r.t_events[0]) == 1 and len(r.t_events[1]) == 1
    assert abs(r.y_events[0][0][0] - 0.9) < 1e-9 and abs(r.y_events[1][0][2] - 0.5) < 1e-9
    assert r.t[-1] == r.t_events[1][0]                       # the last point is the terminal event
    # sol(t) reproduces the step ends and interpolates in between
    y, yp = r.sol(r.t[5])
    assert np.allclose(y, r.y[:, 5], rtol=1e-12, atol=1e-15)
    tt = np.array([r.t[3], 0.5 * (r.t[3] + r.t[4]), r.t[4]])
    Y, Yp = r.sol(tt[[2, 0, 1]])                             # any order, repeated times allowed
    assert Y.shape == (3, 3) and np.allclose(Y[:, 1], r.y[:, 3], rtol=1e-12, atol=1e-15)
    assert (np.abs(Y.sum(axis=0) - 1.0) < 1e-6).all()
    assert _lib.launch_count() - n0 == 1                     # sol(t) costs no launch
    # the last segment ends at the terminal event: its interpolant is the full step's, cut short
    ye, _ = r.sol(r.t_events[1][0])
    assert abs(ye[2] - 0.5) < 1e-9
    # BDF: the Newton-form interpolant of every step
    rb = solve_dae("robertson", (0.0, 1e5), method="BDF", rtol=1e-6, atol=1e-6, dense_output=True)
    yb, ypb = rb.sol(rb.t[7])
    assert np.allclose(yb, rb.y[:, 7], rtol=1e-12, atol=1e-15)
    tm = 0.5 * (rb.t[20] + rb.t[21])
    assert abs(rb.sol(tm)[0].sum() - 1.0) < 1e-6
