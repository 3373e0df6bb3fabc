"""CUDA path against the LIVE, unmodified Python reference on the GPU box.

oracle/make_ref.sh copies the reference's eight hot-path files byte for byte into oracle/_ref/
(git-ignored, travels to the GPU box, SHA-256 sums verified at import).  Here the reference's own
solve_dae integrates, on a multiprocessing pool over the host cores, the same seeded ensembles the
CUDA path integrates in one launch: >= 1024 instances of config 2 and 128 instances each of
configs 3, 4 and 5 (BASELINE.json), per instance.

Bands (tests/parity.py): the north_star band, or twice the distance of the reference from its own
TWIN (the same ensemble with every parameter multiplied by 1 + 1e-15, also integrated live),
whichever is larger, quantile by quantile.  For config 2 the twin is not used: the unwidened band
has to hold for every instance."""
import os

import numpy as np
import pytest

from oracle import reference as REF
from parity import STEP_BAND, instance_ratios, quantile_check

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(not REF.available(), reason="oracle/_ref is missing: run oracle/make_ref.sh")]

CASES = [
    # name, problem, method, N, sigma, seed, rtol, atol, options, use_twin
    ("C2 robertson radau3", "robertson", "Radau", 1024, 0.1, 0, 1e-6, 1e-6, {}, False),
    ("C3 akzo bdf", "akzo_nobel", "BDF", 128, 0.05, 0, 1e-8, 1e-8, {}, True),
    ("C3 akzo radau3", "akzo_nobel", "Radau", 128, 0.05, 0, 1e-8, 1e-8, {}, True),
    ("C4 andrews radau5", "andrews", "Radau", 128, 0.05, 2, 1e-4, 1e-4, {"stages": 5}, True),
    ("C5 transamp radau3", "transamp", "Radau", 128, 0.02, 1, 1e-10, 1e-6, {}, True),
    ("C5 transamp bdf", "transamp", "BDF", 128, 0.02, 1, 1e-10, 1e-6, {}, True),
]


@pytest.mark.parametrize("name,problem,method,N,sigma,seed,rtol,atol,opts,use_twin", CASES, ids=[c[0] for c in CASES])
def test_cuda_matches_live_reference(name, problem, method, N, sigma, seed, rtol, atol, opts, use_twin):
    import torch
    from scipy_dae_b200 import solve_dae_batched, perturbed_params, initial_conditions, get_problem
    pr = get_problem(problem)
    params = perturbed_params(problem, N, sigma, seed)
    y0, yp0 = initial_conditions(problem, params)
    r = solve_dae_batched(problem, pr.t_span, y0, yp0, params, method=method, rtol=rtol, atol=atol, **opts)
    torch.cuda.synchronize()
    r = r.numpy()
    ref = REF.solve_ensemble(problem, params, y0, yp0, pr.t_span, method=method, rtol=rtol, atol=atol, **opts)
    twin = None
    if use_twin:
        pt = params * (1.0 + 1e-15)
        yt, ypt = initial_conditions(problem, pt)
        twin = REF.solve_ensemble(problem, pt, yt, ypt, pr.t_span, method=method, rtol=rtol, atol=atol, **opts)

    assert np.array_equal(r.status, ref["status"]) or (twin is not None and
        np.array_equal(r.status[twin["status"] == ref["status"]], ref["status"][twin["status"] == ref["status"]]))
    ok = (ref["status"] == 0) & (r.status == 0)
    assert ok.mean() > 0.95
    fails = []
    d_state = d_acc = d_rej = None
    rel = lambda a, b: np.abs(np.asarray(a, dtype=float) - b) / np.maximum(b, 1.0)      # noqa: E731
    rej = lambda x: np.asarray(x["n_err_rejected"], dtype=float) + np.asarray(x["n_newton_fail"], dtype=float)  # noqa: E731
    acc_ref = ref["n_accepted"].astype(float)
    attempts = acc_ref + rej(ref)
    if twin is not None:
        ok &= twin["status"] == 0
        d_state = instance_ratios(twin["y_final"][ok], ref["y_final"][ok], rtol, atol)
        d_acc = rel(twin["n_accepted"][ok], acc_ref[ok])
        d_rej = np.abs(rej(twin) - rej(ref))[ok] / attempts[ok]
    rs = instance_ratios(r.y_final[ok], ref["y_final"][ok], rtol, atol)
    fails += quantile_check(rs, d_state, 1.0, "y_final / band")
    fails += quantile_check(rel(r.n_accepted[ok], acc_ref[ok]), d_acc, STEP_BAND + 2.0 / acc_ref[ok], "accepted steps (relative)")
    rej_cuda = r.n_err_rejected.astype(float) + r.n_newton_fail
    fails += quantile_check(np.abs(rej_cuda - rej(ref))[ok] / attempts[ok], d_rej, STEP_BAND + 2.0 / attempts[ok],
                            "rejected attempts / attempts")
    # ensemble means of the work counters: no systematic bias against the reference (1 %; with a
    # twin: or twice the twin's own deviation of the mean)
    for key in ("n_accepted", "nfev", "njev", "nlu"):
        a, b = float(np.mean(getattr(r, key)[ok])), float(np.mean(ref[key][ok]))
        allow = 0.01
        if twin is not None:
            allow = max(allow, 2.0 * abs(float(np.mean(twin[key][ok])) - b) / b)
        if abs(a - b) > allow * b:
            fails.append(f"mean {key}: {a:.2f} vs reference {b:.2f} (allowed {100 * allow:.2f} %)")
    print(f"{name}: N={N} ok={int(ok.sum())} state q50/q90/max = {np.quantile(rs, .5):.3g}/{np.quantile(rs, .9):.3g}/{rs.max():.3g}"
          + (f" (twin {np.quantile(d_state, .5):.3g}/{np.quantile(d_state, .9):.3g}/{d_state.max():.3g})" if twin is not None else "")
          + f" mean nlu {np.mean(r.nlu[ok]):.2f} vs {np.mean(ref['nlu'][ok]):.2f}")
    assert not fails, "; ".join(fails)
