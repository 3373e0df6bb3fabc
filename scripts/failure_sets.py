"""Failure sets at scale (GPU box): which Andrews instances end with `step size too small` on the
CUDA path and in the CPU oracle, and the state-ratio distribution of the rest.  TEST INFRASTRUCTURE."""
import sys, os
sys.path.insert(0, '/root/repo')
import numpy as np
from oracle import oracle as O
from scipy_dae_b200 import solve_dae_batched, perturbed_params, initial_conditions, get_problem
for problem, N, sigma, rtol, atol, opts in (("andrews", 8192, 0.05, 1e-4, 1e-4, {"stages": 5}), ("andrews", 4096, 0.05, 1e-6, 1e-6, {"stages": 3})):
    pr = get_problem(problem)
    P = perturbed_params(problem, N, sigma, seed=3)
    y0, yp0 = initial_conditions(problem, P)
    r = solve_dae_batched(problem, pr.t_span, y0, yp0, P, method="Radau", rtol=rtol, atol=atol, **opts).numpy()
    ref = O.solve_ensemble(problem, P, y0, yp0, pr.t_span, method="Radau", rtol=rtol, atol=atol, nthreads=16, **opts)
    fg, fo = np.nonzero(r.status != 0)[0], np.nonzero(ref["status"] != 0)[0]
    print(problem, opts, "gpu failures", len(fg), r.status[fg][:8], "oracle failures", len(fo), ref["status"][fo][:8], "common", len(set(fg) & set(fo)))
    ok = (r.status == 0) & (ref["status"] == 0)
    ratio = (np.abs(r.y_final - ref["y_final"]) / np.maximum(10 * rtol * np.abs(ref["y_final"]), atol)).max(axis=1)[ok]
    print("   ratio median %.3g p99 %.3g max %.3g; acc mean gpu %.2f oracle %.2f" % (np.median(ratio), np.quantile(ratio, 0.99), ratio.max(), r.n_accepted[ok].mean(), ref["n_accepted"][ok].mean()))
