#!/usr/bin/env python3
"""One tiny launch of every kernel family (for compute-sanitizer runs on the GPU box):
    compute-sanitizer --tool memcheck python scripts/sanity_all.py
Prints one line per launch; exits non-zero if a status is unexpected."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from scipy_dae_b200 import (solve_dae_batched, consistent_initial_conditions_batched, perturbed_params,  # noqa: E402
                            initial_conditions, get_problem, ThresholdEvent)

bad = 0
for problem, N, tspan_scale in (("robertson", 70, 1.0), ("akzo_nobel", 12, 1.0), ("transamp", 9, 0.02), ("andrews", 3, 0.2)):
    pr = get_problem(problem)
    P = perturbed_params(problem, N, 0.05, seed=1)
    y0, yp0 = initial_conditions(problem, P)
    t0, t1 = pr.t_span
    t1 = t0 + (t1 - t0) * tspan_scale
    te = np.linspace(t0, t1, 5)[1:]
    for method in ("Radau", "BDF"):
        for kw in (dict(), dict(t_eval=te, record_steps=8,
                                events=[ThresholdEvent(component=0, value=float(y0[0, 0]) * 0.999 - 1e-3)])):
            if problem == "andrews" and method == "BDF" and kw:
                continue
            for jac in ((None, "analytic") if pr.jacobian is not None and not kw else (None,)):
                r = solve_dae_batched(problem, (t0, t1), y0, yp0, P, method=method, rtol=1e-5, atol=1e-7,
                                      jac=jac, **kw).numpy()
                ok = bool(np.isin(r.status, (0, 1)).all())
                bad += not ok
                print(problem, method, "opts" if kw else "lean", jac, "status", np.unique(r.status), "acc", int(r.n_accepted.mean()))
    g = y0 * (1 + 1e-3), yp0 + 1e-3
    c = consistent_initial_conditions_batched(problem, t0, g[0], g[1], P).status.cpu().numpy()
    print(problem, "cic status", np.unique(c))
    bad += not bool((c == 0).all())
r5 = solve_dae_batched("robertson", (0.0, 1e3), method="Radau", stages=5, rtol=1e-6, atol=1e-8).numpy()
print("robertson Radau5 status", r5.status)
sys.exit(1 if bad else 0)
