#!/usr/bin/env python3
"""Parity report (GPU box): CUDA path vs the CPU oracle on seeded ensembles, per config.
Prints the distribution of |y - y_ref| / max(10 rtol |y_ref|, atol) and of the step-count
differences.  TEST INFRASTRUCTURE (imports oracle/)."""
import json
import os
import sys
import time

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)

from oracle import oracle as O                                              # noqa: E402
from scipy_dae_b200 import (solve_dae_batched, perturbed_params,            # noqa: E402
                            initial_conditions, get_problem)

CASES = [
    ("C2 robertson Radau3", "robertson", "Radau", 8192, 0.1, 1e-6, 1e-6, {}),
    ("C2 robertson Radau3 s0.5", "robertson", "Radau", 4096, 0.5, 1e-6, 1e-6, {}),
    ("robertson Radau5", "robertson", "Radau", 2048, 0.1, 1e-6, 1e-6, {"stages": 5}),
    ("robertson Radau3 tight", "robertson", "Radau", 1024, 0.1, 1e-9, 1e-12, {}),
    ("robertson BDF", "robertson", "BDF", 4096, 0.1, 1e-6, 1e-6, {}),
    ("C3 akzo BDF", "akzo_nobel", "BDF", 2048, 0.05, 1e-8, 1e-8, {}),
    ("akzo Radau3", "akzo_nobel", "Radau", 1024, 0.05, 1e-8, 1e-8, {}),
    ("C4 andrews Radau5", "andrews", "Radau", 128, 0.05, 1e-4, 1e-4, {"stages": 5}),
    ("andrews Radau3 1e-6", "andrews", "Radau", 64, 0.05, 1e-6, 1e-6, {"stages": 3}),
    ("C5 transamp Radau3", "transamp", "Radau", 128, 0.02, 1e-10, 1e-6, {}),
    ("C5 transamp BDF", "transamp", "BDF", 128, 0.02, 1e-10, 1e-6, {}),
    # ensemble MEANS of the work counters at scale (the group kernel was biased by +2.4 % in nlu)
    ("big akzo Radau3 131072", "akzo_nobel", "Radau", 131072, 0.05, 1e-8, 1e-8, {}),
    ("big akzo BDF 131072", "akzo_nobel", "BDF", 131072, 0.05, 1e-8, 1e-8, {}),
    ("big andrews Radau5 4096", "andrews", "Radau", 4096, 0.05, 1e-4, 1e-4, {"stages": 5}),
    ("big transamp Radau3 2048", "transamp", "Radau", 2048, 0.02, 1e-10, 1e-6, {}),
    ("big transamp BDF 2048", "transamp", "BDF", 2048, 0.02, 1e-10, 1e-6, {}),
]


def main():
    import torch
    sel = sys.argv[1:]
    rows = []
    for name, problem, method, N, sigma, rtol, atol, opts in CASES:
        if (sel and not any(s in name for s in sel)) or (not sel and name.startswith("big ")):
            continue
        pr = get_problem(problem)
        params = perturbed_params(problem, N, sigma, seed=11)
        y0, yp0 = initial_conditions(problem, params)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        r = solve_dae_batched(problem, pr.t_span, y0, yp0, params, method=method, rtol=rtol, atol=atol, **opts).numpy()
        tg = time.perf_counter() - t0
        t0 = time.perf_counter()
        ref = O.solve_ensemble(problem, params, y0, yp0, pr.t_span, method=method, rtol=rtol, atol=atol,
                               nthreads=os.cpu_count() or 1, **opts)
        tc = time.perf_counter() - t0
        ok = (ref["status"] == 0) & (r.status == 0)
        ratio = (np.abs(r.y_final - ref["y_final"]) / np.maximum(10 * rtol * np.abs(ref["y_final"]), atol)).max(axis=1)[ok]
        dacc = (r.n_accepted - ref["n_accepted"])[ok] / ref["n_accepted"][ok]
        row = dict(case=name, N=N, status_equal=bool((r.status == ref["status"]).all()),
                   n_failed_ref=int((ref["status"] != 0).sum()),
                   ratio_median=float(np.median(ratio)), ratio_p99=float(np.quantile(ratio, 0.99)),
                   ratio_max=float(ratio.max()), frac_in_band=float((ratio <= 1).mean()),
                   acc_rel_diff_max=float(np.abs(dacc).max()), acc_within_5pct=float((np.abs(dacc) <= 0.05).mean()),
                   acc_mean_gpu=float(r.n_accepted[ok].mean()), acc_mean_ref=float(ref["n_accepted"][ok].mean()),
                   nfev_mean_gpu=float(r.nfev[ok].mean()), nfev_mean_ref=float(ref["nfev"][ok].mean()),
                   njev_mean_gpu=float(r.njev[ok].mean()), njev_mean_ref=float(ref["njev"][ok].mean()),
                   nlu_mean_gpu=float(r.nlu[ok].mean()), nlu_mean_ref=float(ref["nlu"][ok].mean()),
                   nfail_mean_gpu=float(r.n_newton_fail[ok].mean()), nfail_mean_ref=float(ref["n_newton_fail"][ok].mean()),
                   nrej_mean_gpu=float(r.n_err_rejected[ok].mean()), nrej_mean_ref=float(ref["n_err_rejected"][ok].mean()),
                   niter_mean_gpu=float(r.n_newton_iter[ok].mean()), niter_mean_ref=float(ref["n_newton_iter"][ok].mean()),
                   gpu_s=round(tg, 3), cpu_s=round(tc, 3))
        rows.append(row)
        print(json.dumps(row), flush=True)
    return rows


if __name__ == "__main__":
    main()
