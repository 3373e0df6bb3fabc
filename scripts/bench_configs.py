#!/usr/bin/env python3
"""Throughput of the BASELINE.json configs other than the headline (C3 Akzo Nobel BDF, C4 Andrews
Radau s=5, C5 transistor amplifier Radau s=3 / BDF) on one B200, beside the CPU oracle on the host
cores.  These are parity-test configurations, not bench.py lines; the numbers go to
profiles/.  TEST INFRASTRUCTURE (imports oracle/ for the CPU column).

usage: bench_configs.py [substring ...] [--regime thread] [--scale K]"""
import json
import os
import sys
import time

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)

CASES = [
    # name, problem, method, N_gpu, N_cpu, sigma, rtol, atol, opts
    # the headline ensemble with the other method (thread-per-system BDF kernel, n <= 4)
    ("C2 robertson BDF", "robertson", "BDF", 1 << 20, 16384, 0.1, 1e-6, 1e-6, {}),
    ("C3 akzo BDF", "akzo_nobel", "BDF", 1 << 17, 4096, 0.05, 1e-8, 1e-8, {}),
    ("C3 akzo Radau3", "akzo_nobel", "Radau", 1 << 17, 4096, 0.05, 1e-8, 1e-8, {}),
    ("C4 andrews Radau5", "andrews", "Radau", 1 << 13, 256, 0.05, 1e-4, 1e-4, {"stages": 5}),
    ("C4 andrews Radau3", "andrews", "Radau", 1 << 13, 128, 0.05, 1e-6, 1e-6, {"stages": 3}),
    ("C5 transamp Radau3", "transamp", "Radau", 1 << 14, 128, 0.02, 1e-10, 1e-6, {}),
    ("C5 transamp BDF", "transamp", "BDF", 1 << 14, 128, 0.02, 1e-10, 1e-6, {}),
]


def main():
    args = sys.argv[1:]
    scale = 1.0
    if "--scale" in args:
        i = args.index("--scale")
        scale = float(args[i + 1])
        del args[i:i + 2]
    if "--regime" in args:
        i = args.index("--regime")
        os.environ["B200DAE_REGIME"] = args[i + 1]
        del args[i:i + 2]
    no_cpu = "--no-cpu" in args
    args = [a for a in args if a != "--no-cpu"]
    import torch
    from oracle import oracle as O
    from scipy_dae_b200 import solve_dae_batched, perturbed_params, initial_conditions, get_problem
    for name, problem, method, n_gpu, n_cpu, sigma, rtol, atol, opts in CASES:
        if args and not any(s in name for s in args):
            continue
        pr = get_problem(problem)
        n_gpu = max(64, int(n_gpu * scale))
        params = perturbed_params(problem, n_gpu, sigma, seed=3)
        y0, yp0 = initial_conditions(problem, params)
        dev = torch.device("cuda:0")
        pd, yd, ypd = (torch.from_numpy(np.ascontiguousarray(a)).to(dev) for a in (params, y0, yp0))
        times = []
        for it in range(3):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            e0.record()
            r = solve_dae_batched(problem, pr.t_span, yd, ypd, pd, method=method, rtol=rtol, atol=atol,
                                  check_finite=False, **opts)
            e1.record()
            torch.cuda.synchronize()
            times.append(e0.elapsed_time(e1))
        ms = min(times[1:])
        status = r.status.cpu().numpy()
        row = dict(case=name, regime=os.environ.get("B200DAE_REGIME", "group"), N=n_gpu, ms=round(ms, 3),
                   solves_per_s=round(n_gpu / ms * 1e3, 1), n_failed=int((status != 0).sum()),
                   acc_mean=float(r.n_accepted.double().mean()), nfev_mean=float(r.nfev.double().mean()),
                   nlu_mean=float(r.nlu.double().mean()))
        if not no_cpu:
            t0 = time.perf_counter()
            ref = O.solve_ensemble(problem, params[:n_cpu], y0[:n_cpu], yp0[:n_cpu], pr.t_span, method=method,
                                   rtol=rtol, atol=atol, nthreads=os.cpu_count() or 1, **opts)
            tc = time.perf_counter() - t0
            row.update(cpu_port_solves_per_s=round(n_cpu / tc, 1), cpu_cores=os.cpu_count(), cpu_sample=n_cpu,
                       speedup_vs_cpu_port=round(n_gpu / ms * 1e3 / (n_cpu / tc), 1))
        print(json.dumps(row), flush=True)


if __name__ == "__main__":
    main()
