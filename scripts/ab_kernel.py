"""Kernel A/B timing of the headline launch (2^20 Robertson instances, Radau s = 3, lean kernel)
for whichever library B200DAE_LIB selects (scripts/build_variant.py).  Prints one JSON line:
event time per launch (min / median of --reps, L2 flushed in between), the mean work counters and
the number of instances that did not finish, so a variant that is faster because it does
different work shows up at once.

    B200DAE_LIB=scipy_dae_b200/variants/libb200dae_fd2.so python scripts/ab_kernel.py [--jac analytic]
"""
from __future__ import annotations

import argparse
import ctypes
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--instances", type=int, default=1 << 20)
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--jac", default="fd", choices=["fd", "analytic"])
    ap.add_argument("--sigma", type=float, default=0.1)
    args = ap.parse_args()

    import torch
    from scipy_dae_b200 import _lib, perturbed_params, initial_conditions
    from scipy_dae_b200.constants import radau_tables_c

    dev = torch.device("cuda:0")
    L = _lib.load()
    pid, n, npar = _lib.problem_lookup("robertson")
    N = args.instances
    params = perturbed_params("robertson", N, args.sigma, seed=0)
    y0, yp0 = initial_conditions("robertson", params)
    d_p = torch.from_numpy(params).to(dev).t().contiguous()
    d_y0 = torch.from_numpy(y0).to(dev).t().contiguous()
    d_yp0 = torch.from_numpy(yp0).to(dev).t().contiguous()
    f64 = dict(dtype=torch.float64, device=dev)
    y_final = torch.empty((n, N), **f64)
    yp_final = torch.empty((n, N), **f64)
    t_final = torch.empty((N,), **f64)
    status = torch.empty((N,), dtype=torch.int32, device=dev)
    counters = torch.empty((8, N), dtype=torch.int32, device=dev)
    out = _lib.Outputs()
    out.y_final, out.yp_final, out.t_final = y_final.data_ptr(), yp_final.data_ptr(), t_final.data_ptr()
    out.status, out.counters = status.data_ptr(), counters.data_ptr()
    ro = _lib.RadauOpts(rtol=1e-6, max_step=float("inf"), first_step=0.0, continuous_error_weight=0.0,
                        jac_recompute_rate=1e-3, deadband_lo=1.0, deadband_hi=1.2, newton_max_iter=0,
                        newton_iter_embedded=1, max_steps=0, analytic_jac=1 if args.jac == "analytic" else 0)
    tb = radau_tables_c(3)
    atol = (ctypes.c_double * n)(*([1e-6] * n))
    stream = torch.cuda.current_stream(dev)

    def launch():
        _lib.check(L.b200dae_solve_radau(pid, N, N, d_y0.data_ptr(), d_yp0.data_ptr(), d_p.data_ptr(),
                                         0.0, 1e7, atol, ctypes.byref(ro), ctypes.byref(tb),
                                         ctypes.byref(out), ctypes.c_void_p(stream.cuda_stream)))

    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    for _ in range(3):
        launch()
    torch.cuda.synchronize()
    times = []
    for _ in range(args.reps):
        flush.fill_(1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        launch()
        e1.record()
        torch.cuda.synchronize()
        times.append(e0.elapsed_time(e1))
    c = counters.double().mean(dim=1).cpu().numpy()
    yf = y_final.t().cpu().numpy()
    print(json.dumps({"lib": os.path.basename(_lib.LIB_PATH), "jac": args.jac, "ms_min": min(times),
                      "ms_med": float(np.median(times)),
                      "unfinished": int((status != 0).sum()),
                      "counters_mean": {k: round(float(v), 3) for k, v in zip(_lib.COUNTER_NAMES, c)},
                      "y_checksum": [float(x) for x in yf.sum(axis=0)]}), flush=True)


if __name__ == "__main__":
    main()
