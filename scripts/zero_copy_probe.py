"""Feasibility probe: the headline launch with its inputs and outputs in PINNED HOST memory (the kernel
reads / writes them over PCIe through UVA, no separate copies).  Prints the event time per launch for
device-resident and for host-resident buffers, SoA layout in both cases.

    python scripts/zero_copy_probe.py [--instances N]
"""
import argparse
import ctypes
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--instances", type=int, default=1 << 20)
    args = ap.parse_args()
    import torch
    from scipy_dae_b200 import _lib, perturbed_params, initial_conditions
    from scipy_dae_b200.constants import radau_tables_c
    dev = torch.device("cuda:0")
    L = _lib.load()
    pid, n, npar = _lib.problem_lookup("robertson")
    N = args.instances
    params = perturbed_params("robertson", N, 0.1, seed=0)
    y0, yp0 = initial_conditions("robertson", params)
    ro = _lib.RadauOpts(rtol=1e-6, max_step=float("inf"), first_step=0.0, continuous_error_weight=0.0,
                        jac_recompute_rate=1e-3, deadband_lo=1.0, deadband_hi=1.2, newton_max_iter=0,
                        newton_iter_embedded=1, max_steps=0, analytic_jac=0)
    tb = radau_tables_c(3)
    atol = (ctypes.c_double * n)(*([1e-6] * n))
    stream = torch.cuda.current_stream(dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def buffers(where):
        mk = (lambda a: torch.from_numpy(np.ascontiguousarray(a.T)).pin_memory()) if where == "host" else \
             (lambda a: torch.from_numpy(np.ascontiguousarray(a.T)).to(dev))
        empty = (lambda shape, dt: torch.empty(shape, dtype=dt).pin_memory()) if where == "host" else \
                (lambda shape, dt: torch.empty(shape, dtype=dt, device=dev))
        d = dict(y0=mk(y0), yp0=mk(yp0), p=mk(params), y=empty((n, N), torch.float64), yp=empty((n, N), torch.float64),
                 t=empty((N,), torch.float64), st=empty((N,), torch.int32), c=empty((8, N), torch.int32))
        out = _lib.Outputs()
        out.y_final, out.yp_final, out.t_final = d["y"].data_ptr(), d["yp"].data_ptr(), d["t"].data_ptr()
        out.status, out.counters = d["st"].data_ptr(), d["c"].data_ptr()
        d["out"] = out
        return d

    res = {}
    for where in ("device", "host"):
        b = buffers(where)

        def launch():
            _lib.check(L.b200dae_solve_radau(pid, N, N, b["y0"].data_ptr(), b["yp0"].data_ptr(), b["p"].data_ptr(),
                                             0.0, 1e7, atol, ctypes.byref(ro), ctypes.byref(tb), ctypes.byref(b["out"]),
                                             ctypes.c_void_p(stream.cuda_stream)))
        for _ in range(2):
            launch()
        torch.cuda.synchronize()
        ts, wall = [], []
        for _ in range(5):
            flush.fill_(1)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            t0 = time.perf_counter()
            e0.record(); launch(); e1.record()
            torch.cuda.synchronize()
            wall.append((time.perf_counter() - t0) * 1e3)
            ts.append(e0.elapsed_time(e1))
        st = b["st"] if where == "host" else b["st"].cpu()
        res[where] = dict(ms_min=min(ts), ms_med=float(np.median(ts)), wall_ms_min=min(wall), finished=int((st == 0).sum()),
                          y_checksum=float((b["y"] if where == "host" else b["y"].cpu()).double().sum()))
    # instance-major pinned host arrays (the shapes of the reference's API) through io_layout = 1
    hp, hy, hyp = (torch.from_numpy(np.ascontiguousarray(a)).pin_memory() for a in (params, y0, yp0))
    from scipy_dae_b200 import alloc_host_result, solve_dae_batched_host
    ob = alloc_host_result(N, n)
    out = _lib.Outputs()
    out.y_final, out.yp_final, out.t_final = ob.y_final.data_ptr(), ob.yp_final.data_ptr(), ob.t_final.data_ptr()
    out.status, out.counters = ob.status.data_ptr(), ob.counters.data_ptr()
    ro.io_layout = 1

    def launch_aos():
        _lib.check(L.b200dae_solve_radau(pid, N, N, hy.data_ptr(), hyp.data_ptr(), hp.data_ptr(), 0.0, 1e7, atol,
                                         ctypes.byref(ro), ctypes.byref(tb), ctypes.byref(out), ctypes.c_void_p(stream.cuda_stream)))
    for _ in range(2):
        launch_aos()
    torch.cuda.synchronize()
    ts, wall = [], []
    for _ in range(5):
        flush.fill_(1)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record(); launch_aos(); e1.record()
        torch.cuda.synchronize()
        wall.append((time.perf_counter() - t0) * 1e3)
        ts.append(e0.elapsed_time(e1))
    res["host_aos_cabi"] = dict(ms_min=min(ts), wall_ms_min=min(wall), finished=int((ob.status == 0).sum()),
                                y_checksum=float(ob.y_final.double().sum()))
    for cf in (True, False):
        wall = []
        for _ in range(6):
            flush.fill_(1)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            solve_dae_batched_host("robertson", (0.0, 1e7), hy, hyp, hp, rtol=1e-6, atol=1e-6, out=ob, zero_copy=True,
                                   check_finite=cf)
            wall.append((time.perf_counter() - t0) * 1e3)
        res["host_aos_api_check_finite_%s" % cf] = dict(wall_ms_min=min(wall[1:]), wall_ms_med=float(np.median(wall[1:])))
    print(json.dumps(res))


if __name__ == "__main__":
    main()
