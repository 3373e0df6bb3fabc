"""Timeline of one host-pipeline call (solve_dae_batched_host, graph replay or eager queueing) from torch.profiler's
CUDA activity records: per stream, when each copy / kernel started and how long it ran.  nsys is
not in the image; this is enough to see what the pipeline leaves exposed.

    python scripts/e2e_timeline.py [chunks] [streams] [graph=1]"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    chunks = int(sys.argv[1]) if len(sys.argv) > 1 else 8
    nstr = int(sys.argv[2]) if len(sys.argv) > 2 else 3
    graph = bool(int(sys.argv[3])) if len(sys.argv) > 3 else True
    import torch
    from torch.profiler import profile, ProfilerActivity
    from scipy_dae_b200 import solve_dae_batched_host, alloc_host_result, perturbed_params, initial_conditions
    N = 1 << 20
    params = perturbed_params("robertson", N, 0.1, seed=0)
    y0, yp0 = initial_conditions("robertson", params)
    hp, hy, hyp = (torch.from_numpy(a).pin_memory() for a in (params, y0, yp0))
    out = alloc_host_result(N, 3)

    def step():
        return solve_dae_batched_host("robertson", (0.0, 1e7), hy, hyp, hp, method="Radau", rtol=1e-6, atol=1e-6,
                                      chunks=chunks, n_streams=nstr, out=out, use_graph=graph)
    for _ in range(3):
        step()
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        step()
        torch.cuda.synchronize()
    ev = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
    ev.sort(key=lambda e: e.time_range.start)
    t0 = ev[0].time_range.start
    print(f"{len(ev)} device activities, span {(max(e.time_range.end for e in ev) - t0) / 1e3:.3f} ms")
    busy = []
    for e in ev:
        name = e.name[:48]
        big = e.time_range.end - e.time_range.start > 50
        if big:
            print(f"{(e.time_range.start - t0) / 1e3:8.3f} ms  +{(e.time_range.end - e.time_range.start) / 1e3:7.3f}  {name}")
        if "radau_kernel" in e.name:
            busy.append((e.time_range.start - t0, e.time_range.end - t0))
    busy.sort()
    covered, cur_s, cur_e = 0.0, None, None
    for s, e in busy:
        if cur_e is None or s > cur_e:
            if cur_e is not None:
                covered += cur_e - cur_s
            cur_s, cur_e = s, e
        else:
            cur_e = max(cur_e, e)
    covered += (cur_e - cur_s) if cur_e is not None else 0.0
    print(f"integrator kernels: first start {busy[0][0] / 1e3:.3f} ms, last end {max(e for _, e in busy) / 1e3:.3f} ms, "
          f"union {covered / 1e3:.3f} ms, sum {sum(e - s for s, e in busy) / 1e3:.3f} ms")


if __name__ == "__main__":
    main()
