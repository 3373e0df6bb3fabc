"""Host-facing batched entry point: inputs and results live in HOST memory.

``solve_dae_batched`` works on device tensors and returns device tensors.  A caller that holds
the ensemble in host arrays (the reference's world: NumPy) pays, per call, the host->device copy
of ``y0, yp0, params``, the AoS->SoA relayout, the launch and the device->host copy of the
results.  ``solve_dae_batched_host`` pipelines those stages over chunks of the instance axis on a
few CUDA streams, so the copies of chunk ``c+1`` and ``c-1`` run on the copy engines while the
integrator kernel of chunk ``c`` occupies the SMs (instances are independent, SURVEY.md §8e, so
any split of the batch gives the same per-instance results).

The per-instance semantics are those of ``solve_dae_batched`` (same C-ABI call per chunk).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional

import numpy as np
import torch

from . import _lib
from .batched import BatchedDaeResult, solve_dae_batched
from .problems import get_problem

_COUNTERS = ("nfev", "njev", "nlu", "nsolve", "n_accepted", "n_err_rejected", "n_newton_fail",
             "n_newton_iter")


@dataclass
class HostResultBuffers:
    """Pinned host tensors that receive the results of one ensemble (reusable across calls)."""
    t_final: torch.Tensor     # [N]
    y_final: torch.Tensor     # [N, n]
    yp_final: torch.Tensor    # [N, n]
    status: torch.Tensor      # [N] int32
    counters: torch.Tensor    # [N, 8] int32

    @property
    def nbytes(self) -> int:
        return sum(t.numel() * t.element_size() for t in
                   (self.t_final, self.y_final, self.yp_final, self.status, self.counters))


def alloc_host_result(N: int, n: int, pin: bool = True) -> HostResultBuffers:
    mk = lambda shape, dt: torch.empty(shape, dtype=dt, pin_memory=pin)        # noqa: E731
    return HostResultBuffers(mk((N,), torch.float64), mk((N, n), torch.float64), mk((N, n), torch.float64),
                             mk((N,), torch.int32), mk((N, len(_COUNTERS)), torch.int32))


def _as_host(x, name):
    if isinstance(x, torch.Tensor):
        if x.is_cuda:
            raise ValueError(f"`{name}` is a device tensor: use solve_dae_batched")
        t = x
    else:
        t = torch.from_numpy(np.ascontiguousarray(np.asarray(x), dtype=np.float64))
    if t.dtype != torch.float64:
        t = t.to(torch.float64)
    if t.ndim != 2:
        raise ValueError(f"`{name}` must have shape (N, width)")
    return t.contiguous()


def _enqueue(prob, t_span, h_y0, h_yp0, h_p, method, rtol, atol, device, bounds, streams, out,
             check_finite, options):
    """Queue the whole pipeline (no host synchronisation): per chunk, on its stream, H2D ->
    relayout -> integrator kernel -> relayout -> D2H.  Returns the device flag `inputs finite`
    (or None) and the tensors that must outlive the queued work."""
    main = torch.cuda.current_stream(device)
    for s in streams:
        s.wait_stream(main)
    flags, keep = [], []
    for c in range(len(bounds) - 1):
        lo, hi = int(bounds[c]), int(bounds[c + 1])
        if hi <= lo:
            continue
        s = streams[c % len(streams)]
        with torch.cuda.stream(s):
            r = solve_dae_batched(prob, t_span, h_y0[lo:hi], h_yp0[lo:hi], h_p[lo:hi], method=method,
                                  rtol=rtol, atol=atol, device=device, stream=s, check_finite=False,
                                  **options)
            if check_finite:
                yd, ypd, _ = r.extra["_keepalive"]
                flags.append(torch.isfinite(yd).all() & torch.isfinite(ypd).all())
            out.y_final[lo:hi].copy_(r.y_final.contiguous(), non_blocking=True)
            out.yp_final[lo:hi].copy_(r.yp_final.contiguous(), non_blocking=True)
            out.t_final[lo:hi].copy_(r.t_final, non_blocking=True)
            out.status[lo:hi].copy_(r.status, non_blocking=True)
            cnt = torch.stack([getattr(r, k) for k in _COUNTERS], dim=1)
            out.counters[lo:hi].copy_(cnt, non_blocking=True)
            keep.append(r)
    for s in streams:
        main.wait_stream(s)
    ok = torch.stack(flags).all() if flags else None
    return ok, keep


_GRAPHS: dict = {}


def solve_dae_batched_host(problem, t_span, y0, yp0, params, method="Radau", rtol=1e-3, atol=1e-6, *,
                           device=None, chunks: int = 8, n_streams: int = 3,
                           out: Optional[HostResultBuffers] = None, check_finite: bool = True,
                           use_graph: bool = False, **options) -> BatchedDaeResult:
    """Integrate ``N`` instances held in host memory; returns host (pinned) tensors.

    y0, yp0 : [N, n], params : [N, p]   NumPy arrays or CPU tensors (pinned memory makes the
                                        copies asynchronous)
    chunks, n_streams                   pipeline shape; ``chunks=1`` is the unpipelined call
    out                                 reusable result buffers (``alloc_host_result``)
    use_graph                           capture the whole pipeline (copies, relayouts, kernels) in
                                        ONE CUDA graph, cached for this set of pinned input /
                                        output buffers and options; later calls with the same
                                        buffers replay it with a single launch, so host-side
                                        jitter cannot stall the queue between chunks.  Needs
                                        pinned ``y0, yp0, params`` tensors and ``out``.

    ``t_eval`` / ``record_steps`` / ``events`` are not offered here (their outputs are much larger
    than the inputs; use ``solve_dae_batched`` and copy what is needed).
    """
    for k in ("t_eval", "record_steps", "events"):
        if options.get(k):
            raise NotImplementedError(f"`{k}` is not available on the host pipeline; use solve_dae_batched")
    if not torch.cuda.is_available():
        raise _lib.B200DaeError("scipy_dae_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
    prob = get_problem(problem)
    _, n, npar = _lib.problem_lookup(prob.name)
    device = torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}")
    h_p, h_y0, h_yp0 = _as_host(params, "params"), _as_host(y0, "y0"), _as_host(yp0, "yp0")
    N = h_y0.shape[0]
    if h_yp0.shape != h_y0.shape:
        raise ValueError("`y0` and `yp0` must be of same shape.")
    if h_y0.shape[1] != n or h_p.shape != (N, npar):
        raise ValueError(f"expected y0/yp0 of shape ({N}, {n}) and params of shape ({N}, {npar})")
    if use_graph and (out is None or not all(t.is_pinned() for t in (h_p, h_y0, h_yp0))):
        raise ValueError("use_graph=True needs pinned y0 / yp0 / params tensors and `out` buffers "
                         "(a CUDA graph replays fixed addresses)")
    if out is None:
        out = alloc_host_result(N, n)
    chunks = max(1, min(int(chunks), N)) if N > 0 else 1
    bounds = np.linspace(0, N, chunks + 1).astype(np.int64)
    nstr = max(1, min(n_streams, chunks))
    with torch.cuda.device(device):
        main = torch.cuda.current_stream(device)
        if not use_graph:
            streams = [torch.cuda.Stream(device=device) for _ in range(nstr)]
            ok, keep = _enqueue(prob, t_span, h_y0, h_yp0, h_p, method, rtol, atol, device, bounds,
                                streams, out, check_finite, options)
            main.synchronize()
        else:
            key = (prob.name, tuple(map(float, t_span)), method, float(rtol), np.asarray(atol, dtype=float).tobytes(),
                   str(device), chunks, nstr, check_finite, tuple(sorted((k, repr(v)) for k, v in options.items())),
                   h_y0.data_ptr(), h_yp0.data_ptr(), h_p.data_ptr(), N,
                   out.y_final.data_ptr(), out.yp_final.data_ptr(), out.t_final.data_ptr(),
                   out.status.data_ptr(), out.counters.data_ptr())
            entry = _GRAPHS.get(key)
            if entry is None:
                streams = [torch.cuda.Stream(device=device) for _ in range(nstr)]
                # one eager pass first: warms the allocator and any per-kernel attribute set-up
                _enqueue(prob, t_span, h_y0, h_yp0, h_p, method, rtol, atol, device, bounds, streams, out,
                         check_finite, options)
                main.synchronize()
                graph = torch.cuda.CUDAGraph()
                with torch.cuda.graph(graph):
                    ok, keep = _enqueue(prob, t_span, h_y0, h_yp0, h_p, method, rtol, atol, device, bounds,
                                        streams, out, check_finite, options)
                entry = (graph, ok, keep, streams, (h_y0, h_yp0, h_p, out))
                _GRAPHS[key] = entry
            graph, ok = entry[0], entry[1]
            graph.replay()
            torch.cuda.current_stream(device).synchronize()
        if ok is not None and not bool(ok):
            raise ValueError("All components of the initial state `y0` / `yp0` must be finite.")
    cn = out.counters
    return BatchedDaeResult(
        t_final=out.t_final, y_final=out.y_final, yp_final=out.yp_final, status=out.status,
        nfev=cn[:, 0], njev=cn[:, 1], nlu=cn[:, 2], nsolve=cn[:, 3], n_accepted=cn[:, 4],
        n_err_rejected=cn[:, 5], n_newton_fail=cn[:, 6], n_newton_iter=cn[:, 7],
        method=method, problem=prob.name, extra={"chunks": chunks, "streams": nstr, "graph": bool(use_graph)})
