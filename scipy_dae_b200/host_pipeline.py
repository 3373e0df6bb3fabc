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


_FLAG = None


def _bad_input_flag():
    """One page-locked int32 the zero-copy launch can raise (b200dae_outputs.bad_input)."""
    global _FLAG
    if _FLAG is None:
        _FLAG = torch.zeros(1, dtype=torch.int32).pin_memory()
    return _FLAG


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


class _Streams:
    """Streams of one pipeline: `feed` (host->device copies + AoS->SoA relayout of every chunk) and
    `drain` (SoA->AoS relayout + device->host copies) have HIGH priority, the `compute` streams
    that carry the integrator kernels the default one.

    Why priorities (measured with scripts/e2e_timeline.py): the small relayout kernels that follow
    the integrator kernel of chunk c used to queue behind every pending block of the integrator
    kernels of chunks c+1, c+2 (the block scheduler serves kernels in launch order), so the
    device->host copy of chunk 0 only started after 4 ms, the stream's next host->device copy
    waited behind it, and the SMs idled 0.3 ms between chunks.  On a high-priority stream those
    kernels are dispatched as soon as any block slot frees up."""

    def __init__(self, device, n_compute: int):
        lo, hi = 0, -1                                    # CUDA: lower number = higher priority
        self.feed = torch.cuda.Stream(device=device, priority=hi)
        self.drain = torch.cuda.Stream(device=device, priority=hi)
        self.compute = [torch.cuda.Stream(device=device, priority=lo) for _ in range(n_compute)]

    def all(self):
        return [self.feed, self.drain, *self.compute]


def _enqueue(prob, t_span, h_y0, h_yp0, h_p, method, rtol, atol, device, bounds, streams, out,
             check_finite, options):
    """Queue the whole pipeline (no host synchronisation).  feed: every chunk's H2D + relayout, in
    chunk order, not waiting for anything; compute[c % k]: the integrator kernel of chunk c once its
    inputs are there; drain: relayout + D2H of chunk c once its kernel is done.  Returns the
    device flag `inputs finite` (or None) and the tensors that must outlive the queued work."""
    main = torch.cuda.current_stream(device)
    for s in streams.all():
        s.wait_stream(main)
    flags, keep, staged = [], [], []
    with torch.cuda.stream(streams.feed):
        for c in range(len(bounds) - 1):
            lo, hi = int(bounds[c]), int(bounds[c + 1])
            if hi <= lo:
                continue
            soa = [h[lo:hi].to(device, non_blocking=True).t().contiguous() for h in (h_y0, h_yp0, h_p)]
            if check_finite:
                flags.append(torch.isfinite(soa[0]).all() & torch.isfinite(soa[1]).all())
            ev = torch.cuda.Event()
            ev.record(streams.feed)
            staged.append((lo, hi, soa, ev))
    keep.append(staged)
    for c, (lo, hi, soa, ev) in enumerate(staged):
        s = streams.compute[c % len(streams.compute)]
        s.wait_event(ev)
        with torch.cuda.stream(s):
            # [N, width] views of the SoA tensors: solve_dae_batched's own relayout is then a no-op
            r = solve_dae_batched(prob, t_span, soa[0].t(), soa[1].t(), soa[2].t(), method=method,
                                  rtol=rtol, atol=atol, device=device, stream=s, check_finite=False,
                                  **options)
            done = torch.cuda.Event()
            done.record(s)
        streams.drain.wait_event(done)
        with torch.cuda.stream(streams.drain):
            out.y_final[lo:hi].copy_(r.y_final.contiguous(), non_blocking=True)
            out.yp_final[lo:hi].copy_(r.yp_final.contiguous(), non_blocking=True)
            out.t_final[lo:hi].copy_(r.t_final, non_blocking=True)
            out.status[lo:hi].copy_(r.status, non_blocking=True)
            cnt = torch.stack([getattr(r, k) for k in _COUNTERS], dim=1)
            out.counters[lo:hi].copy_(cnt, non_blocking=True)
        keep.append((r, cnt, done))
    for s in streams.all():
        main.wait_stream(s)
    ok = torch.stack(flags).all() if flags else None
    if ok is not None:
        keep.append(flags)
    return ok, keep


def _chunk_bounds(N: int, chunks: int, taper: bool) -> np.ndarray:
    """Chunk boundaries on the instance axis.  With `taper` the first and the last chunk are a
    quarter (and their neighbours three quarters) of the interior ones: the host->device copy
    of the first chunk and the device->host copy of the last one are the only transfers no kernel
    hides, so they are kept short."""
    if not taper or chunks < 5:
        return np.linspace(0, N, chunks + 1).astype(np.int64)
    w = np.ones(chunks)
    w[0] = w[-1] = 0.25
    w[1] = w[-2] = 0.75
    edges = np.concatenate([[0.0], np.cumsum(w) / w.sum()])
    b = np.round(edges * N).astype(np.int64)
    b[-1] = N
    return b


_GRAPHS: dict = {}


def solve_dae_batched_host(problem, t_span, y0, yp0, params, method="Radau", rtol=1e-3, atol=1e-6, *,
                           device=None, chunks: int = 16, n_streams: int = 3,
                           out: Optional[HostResultBuffers] = None, check_finite: bool = True,
                           use_graph: bool = False, taper: bool = True, zero_copy: bool = False,
                           **options) -> BatchedDaeResult:
    """Integrate ``N`` instances held in host memory; returns host (pinned) tensors.

    y0, yp0 : [N, n], params : [N, p]   NumPy arrays or CPU tensors (pinned memory makes the
                                        copies asynchronous)
    chunks, n_streams                   pipeline shape (``n_streams`` compute streams besides the
                                        feed and drain streams); ``chunks=1`` is the unpipelined call
    taper                               short first / last chunks (see ``_chunk_bounds``)
    out                                 reusable result buffers (``alloc_host_result``)
    use_graph                           capture the whole pipeline (copies, relayouts, kernels) in
                                        ONE CUDA graph, cached for this set of pinned input /
                                        output buffers and options; later calls with the same
                                        buffers replay it with a single launch, so host-side
                                        jitter cannot stall the queue between chunks.  Needs
                                        pinned ``y0, yp0, params`` tensors and ``out``.

    zero_copy                           ONE launch that reads the pinned ``(N, n)`` / ``(N, p)`` inputs and writes
                                        the pinned result buffers in place over PCIe (C-ABI ``io_layout = 1``):
                                        no staging copies, no relayouts, no chunking -- a thread's nine input
                                        loads and its result stores hide behind the other resident warps'
                                        arithmetic (measured: 10.74 ms for 2^20 Robertson instances against
                                        10.66 ms with device-resident buffers and 11.20 ms through the chunked
                                        pipeline).  Needs pinned ``y0, yp0, params`` tensors; non-finite initial
                                        values are found from the per-instance status afterwards.

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
    if zero_copy and N > 0:
        if not all(t.is_pinned() for t in (h_p, h_y0, h_yp0)):
            raise ValueError("zero_copy=True needs pinned y0 / yp0 / params tensors (and page-locked `out` buffers)")
        flag = _bad_input_flag()
        flag[0] = 0
        with torch.cuda.device(device):
            r = solve_dae_batched(prob, t_span, h_y0, h_yp0, h_p, method=method, rtol=rtol, atol=atol, device=device,
                                  check_finite=False,
                                  _host_io={"y0": h_y0, "yp0": h_yp0, "params": h_p, "out": out, "flag": flag}, **options)
            torch.cuda.current_stream(device).synchronize()
        if check_finite:
            # A non-finite initial value makes the first iteration matrix non-finite: status -2 before the
            # first accepted step, for which the kernels raise ONE flag word (b200dae_outputs: bad_input), so
            # that the host does not have to scan status[] (0.3 ms for 2^20 instances).
            if bool(flag[0]):
                idx = np.nonzero(out.status.numpy() == _lib.ST_NONFINITE)[0]
                if not (np.isfinite(h_y0.numpy()[idx]).all() and np.isfinite(h_yp0.numpy()[idx]).all()):
                    raise ValueError("All components of the initial state `y0` / `yp0` must be finite.")
        r.extra.update({"chunks": 1, "streams": 1, "graph": False, "zero_copy": True})
        return r
    chunks = max(1, min(int(chunks), N)) if N > 0 else 1
    bounds = _chunk_bounds(N, chunks, taper)
    nstr = max(1, min(n_streams, chunks))
    with torch.cuda.device(device):
        main = torch.cuda.current_stream(device)
        if not use_graph:
            streams = _Streams(device, nstr)
            ok, keep = _enqueue(prob, t_span, h_y0, h_yp0, h_p, method, rtol, atol, device, bounds,
                                streams, out, check_finite, options)
            main.synchronize()
        else:
            key = (prob.name, tuple(map(float, t_span)), method, float(rtol), np.asarray(atol, dtype=float).tobytes(),
                   str(device), chunks, nstr, bool(taper), check_finite, tuple(sorted((k, repr(v)) for k, v in options.items())),
                   h_y0.data_ptr(), h_yp0.data_ptr(), h_p.data_ptr(), N,
                   out.y_final.data_ptr(), out.yp_final.data_ptr(), out.t_final.data_ptr(),
                   out.status.data_ptr(), out.counters.data_ptr())
            entry = _GRAPHS.get(key)
            if entry is None:
                streams = _Streams(device, nstr)
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
