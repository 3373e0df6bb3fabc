"""Multi-GPU sharding of an ensemble: one process per GPU (torch.distributed), contiguous index
ranges per rank, NO collective on the hot path, one gather of the results.

The reference has no distributed layer; instances are independent, so rank r
integrates instances [r*ceil(N/W), min(N, (r+1)*ceil(N/W))) on its own device with the same
single-launch kernel and the per-rank results are gathered once (NCCL over NVLink on the GPU
box, gloo in the CPU tests).
"""
from __future__ import annotations

from typing import Callable, Optional

import numpy as np
import torch
import torch.distributed as dist


def shard_bounds(N: int, world_size: int, rank: int):
    """Contiguous block partition: instance i -> rank i // ceil(N / world_size)."""
    per = -(-N // world_size) if N > 0 else 0
    lo = min(N, rank * per)
    hi = min(N, lo + per)
    return lo, hi, per


GATHER_FIELDS = ("t_final", "y_final", "yp_final", "status", "nfev", "njev", "nlu", "nsolve",
                 "n_accepted", "n_err_rejected", "n_newton_fail", "n_newton_iter")


def gather_shards(local: dict, N: int, per: int, group=None, dst: int = 0):
    """Gather per-rank result tensors (leading axis = local instances, padded to ``per``) onto
    ``dst``; returns a dict of ``[N, ...]`` tensors on ``dst`` and ``None`` elsewhere."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    out = {}
    for key in GATHER_FIELDS:
        x = local[key]
        pad_shape = (per,) + tuple(x.shape[1:])
        buf = torch.zeros(pad_shape, dtype=x.dtype, device=x.device)
        buf[: x.shape[0]] = x
        if rank == dst:
            parts = [torch.empty_like(buf) for _ in range(world)]
            dist.gather(buf, parts, dst=dst, group=group)
            out[key] = torch.cat(parts, dim=0)[:N]
        else:
            dist.gather(buf, None, dst=dst, group=group)
    return out if rank == dst else None


def solve_dae_sharded(problem, t_span, y0, yp0, params, method="Radau", t_eval=None, rtol=1e-3,
                      atol=1e-6, *, group=None, device=None, gather: bool = True,
                      _solve: Optional[Callable] = None, **options):
    """Every rank passes the same full-size host arrays; rank r integrates its block on its own
    GPU; rank 0 receives the gathered ``[N, ...]`` results (dict), other ranks ``None``.

    ``_solve(problem, t_span, y0, yp0, params, method=..., ...) -> dict of tensors`` replaces the
    CUDA solve in the CPU (gloo) tests of the sharding logic.
    """
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    params = np.atleast_2d(np.asarray(params, dtype=float))
    N = params.shape[0]
    lo, hi, per = shard_bounds(N, world, rank)
    y0 = np.atleast_2d(np.asarray(y0, dtype=float))
    yp0 = np.atleast_2d(np.asarray(yp0, dtype=float))
    if _solve is None:
        from .batched import solve_dae_batched

        def _solve(problem, t_span, y0, yp0, params, **kw):
            r = solve_dae_batched(problem, t_span, y0, yp0, params, device=device, **kw)
            return {k: getattr(r, k).contiguous() for k in GATHER_FIELDS}
    if hi > lo:
        local = _solve(problem, t_span, y0[lo:hi], yp0[lo:hi], params[lo:hi], method=method,
                       t_eval=t_eval, rtol=rtol, atol=atol, **options)
    else:
        local = None
    if world == 1 or not gather:
        return local
    if local is None:   # empty shard: fabricate zero-length tensors with the right trailing shapes
        n = y0.shape[1]
        # on the device the other ranks' results live on: a CPU tensor in an NCCL gather errors or hangs
        if device is not None:
            dev = torch.device(device)
        elif torch.cuda.is_available() and (not dist.is_initialized() or dist.get_backend(group) == "nccl"):
            dev = torch.device("cuda", torch.cuda.current_device())
        else:
            dev = torch.device("cpu")
        local = {k: torch.zeros((0, n) if k in ("y_final", "yp_final") else (0,),
                                dtype=torch.float64 if k in ("t_final", "y_final", "yp_final") else torch.int32,
                                device=dev) for k in GATHER_FIELDS}
    return gather_shards(local, N, per, group=group)
