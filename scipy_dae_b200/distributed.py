"""Multi-GPU sharding of an ensemble: one process per GPU (torch.distributed), contiguous index
ranges per rank, NO collective on the hot path, one gather of the results.

The reference has no distributed layer; instances are independent, so rank r
integrates instances [r*ceil(N/W), min(N, (r+1)*ceil(N/W))) on its own device with the same
single-launch kernel and the per-rank results are gathered once.  Two gathers:

* ``gather="collective"``: ``torch.distributed.gather`` of the device tensors onto rank 0 (NCCL over
  NVLink on the GPU box, gloo in the CPU tests).  The first NCCL gather of a process also builds the
  peer-to-peer channels (0.2 / 0.5 / 1.7 s at 2 / 4 / 8 GPUs, measured), 100 x the 12 ms job;
* ``gather="host"`` (``SharedHostResult``): the result arrays of the WHOLE ensemble live in one POSIX
  shared-memory allocation that every rank of the node maps and page-locks, and each GPU writes its
  shard straight into its slice (device->host DMA) -- the single gather without any collective
  or channel set-up; one barrier tells rank 0 that all shards have landed.
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


class SharedHostResult:
    """Host result arrays ``[N, ...]`` of one ensemble in POSIX shared memory (``/dev/shm``), mapped
    by every rank of the node and page-locked where a CUDA device is present, so that
    ``solve_dae_batched_host(out=...)`` of rank r DMA-writes its shard ``[lo:hi]`` into them.

    Rank 0 creates the file, the others map it after a barrier; ``close()`` unpins / unlinks.
    The tail of the allocation holds one arrival counter per rank: ``arrive()`` / ``wait_all()`` are
    the barrier that closes a gather, in host shared memory (microseconds; an NCCL barrier is a
    device all-reduce plus a synchronisation, 0.1-0.2 ms of a 1.5 ms job at 8 GPUs)."""

    FIELDS = (("t_final", torch.float64, ()), ("y_final", torch.float64, ("n",)), ("yp_final", torch.float64, ("n",)),
              ("status", torch.int32, ()), ("counters", torch.int32, (8,)))

    def __init__(self, name: str, N: int, n: int, group=None, pin: Optional[bool] = None):
        import os
        self.N, self.n, self.group = int(N), int(n), group
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.epoch = 0
        self.path = os.path.join("/dev/shm", name)
        shapes, off = [], 0
        for key, dt, tail in self.FIELDS:
            shape = (self.N,) + tuple(self.n if d == "n" else d for d in tail)
            nbytes = int(np.prod(shape)) * torch.empty((), dtype=dt).element_size()
            shapes.append((key, dt, shape, off, nbytes))
            off += (nbytes + 255) // 256 * 256
        flags_off = off
        off += (8 * self.world + 255) // 256 * 256
        self.nbytes = max(off, 256)
        if self.rank == 0:
            with open(self.path, "wb") as fh:
                fh.truncate(self.nbytes)
        if dist.is_initialized():
            dist.barrier(group)
        self._raw = torch.from_file(self.path, shared=True, size=self.nbytes, dtype=torch.uint8)
        self._pinned = False
        if pin if pin is not None else torch.cuda.is_available():
            rc = torch.cuda.cudart().cudaHostRegister(self._raw.data_ptr(), self.nbytes, 0)
            if int(rc) != 0:
                raise RuntimeError(f"cudaHostRegister failed ({rc})")
            self._pinned = True
        for key, dt, shape, o, nbytes in shapes:
            setattr(self, key, self._raw[o:o + nbytes].view(dt).view(shape))
        self._flags = self._raw[flags_off:flags_off + 8 * self.world].view(torch.int64)
        self._flags_np = self._flags.numpy()            # plain loads / stores on the shared mapping

    def arrive(self):
        """This rank's shard of the current gather has landed in the shared arrays."""
        self.epoch += 1
        self._flags_np[self.rank] = self.epoch

    def wait_all(self, timeout_s: float = 120.0):
        """Rank 0: spin until every rank has arrived at the current epoch."""
        import time
        t0 = time.perf_counter()
        while int(self._flags_np.min()) < self.epoch:
            if time.perf_counter() - t0 > timeout_s:
                raise TimeoutError("a rank did not deliver its shard")

    def shard(self, lo: int, hi: int):
        """The slice [lo:hi] as the result buffers the host pipeline writes into."""
        from .host_pipeline import HostResultBuffers
        return HostResultBuffers(self.t_final[lo:hi], self.y_final[lo:hi], self.yp_final[lo:hi],
                                 self.status[lo:hi], self.counters[lo:hi])

    def as_dict(self):
        cn = self.counters
        out = dict(t_final=self.t_final, y_final=self.y_final, yp_final=self.yp_final, status=self.status)
        for i, k in enumerate(("nfev", "njev", "nlu", "nsolve", "n_accepted", "n_err_rejected", "n_newton_fail",
                               "n_newton_iter")):
            out[k] = cn[:, i]
        return out

    def close(self):
        import os
        if self._pinned:
            torch.cuda.cudart().cudaHostUnregister(self._raw.data_ptr())
            self._pinned = False
        if dist.is_initialized():
            dist.barrier(self.group)
        if self.rank == 0 and os.path.exists(self.path):
            os.unlink(self.path)


def solve_dae_sharded(problem, t_span, y0, yp0, params, method="Radau", t_eval=None, rtol=1e-3,
                      atol=1e-6, *, group=None, device=None, gather=True, host_result: "SharedHostResult" = None,
                      _solve: Optional[Callable] = None, _solve_host: Optional[Callable] = None, **options):
    """Every rank passes the same full-size host arrays; rank r integrates its block on its own
    GPU; rank 0 receives the gathered ``[N, ...]`` results (dict), other ranks ``None``.

    gather : True / "collective"  one ``dist.gather`` of the device results onto rank 0;
             "host"               every rank runs the host pipeline (``solve_dae_batched_host``) on
                                  its shard with ``host_result.shard(lo, hi)`` as output buffers, i.e.
                                  its GPU DMA-writes the shard into the node-wide shared result
                                  arrays; one barrier; rank 0 returns views of them.  Inputs that
                                  are pinned tensors are copied asynchronously;
             False                no gather, every rank keeps its local result.

    ``_solve(problem, t_span, y0, yp0, params, method=..., ...) -> dict of tensors`` replaces the
    CUDA solve in the CPU (gloo) tests of the sharding logic, ``_solve_host(..., out=buffers)`` the
    host pipeline.
    """
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    as2d = lambda a: a if isinstance(a, torch.Tensor) and a.ndim == 2 else np.atleast_2d(np.asarray(a, dtype=float))   # noqa: E731
    params, y0, yp0 = as2d(params), as2d(y0), as2d(yp0)
    N = params.shape[0]
    lo, hi, per = shard_bounds(N, world, rank)
    if gather == "host":
        if host_result is None or host_result.N != N:
            raise ValueError('gather="host" needs a SharedHostResult for N instances (create it once, reuse it)')
        if t_eval is not None:
            raise NotImplementedError('gather="host" returns final states; use the collective gather with t_eval')
        if hi > lo:
            if _solve_host is None:
                from .host_pipeline import solve_dae_batched_host as _solve_host
                options.setdefault("device", device)
            _solve_host(problem, t_span, y0[lo:hi], yp0[lo:hi], params[lo:hi], method=method, rtol=rtol, atol=atol,
                        out=host_result.shard(lo, hi), **options)
        # the barrier that closes the gather lives in the same shared memory: every rank marks its
        # arrival (its stream is synchronised: the shard is in host memory), rank 0 waits for all
        host_result.arrive()
        if rank == 0:
            host_result.wait_all()
        return host_result.as_dict() if rank == 0 else None
    to_np = lambda a: a.numpy() if isinstance(a, torch.Tensor) else a        # noqa: E731
    params, y0, yp0 = to_np(params), to_np(y0), to_np(yp0)
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
