"""world_size-2 gloo test of the sharding / gather logic (CPU).  The CUDA solve is replaced by the
CPU oracle through the `_solve` hook; what is under test is shard_bounds, the per-rank slicing
and the single gather, i.e. exactly the host logic bench.py / solve_dae_sharded run at N > 1."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import REPO


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _oracle_solve(problem, t_span, y0, yp0, params, method="Radau", t_eval=None, rtol=1e-3, atol=1e-6, **opts):
    from oracle import oracle as O
    r = O.solve_ensemble(problem, params, y0, yp0, t_span, method=method, rtol=rtol, atol=atol, **opts)
    out = {k: torch.from_numpy(np.ascontiguousarray(r[k])) for k in ("t_final", "y_final", "yp_final")}
    out["status"] = torch.from_numpy(r["status"].astype(np.int32))
    for k in ("nfev", "njev", "nlu", "nsolve", "n_accepted", "n_err_rejected", "n_newton_fail", "n_newton_iter"):
        out[k] = torch.from_numpy(r[k].astype(np.int32))
    return out


def _oracle_solve_host(problem, t_span, y0, yp0, params, method="Radau", rtol=1e-3, atol=1e-6, out=None, **opts):
    """Stand-in for solve_dae_batched_host: fills the rank's slice of the shared result arrays."""
    opts.pop("device", None)
    r = _oracle_solve(problem, t_span, np.asarray(y0), np.asarray(yp0), np.asarray(params), method=method, rtol=rtol, atol=atol, **opts)
    out.t_final.copy_(r["t_final"]); out.y_final.copy_(r["y_final"]); out.yp_final.copy_(r["yp_final"])
    out.status.copy_(r["status"])
    for i, k in enumerate(("nfev", "njev", "nlu", "nsolve", "n_accepted", "n_err_rejected", "n_newton_fail", "n_newton_iter")):
        out.counters[:, i] = r[k]


def _worker_host(rank, world, port, N, q):
    """gather="host": every rank writes its shard into the node-wide shared-memory result arrays."""
    import sys
    sys.path.insert(0, REPO)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from scipy_dae_b200.distributed import solve_dae_sharded, SharedHostResult
    from scipy_dae_b200.problems import perturbed_params, initial_conditions
    params = perturbed_params("robertson", N, 0.1, seed=3)
    y0, yp0 = initial_conditions("robertson", params)
    shared = SharedHostResult(f"b200dae_test_{port}", N, 3)
    for rep in range(2):                      # the shared arrays are reusable
        res = solve_dae_sharded("robertson", (0.0, 1e3), torch.from_numpy(y0), torch.from_numpy(yp0), torch.from_numpy(params),
                                method="Radau", rtol=1e-6, atol=1e-6, gather="host", host_result=shared,
                                _solve_host=_oracle_solve_host)
    if rank == 0:
        q.put({k: v.numpy().copy() for k, v in res.items()})
    else:
        assert res is None
    shared.close()
    dist.destroy_process_group()


def _worker(rank, world, port, N, q):
    import sys
    sys.path.insert(0, REPO)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from scipy_dae_b200.distributed import solve_dae_sharded
    from scipy_dae_b200.problems import perturbed_params, initial_conditions
    params = perturbed_params("robertson", N, 0.1, seed=3)
    y0, yp0 = initial_conditions("robertson", params)
    res = solve_dae_sharded("robertson", (0.0, 1e3), y0, yp0, params, method="Radau", rtol=1e-6,
                            atol=1e-6, _solve=_oracle_solve)
    if rank == 0:
        q.put({k: v.numpy() for k, v in res.items()})
    else:
        assert res is None
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("N,worker", [(10, _worker), (7, _worker), (1, _worker), (9, _worker_host), (1, _worker_host)])
def test_sharded_solve_matches_single_process(N, worker):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=worker, args=(r, 2, port, N, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = q.get(timeout=120)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    from scipy_dae_b200.problems import perturbed_params, initial_conditions
    params = perturbed_params("robertson", N, 0.1, seed=3)
    y0, yp0 = initial_conditions("robertson", params)
    ref = _oracle_solve("robertson", (0.0, 1e3), y0, yp0, params, rtol=1e-6, atol=1e-6)
    for k, v in ref.items():
        assert got[k].shape == tuple(v.shape), k
        assert np.array_equal(got[k], v.numpy()), k


def test_shard_bounds_cover_and_balance():
    from scipy_dae_b200.distributed import shard_bounds
    for N in (0, 1, 7, 8, 9, 1 << 20, (1 << 20) + 3):
        for W in (1, 2, 4, 8):
            spans = [shard_bounds(N, W, r) for r in range(W)]
            assert spans[0][0] == 0 and spans[-1][1] == N
            for a, b in zip(spans[:-1], spans[1:]):
                assert a[1] == b[0]
            sizes = [hi - lo for lo, hi, _ in spans]
            assert max(sizes) <= -(-N // W) if N else True
            for lo, hi, per in spans:        # instance i -> rank i // ceil(N/W)
                assert all(i // per == r for r, (l2, h2, _) in enumerate(spans) for i in (l2, h2 - 1) if h2 > l2)
