"""DaeSolver adapters: the B200 path behind the reference's plug-in API
``solve_dae(..., method=<DaeSolver subclass>)``.

CPU: the adapters subclass the REAL reference DaeSolver and run through the reference's own
solve_dae driver (dae.py:446-518) with the CPU oracle injected as the solve hook, where the
reference is importable (/root/reference, i.e. the build container).
GPU: the same adapters on a stand-in base class that restates DaeSolver.step / dense_output
(base.py:361-411) drive the CUDA path, because the GPU box has no /root/reference."""
import os
import sys
import types

import numpy as np
import pytest

REF = "/root/reference"


def _oracle_solve(O):
    def solve(problem, t_span, y0, yp0, params, method, t_eval=None, record_steps=0, **options):
        rtol, atol = options.pop("rtol"), options.pop("atol")
        if t_eval is not None:
            r = O.solve_ensemble(problem.name, params[None, :], y0[None, :], yp0[None, :], t_span, method=method,
                                 rtol=rtol, atol=atol, t_eval=np.asarray(t_eval, dtype=float), **options)
            return types.SimpleNamespace(y=np.swapaxes(r["y_eval"], 1, 2), yp=np.swapaxes(r["yp_eval"], 1, 2))
        tr = O.solve_trace(problem.name, params, y0, yp0, t_span, method=method, rtol=rtol, atol=atol,
                           cap=record_steps, **options)
        k = len(tr["t_steps"])
        return types.SimpleNamespace(
            n_steps=np.array([k]), n_accepted=np.array([tr["n_accepted"]]), t_steps=tr["t_steps"][None, :],
            y_steps=tr["y_steps"][None], yp_steps=tr["yp_steps"][None], status=np.array([tr["status"]]),
            nfev=np.array([tr["nfev"]]), njev=np.array([tr["njev"]]), nlu=np.array([tr["nlu"]]))
    return solve


@pytest.mark.skipif(not os.path.isdir(REF), reason="the reference is only mounted in the build container")
@pytest.mark.parametrize("method", ["Radau", "BDF"])
def test_adapter_through_the_reference_driver(method, oracle):
    sys.path.insert(0, REF)
    try:
        from scipy_dae.integrate import solve_dae
        from scipy_dae.integrate._dae.base import DaeSolver, DAEDenseOutput
    finally:
        sys.path.remove(REF)
    from scipy_dae_b200.adapter import make_dae_solver_adapters
    from scipy_dae_b200.problems import get_problem
    B200Radau, B200BDF = make_dae_solver_adapters(DaeSolver, DAEDenseOutput, solve=_oracle_solve(oracle))
    cls = B200Radau if method == "Radau" else B200BDF
    pr = get_problem("robertson")
    y0, yp0 = pr.initial(pr.default_params)
    # the reference's own solver for comparison, same driver
    ref = solve_dae(pr.residual, (0.0, 1e7), y0, yp0, method=method, rtol=1e-6, atol=1e-6, args=pr.default_params)
    sol = solve_dae(pr, (0.0, 1e7), y0, yp0, method=cls, rtol=1e-6, atol=1e-6, params=pr.default_params)
    assert sol.status == 0 and sol.success
    # the oracle behind the hook follows the reference's step sequence up to the drift documented
    # in DESIGN.md §4 (Radau: identical counters; BDF: within 2 %)
    for a, b in ((sol.nfev, ref.nfev), (sol.njev, ref.njev), (sol.nlu, ref.nlu), (sol.t.size, ref.t.size)):
        assert abs(a - b) <= 0.02 * b + 1, (a, b)
    band = np.maximum(1e-5 * np.abs(ref.y[:, -1]), 1e-6) * (4.0 if method == "BDF" else 1.0)
    assert sol.t[-1] == ref.t[-1] == 1e7 and (np.abs(sol.y[:, -1] - ref.y[:, -1]) <= band).all()
    if method == "Radau":
        assert sol.t.shape == ref.t.shape and np.allclose(sol.t[:40], ref.t[:40], rtol=1e-3)
    # t_eval and dense_output go through DaeSolver.dense_output() of every step
    te = np.logspace(-5, 7, 25)
    ref2 = solve_dae(pr.residual, (0.0, 1e7), y0, yp0, method=method, rtol=1e-6, atol=1e-6, args=pr.default_params,
                     t_eval=te, dense_output=True)
    sol2 = solve_dae(pr, (0.0, 1e7), y0, yp0, method=cls, rtol=1e-6, atol=1e-6, params=pr.default_params,
                     t_eval=te, dense_output=True)
    slack = 4.0 if method == "BDF" else 1.0
    assert (np.abs(sol2.y - ref2.y) <= slack * np.maximum(1e-5 * np.abs(ref2.y), 1e-6)).all()
    tt = np.array([3.0, 0.02, 5e5])
    a, b = sol2.sol(tt)[0], ref2.sol(tt)[0]
    assert a.shape == b.shape == (3, 3) and (np.abs(a - b) <= slack * np.maximum(1e-5 * np.abs(b), 1e-6)).all()
    with pytest.raises(TypeError):
        solve_dae(pr.residual, (0.0, 1.0), y0, yp0, method=cls, rtol=1e-6, atol=1e-6)


class _StandInDaeSolver:
    """DaeSolver.step / dense_output / step_size restated (base.py:355-411) for the GPU box."""
    @property
    def step_size(self):
        return None if self.t_old is None else np.abs(self.t - self.t_old)

    def step(self):
        if self.status != "running":
            raise RuntimeError("Attempt to step on a failed or finished solver.")
        if self.n == 0 or self.t == self.t_bound:
            self.t_old, self.t, self.status = self.t, self.t_bound, "finished"
            return None
        t = self.t
        success, message = self._step_impl()
        if not success:
            self.status = "failed"
        else:
            self.t_old = t
            if self.direction * (self.t - self.t_bound) >= 0:
                self.status = "finished"
        return message

    def dense_output(self):
        if self.t_old is None:
            raise RuntimeError("Dense output is available after a successful step was made.")
        return self._dense_output_impl()


class _StandInDenseOutput:
    def __init__(self, t_old, t):
        self.t_old, self.t, self.t_min, self.t_max = t_old, t, min(t, t_old), max(t, t_old)

    def __call__(self, t):
        return self._call_impl(np.asarray(t))


@pytest.mark.gpu
@pytest.mark.parametrize("method", ["Radau", "BDF"])
def test_adapter_drives_the_cuda_path(method):
    from conftest import load_golden
    from scipy_dae_b200.adapter import make_dae_solver_adapters
    from scipy_dae_b200.problems import get_problem
    Radau, BDF = make_dae_solver_adapters(_StandInDaeSolver, _StandInDenseOutput)
    g, meta = load_golden("robertson_radau3_c1" if method == "Radau" else "robertson_bdf_c1")
    pr = get_problem("robertson")
    solver = (Radau if method == "Radau" else BDF)(pr, 0.0, g["y0"][0], g["yp0"][0], 1e7, rtol=1e-6, atol=1e-6,
                                                 params=g["params"][0])
    ts, ys = [solver.t], [solver.y]
    mids = []
    while solver.status == "running":                    # the loop of dae.py:446-518
        assert solver.step() is None
        ts.append(solver.t); ys.append(solver.y)
        if len(ts) in (5, 30):
            tm = 0.5 * (solver.t_old + solver.t)
            y_mid, yp_mid = solver.dense_output()(tm)
            mids.append((tm, y_mid))
    assert solver.status == "finished" and solver.t == 1e7
    assert abs(len(ts) - len(g["t_steps"])) <= 0.05 * len(g["t_steps"]) + 2
    band = np.maximum(1e-5 * np.abs(g["y_final"][0]), 1e-6) * (4.0 if method == "BDF" else 1.0)
    assert (np.abs(ys[-1] - g["y_final"][0]) <= band).all()
    assert abs(solver.nfev - g["nfev"][0]) <= 0.05 * g["nfev"][0]
    for tm, y_mid in mids:
        assert abs(y_mid.sum() - 1.0) < 1e-5                  # interpolant respects the constraint
