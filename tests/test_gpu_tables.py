"""The Radau launcher rejects tables that are not in the reference's normalisation (the kernel
skips the multiplications by the unit last row of T and tests F for finiteness through TI[0])."""
import ctypes

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _call(tb):
    from scipy_dae_b200 import _lib
    L = _lib.load()
    pid, n, npar = _lib.problem_lookup("robertson")
    N = 32
    dev = torch.device("cuda:0")
    f64 = dict(dtype=torch.float64, device=dev)
    p = torch.tensor([[0.04], [1e4], [3e7]], **f64).repeat(1, N).contiguous()
    y0 = torch.tensor([[1.0], [0.0], [0.0]], **f64).repeat(1, N).contiguous()
    yp0 = torch.tensor([[-0.04], [0.04], [0.0]], **f64).repeat(1, N).contiguous()
    y_final, yp_final = torch.empty((n, N), **f64), torch.empty((n, N), **f64)
    t_final = torch.empty((N,), **f64)
    status = torch.empty((N,), dtype=torch.int32, device=dev)
    counters = torch.empty((8, N), dtype=torch.int32, device=dev)
    out = _lib.Outputs()
    out.y_final, out.yp_final, out.t_final = y_final.data_ptr(), yp_final.data_ptr(), t_final.data_ptr()
    out.status, out.counters = status.data_ptr(), counters.data_ptr()
    ro = _lib.RadauOpts(rtol=1e-6, max_step=float("inf"), first_step=0.0, continuous_error_weight=0.0,
                        jac_recompute_rate=1e-3, deadband_lo=1.0, deadband_hi=1.2, newton_max_iter=0,
                        newton_iter_embedded=1, max_steps=0, analytic_jac=0)
    atol = (ctypes.c_double * n)(*([1e-6] * n))
    rc = L.b200dae_solve_radau(pid, N, N, y0.data_ptr(), yp0.data_ptr(), p.data_ptr(), 0.0, 1.0, atol,
                               ctypes.byref(ro), ctypes.byref(tb), ctypes.byref(out),
                               ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()
    return rc, status.cpu().numpy()


def test_reference_tables_are_accepted_and_tampered_ones_rejected():
    from scipy_dae_b200 import _lib
    from scipy_dae_b200.constants import radau_tables_c
    rc, status = _call(radau_tables_c(3))
    assert rc == 0 and (status == 0).all()
    tb = radau_tables_c(3)
    tb.T[2 * 3 + 1] = 0.25                     # last row of T no longer (1, 0, 1)
    rc, _ = _call(tb)
    assert rc != 0
    with pytest.raises(ValueError, match="normalisation"):
        _lib.check(rc)
    tb = radau_tables_c(3)
    tb.TI[1] = 0.0                             # a zero in the first row of TI
    rc, _ = _call(tb)
    assert rc != 0
