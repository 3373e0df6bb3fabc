"""Accuracy and special cases of the branch-free FP64 reciprocal / sqrt (csrc/fastmath.cuh)."""
import ctypes

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _run(x):
    from scipy_dae_b200 import _lib
    L = _lib.load()
    xd = torch.from_numpy(x).cuda()
    r = torch.empty_like(xd)
    s = torch.empty_like(xd)
    _lib.check(L.b200dae_selftest_fastmath(xd.numel(), xd.data_ptr(), r.data_ptr(), s.data_ptr(),
                                           ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)))
    torch.cuda.synchronize()
    return r.cpu().numpy(), s.cpu().numpy()


def test_accuracy_within_one_ulp():
    rng = np.random.default_rng(0)
    x = np.concatenate([np.exp(rng.uniform(-300, 300, 200000)), -np.exp(rng.uniform(-300, 300, 50000)),
                        rng.uniform(0.5, 2.0, 100000), np.exp(rng.uniform(-700, -600, 20000))])
    r, s = _run(x)
    ref_r = 1.0 / x
    assert (np.abs(r - ref_r) <= 1.0 * np.spacing(np.abs(ref_r))).all()
    pos = x > 0
    ref_s = np.sqrt(x[pos])
    assert (np.abs(s[pos] - ref_s) <= 1.0 * np.spacing(ref_s)).all()
    assert np.isnan(s[~pos]).all()


def test_special_values():
    x = np.array([0.0, -0.0, np.inf, -np.inf, np.nan, 5e-324, 1e-310, 1e300, 4.0, 1.0])
    r, s = _run(x)
    with np.errstate(divide="ignore", invalid="ignore"):
        assert np.array_equal(r[:5], 1.0 / x[:5], equal_nan=True)
        assert np.isinf(r[5]) and np.isinf(r[6])          # denormals reciprocate to inf (flushed seed)
        assert np.array_equal(s[[0, 1, 2]], np.sqrt(x[[0, 1, 2]]))
        assert np.isnan(s[3]) and np.isnan(s[4])
        assert np.allclose(s[5:], np.sqrt(x[5:]), rtol=4e-16, atol=0)
    assert r[8] == 0.25 and s[8] == 2.0 and r[9] == 1.0 and s[9] == 1.0
