"""Radau IIA / BDF tables against the reference (radau_constants radau.py:51-134, bdf.py:242-252).

The reference's tables carry ~1e-14 of eigen-solver noise that the step sequences are sensitive to
(SURVEY.md App. E), so `scipy_dae_b200.constants.radau_tables` derives them with the same
NumPy / SciPy primitives.  Pinned here against values frozen from the unmodified reference
(oracle/gen_golden_constants.py -> tests/golden/constants/radau_constants_ref.npz): bit for bit where this
machine's LAPACK reproduces the frozen eigen-decomposition (the build container and the GPU box share
one image), to 1e-13 otherwise; and bit for bit against the LIVE reference wherever it is importable."""
import os

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
FIELDS = {"A": "A", "c": "c", "T": "T", "TI": "TI", "P": "P", "P2": "P2", "b_hat1_implicit": "b_hat1_implicit",
          "v_implicit": "v_implicit", "v_explicit": "v_explicit", "MU_REAL": "mu_real", "MU_COMPLEX": "mu_complex",
          "b_hat_implicit": "b_hat_implicit", "b_hat_explicit": "b_hat_explicit", "p": "order"}


@pytest.mark.parametrize("s", [1, 3, 5, 7])
def test_radau_tables_match_the_frozen_reference_tables(s):
    from scipy_dae_b200.constants import radau_tables
    ref = np.load(os.path.join(HERE, "golden", "constants", "radau_constants_ref.npz"))
    tb = radau_tables(s)
    exact = True
    for rname, ours in FIELDS.items():
        a, b = np.asarray(getattr(tb, ours)), ref[f"s{s}_{rname}"]
        assert a.shape == b.shape, rname
        assert np.allclose(a, b, rtol=1e-13, atol=1e-13), rname
        exact = exact and np.array_equal(a, b)
    same_lapack = np.array_equal(np.asarray(tb.T), ref[f"s{s}_T"])
    assert exact or not same_lapack      # same eigenvectors => every derived table is identical too


@pytest.mark.parametrize("s", [3, 5, 7])
def test_radau_tables_are_bit_identical_to_the_live_reference(s):
    from oracle import reference as REF
    if not REF.available():
        pytest.skip("the reference is not importable here")
    REF.import_reference()
    from scipy_dae.integrate._dae.radau import radau_constants
    from scipy_dae_b200.constants import radau_tables
    names = ("A", "A_inv", "c", "T", "TI", "P", "P2", "b_hat1_implicit", "v_implicit", "v_explicit",
             "MU_REAL", "MU_COMPLEX", "b_hat_implicit", "b_hat_explicit", "p")
    live = dict(zip(names, radau_constants(s)))
    tb = radau_tables(s)
    for rname, ours in FIELDS.items():
        assert np.array_equal(np.asarray(getattr(tb, ours)), np.asarray(live[rname])), rname


def test_bdf_tables_match_the_reference_formulas():
    """bdf.py:242-252: kappa per NDF strategy, gamma_k = sum 1/j, alpha = (1 - kappa) gamma,
    error_const = kappa gamma + 1 / (k + 1)."""
    from scipy_dae_b200.constants import bdf_tables
    kappas = {"stability": [0, -37 / 200, -1 / 9, 0.0834, 0.0665, 0.0551, 0.0464],
              "efficiency": [0, -37 / 200, -1 / 9, -0.0823, -0.0415, 0, 0],
              None: [0, -37 / 200, -1 / 9, 0, 0, 0, 0]}
    for strategy, kappa in kappas.items():
        for max_order in (1, 3, 5, 6):
            tb = bdf_tables(max_order, strategy)
            k = np.asarray(kappa[:max_order + 1], dtype=float)
            gamma = np.hstack((0, np.cumsum(1 / np.arange(1, max_order + 1))))
            assert np.array_equal(np.asarray(tb.kappa)[:max_order + 1], k)
            assert np.array_equal(np.asarray(tb.gamma)[:max_order + 1], gamma)
            assert np.array_equal(np.asarray(tb.alpha)[:max_order + 1], (1 - k) * gamma)
            assert np.array_equal(np.asarray(tb.error_const)[:max_order + 1], k * gamma + 1 / np.arange(1, max_order + 2))
