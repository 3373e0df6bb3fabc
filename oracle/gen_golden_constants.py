"""Freeze the reference's Radau IIA tables (radau_constants, radau.py:51-134) for s = 1, 3, 5, 7
into tests/golden/constants/radau_constants_ref.npz — TEST INFRASTRUCTURE, run in the build container where
the unmodified reference is importable:

    python oracle/gen_golden_constants.py

tests/test_constants.py pins scipy_dae_b200.constants.radau_tables against this file (and, where
the reference travels as oracle/_ref, bit for bit against the live function)."""
import os
import sys

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
from oracle import reference as REF   # noqa: E402

NAMES = ("A", "A_inv", "c", "T", "TI", "P", "P2", "b_hat1_implicit", "v_implicit", "v_explicit",
         "MU_REAL", "MU_COMPLEX", "b_hat_implicit", "b_hat_explicit", "p")


def main():
    REF.import_reference()
    from scipy_dae.integrate._dae.radau import radau_constants
    out = {}
    for s in (1, 3, 5, 7):
        for name, val in zip(NAMES, radau_constants(s)):
            out[f"s{s}_{name}"] = np.asarray(val)
    path = os.path.join(REPO, "tests", "golden", "constants", "radau_constants_ref.npz")
    np.savez(path, **out)
    print(path, len(out), "arrays")


if __name__ == "__main__":
    main()
