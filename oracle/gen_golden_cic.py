#!/usr/bin/env python3
"""Mint golden vectors for consistent_initial_conditions from the UNMODIFIED reference
(oracle/_ref or /root/reference, scipy_dae/integrate/_dae/common.py:186-372).  TEST INFRASTRUCTURE.

    python oracle/gen_golden_cic.py

Writes tests/golden/cic/<case>.npz: the inconsistent guesses (y0_in, yp0_in), the reference's
outputs (y0, yp0, f) and a status per instance (0 ok, -1 RuntimeError "Convergence failed.",
-2 ValueError "Too many fixed components", -3 ValueError "Index greater than one.")."""
import json
import os
import sys

import numpy as np
import scipy

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
from oracle.reference import import_reference, ref_source  # noqa: E402

import_reference()
from scipy_dae.integrate import consistent_initial_conditions  # noqa: E402  (the unmodified reference)
from scipy_dae_b200.problems import get_problem, perturbed_params, initial_conditions  # noqa: E402

OUT = os.path.join(REPO, "tests", "golden", "cic")


def run(name, problem, N, sigma, seed, rel_y, rel_yp, jac=None, fixed_y0=None, fixed_yp0=None, **kw):
    pr = get_problem(problem)
    rng = np.random.default_rng(seed + 100)
    P = perturbed_params(problem, N, sigma, seed)
    Y0, YP0 = initial_conditions(problem, P)
    Y0 = Y0 * (1 + rel_y * rng.standard_normal(Y0.shape))
    YP0 = YP0 + rel_yp * rng.standard_normal(YP0.shape) * np.maximum(np.abs(YP0), 1e-2)
    ys, yps, fs, st = [], [], [], []
    for i in range(N):
        fun = lambda t, y, yp, i=i: pr.residual(t, y, yp, *P[i])          # noqa: E731
        j = (lambda t, y, yp, i=i: pr.jacobian(t, y, yp, *P[i])) if jac == "analytic" else None
        try:
            y, yp, f = consistent_initial_conditions(fun, 0.0, Y0[i].copy(), YP0[i].copy(), jac=j,
                                                     fixed_y0=fixed_y0, fixed_yp0=fixed_yp0, **kw)
            code = 0
        except RuntimeError:
            y, yp, f, code = (np.full(pr.n, np.nan),) * 3 + (-1,)
        except ValueError as e:
            y, yp, f, code = (np.full(pr.n, np.nan),) * 3 + (-2 if "Too many" in str(e) else -3,)
        ys.append(y); yps.append(yp); fs.append(f); st.append(code)
    meta = dict(case=name, problem=pr.name, jac=jac, fixed_y0=fixed_y0, fixed_yp0=fixed_yp0, options=kw,
                reference="JonasBreuling/scipy_dae v0.1.0 (" + ref_source() + ")", scipy=scipy.__version__,
                numpy=np.__version__, generator="oracle/gen_golden_cic.py")
    os.makedirs(OUT, exist_ok=True)
    np.savez_compressed(os.path.join(OUT, name + ".npz"), params=P, y0_in=Y0, yp0_in=YP0, y0=np.stack(ys),
                        yp0=np.stack(yps), f=np.stack(fs), status=np.array(st), meta=np.array(json.dumps(meta)))
    print(name, "status", st, "max|f|", np.nanmax(np.abs(np.stack(fs))))


def main():
    run("robertson_free", "robertson", 8, 0.1, 0, 1e-3, 1e-3)
    run("robertson_fixed_y01", "robertson", 8, 0.1, 1, 1e-3, 1e-2, fixed_y0=[0, 1])
    run("robertson_fixed_all_y", "robertson", 4, 0.1, 2, 1e-3, 1e-3, fixed_y0=[0, 1, 2])      # -2 quirk
    run("robertson_fixed_yp", "robertson", 4, 0.1, 3, 1e-3, 1e-3, fixed_yp0=[0, 1, 2])
    run("robertson_jac", "robertson", 8, 0.1, 4, 1e-2, 1e-2, jac="analytic")
    run("akzo_free", "akzo_nobel", 8, 0.05, 0, 1e-3, 1e-3)
    run("akzo_fixed_y", "akzo_nobel", 8, 0.05, 1, 1e-3, 1e-3, fixed_y0=[0, 1, 2, 3, 4])
    run("akzo_jac_tight", "akzo_nobel", 4, 0.05, 2, 1e-2, 1e-2, jac="analytic", rtol=1e-12, atol=1e-12)
    run("transamp_free", "transamp", 6, 0.02, 0, 1e-3, 1e-3)
    run("transamp_jac", "transamp", 6, 0.02, 3, 1e-3, 1e-3, jac="analytic")
    run("transamp_fixed_y", "transamp", 6, 0.02, 1, 0.0, 1e-1, fixed_y0=list(range(8)))
    run("andrews_free", "andrews", 4, 0.05, 2, 1e-3, 1e-3)
    run("andrews_big_guess_error", "andrews", 4, 0.05, 3, 5e-2, 1e-1)


if __name__ == "__main__":
    main()
