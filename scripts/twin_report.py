#!/usr/bin/env python3
"""Per golden case: distance of a path (the CPU oracle here, or the CUDA path with --cuda) from the
reference beside the distance of the reference from its own twins (tests/parity.py), as JSON lines."""
import json
import os
import sys

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.join(REPO, "tests"))
from conftest import golden_cases, load_golden  # noqa: E402
from parity import instance_ratios  # noqa: E402


def main():
    cuda = "--cuda" in sys.argv
    if cuda:
        from scipy_dae_b200 import solve_dae_batched
    else:
        from oracle import oracle as O
    for case in golden_cases():
        g, meta = load_golden(case)
        te = g.get("t_eval")
        if cuda:
            r = solve_dae_batched(meta["problem"], meta["t_span"], g["y0"], g["yp0"], g["params"], method=meta["method"],
                                  rtol=meta["rtol"], atol=meta["atol"], t_eval=te, **meta["options"]).numpy()
            res = dict(status=r.status, y_final=r.y_final, n_accepted=r.n_accepted, t_final=r.t_final)
        else:
            res = O.solve_ensemble(meta["problem"], g["params"], g["y0"], g["yp0"], meta["t_span"], method=meta["method"],
                                   rtol=meta["rtol"], atol=meta["atol"], t_eval=te, nthreads=8, **meta["options"])
        n = g["y0"].shape[1]
        atol = np.broadcast_to(np.asarray(meta["atol"], dtype=float), (n,))
        ok = (g["status"] == 0) & (np.asarray(res["status"]) == 0)
        line = dict(case=case, path="cuda" if cuda else "oracle", N=int(g["status"].size), status_ref=g["status"].tolist()[:6],
                    status=np.asarray(res["status"]).tolist()[:6])
        if ok.any():
            r = instance_ratios(np.asarray(res["y_final"])[ok], g["y_final"][ok], meta["rtol"], atol)
            line["state_max"] = float(r.max())
            line["acc_rel_max"] = float((np.abs(np.asarray(res["n_accepted"])[ok] - g["n_accepted"][ok]) / g["n_accepted"][ok]).max())
            if "twin_y_final" in g:
                same = g["twin_status"] == g["status"][None]
                d = np.stack([instance_ratios(g["twin_y_final"][k][ok], g["y_final"][ok], meta["rtol"], atol) for k in range(same.shape[0])])
                line["twin_state_max"] = float(np.where(same[:, ok], d, 0).max())
                da = np.abs(g["twin_n_accepted"][:, ok] - g["n_accepted"][ok][None]) / g["n_accepted"][ok][None]
                line["twin_acc_rel_max"] = float(np.where(same[:, ok], da, 0).max())
        if (g["status"] == -1).any():
            line["t_final_ref"] = g["t_final"].tolist()[:6]
            line["t_final"] = np.asarray(res["t_final"]).tolist()[:6]
            if "twin_t_final" in g:
                line["twin_t_final"] = g["twin_t_final"][:, :3].tolist()
        print(json.dumps(line), flush=True)


if __name__ == "__main__":
    main()
