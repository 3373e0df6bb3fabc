"""Per-step relative drift of the CUDA step sequence against the golden trace of config 1
(run on the GPU box: python scripts/step_drift.py)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from conftest import load_golden
from scipy_dae_b200 import solve_dae_batched

g, meta = load_golden("robertson_radau3_c1")
r = solve_dae_batched("robertson", meta["t_span"], g["y0"], g["yp0"], g["params"],
                      method="Radau", rtol=1e-6, atol=1e-6, record_steps=128).numpy()
k = int(r.n_steps[0])
t, tg = r.t_steps[0, :k], g["t_steps"]
for i in range(1, min(k, len(tg))):
    print(i, f"{tg[i]:.6e}", f"{abs(t[i] - tg[i]) / tg[i]:.2e}",
          f"{np.max(np.abs(r.y_steps[0, :, i] - g['y_steps'][:, i]) / np.maximum(np.abs(g['y_steps'][:, i]), 1e-30)):.2e}")
