#!/usr/bin/env python3
"""bench.py — headline benchmark: ensemble solves/sec (Robertson index-1 DAE, Radau IIA s=3,
rtol=atol=1e-6, t in [0, 1e7], k_i = k * exp(0.1 u), u ~ U(-1,1)^3) on N B200s.

    python bench.py --gpus N --steps K --warmup W            # our arm (one rank per GPU under torchrun)
    python bench.py --impl reference --gpus N --steps K ...  # the CPU arm (rank 0 only)

One "step" = one pass of the hot path over one batch: every rank integrates its own
`--instances` (default 2^20) instances from t0 to t_bound in ONE kernel launch (weak scaling:
per-GPU work is fixed, no collective on the data path).

Timed regions
  value : inputs already resident in HBM (SoA); CUDA events around each launch on the launching
          stream; L2 flushed between timed iterations; max over ranks.
  e2e   : the public API call a user makes (solve_dae_batched) with HOST (pinned) inputs, every
          step including the host->device copies, the AoS->SoA relayout, the launch and the
          device->host read-back of states, status and counters.
  cpu_baseline : the UNMODIFIED Python reference (oracle/_ref) over the first 16 384 instances of the
          same ensemble on a multiprocessing.Pool(os.cpu_count()) (kind "reference"); beside it
          cpu_baseline_port, the plain-C oracle port on the same cores.
Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

REPO = os.path.dirname(os.path.abspath(__file__))
if REPO not in sys.path:
    sys.path.insert(0, REPO)

WORKLOAD = dict(problem="robertson", method="Radau", stages=3, rtol=1e-6, atol=1e-6,
                t_span=(0.0, 1e7), sigma=0.1)
METRIC = "ensemble solves/sec (Robertson, Radau s=3)"
UNIT = "solves/s"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--instances", type=int, default=1 << 20, help="instances per GPU")
    ap.add_argument("--cpu-sample", type=int, default=0, help="instances of the CPU sample (0 = auto)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--jac", default="fd", choices=["fd", "analytic"],
                    help="fd = finite-difference Jacobian (the headline, jac=None in the reference); "
                         "analytic = the functor's analytic Jacobians (a `jac` callable in the reference)")
    ap.add_argument("--t-eval-points", type=int, nargs="*", default=[64, 1000],
                    help="also time the ensemble with dense output at these many t_eval points (none = skip); "
                         "1000 log-spaced points is the reference README's own configuration (README.md:54)")
    ap.add_argument("--no-configs", action="store_true", help="skip the BASELINE configs 3-5 (run at --gpus 1 only)")
    ap.add_argument("--no-strong", action="store_true", help="skip the strong-scaling measurement")
    ap.add_argument("--strong-total", type=int, default=1 << 20,
                    help="instances of the strong-scaling job (BASELINE configs[1]: 1M instances sharded over the GPUs)")
    ap.add_argument("--e2e-no-graph", action="store_true",
                    help="queue the host pipeline eagerly instead of replaying its CUDA graph")
    ap.add_argument("--e2e-mode", default="pipeline", choices=["zero_copy", "pipeline"],
                    help="host-facing call reported as e2e: the chunked H2D / kernel / D2H pipeline (default) or ONE "
                         "launch working on the pinned host arrays in place over PCIe; both are timed every run "
                         "(e2e_pipeline, e2e_zero_copy)")
    ap.add_argument("--strong-mode", default="auto", choices=["auto", "zero_copy", "pipeline"],
                    help="host path of the strong-scaling job: zero-copy launches, the chunked pipeline, or auto = "
                         "zero-copy up to 4 GPUs (measured 5.93 vs 6.15 ms at 2 GPUs) and the pipeline at 8, where "
                         "eight GPUs reading and writing host memory in small transactions contend (2.57 vs 2.22 ms)")
    ap.add_argument("--e2e-chunks", type=int, default=16, help="pipeline chunks of the host-facing call")
    ap.add_argument("--e2e-no-taper", action="store_true", help="equal chunk sizes in the host pipeline")
    ap.add_argument("--e2e-streams", type=int, default=3, help="compute streams of the host pipeline")
    return ap.parse_args()


def workload_config(n_per_gpu, n_gpus, jac="fd"):
    return {"workload": f"robertson_dae_ensemble_N{n_per_gpu}_per_gpu", "problem": "robertson",
            "method": "Radau IIA s=3, " + ("FD Jacobian" if jac == "fd" else "analytic Jacobian"), "rtol": 1e-6, "atol": 1e-6, "t_span": [0.0, 1e7],
            "instances_per_gpu": n_per_gpu, "instances_total": n_per_gpu * n_gpus,
            "param_perturbation": "k*exp(0.1*U(-1,1)), default_rng(seed=rank)",
            "t_eval": None, "parallelism": f"instance-sharded x{n_gpus}, no collective",
            "l2": "256 MiB buffer written between timed iterations (L2 flush)"}


# --------------------------------------------------------------------------------------------
# algorithmic flop model (DESIGN.md "Roofline"): counted from the per-instance
# work counters the kernel returns, not from hardware
# --------------------------------------------------------------------------------------------
def algorithmic_flops(cnt: dict, n: int, s: int, c_f: int, fd: int = 1) -> float:
    """cnt: dict of SUMS over instances of nfev, njev, nlu, nsolve, n_newton_iter, n_accepted."""
    ncs = s // 2
    n_lu_sets = cnt["nlu"] / (1 + ncs)
    n_f = cnt["nfev"] + cnt["njev"] * (2 * n + 1) * fd       # every residual evaluation incl. FD columns
    it = cnt["n_newton_iter"]
    solves_real = it + cnt["n_accepted"] + cnt.get("n_err_rejected", 0)   # + error-estimate solves
    solves_cplx = it * ncs
    fl = n_f * c_f
    fl += n_lu_sets * ((2.0 / 3.0) * n ** 3 + ncs * (8.0 / 3.0) * n ** 3)  # real + complex LU
    fl += n_lu_sets * (2 * n * n + ncs * 4 * n * n)                        # matrix assembly
    fl += solves_real * 2 * n * n + solves_cplx * 8 * n * n                # triangular solves
    fl += it * (3 * 2 * s * s * n + 8 * s * n)                             # TI/T/A transforms, updates, norm
    fl += cnt["njev"] * 2 * n * 2 * n * fd                                 # FD quotients
    return float(fl)


def algorithmic_flops_bdf(cnt: dict, n: int, c_f: int, fd: int = 1) -> float:
    """BDF / NDF twin of algorithmic_flops: residuals, LUs of Jy + c Jyp, Newton
    solves and updates, predictor / difference-array work per attempt and change_D rescalings
    (bdf.py:14-29: two q x q by q x n products, q = order + 1; the number of rescalings is not a
    counter, it is taken as one per LU refresh plus one per rejection, at the mean order 3)."""
    attempts = cnt["n_accepted"] + cnt.get("n_err_rejected", 0) + cnt.get("n_newton_fail", 0)
    n_f = cnt["nfev"] + cnt["njev"] * (2 * n + 1) * fd
    fl = n_f * c_f
    fl += cnt["nlu"] * ((2.0 / 3.0) * n ** 3 + 2 * n * n)            # LU + assembly
    fl += cnt["nsolve"] * 2 * n * n                                   # triangular solves
    fl += cnt["n_newton_iter"] * 7 * n                                # yp = c d + psi, y / d updates, norm
    fl += attempts * 13 * n                                           # predictor, psi, error scale
    fl += cnt["n_accepted"] * 8 * n                                   # difference update, order-selection norms
    fl += (cnt["nlu"] + cnt.get("n_err_rejected", 0)) * 4 * 16 * n    # change_D at q = 4
    fl += cnt["njev"] * 2 * n * 2 * n * fd                            # FD quotients
    return float(fl)


# BASELINE.json configs 3-5 (configs[2..4]): (label, problem, method, instances, sigma, rtol, atol, options, oracle sample)
# Ensemble sizes fill the GPU for several rounds of resident systems (amplifier: 1776 warps x 4 systems
# resident, one system runs ~90 ms: 16 384 instances were 2.3 rounds and measured 66 k solves/s where
# 65 536 measure 81 k; profiles/r02_config_sizes.jsonl).
CONFIGS = [
    ("C3 Akzo Nobel, BDF/NDF, rtol=atol=1e-8", "akzo_nobel", "BDF", 1 << 17, 0.05, 1e-8, 1e-8, {}, 512),
    ("C3 Akzo Nobel, Radau s=3, rtol=atol=1e-8", "akzo_nobel", "Radau", 1 << 17, 0.05, 1e-8, 1e-8, {}, 512),
    ("C4 Andrews squeezer (n=26), Radau s=5, rtol=atol=1e-4", "andrews", "Radau", 1 << 14, 0.05, 1e-4, 1e-4, {"stages": 5}, 64),
    ("C5 transistor amplifier (n=8), Radau s=3, rtol=1e-10", "transamp", "Radau", 1 << 16, 0.02, 1e-10, 1e-6, {}, 32),
    ("C5 transistor amplifier (n=8), BDF, rtol=1e-10", "transamp", "BDF", 1 << 16, 0.02, 1e-10, 1e-6, {}, 32),
]


def run_configs(dev, fp64_peak_tflops, nthreads):
    """Configs 3-5 through the product API (device-resident inputs, CUDA events around the call that
    launches ONE kernel), each with the roofline fraction of the same algorithmic flop model and a
    parity sample against the CPU oracle on the first instances (north_star band, step counts)."""
    import torch
    from oracle import oracle as O
    from scipy_dae_b200 import solve_dae_batched, perturbed_params, initial_conditions, get_problem
    rows = []
    for label, problem, method, N, sigma, rtol, atol, opts, n_par in CONFIGS:
        pr = get_problem(problem)
        params = perturbed_params(problem, N, sigma, seed=0)
        y0, yp0 = initial_conditions(problem, params)
        pd, yd, ypd = (torch.from_numpy(np.ascontiguousarray(a)).to(dev) for a in (params, y0, yp0))
        times, r = [], None
        for it in range(3):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize(dev)
            e0.record()
            r = solve_dae_batched(problem, pr.t_span, yd, ypd, pd, method=method, rtol=rtol, atol=atol,
                                  check_finite=False, device=dev, **opts)
            e1.record()
            torch.cuda.synchronize(dev)
            times.append(e0.elapsed_time(e1))
        ms = float(min(times[1:]))
        ok = r.status == 0
        cnt = {k: float(getattr(r, k)[ok].double().sum()) for k in
               ("nfev", "njev", "nlu", "nsolve", "n_accepted", "n_err_rejected", "n_newton_fail", "n_newton_iter")}
        s = int(opts.get("stages", 3))
        flops = (algorithmic_flops(cnt, pr.n, s, pr.flops) if method == "Radau"
                 else algorithmic_flops_bdf(cnt, pr.n, pr.flops))
        ach = flops / (ms * 1e-3) / 1e12
        ref = O.solve_ensemble(problem, params[:n_par], y0[:n_par], yp0[:n_par], pr.t_span, method=method, rtol=rtol,
                               atol=atol, nthreads=nthreads, **opts)
        rn = r.numpy()
        both = (ref["status"] == 0) & (rn.status[:n_par] == 0)
        ratio = (np.abs(rn.y_final[:n_par] - ref["y_final"]) / np.maximum(10 * rtol * np.abs(ref["y_final"]), atol)).max(axis=1)[both]
        dacc = np.abs(rn.n_accepted[:n_par] - ref["n_accepted"])[both] / ref["n_accepted"][both]
        rows.append({"config": label, "problem": problem, "method": method, "options": opts, "instances": N,
                     "kernel": (("radau" if method == "Radau" else "bdf") + ("" if pr.n <= 6 else "_group") + f"_kernel<{problem}>"),
                     "ms": ms, "value": N / (ms * 1e-3), "unit": UNIT, "n_unfinished": int((~ok).sum()),
                     "work_per_solve": {k: v / max(int(ok.sum()), 1) for k, v in cnt.items()},
                     "roofline": {"bound": "fp64", "achieved": ach, "peak": fp64_peak_tflops, "unit": "TFLOP/s",
                                  "frac": ach / fp64_peak_tflops, "flops_per_solve": flops / max(int(ok.sum()), 1)},
                     "parity_vs_oracle": {"sample": int(n_par), "status_equal": bool((rn.status[:n_par] == ref["status"]).all()),
                                          "state_ratio_median": float(np.median(ratio)), "state_ratio_max": float(ratio.max()),
                                          "frac_in_band": float((ratio <= 1).mean()),
                                          "accepted_rel_diff_max": float(dacc.max()),
                                          "frac_accepted_within_5pct": float((dacc <= 0.05).mean())}})
        del pd, yd, ypd, r
    return rows


def ncu_traffic(n_per_gpu: int):
    """dram__bytes_read.sum + dram__bytes_write.sum of the dominant kernel from the committed
    `ncu --set full` capture (profiles/), per launch, if it was taken at this batch size."""
    try:
        d = json.load(open(os.path.join(REPO, "profiles", "r02_headline_traffic.json")))
    except (OSError, ValueError):
        return None
    if int(d.get("instances", -1)) != int(n_per_gpu):
        return None
    return int(d["dram_bytes_read"]) + int(d["dram_bytes_write"])


def compulsory_bytes(n: int, n_params: int) -> int:
    """HBM bytes one instance must move without t_eval: params, y0, yp0 in; y, yp, t, status, 8 counters out."""
    return 8 * (n_params + 2 * n) + 8 * (2 * n + 1) + 4 + 4 * 8


# --------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                 "-lms", "100", "-i", str(self.index)],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, smax, power, reasons = [], [], [], set()
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        for ln in self.lines:
            parts = [p.strip() for p in ln.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0])); smax.append(float(parts[1])); power.append(float(parts[2]))
            except ValueError:
                continue
            for name, v in zip(names, parts[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None,
                "sm_max_mhz": float(max(smax)) if smax else None,
                "power_w_max": float(max(power)) if power else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# --------------------------------------------------------------------------------------------
def bind_to_gpu_numa_node(index: int):
    """Run this rank (and allocate its pinned buffers) on the CPUs NVML reports as local to the
    GPU: host<->device copies from the far socket were 2-4x slower and erratic in the e2e timing."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(index)
        words = (os.cpu_count() + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, words)
        cpus = [64 * w + b for w, m in enumerate(mask) for b in range(64) if (m >> b) & 1]
        allowed = sorted(set(cpus) & os.sched_getaffinity(0))
        if allowed:
            os.sched_setaffinity(0, allowed)
            return f"{len(allowed)} cpus local to gpu {index}"
    except Exception as exc:            # noqa: BLE001 - measurement hygiene only
        return f"unchanged ({type(exc).__name__})"
    return "unchanged"


def cpu_port_run(params, y0, yp0, nthreads):
    from oracle import oracle as O
    t = time.perf_counter()
    r = O.solve_ensemble(WORKLOAD["problem"], params, y0, yp0, WORKLOAD["t_span"], method="Radau",
                         rtol=WORKLOAD["rtol"], atol=WORKLOAD["atol"], nthreads=nthreads)
    dt = time.perf_counter() - t
    assert (r["status"] == 0).all()
    return dt, r


def reference_pool_run(sample: int, budget_s: float, steps: int = 1, warmup: int = 0):
    """The UNMODIFIED Python reference (oracle/_ref, copied by oracle/make_ref.sh) over the first
    instances of the rank-0 ensemble: solve_dae(F, ..., args=(k1, k2, k3)) on a
    multiprocessing.Pool(os.cpu_count()) with BLAS threads pinned to 1.  Runs as its
    own process: the thread-count environment must be in place before NumPy loads, and a process
    that holds a CUDA context must not fork."""
    env = dict(os.environ, OMP_NUM_THREADS="1", OPENBLAS_NUM_THREADS="1", MKL_NUM_THREADS="1")
    env.pop("CUDA_VISIBLE_DEVICES", None)
    cmd = [sys.executable, "-m", "oracle.reference", "--bench", "--problem", WORKLOAD["problem"],
           "--method", WORKLOAD["method"], "--rtol", repr(WORKLOAD["rtol"]), "--atol", repr(WORKLOAD["atol"]),
           "--t-span", repr(WORKLOAD["t_span"][0]), repr(WORKLOAD["t_span"][1]), "--sigma", repr(WORKLOAD["sigma"]),
           "--seed", "0", "--sample", str(sample), "--budget-s", repr(budget_s), "--steps", str(steps),
           "--warmup", str(warmup)]
    r = subprocess.run(cmd, cwd=REPO, env=env, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"oracle.reference failed: {r.stderr[-2000:]}")
    return json.loads(r.stdout.strip().splitlines()[-1])


def cpu_baseline(sample: int, nthreads: int):
    """The reference on the host cores of this box: the first 16 384 instances of the rank-0
    ensemble, bounded to about 30 s of wall time."""
    d = reference_pool_run(sample if sample > 0 else 16384, 30.0)
    return {"value": d["value"], "unit": UNIT, "cores": d["cores"], "kind": "reference",
            "sample": d["sample"], "instances": d["instances"], "source": d["source"],
            "all_finished": d["all_finished"]}


def cpu_baseline_port(nthreads: int):
    """The plain-C oracle port (oracle/dae_oracle.c) on the same host cores: the stronger CPU
    baseline (compiled code, POSIX threads), reported beside the reference."""
    from scipy_dae_b200 import perturbed_params, initial_conditions
    probe = 2048
    p = perturbed_params("robertson", probe, WORKLOAD["sigma"], 0)
    y0, yp0 = initial_conditions("robertson", p)
    dt, _ = cpu_port_run(p, y0, yp0, nthreads)
    sample = int(min(1 << 20, max(4096, 8.0 * probe / dt)))           # ~8 s of CPU work
    p = perturbed_params("robertson", sample, WORKLOAD["sigma"], 0)
    y0, yp0 = initial_conditions("robertson", p)
    dt, _ = cpu_port_run(p, y0, yp0, nthreads)
    return {"value": sample / dt, "unit": UNIT, "cores": nthreads, "kind": "port",
            "sample": f"first {sample} instances of the rank-0 ensemble, oracle/dae_oracle.c "
                      f"(plain-C port of RadauDAE), {nthreads} POSIX threads, {dt:.2f} s"}


def run_reference(args):
    """--impl reference: the reference's own CPU implementation of the path (the unmodified Python
    solve_dae from oracle/_ref) on this box's host cores, each step a bounded sample of the workload."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    budget_s = 150.0 / max(1, args.steps + args.warmup)
    d = reference_pool_run(min(args.instances, args.cpu_sample if args.cpu_sample > 0 else 16384), budget_s,
                           steps=args.steps, warmup=args.warmup)
    value = d["value"]
    cfg = workload_config(args.instances, args.gpus)
    cfg["reference_sample"] = (f"{d['instances']} instances per step (bounded sample of the "
                               f"{args.instances}-instance workload: the first indices of the seed-0 ensemble)")
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * d["seconds_per_step"], "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "impl": "reference", "config": cfg,
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": d["cores"], "kind": "reference",
                             "sample": d["sample"], "instances": d["instances"], "source": d["source"]},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    emit(line)


# --------------------------------------------------------------------------------------------
_JSON_FD = None


def claim_stdout():
    """The contract is ONE JSON line on stdout.  Libraries write there too (NCCL prints its version
    line on the first communicator), so file descriptor 1 is pointed at stderr for the whole run and
    the JSON line goes to a duplicate of the original stdout."""
    global _JSON_FD
    if _JSON_FD is None:
        sys.stdout.flush()
        _JSON_FD = os.dup(1)
        os.dup2(2, 1)


def emit(line: dict):
    sys.stdout.flush()
    os.write(_JSON_FD if _JSON_FD is not None else 1, (json.dumps(line) + "\n").encode())


def main():
    args = parse_args()
    claim_stdout()
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    from scipy_dae_b200 import _lib, solve_dae_batched, perturbed_params, initial_conditions
    from scipy_dae_b200.constants import radau_tables_c
    from scipy_dae_b200.problems import get_problem

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: scipy_dae_b200 has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device(f"cuda:{local_rank}")
    affinity = bind_to_gpu_numa_node(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    L = _lib.load()
    prob = get_problem("robertson")
    pid, n, npar = _lib.problem_lookup("robertson")
    Nper = args.instances
    s = WORKLOAD["stages"]

    # ---- synthetic ensemble of this rank; rank 0 holds the config-2 ensemble (seed 0) ---------
    params = perturbed_params("robertson", Nper, WORKLOAD["sigma"], seed=rank)
    y0, yp0 = initial_conditions("robertson", params)
    h_params = torch.from_numpy(params).pin_memory()
    h_y0 = torch.from_numpy(y0).pin_memory()
    h_yp0 = torch.from_numpy(yp0).pin_memory()

    # ---- device-resident SoA inputs and outputs for the kernel-only measurement ---------------
    d_p = h_params.to(dev).t().contiguous()
    d_y0 = h_y0.to(dev).t().contiguous()
    d_yp0 = h_yp0.to(dev).t().contiguous()
    f64 = dict(dtype=torch.float64, device=dev)
    y_final = torch.empty((n, Nper), **f64)
    yp_final = torch.empty((n, Nper), **f64)
    t_final = torch.empty((Nper,), **f64)
    status = torch.empty((Nper,), dtype=torch.int32, device=dev)
    counters = torch.empty((8, Nper), dtype=torch.int32, device=dev)
    out = _lib.Outputs()
    out.y_final, out.yp_final, out.t_final = y_final.data_ptr(), yp_final.data_ptr(), t_final.data_ptr()
    out.status, out.counters = status.data_ptr(), counters.data_ptr()
    ro = _lib.RadauOpts(rtol=WORKLOAD["rtol"], max_step=float("inf"), first_step=0.0,
                        continuous_error_weight=0.0, jac_recompute_rate=1e-3, deadband_lo=1.0,
                        deadband_hi=1.2, newton_max_iter=0, newton_iter_embedded=1, max_steps=0,
                        analytic_jac=1 if args.jac == "analytic" else 0)
    tb = radau_tables_c(s)
    atol = (ctypes.c_double * n)(*([WORKLOAD["atol"]] * n))
    stream = torch.cuda.current_stream(dev)

    def launch():
        rc = L.b200dae_solve_radau(pid, Nper, Nper, d_y0.data_ptr(), d_yp0.data_ptr(), d_p.data_ptr(),
                                   WORKLOAD["t_span"][0], WORKLOAD["t_span"][1], atol,
                                   ctypes.byref(ro), ctypes.byref(tb), ctypes.byref(out),
                                   ctypes.c_void_p(stream.cuda_stream))
        _lib.check(rc)

    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    # ---- warm-up ----------------------------------------------------------------------------------
    for _ in range(max(args.warmup, 3)):
        launch()
    torch.cuda.synchronize(dev)
    assert bool((status == 0).all()), "bench ensemble did not finish everywhere"

    # ---- value: K launches, inputs resident, events per launch, L2 flushed in between ---------------
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
        time.sleep(0.3)
    launches0 = _lib.launch_count()
    barrier()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    for a, b in ev:
        flush.fill_(1)
        a.record(stream)
        launch()
        b.record(stream)
    barrier()
    gpu_launches = _lib.launch_count() - launches0
    kernel_ms = [a.elapsed_time(b) for a, b in ev]
    total_ms = float(sum(kernel_ms))

    # ---- e2e: public API, host (pinned) buffers, H2D + relayout + launch + D2H every step -----------
    # solve_dae_batched_host pipelines the copies of neighbouring chunks with the kernel of the
    # current one; the timed region still starts with the inputs in host memory and ends with
    # the states, status and counters of every instance in host memory.
    from scipy_dae_b200 import solve_dae_batched_host, alloc_host_result
    host_out = alloc_host_result(Nper, n)

    e2e_graph = not args.e2e_no_graph

    def e2e_pipeline_step():
        return solve_dae_batched_host("robertson", WORKLOAD["t_span"], h_y0, h_yp0, h_params, method="Radau",
                                      rtol=WORKLOAD["rtol"], atol=WORKLOAD["atol"], device=dev,
                                      chunks=args.e2e_chunks, n_streams=args.e2e_streams,
                                      taper=not args.e2e_no_taper, out=host_out, use_graph=e2e_graph,
                                      jac="analytic" if args.jac == "analytic" else None)

    def e2e_zero_copy_step():
        return solve_dae_batched_host("robertson", WORKLOAD["t_span"], h_y0, h_yp0, h_params, method="Radau",
                                      rtol=WORKLOAD["rtol"], atol=WORKLOAD["atol"], device=dev, out=host_out,
                                      zero_copy=True, jac="analytic" if args.jac == "analytic" else None)

    try:
        e2e_pipeline_step()
    except Exception as exc:          # noqa: BLE001 - graph capture is an optimisation, not a requirement
        if not e2e_graph:
            raise
        print(f"bench.py: CUDA-graph capture of the host pipeline failed ({type(exc).__name__}: {exc}); "
              "timing the eagerly queued pipeline instead", file=sys.stderr)
        e2e_graph = False
        torch.cuda.synchronize(dev)

    def time_e2e(step):
        for _ in range(2):
            step()
        barrier()
        ms = []
        for _ in range(args.steps):
            flush.fill_(1)
            torch.cuda.synchronize(dev)
            t0 = time.perf_counter()
            step()
            torch.cuda.synchronize(dev)
            ms.append(1e3 * (time.perf_counter() - t0))
        barrier()
        return ms

    pipe_ms = time_e2e(e2e_pipeline_step)
    zc_ms = time_e2e(e2e_zero_copy_step)
    e2e_ms = zc_ms if args.e2e_mode == "zero_copy" else pipe_ms
    e2e_total_ms = float(sum(e2e_ms))
    clocks = sampler.stop() if rank == 0 else None
    h2d = 8 * Nper * (npar + 2 * n)
    d2h = host_out.nbytes
    assert bool((host_out.status == 0).all()) and torch.equal(host_out.y_final, y_final.t().cpu())

    # ---- work counters -> algorithmic flops / bytes of the dominant (only) kernel --------------------
    csum = counters.to(torch.int64).sum(dim=1).cpu().numpy()
    cnt = dict(nfev=csum[0], njev=csum[1], nlu=csum[2], nsolve=csum[3], n_accepted=csum[4],
               n_err_rejected=csum[5], n_newton_fail=csum[6], n_newton_iter=csum[7])
    flops_per_launch = algorithmic_flops(cnt, n, s, 15, fd=0 if args.jac == "analytic" else 1)
    bytes_per_launch = compulsory_bytes(n, npar) * Nper

    # ---- FP64 pipe peak of THIS device (MEASURED_PEAKS.json has no FP64 figure) ----------------------
    sink = torch.empty(148 * 8 * 256, **f64)
    iters = 1 << 16
    for _ in range(2):
        L.b200dae_fp64_peak_kernel(148 * 8, 256, iters, sink.data_ptr(), ctypes.c_void_p(stream.cuda_stream))
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(stream)
    L.b200dae_fp64_peak_kernel(148 * 8, 256, iters, sink.data_ptr(), ctypes.c_void_p(stream.cuda_stream))
    b.record(stream)
    torch.cuda.synchronize(dev)
    fp64_peak_tflops = 148 * 8 * 256 * iters * 16 / (a.elapsed_time(b) * 1e-3) / 1e12

    # ---- secondary: the same ensemble WITH t_eval output (the HBM-heavier mode) --------
    te_modes = []
    if rank == 0:
        for m, layout in [(int(x), lay) for x in (args.t_eval_points or []) if int(x) > 0 for lay in ("soa", "instance")]:
            te = torch.from_numpy(np.logspace(-6, 7, m)).to(dev)
            try:
                y_eval = torch.empty((m, n, Nper) if layout == "soa" else (Nper, m, n), **f64)
                yp_eval = torch.empty_like(y_eval)
            except RuntimeError as exc:          # not enough free HBM for 2 x 8 m n N bytes
                te_modes.append({"t_eval_points": m, "skipped": f"{type(exc).__name__}"})
                continue
            nev = torch.empty((Nper,), dtype=torch.int32, device=dev)
            out2 = _lib.Outputs()
            out2.y_final, out2.yp_final, out2.t_final = y_final.data_ptr(), yp_final.data_ptr(), t_final.data_ptr()
            out2.status, out2.counters = status.data_ptr(), counters.data_ptr()
            out2.t_eval, out2.m = te.data_ptr(), m
            out2.y_eval, out2.yp_eval, out2.n_eval_valid = y_eval.data_ptr(), yp_eval.data_ptr(), nev.data_ptr()
            if layout == "instance":          # [N][m][n]: every instance's rows one contiguous stream
                out2.eval_stride_row, out2.eval_stride_comp, out2.eval_stride_inst = n, 1, m * n

            def launch_te():
                _lib.check(L.b200dae_solve_radau(pid, Nper, Nper, d_y0.data_ptr(), d_yp0.data_ptr(), d_p.data_ptr(),
                                                 WORKLOAD["t_span"][0], WORKLOAD["t_span"][1], atol, ctypes.byref(ro),
                                                 ctypes.byref(tb), ctypes.byref(out2), ctypes.c_void_p(stream.cuda_stream)))
            launch_te()
            torch.cuda.synchronize(dev)
            ts = []
            for _ in range(2):
                flush.fill_(1)
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record(stream); launch_te(); b.record(stream)
                torch.cuda.synchronize(dev)
                ts.append(a.elapsed_time(b))
            assert bool((nev == m).all())
            ms = float(np.mean(ts))
            nbytes = Nper * (compulsory_bytes(n, npar) + 16 * n * m)
            te_modes.append({"t_eval_points": m, "layout": "[N][m][n]" if layout == "instance" else "[m][n][N]",
                             "ms_per_launch": ms, "solves_per_s": Nper / (ms * 1e-3),
                             "hbm_bytes_per_launch": nbytes, "hbm_gbs": nbytes / (ms * 1e-3) / 1e9,
                             "note": "same ensemble with dense output at m log-spaced points in [1e-6, 1e7] written to "
                                     "HBM (y and yp); lean kernel twin with t_eval output"})
            del y_eval, yp_eval, nev
            torch.cuda.empty_cache()
    te_mode = te_modes[0] if te_modes else None

    # ---- secondary: smaller ensembles (2^16, 2^18 for the size curve) and the
    # divergence stress sigma = 0.5 (rate constants spread over e^-0.5 .. e^0.5), same kernel ----------------
    variants = []
    if rank == 0 and world == 1:
        def timed(fn, reps=3):
            ts = []
            for _ in range(reps):
                flush.fill_(1)
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record(stream); fn(); b.record(stream)
                torch.cuda.synchronize(dev)
                ts.append(a.elapsed_time(b))
            return float(min(ts))
        for Ns in (1 << 16, 1 << 18):
            if Ns >= Nper:
                continue
            def launch_ns(Ns=Ns):                   # the first Ns instances of the resident SoA arrays (ld = Nper)
                _lib.check(L.b200dae_solve_radau(pid, Ns, Nper, d_y0.data_ptr(), d_yp0.data_ptr(), d_p.data_ptr(),
                                                 WORKLOAD["t_span"][0], WORKLOAD["t_span"][1], atol, ctypes.byref(ro),
                                                 ctypes.byref(tb), ctypes.byref(out), ctypes.c_void_p(stream.cuda_stream)))
            launch_ns()
            ms = timed(launch_ns)
            variants.append({"instances": Ns, "sigma": WORKLOAD["sigma"], "ms_per_launch": ms, "solves_per_s": Ns / (ms * 1e-3)})
        p5 = perturbed_params("robertson", Nper, 0.5, seed=0)
        y5, yp5 = initial_conditions("robertson", p5)
        d5 = [torch.from_numpy(np.ascontiguousarray(a)).to(dev).t().contiguous() for a in (y5, yp5, p5)]

        def launch_s5():
            _lib.check(L.b200dae_solve_radau(pid, Nper, Nper, d5[0].data_ptr(), d5[1].data_ptr(), d5[2].data_ptr(),
                                             WORKLOAD["t_span"][0], WORKLOAD["t_span"][1], atol, ctypes.byref(ro),
                                             ctypes.byref(tb), ctypes.byref(out), ctypes.c_void_p(stream.cuda_stream)))
        launch_s5()
        ms = timed(launch_s5)
        torch.cuda.synchronize(dev)
        variants.append({"instances": Nper, "sigma": 0.5, "ms_per_launch": ms, "solves_per_s": Nper / (ms * 1e-3),
                         "n_unfinished": int((status != 0).sum()),
                         "accepted_steps_mean": float(counters[4].double().mean()),
                         "accepted_steps_min_max": [int(counters[4].min()), int(counters[4].max())],
                         "note": "divergence stress: lanes of a warp integrate systems whose rate constants differ by up to e"})
        del d5
        launch()                                    # leave the sigma = 0.1 results in the output buffers
        torch.cuda.synchronize(dev)

    # ---- max over ranks ----------------------------------------------------------------------------------
    t_all = torch.tensor([total_ms, e2e_total_ms, float(sum(pipe_ms)), float(sum(zc_ms))], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t_all, op=dist.ReduceOp.MAX)
    total_ms, e2e_total_ms, pipe_all_ms, zc_all_ms = (float(x) for x in t_all)

    # ---- strong scaling: BASELINE configs[1] literally -- ONE ensemble of --strong-total instances (2^20)
    # sharded over the N GPUs through the product API (distributed.solve_dae_sharded), host inputs to
    # host results, with the single gather INSIDE the timed region: every rank's host pipeline DMA-
    # writes its shard into the node-wide shared, page-locked result arrays (distributed.SharedHostResult;
    # no collective, no NCCL channel set-up), one barrier, done.  Max over ranks.
    strong = None
    if not args.no_strong:
        from scipy_dae_b200.distributed import SharedHostResult, solve_dae_sharded, shard_bounds
        Nt = int(args.strong_total)
        lo, hi, per = shard_bounds(Nt, world, rank)
        # the seed-0 ensemble (rank 0's weak-scaling ensemble when Nt == Nper) in pinned host memory;
        # solve_dae_sharded slices a rank's block out of it (a rank never touches the other blocks)
        p_all = params if (rank == 0 and Nt == Nper) else perturbed_params("robertson", Nt, WORKLOAD["sigma"], seed=0)
        y_all, yp_all = initial_conditions("robertson", p_all)
        hs_p, hs_y0, hs_yp0 = (torch.from_numpy(np.ascontiguousarray(a)).pin_memory() for a in (p_all, y_all, yp_all))
        shared = SharedHostResult(f"b200dae_bench_{os.environ.get('MASTER_PORT', '0')}_{os.getppid()}", Nt, n)
        chunks = max(2, args.e2e_chunks // max(1, min(world, 4)))

        strong_zc = args.strong_mode == "zero_copy" or (args.strong_mode == "auto" and world <= 4)

        def strong_step():
            if strong_zc:       # every rank's launch writes its shard straight into the shared pinned arrays
                return solve_dae_sharded("robertson", WORKLOAD["t_span"], hs_y0, hs_yp0, hs_p, method="Radau",
                                         rtol=WORKLOAD["rtol"], atol=WORKLOAD["atol"], gather="host", host_result=shared,
                                         device=dev, zero_copy=True, jac="analytic" if args.jac == "analytic" else None)
            return solve_dae_sharded("robertson", WORKLOAD["t_span"], hs_y0, hs_yp0, hs_p, method="Radau",
                                     rtol=WORKLOAD["rtol"], atol=WORKLOAD["atol"], gather="host", host_result=shared,
                                     device=dev, chunks=chunks, n_streams=args.e2e_streams, use_graph=True,
                                     jac="analytic" if args.jac == "analytic" else None)
        for _ in range(3):
            strong_step()
        barrier()
        s_ms = []
        for _ in range(args.steps):
            flush.fill_(1)
            barrier()
            t0 = time.perf_counter()
            res_s = strong_step()                      # ends with the barrier that closes the gather
            s_ms.append(1e3 * (time.perf_counter() - t0))
        t_s = torch.tensor(s_ms, dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t_s, op=dist.ReduceOp.MAX)
        s_ms = [float(x) for x in t_s]
        if rank == 0:
            ok_all = bool((res_s["status"] == 0).all())
            same = bool(torch.equal(res_s["y_final"], host_out.y_final)) if Nt == Nper else None
            strong = {"scaling": "strong", "instances_total": Nt, "n_gpus": world, "instances_per_gpu": per,
                      "value": Nt * len(s_ms) / (sum(s_ms) * 1e-3), "unit": UNIT, "ms_per_step": float(np.mean(s_ms)),
                      "ms_each": [round(x, 3) for x in s_ms], "all_finished": ok_all,
                      "matches_single_gpu_results": same,
                      "api": ("distributed.solve_dae_sharded(gather='host', zero_copy=True): every rank's ONE launch reads its "
                              "block of the pinned host inputs and writes its shard of the shared pinned host result arrays "
                              "in place over PCIe -> arrival counters" if strong_zc else
                              f"distributed.solve_dae_sharded(gather='host', chunks={chunks}, use_graph=True): host inputs -> "
                              "per-rank host pipeline -> shards DMA-written into shared pinned host arrays -> barrier"),
                      "h2d_bytes_per_step": 8 * Nt * (npar + 2 * n), "d2h_bytes_per_step": int(shared.nbytes),
                      "gather": "inside the timed region (no collective)"}
        shared.close()
        del hs_p, hs_y0, hs_yp0

    # the collective gather for comparison (rank 0 receives every rank's final states), timed apart
    gather_ms = gather_first_ms = None
    if world > 1:
        parts = [torch.empty_like(y_final) for _ in range(world)] if rank == 0 else None
        # the first NCCL gather also sets up the peer-to-peer channels of every (rank, 0) pair
        # (0.16 / 0.48 / 1.8 s at 2 / 4 / 8 GPUs, once per process): reported apart from the steady one
        barrier()
        t0 = time.perf_counter()
        dist.gather(y_final, parts, dst=0)
        torch.cuda.synchronize(dev)
        gather_first_ms = 1e3 * (time.perf_counter() - t0)
        barrier()
        t0 = time.perf_counter()
        dist.gather(y_final, parts, dst=0)
        torch.cuda.synchronize(dev)
        gather_ms = 1e3 * (time.perf_counter() - t0)

    if rank == 0:
        n_total = Nper * world
        value = n_total * args.steps / (total_ms * 1e-3)
        e2e_value = n_total * args.steps / (e2e_total_ms * 1e-3)
        kernel_s = (sum(kernel_ms) / len(kernel_ms)) * 1e-3
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(REPO, "MEASURED_PEAKS.json")))
        except OSError:
            pass
        hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
        for tm in te_modes:
            if "hbm_gbs" in tm:
                tm["hbm_frac"] = tm["hbm_gbs"] / hbm_peak
        ach_tflops = flops_per_launch / kernel_s / 1e12
        ach_gbs = bytes_per_launch / kernel_s / 1e9
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": total_ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic", "config": workload_config(Nper, world, args.jac),
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": d2h, "ms_per_step": e2e_total_ms / args.steps,
                    "timer": "host perf_counter bracketed by cuda synchronize",
                    "api": ("solve_dae_batched_host(zero_copy=True): ONE launch that reads the pinned (N, n) host inputs "
                            "and writes the pinned host results in place over PCIe (io_layout = 1), check_finite on"
                            if args.e2e_mode == "zero_copy" else
                            f"solve_dae_batched_host(chunks={args.e2e_chunks}, n_streams={args.e2e_streams}, "
                            f"taper={not args.e2e_no_taper}, use_graph={e2e_graph}), pinned host inputs and outputs"),
                    "ms_each": [round(x, 3) for x in e2e_ms], "cpu_affinity": affinity},
            "e2e_pipeline": {"value": Nper * world * args.steps / (pipe_all_ms * 1e-3), "unit": UNIT,
                             "ms_per_step": pipe_all_ms / args.steps,
                             "api": f"solve_dae_batched_host(chunks={args.e2e_chunks}, n_streams={args.e2e_streams}, "
                                    f"taper={not args.e2e_no_taper}, use_graph={e2e_graph}): chunked H2D / relayout / kernel / D2H pipeline"},
            "e2e_zero_copy": {"value": Nper * world * args.steps / (zc_all_ms * 1e-3), "unit": UNIT,
                              "ms_per_step": zc_all_ms / args.steps},
            "gpu_launches": int(gpu_launches),
            "clocks": clocks,
            "roofline": {
                "bound": "fp64",
                "kernel": "radau_kernel<Robertson,3,64 threads,6 blocks/SM,lean>",
                "achieved": ach_tflops, "peak": fp64_peak_tflops, "unit": "TFLOP/s",
                "frac": ach_tflops / fp64_peak_tflops,
                "peak_source": "DFMA-chain microbenchmark run in this process (b200dae_fp64_peak_kernel); "
                               "MEASURED_PEAKS.json has no FP64 figure",
                "flops_per_launch": flops_per_launch, "flops_per_solve": flops_per_launch / Nper,
                "kernel_ms": kernel_s * 1e3,
                "traffic": ncu_traffic(Nper),
                "hbm": {"achieved": ach_gbs, "peak": hbm_peak, "unit": "GB/s", "frac": ach_gbs / hbm_peak,
                        "bytes_per_launch": bytes_per_launch,
                        "peak_source": "MEASURED_PEAKS.json" if peaks else "fallback"},
            },
            "work_per_solve": {k: float(v) / Nper for k, v in cnt.items()},
            "t_eval_mode": te_mode, "t_eval_modes": te_modes,
            "ensemble_variants": variants,
            "strong": strong,
            "gather_ms": gather_ms, "gather_first_ms": gather_first_ms,
        }
        if world == 1 and not args.no_configs:
            line["configs"] = run_configs(dev, fp64_peak_tflops, os.cpu_count() or 1)
        if not args.no_cpu_baseline and world == 1:
            line["cpu_baseline"] = cpu_baseline(args.cpu_sample, os.cpu_count() or 1)
            line["cpu_baseline_port"] = cpu_baseline_port(os.cpu_count() or 1)
        emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
