// radau.cuh — the whole adaptive Radau IIA integration of one system per thread.
//
// One launch replaces, for N independent instances, RadauDAE.__init__ (radau.py:393-463,
// base.py:160-242), the `while status is None: solver.step()` loop of solve_dae with its t_eval
// slicing (dae.py:446-518) and RadauDAE._step_impl (radau.py:465-659) with everything below it:
// solve_collocation_system (radau.py:137-195), predict_factor (radau.py:198-240), the FD Jacobian
// (fd_jacobian.cuh), the real + complex LUs (dense_lu.cuh), the embedded error estimate
// (radau.py:568-594) and the collocation-polynomial dense output (radau.py:661-699).
//
// Control flow.  The reference nests four loops (steps > attempts > Jacobian refresh > Newton).
// Here the three outer ones are flattened into ONE loop whose trip is "one simplified-Newton
// solve", with the FD Jacobian, the LU factorisations, the error estimate and the accept phase
// each appearing at exactly one code site, guarded by per-lane flags.  Lanes of a warp that are
// at different points of their own step sequence (a halved step here, a Jacobian refresh there)
// therefore still execute the same phase code together instead of serialising through
// different call sites; only the trip counts differ.  Every quirk of the reference that moves
// step sequences is kept and marked "quirk".
#pragma once
#include "rt_compat.cuh"
#include "../../include/b200dae.h"
#include "fastmath.cuh"
#include "events.cuh"
#include "fd_jacobian.cuh"

namespace b200dae {

struct RadauLaunch {
    b200dae_radau_tables tb;
    double t0, t_bound, dir;
    double rtol, max_step, first_step;
    double cew;                 // continuous_error_weight
    double jac_rate, db_lo, db_hi;
    double newton_tol;
    double inv_mu_real;         // 1 / tb.mu_real, formed once on the host
    double atol[B200DAE_MAX_N];
    double thr[B200DAE_MAX_N];  // atol / rtol (num_jac threshold, base.py:277)
    int32_t newton_max_iter, newton_iter_embedded;
    int64_t max_steps;
    int32_t analytic_jac;       // host-side dispatch only (kJac template parameter)
    int64_t N, ld;
    const double *y0, *yp0, *params;
    b200dae_outputs out;
    EventLaunch ev;             // ev.spec.n_events == 0: no events
    double* jscratch;           // group kernels with the Jacobians in global memory (RadauGroupLayout: kGJ): N blocks, else NULL
    int32_t io_aos;             // 1: y0 / yp0 / params / y_final / yp_final / counters are instance-major ([ld][n], [ld][p], [ld][8])
};

// Index of (component j of instance i) in the main input / output arrays: [n][ld] (default) or, with
// io_aos, [ld][n] -- the layout of the reference's own arrays, which lets a launch read and write
// pinned HOST arrays of shape (N, n) in place (host_pipeline.py: zero_copy).  Prologue / epilogue only.
__device__ __forceinline__ int64_t io_index(int aos, int64_t ld, int width, int j, int64_t inst) {
    return aos ? inst * width + j : (int64_t)j * ld + inst;
}

__device__ __forceinline__ bool finite_d(double x) { return fabs(x) <= 1.7976931348623157e308; }

template <int kBits = 5>
__device__ __forceinline__ double powi(double x, int k) { return powi_bits<kBits>(x, k); }   // 0 <= k < 2^kBits

// predict_factor, radau.py:198-240 (KAPPA = 1).  inv_sp1 = 1 / (s + 1).
template <int S>
__device__ __forceinline__ double root_sp1(double x) {      // x ** (1 / (S + 1))
    if constexpr (S == 3) return fast_sqrt(fast_sqrt(x));
    else if constexpr (S == 7) return fast_sqrt(fast_sqrt(fast_sqrt(x)));
    else return pow(x, 1.0 / (S + 1));
}

template <int S>
__device__ __forceinline__ double predict_factor(double h_abs, double h_abs_old, double err,
                                                 double err_old, bool have_old) {
    double mult = 1.0;
    // err_old / err and h_abs / h_abs_old share one reciprocal each; err == 0 -> factor = inf -> 1 + pi/2
    const double r_err = fast_rcp(err);
    if (have_old && err != 0.0) mult = h_abs * fast_rcp(h_abs_old) * root_sp1<S>(err_old * r_err);
    const double factor = min_sel(mult, 1.0) * root_sp1<S>(r_err);     // min(1, nan) = 1 like Python's min
    return 1.0 + atan(factor - 1.0);
}

#ifndef B200DAE_PF_ROOTS
#define B200DAE_PF_ROOTS 1
#endif
#ifndef B200DAE_PIVOT_ASM
#define B200DAE_PIVOT_ASM 1
#endif
#ifndef B200DAE_FIN_LATE
#define B200DAE_FIN_LATE 1
#endif

#ifndef B200DAE_EVAL_ALIGNED
#define B200DAE_EVAL_ALIGNED 1
#endif
// B200DAE_PF_ROOTS = 1 (s = 3): err ** (1/4) and err ** (-1/4) from one sqrt and one coupled
// sqrt / rsqrt, and (err_old / err) ** (1/4) as err_old ** (1/4) * err ** (-1/4) with the first
// factor carried over from the step that produced err_old: 2 root evaluations per attempt
// instead of 4 and no reciprocal.  `q` returns err ** (1/4) for the caller to keep.
template <int S>
__device__ __forceinline__ double predict_factor(double h_abs, double h_abs_old, double err,
                                                 double err_old, bool have_old, double q_old, double& q) {
    double mult = 1.0;
    if constexpr (S == 3 && B200DAE_PF_ROOTS) {
        double rq;                                   // err ** (-1/4); err == 0 -> inf -> 1 + pi/2
        q = fast_sqrt_rsqrt<false>(fast_sqrt<false>(err), rq);   // err < 1e-200 acts like err = 0: same limit
        if (have_old && err != 0.0) mult = h_abs * fast_rcp(h_abs_old) * (q_old * rq);
        const double factor = min_sel(mult, 1.0) * rq;
        return 1.0 + atan(factor - 1.0);
    }
    q = 0.0;
    return predict_factor<S>(h_abs, h_abs_old, err, err_old, have_old);
}

// kLean: the launch uses the default error estimator (newton_iter_embedded = 1,
// continuous_error_weight = 0) and asks for neither t_eval output nor a step trace.  Those options
// are then compiled out, which shrinks the hot kernel (instruction-cache misses were 11 % of the
// warp stalls of the all-options kernel).
// B200DAE_REG_MAXNREG (experiments): an explicit register cap instead of the one ptxas derives
// from the launch bounds (which rounds 7 blocks x 64 threads down to 128 registers).
#ifdef B200DAE_REG_MAXNREG
#define B200DAE_RADAU_BOUNDS(T, MB) __maxnreg__(B200DAE_REG_MAXNREG)
#else
#define B200DAE_RADAU_BOUNDS(T, MB) __launch_bounds__(T, MB)
#endif
// kEval: t_eval output compiled in (always in the full-options kernel; the lean kernel comes with
// and without it, so that dense output at t_eval does not pull in the other option paths).
// kPowBits: bits of the exponent newton_max_iter - k in the rate test (3 when newton_max_iter <= 8,
// the launcher's choice for the default 7: two squarings and two select-multiplies less per iteration).
// kIoStage: the instance-major I/O of the lean kernel staged through shared memory (see below); its own
// instantiation, so that the default launches carry none of it (the staging indices cost 0.045 ms of the
// 10.7 ms headline launch when compiled into the same kernel).
template <class P, int S, int kThreads, int kMinBlocks, bool kSharedCold, bool kLean, bool kJac = false,
          bool kEval = !kLean, int kPowBits = 5, bool kIoStage = false>
__global__ void B200DAE_RADAU_BOUNDS(kThreads, kMinBlocks)
radau_kernel(const __grid_constant__ RadauLaunch L) {
    constexpr int N = P::N, NP = P::NP, NCS = S / 2;
    constexpr int U = Regime<N>::kUnroll;
    // Instance-major I/O (io_aos) of the lean kernel goes THROUGH SHARED MEMORY: the block's inputs are
    // one contiguous run of kThreads * n doubles per array, loaded with consecutive lanes on consecutive
    // words and picked up per thread from shared memory (results the other way round).  Per-thread
    // accesses of 8 bytes at a 24-byte stride are fine in HBM but not over PCIe, which is where this
    // layout matters (zero-copy launches on pinned host arrays): 11.48 ms per 2^20 Robertson instances
    // with direct indexing, 10.7x ms staged (SoA host arrays: 10.74 ms).  Threads past the end of the
    // ensemble stay alive for the warp barriers and store nothing.
    constexpr bool kStage = kIoStage && kSharedCold && kLean && !kEval;
    const int64_t inst0 = (int64_t)blockIdx.x * kThreads;
    const bool staged = kStage && L.io_aos != 0;
    if (!staged && inst0 + threadIdx.x >= L.N) return;
    const bool valid = inst0 + threadIdx.x < L.N;
    const int64_t inst = valid ? inst0 + threadIdx.x : L.N - 1;
    const int64_t ld = L.ld;
    const b200dae_radau_tables& tb = L.tb;

    // "Cold" per-instance state (touched once per step or per LU, not per Newton iteration):
    // Jacobians, FD step factors, dense-output coefficients, complex LU factors.  With
    // kSharedCold they live in shared memory [element][thread], which frees ~100 registers and
    // lets a third warp per scheduler hide the 8.7-cycle DFMA latency.
    constexpr int kCold = 2 * N * N + 2 * N + N * S + 2 * NCS * N * N;
    static_assert(!kStage || kCold >= 2 * N + NP + 4, "the staging of instance-major I/O reuses the warp's cold-state slots");
    // up to 48 KB the block's cold state is a static array; beyond that (4 < n <= B200DAE_REG_MAXN:
    // 32-thread blocks) it is the launch's dynamic shared memory
    constexpr bool kDynCold = kSharedCold && ((size_t)kCold * kThreads * sizeof(double) > 48 * 1024);
    extern __shared__ __align__(16) double s_cold_dyn[];
    __shared__ double s_cold_static[(kSharedCold && !kDynCold) ? kCold * kThreads : 1];
    double* const s_cold = kDynCold ? s_cold_dyn : s_cold_static;

    // ---- per-instance state -------------------------------------------------------------
    double p[NP > 0 ? NP : 1], y[N], yp[N];
    // staging area of a WARP: its own 32 slots of the first cold-state elements (the cold block is laid
    // out [element][thread]), so that only the warp's own lanes ever touch it: __syncwarp() suffices and a
    // warp retires without waiting for the other warp of its block
    const int wbase = (int)(threadIdx.x & ~31u);
    const int64_t winst0 = inst0 + wbase;                  // first instance of this warp
    const int wn = (int)((L.N - winst0 < 32) ? ((L.N - winst0 > 0) ? (L.N - winst0) : 0) : 32);
    auto stage = [&](int i) -> double& { return s_cold[(i >> 5) * kThreads + wbase + (i & 31)]; };
    const int wl = (int)(threadIdx.x & 31u);
    if (staged) {
        constexpr int oP = 0, oY = 32 * NP, oYp = oY + 32 * N;
        for (int e = wl; e < wn * NP; e += 32) stage(oP + e) = L.params[winst0 * NP + e];
        for (int e = wl; e < wn * N; e += 32) { stage(oY + e) = L.y0[winst0 * N + e]; stage(oYp + e) = L.yp0[winst0 * N + e]; }
        __syncwarp();
        const int me = valid ? wl : 0;
#pragma unroll U
        for (int j = 0; j < NP; ++j) p[j] = stage(oP + me * NP + j);
#pragma unroll U
        for (int j = 0; j < N; ++j) { y[j] = stage(oY + me * N + j); yp[j] = stage(oYp + me * N + j); }
        __syncwarp();                                      // the slots are the cold state from here on
    } else {
#pragma unroll U
        for (int j = 0; j < NP; ++j) p[j] = L.params[io_index(L.io_aos, ld, NP, j, inst)];
#pragma unroll U
        for (int j = 0; j < N; ++j) { y[j] = L.y0[io_index(L.io_aos, ld, N, j, inst)]; yp[j] = L.yp0[io_index(L.io_aos, ld, N, j, inst)]; }
    }

    double r_cold[kSharedCold ? 1 : kCold];
    using V = Vec<kSharedCold, kThreads>;
    const V Jy{kSharedCold ? (s_cold + threadIdx.x) : r_cold};
    const V Jyp = Jy + N * N, factor_y = Jyp + N * N, factor_yp = factor_y + N;
    const V ZP = factor_yp + N;                  // [N][S], dense output of the last accepted step
    const V LUcr0 = ZP + N * S, LUci0 = LUcr0 + NCS * N * N;
    // (Tried on B200: Jy / Jyp and the re / im parts of the complex factors interleaved pairwise,
    // [entry][thread][2], so that one LDS.128 / STS.128 moves a pair: 48 of the 103 + 66 shared
    // accesses became 128-bit, 32 instructions less, and the launch got slower, 12.49 ms instead of
    // 12.30 ms per 2^20 solves -- a 128-bit access takes four wavefronts and holds its
    // short-scoreboard slot longer.)
    // (Tried on B200, round 2: the real factors in shared memory as well, to free 18 registers for
    // the scheduler: 11.03 instead of 10.87 ms per 2^20 solves; with 7 blocks of 144 registers 14.8 ms;
    // 128-thread blocks x 3: 13.8 ms.)
    double LUr[N * N];
    int pivr[N], pivc[NCS > 0 ? NCS : 1][N];
    double Y[S][N], Yp[S][N], W[S][N], G[S][N];
    double h_inv_scale[N];           // h / scale of the current attempt
#pragma unroll U
    for (int j = 0; j < N; ++j) factor_y[j] = factor_yp[j] = kNjInitFactor;

    const double dir = L.dir, rtol = L.rtol;
    double t = L.t0, t_old = L.t0, h_old = 1.0;
    // select_initial_step, common.py:129-183 (or first_step, base.py:172-176)
    double h_abs;
    if (L.first_step > 0.0) {
        h_abs = L.first_step;
    } else {
        double e = 0.0;
#pragma unroll U
        for (int j = 0; j < N; ++j) {
            const double v = fabs(yp[j] * fast_rcp(fmax(fabs(y[j]), L.thr[j])));
            if (v > e || v != v) e = v;
        }
        const double rh = e * fast_rcp(fast_sqrt(rtol) * 0.8);
        h_abs = fmin(L.max_step, fabs(L.t_bound - L.t0));
        if (h_abs * rh > 1.0) h_abs = fast_rcp(rh);
        h_abs = fmax(h_abs, 16.0 * kEps * fabs(L.t0));
    }
    double h_abs_old = 0.0, err_old = 0.0, q_old = 0.0;   // q_old = err_old ** (1/4) (B200DAE_PF_ROOTS)
    bool have_old = false, have_sol = false;
    bool need_jac = true;            // first Jacobian in the constructor, base.py:297
    bool current_jac = true;         // radau.py:460 (quirk 7)
    bool lu_valid = false, step_begin = true, retry = false;
    int32_t nfev = 0, njev = 0, nlu = 0, nsolve = 0, nacc = 0, nrej = 0, nfail = 0, niter = 0;
    int32_t status = valid ? B200DAE_ST_RUNNING : B200DAE_ST_FINISHED;   // threads past the end (staged I/O only) skip the loop
    int64_t ie = 0;                  // next t_eval index
    int32_t nst = 0;                 // rows written to the step trace
    const int newton_max_iter = L.newton_max_iter;

    if (!kLean && L.out.t_steps != nullptr && L.out.steps_cap > 0) {
        L.out.t_steps[inst] = t;
#pragma unroll U
        for (int j = 0; j < N; ++j) {
            if (L.out.y_steps) L.out.y_steps[j * ld + inst] = y[j];
            if (L.out.yp_steps) L.out.yp_steps[j * ld + inst] = yp[j];
        }
        nst = 1;
    }
    if (L.t0 == L.t_bound) { status = B200DAE_ST_FINISHED; need_jac = false; njev = 1; }  // base.py:376-381
    EventState evs;                  // g = event(t0, y0, yp0), dae.py:436
    if (!kLean) {
#pragma unroll
        for (int k = 0; k < B200DAE_MAX_EVENTS; ++k) {
            evs.count[k] = evs.found[k] = 0;
            evs.g[k] = (k < L.ev.spec.n_events) ? event_quantity<P, N>(L.ev.spec, k, t, y, yp, p) - L.ev.spec.value[k] : 0.0;
        }
    }

    // per-step locals that survive across attempts of the same step
    double hl = 0.0, hl_old = 0.0, errl_old = 0.0, ql_old = 0.0, min_step = 0.0, factor = 0.0;
    double h = 0.0, t_new = 0.0;
    bool havel_old = false, have_factor = false;

    while (status == B200DAE_ST_RUNNING) {
        // ---- phase J: finite-difference Jacobian at the current (t, y, yp) ----------------
        if (need_jac) {
            if constexpr (kJac) analytic_jacobian<P, N>(t, y, yp, p, Jy, Jyp);   // jac=callable, base.py:299-331
            else fd_jacobian<P, V>(t, y, yp, p, L.thr, factor_y, factor_yp, Jy, Jyp);
            ++njev;
            need_jac = false;
        }
        // ---- phase B: a new step begins (radau.py:489-512) --------------------------------
        if (step_begin) {
            min_step = 10.0 * fabs(next_toward(t, dir) - t);
            if (h_abs > L.max_step) { hl = L.max_step; havel_old = false; }
            else if (h_abs < min_step) { hl = min_step; havel_old = false; }
            else { hl = h_abs; havel_old = have_old; }
            hl_old = h_abs_old; errl_old = err_old; ql_old = q_old;
            have_factor = false;
            step_begin = false;
        }
        // ---- phase A: a new attempt (radau.py:514-532); skipped on a Jacobian-refresh retry,
        //      which restarts Newton from the FAILED iterate (quirk 1: Yp0 is aliased) ---------
        if (!retry) {
            if (hl < min_step) { status = B200DAE_ST_TOO_SMALL_STEP; break; }
            h = hl * dir;
            t_new = t + h;
            if (dir * (t_new - L.t_bound) > 0.0) t_new = L.t_bound;
            h = t_new - t;
            hl = fabs(h);
            if (!have_sol) {
#pragma unroll U
                for (int i = 0; i < S; ++i)
#pragma unroll U
                    for (int j = 0; j < N; ++j) Yp[i][j] = yp[j];
            } else {
                // extrapolate the previous collocation polynomial's derivative (radau.py:531,693)
                const double rho = fast_rcp_nz(h_old);
#pragma unroll
                for (int i = 0; i < S; ++i) {
                    const double x = (fma(h, tb.C[i], t) - t_old) * rho;
                    double w[S];
                    double xk = rho;
#pragma unroll
                    for (int k = 0; k < S; ++k) { w[k] = (k + 1) * xk; xk *= x; }
#pragma unroll U
                    for (int j = 0; j < N; ++j) {
                        double acc = 0.0;
#pragma unroll
                        for (int k = 0; k < S; ++k) acc = fma(ZP[j * S + k], w[k], acc);
                        Yp[i][j] = acc;
                    }
                }
            }
#pragma unroll U
            for (int j = 0; j < N; ++j) h_inv_scale[j] = h * fast_rcp_scale(fma(fabs(y[j]), rtol, L.atol[j]));
        }
        retry = false;

        // ---- phase F: iteration matrices and their LUs (radau.py:536-538) --------------------
        if (!lu_valid) {
            FiniteAcc fin;
            const double cr = h * tb.mu_real;
            if constexpr (kSharedCold && B200DAE_PIVOT_ASM) {
                // pivoted assembly (dense_lu.cuh: first_pivot_order): column 0 first, then the rows in
                // the order of LAPACK's first interchange, read through run-time row offsets
                double mag[N];
                int r[N];
#pragma unroll
                for (int i = 0; i < N; ++i) mag[i] = fabs(fma(cr, Jy[i * N], Jyp[i * N]));
                first_pivot_order<N>(mag, r, pivr[0]);
#pragma unroll
                for (int k = 0; k < N; ++k) {
                    const V jy = Jy + r[k] * N, jyp = Jyp + r[k] * N;
#pragma unroll
                    for (int j = 0; j < N; ++j) {
                        const double v = fma(cr, jy[j], jyp[j]);
                        fin.add(v);
                        LUr[k * N + j] = v;
                    }
                }
                if (!fin.all_finite()) { status = B200DAE_ST_NONFINITE; break; }
                lu_factor<N, true>(LUr, pivr);
#pragma unroll
                for (int c = 0; c < NCS; ++c) {
                    const double mr = h * tb.mu_re[c], mi = h * tb.mu_im[c];
                    double Ar[N * N], Ai[N * N];
                    int rc[N];
#pragma unroll
                    for (int i = 0; i < N; ++i) mag[i] = fabs(fma(mr, Jy[i * N], Jyp[i * N])) + fabs(mi * Jy[i * N]);
                    first_pivot_order<N>(mag, rc, pivc[c][0]);
#pragma unroll
                    for (int k = 0; k < N; ++k) {
                        const V jy = Jy + rc[k] * N, jyp = Jyp + rc[k] * N;
#pragma unroll
                        for (int j = 0; j < N; ++j) {
                            const double a = jy[j];
                            Ar[k * N + j] = fma(mr, a, jyp[j]);
                            Ai[k * N + j] = mi * a;
                        }
                    }
                    zlu_factor<N, double*, double*, true>(Ar, Ai, pivc[c]);
#pragma unroll
                    for (int k = 0; k < N * N; ++k) {
                        (LUcr0 + c * N * N)[k] = Ar[k];
                        (LUci0 + c * N * N)[k] = Ai[k];
                    }
                }
            } else {
#pragma unroll U
#pragma unroll U
            for (int k = 0; k < N * N; ++k) {
                const double v = fma(cr, Jy[k], Jyp[k]);
                fin.add(v);
                LUr[k] = v;
            }
#pragma unroll
            for (int c = 0; c < NCS; ++c) {
                const double mr = h * tb.mu_re[c], mi = h * tb.mu_im[c];
#pragma unroll U
                for (int k = 0; k < N * N; ++k) {
                    const double vr = fma(mr, Jy[k], Jyp[k]), vi = mi * Jy[k];
                    (LUcr0 + c * N * N)[k] = vr;
                    (LUci0 + c * N * N)[k] = vi;
                }
            }
            // lu_factor(check_finite=True) raises.  Checked on the real matrix only: a non-finite
            // entry of Jy or Jyp shows up there (or as inf - inf = NaN), and |mu_complex| < mu_real
            if (!fin.all_finite()) { status = B200DAE_ST_NONFINITE; break; }
            lu_factor<N>(LUr, pivr);
#pragma unroll
            for (int c = 0; c < NCS; ++c) zlu_factor<N>(LUcr0 + c * N * N, LUci0 + c * N * N, pivc[c]);
            }
            nlu += 1 + NCS;
            lu_valid = true;
        }

        // ---- phase N: simplified Newton on the stage derivatives (radau.py:137-195) ----------
        bool converged = false, have_rate = false;
        double rate = 0.0;
        int n_iter = 0;
        {
#pragma unroll
            for (int i = 0; i < S; ++i)
#pragma unroll U
                for (int j = 0; j < N; ++j) {
                    double acc = 0.0;
#pragma unroll
                    for (int l = 0; l < S; ++l) acc = fma(tb.A[i * S + l], Yp[l][j], acc);
                    Y[i][j] = fma(h, acc, y[j]);
                }
            // (Tried, B200 round 1: evaluating the residuals of iteration k+1 speculatively before
            // the norm / rate / convergence chain of iteration k so that the two overlap.  Same
            // results, 15.37 ms instead of 15.42 ms per 2^20 solves: with three warps per scheduler
            // that latency is already covered, so the plain loop is kept.)
            double inv_norm_old = 0.0;         // 1 / dY_norm of the previous iteration
            bool has_old = false;
            int k = 0;
            for (; k < newton_max_iter; ++k) {
#pragma unroll U
                for (int i = 0; i < S; ++i) P::eval(fma(h, tb.C[i], t), Y[i], Yp[i], p, W[i]);
                nfev += S;
                // G = -(TI @ F); real block, then the complex blocks (radau.py:160-174)
#pragma unroll
                for (int i = 0; i < S; ++i)
#pragma unroll U
                    for (int j = 0; j < N; ++j) {
                        double acc = 0.0;
#pragma unroll
                        for (int l = 0; l < S; ++l) acc = fma(tb.TI[i * S + l], W[l][j], acc);
                        G[i][j] = -acc;
                    }
                // "if not np.all(np.isfinite(F)): break" (radau.py:157-158), tested on G[0] = -TI[0] @ F:
                // every entry of TI[0] is non-zero (checked by the launcher), so a non-finite F[l][j]
                // makes G[0][j] non-finite: n values to test instead of s * n
                FiniteAcc finG;
#pragma unroll U
                for (int j = 0; j < N; ++j) finG.add(G[0][j]);
#if !B200DAE_FIN_LATE
                if (!finG.all_finite()) break;
#endif
                lu_solve<N>(LUr, pivr, G[0]);
#pragma unroll
                for (int c = 0; c < NCS; ++c)
                    zlu_solve<N>(LUcr0 + c * N * N, LUci0 + c * N * N, pivc[c], G[2 * c + 1], G[2 * c + 2]);
#if B200DAE_FIN_LATE
                // the solves of a non-finite right-hand side are harmless (the iterate is dropped): testing
                // after them keeps residuals, transformation and solves in one basic block
                if (!finG.all_finite()) break;
#endif
                nsolve += 1 + NCS;
                // dYp = T @ dW; Yp += dYp   (radau.py:176,179)
#pragma unroll
                for (int i = 0; i < S; ++i)
#pragma unroll U
                    for (int j = 0; j < N; ++j) {
                        double acc = 0.0;
                        if (i == S - 1) {
                            // T[s-1] = (1, 0, 1, 0, ..., 1) (checked by the launcher): the same sum
                            // as the fma chain below, without the multiplications
                            acc = G[0][j];
#pragma unroll
                            for (int l = 2; l < S; l += 2) acc += G[l][j];
                        } else {
#pragma unroll
                            for (int l = 0; l < S; ++l) acc = fma(tb.T[i * S + l], G[l][j], acc);
                        }
                        W[i][j] = acc;
                        Yp[i][j] += acc;
                    }
                // dY = h * A @ dYp; Y += dY; ||dY / scale||_rms over all s*n entries (radau.py:177-182)
                double ssi[S];
#pragma unroll
                for (int i = 0; i < S; ++i) {
                    double ss = 0.0;
#pragma unroll U
                    for (int j = 0; j < N; ++j) {
                        double acc = 0.0;
#pragma unroll
                        for (int l = 0; l < S; ++l) acc = fma(tb.A[i * S + l], W[l][j], acc);
                        Y[i][j] = fma(h, acc, Y[i][j]);
                        const double q = acc * h_inv_scale[j];      // dY / scale
                        ss = fma(q, q, ss);
                    }
                    ssi[i] = ss;
                }
                double ss = 0.0;
#pragma unroll
                for (int i = 0; i < S; ++i) ss += ssi[i];
                double inv_norm;
                const double dY_norm = fast_sqrt_rsqrt(ss * (1.0 / (S * N)), inv_norm);
                if (has_old) {
                    // rate = dY_norm / dY_norm_old; the two tests of radau.py:186-191 share
                    // c = dY_norm / (1 - rate).  quirk 2: the rate test breaks AFTER the update.
                    rate = dY_norm * inv_norm_old;
                    have_rate = true;
                    const double c = dY_norm * fast_rcp_nz(1.0 - rate);   // only used when rate < 1
                    if (rate >= 1.0 || powi<kPowBits>(rate, newton_max_iter - k) * c > L.newton_tol) break;
                    if (dY_norm == 0.0 || rate * c < L.newton_tol) { converged = true; break; }
                } else if (dY_norm == 0.0) {
                    converged = true;
                    break;
                }
                // (Tried on B200, round 2: both tests as predicates with ONE exit branch, 11.16 instead of
                // 10.87 ms per 2^20 solves; the loop unrolled by two, 14.1 ms: spills.)
                inv_norm_old = inv_norm;
                has_old = true;
            }
            n_iter = (k == newton_max_iter) ? newton_max_iter : k + 1;
            niter += n_iter;
        }
        if (!converged) {
            ++nfail;
            if (!current_jac) {                    // radau.py:545-552: refresh J at the step START
                need_jac = true;
                current_jac = true;
                lu_valid = false;
                retry = true;
                continue;
            }
            hl *= 0.5;                             // radau.py:554-558
            lu_valid = false;
            continue;
        }

        // ---- phase E: embedded error estimate (radau.py:560-596) ---------------------------
        double err[N];
        {
            const int emb = kLean ? 1 : L.newton_iter_embedded;
            if (emb == 1) {
                // implicit embedded method, one Newton iteration
                double yph[N];
                const double rmu = L.inv_mu_real;
#pragma unroll U
                for (int j = 0; j < N; ++j) {
                    double acc = 0.0;
#pragma unroll
                    for (int l = 0; l < S; ++l) acc = fma(tb.v_implicit[l], Yp[l][j], acc);
                    yph[j] = (acc - tb.b_hat1_implicit * yp[j]) * rmu;
                }
                P::eval(t_new, Y[S - 1], yph, p, err);
                ++nfev;
                FiniteAcc fin;
#pragma unroll U
                for (int j = 0; j < N; ++j) fin.add(err[j]);
                if (!fin.all_finite()) { status = B200DAE_ST_NONFINITE; break; }   // lu_solve(check_finite) raises
                lu_solve<N>(LUr, pivr, err);
                ++nsolve;
                const double sc = -h * tb.mu_real;
#pragma unroll U
                for (int j = 0; j < N; ++j) err[j] *= sc;
            } else if (emb == 0) {
                // explicit embedded method
#pragma unroll U
                for (int j = 0; j < N; ++j) {
                    double acc = 0.0;
#pragma unroll
                    for (int l = 0; l < S; ++l) acc = fma(tb.v_explicit[l], Yp[l][j], acc);
                    err[j] = h * fma(tb.mu_real, yp[j], acc);
                }
            } else {
                // implicit embedded method iterated `emb` times (radau.py:578-587)
                double yh[N], yph0[N], yph[N];
                const double rhm = 1.0 / (h * tb.mu_real);
                bool fin = true;
#pragma unroll U
                for (int j = 0; j < N; ++j) {
                    double acc = 0.0;
#pragma unroll
                    for (int l = 0; l < S; ++l) acc = fma(tb.b_hat_implicit[l], Yp[l][j], acc);
                    yph0[j] = -(y[j] / h + tb.b_hat1_implicit * yp[j] + acc) / tb.mu_real;
                    yh[j] = Y[S - 1][j];
                }
                for (int it = 0; it < emb && fin; ++it) {
#pragma unroll U
                    for (int j = 0; j < N; ++j) yph[j] = fma(yh[j], rhm, yph0[j]);
                    P::eval(t_new, yh, yph, p, err);
                    ++nfev;
#pragma unroll U
                    for (int j = 0; j < N; ++j) fin = fin && finite_d(err[j]);
                    if (!fin) break;
                    lu_solve<N>(LUr, pivr, err);
                    ++nsolve;
#pragma unroll U
                    for (int j = 0; j < N; ++j) yh[j] = fma(-h * tb.mu_real, err[j], yh[j]);
                }
                if (!fin) { status = B200DAE_ST_NONFINITE; break; }
#pragma unroll U
                for (int j = 0; j < N; ++j) err[j] = yh[j] - Y[S - 1][j];
            }
        }
        double error_norm;
        {
            const double w = kLean ? 0.0 : L.cew;
            double ss = 0.0;
#pragma unroll U
            for (int j = 0; j < N; ++j) {
                double e = (1.0 - w) * fabs(err[j]);
                if (w != 0.0) {           // mix with the collocation error (radau.py:566,590-593)
                    double acc = 0.0;
#pragma unroll
                    for (int l = 0; l < S; ++l) acc = fma(tb.P2row0[l], Y[l][j], acc);
                    e = fma(w, pow(fabs(y[j] - acc), (S + 1.0) / S), e);
                }
                const double sc = fma(amax_sel(y[j], Y[S - 1][j]), rtol, L.atol[j]);
                const double q = e * fast_rcp(sc);
                ss = fma(q, q, ss);
            }
            error_norm = fast_sqrt<false>(ss * (1.0 / N));   // only compared with 1 and fed to predict_factor
        }
        const double safety = 0.9 * (2 * newton_max_iter + 1) * fast_rcp_nz((double)(2 * newton_max_iter + n_iter));

        // one code site for predict_factor: the reject path always needs it, the accept path
        // unless a factor is left over from a rejected attempt of this step (quirk 3)
        double q_err;
        const double pf = predict_factor<S>(hl, hl_old, error_norm, errl_old, havel_old, ql_old, q_err);
        if (error_norm > 1.0) {                    // reject (radau.py:598-603); quirk 8: J is kept
            factor = pf;
            have_factor = true;
            hl *= safety * factor;
            lu_valid = false;
            ++nrej;
            continue;
        }

        // ---- phase C: the step is accepted (radau.py:607-657) ---------------------------------
        const bool recompute_jac = (n_iter > 2) && have_rate && (rate > L.jac_rate);   // quirk 6
        // quirk 3: a factor left over from a rejected attempt is reused as is, without safety
        if (!have_factor) factor = safety * pf;
        if (!recompute_jac && L.db_lo <= factor && factor <= L.db_hi) factor = 1.0;   // quirk 5: LUs kept
        else lu_valid = false;
        need_jac = recompute_jac;
        current_jac = recompute_jac;
        h_abs_old = h_abs;                         // quirk 4: the PRE-step proposal
        err_old = error_norm;
        q_old = q_err;
        have_old = true;
        h_abs = hl * factor;

        // dense output ZP = (Y - y_old).T @ P (radau.py:661-666)
#pragma unroll U
        for (int j = 0; j < N; ++j) {
            double z[S];
#pragma unroll
            for (int l = 0; l < S; ++l) z[l] = Y[l][j] - y[j];
#pragma unroll
            for (int k = 0; k < S; ++k) {
                double acc = 0.0;
#pragma unroll
                for (int l = 0; l < S; ++l) acc = fma(z[l], tb.P[l * S + k], acc);
                ZP[j * S + k] = acc;
            }
        }
        t_old = t;
        h_old = h;
        have_sol = true;

        // events on this step's dense output (dae.py:466-490); a terminal one moves the step end
        double t_lim = t_new;
        bool ev_stop = false;
        double y_stop[N], yp_stop[N];
        if (!kLean && L.ev.spec.n_events > 0) {
            const double rho_e = fast_rcp_nz(h);
            auto dense_all = [&](double te, double* ya, double* ypa) {     // RadauDenseOutput, radau.py:681-699
                const double x = (te - t) * rho_e;
                double pw[S], dpw[S];
                double xk = 1.0;
#pragma unroll
                for (int k = 0; k < S; ++k) { dpw[k] = (k + 1) * rho_e * xk; xk *= x; pw[k] = xk; }
#pragma unroll U
                for (int j = 0; j < N; ++j) {
                    double a = 0.0, b = 0.0;
#pragma unroll
                    for (int k = 0; k < S; ++k) { a = fma(ZP[j * S + k], pw[k], a); b = fma(ZP[j * S + k], dpw[k], b); }
                    ya[j] = a + y[j];
                    ypa[j] = b;
                }
            };
            auto dense = [&](double te, int comp, int kind) -> double {
                double ya[N], ypa[N];
                dense_all(te, ya, ypa);
                if constexpr (EventOf<P>::kNum > 0) {
                    if (kind == 2) return EventOf<P>::eval(comp, te, ya, ypa, p);
                }
                return kind == 1 ? pick<N>(ypa, comp) : pick<N>(ya, comp);
            };
            auto record = [&](int k, int idx, double tr) {
                double ya[N], ypa[N];
                dense_all(tr, ya, ypa);
                const int64_t row = (int64_t)k * L.ev.cap + idx;
                L.ev.t_events[row * ld + inst] = tr;
#pragma unroll U
                for (int j = 0; j < N; ++j) {
                    if (L.ev.y_events) L.ev.y_events[(row * N + j) * ld + inst] = ya[j];
                    if (L.ev.yp_events) L.ev.yp_events[(row * N + j) * ld + inst] = ypa[j];
                }
            };
            double gnew[B200DAE_MAX_EVENTS];
#pragma unroll
            for (int k = 0; k < B200DAE_MAX_EVENTS; ++k)
                gnew[k] = event_quantity<P, N>(L.ev.spec, k, t_new, Y[S - 1], Yp[S - 1], p) - L.ev.spec.value[k];
            double ts = t_new;
            if (handle_events(L.ev, evs, t, t_new, dir, gnew, dense, record, ts)) {
                ev_stop = true;
                t_lim = ts;
                dense_all(ts, y_stop, yp_stop);
            }
        }
        // t_eval points passed by this step (dae.py:496-515; RadauDenseOutput radau.py:681-699).
        // The lanes of a warp hold neighbouring instances, so row `r` of y_eval is 32 consecutive doubles
        // for them -- if they store it in the SAME instruction.  Their next-row indices `ie` differ by a
        // row or two (their step ends differ), so a lane-local loop pairs row ie_a of lane a with row
        // ie_b of lane b and every store instruction scatters over several rows.  The loop therefore runs
        // over the warp's common row range and a lane takes part in the rows of its own range
        // [ie, ie_end): lanes that emit row r in this step emit it together, as one 256-byte segment.
        if (kEval && L.out.t_eval != nullptr) {
            const double rho = fast_rcp_nz(h);
            int ie_end = (int)ie;
            while (ie_end < (int)L.out.m) {
                const double te = __ldg(L.out.t_eval + ie_end);
                if (!(dir > 0.0 ? te <= t_lim : te >= t_lim)) break;
                ++ie_end;
            }
            // (instance-major layouts, eval_stride_inst != 1: every lane streams its own rows, nothing to align)
            const int64_t e_sr = L.out.eval_stride_row, e_sc = L.out.eval_stride_comp, e_si = L.out.eval_stride_inst;
            int r_lo = (int)ie, r_hi = ie_end;
#if B200DAE_EVAL_ALIGNED
            if (e_si == 1) {
                const unsigned amask = __activemask();
                r_lo = __reduce_min_sync(amask, (int)ie);
                r_hi = __reduce_max_sync(amask, ie_end);
            }
#endif
            for (int r = r_lo; r < r_hi; ++r) {
                if (r < (int)ie || r >= ie_end) continue;
                const double te = __ldg(L.out.t_eval + r);
                const double x = (te - t) * rho;
                double pw[S], dpw[S];
                double xk = 1.0;
#pragma unroll
                for (int k = 0; k < S; ++k) { dpw[k] = (k + 1) * rho * xk; xk *= x; pw[k] = xk; }
                double* ye = L.out.y_eval + (int64_t)r * e_sr + inst * e_si;
                double* ype = L.out.yp_eval ? L.out.yp_eval + (int64_t)r * e_sr + inst * e_si : nullptr;
#pragma unroll U
                for (int j = 0; j < N; ++j) {
                    double a = 0.0, b = 0.0;
#pragma unroll
                    for (int k = 0; k < S; ++k) { a = fma(ZP[j * S + k], pw[k], a); b = fma(ZP[j * S + k], dpw[k], b); }
                    ye[j * e_sc] = a + y[j];
                    if (ype) ype[j * e_sc] = b;
                }
            }
            ie = ie_end;
        }
        // commit
        t = t_lim;
#pragma unroll U
        for (int j = 0; j < N; ++j) {
            y[j] = ev_stop ? y_stop[j] : Y[S - 1][j];
            yp[j] = ev_stop ? yp_stop[j] : Yp[S - 1][j];
        }
        ++nacc;
        if (!kLean && L.out.t_steps != nullptr && nst < L.out.steps_cap) {
            L.out.t_steps[(int64_t)nst * ld + inst] = t;
#pragma unroll U
            for (int j = 0; j < N; ++j) {
                if (L.out.y_steps) L.out.y_steps[((int64_t)nst * N + j) * ld + inst] = y[j];
                if (L.out.yp_steps) L.out.yp_steps[((int64_t)nst * N + j) * ld + inst] = yp[j];
            }
            if (L.out.dense_coef != nullptr) {     // this step's RadauDenseOutput (radau.py:661-679)
                L.out.h_steps[((int64_t)nst * 2) * ld + inst] = h;
                L.out.h_steps[((int64_t)nst * 2 + 1) * ld + inst] = t_new;
                if (L.out.order_steps) L.out.order_steps[(int64_t)nst * ld + inst] = S;
#pragma unroll U
                for (int k = 0; k < N * S; ++k)
                    L.out.dense_coef[((int64_t)nst * (N * S) + k) * ld + inst] = ZP[k];
            }
            ++nst;
        }
        step_begin = true;
        if (ev_stop) status = B200DAE_ST_EVENT;                                    // dae.py:484-487
        else if (dir * (t - L.t_bound) >= 0.0) status = B200DAE_ST_FINISHED;       // base.py:389-390
        else if (L.max_steps > 0 && nacc >= L.max_steps) status = B200DAE_ST_MAX_STEPS;
        // the reference still evaluates the Jacobian after its last step (radau.py:628-630)
        if (status != B200DAE_ST_RUNNING && need_jac) ++njev;
    }

    // ---- results ----------------------------------------------------------------------------
    if (staged) {
        __syncwarp();                                      // every lane of the warp is done with its cold state
        constexpr int oY = 0, oYp = 32 * N, oC = 64 * N;   // counters: 8 ints = 4 doubles per lane
        if (L.out.bad_input != nullptr && status == B200DAE_ST_NONFINITE && nacc == 0) *L.out.bad_input = 1;
#pragma unroll U
        for (int j = 0; j < N; ++j) { stage(oY + wl * N + j) = y[j]; stage(oYp + wl * N + j) = yp[j]; }
        stage(oC + wl * 4 + 0) = __hiloint2double(njev, nfev);
        stage(oC + wl * 4 + 1) = __hiloint2double(nsolve, nlu);
        stage(oC + wl * 4 + 2) = __hiloint2double(nrej, nacc);
        stage(oC + wl * 4 + 3) = __hiloint2double(niter, nfail);
        __syncwarp();
        for (int e = wl; e < wn * N; e += 32) {
            L.out.y_final[winst0 * N + e] = stage(oY + e);
            if (L.out.yp_final) L.out.yp_final[winst0 * N + e] = stage(oYp + e);
        }
        if (L.out.counters) {
            double* c2 = reinterpret_cast<double*>(L.out.counters + winst0 * 8);      // [instance][8] ints, 32-byte rows
            for (int e = wl; e < wn * 4; e += 32) c2[e] = stage(oC + e);
        }
        if (valid) {
            L.out.t_final[inst] = t;
            L.out.status[inst] = status;
            if (L.out.n_eval_valid) L.out.n_eval_valid[inst] = (int32_t)ie;
            if (L.out.n_steps) L.out.n_steps[inst] = nst;
        }
        return;
    }
    if (L.out.bad_input != nullptr && status == B200DAE_ST_NONFINITE && nacc == 0) *L.out.bad_input = 1;
#pragma unroll U
    for (int j = 0; j < N; ++j) {
        L.out.y_final[io_index(L.io_aos, ld, N, j, inst)] = y[j];
        if (L.out.yp_final) L.out.yp_final[io_index(L.io_aos, ld, N, j, inst)] = yp[j];
    }
    L.out.t_final[inst] = t;
    L.out.status[inst] = status;
    if (L.out.n_eval_valid) L.out.n_eval_valid[inst] = (int32_t)ie;
    if (L.out.n_steps) L.out.n_steps[inst] = nst;
    if (!kLean) {
        for (int k = 0; k < L.ev.spec.n_events; ++k) L.ev.n_found[(int64_t)k * ld + inst] = evs.found[k];
    }
    if (L.out.counters) {
        int32_t* c = L.out.counters + (L.io_aos ? inst * 8 : inst);
        const int64_t cs = L.io_aos ? 1 : ld;
        c[B200DAE_CNT_NFEV * cs] = nfev;
        c[B200DAE_CNT_NJEV * cs] = njev;
        c[B200DAE_CNT_NLU * cs] = nlu;
        c[B200DAE_CNT_NSOLVE * cs] = nsolve;
        c[B200DAE_CNT_ACCEPTED * cs] = nacc;
        c[B200DAE_CNT_ERR_REJECTED * cs] = nrej;
        c[B200DAE_CNT_NEWTON_FAIL * cs] = nfail;
        c[B200DAE_CNT_NEWTON_ITER * cs] = niter;
    }
}

}  // namespace b200dae
