// radau_group.cuh — adaptive Radau IIA integration with G cooperating lanes per system.
//
// Same replacement scope and the same flattened control flow as radau.cuh (one loop trip = one
// simplified-Newton solve; reference: RadauDAE._step_impl radau.py:465-659,
// solve_collocation_system radau.py:137-195, predict_factor radau.py:198-240, dense output
// radau.py:661-699, solve_dae's loop and t_eval slicing dae.py:446-518), for systems too large
// for one thread's registers (n > 4: Akzo Nobel n = 6, transistor amplifier n = 8, Andrews'
// squeezer n = 26).
//
// Work split inside a group (lane j = component j = matrix row j):
//   * stage vectors Y, Yp, the T / TI / A stage transformations, error scales, dense output:
//     lane-local (a lane holds its component of every stage in registers);
//   * norms: butterfly all-reduce over the group (bit-identical in every lane, so every lane
//     takes the same accept / reject / convergence decisions);
//   * LU factorisations and triangular solves: group.cuh (rows in registers / shared memory);
//   * residual evaluations: the S stage residuals of a Newton iteration run on S lanes at once,
//     the 2n columns of the finite-difference Jacobian on n lanes at once (two rounds); the
//     vectors they read are staged in shared memory.
// Shared memory per system: Jy, Jyp (row-major, odd leading dimension), the complex LU factors
// (aliased with the FD scratch: the Jacobian is only ever refreshed when the factors are stale),
// S staged stage vectors, the pivot-row buffer.
#pragma once
#include "rt_compat.cuh"
#include "../../include/b200dae.h"
#include "group.cuh"
#include "radau.cuh"

namespace b200dae {

// kGJ: the Jacobians Jy, Jyp live in GLOBAL memory (a per-instance scratch block that stays in L2:
// written once per Jacobian refresh, read once per factorisation) instead of shared memory.  For
// Andrews' squeezer (n = 26, s = 5) that and two trims (the pivot index buffer aliases the staged
// residuals, which are dead during a factorisation; the real factors need no read-ahead pad because
// the work area follows them) take a system from 43.7 KB to 32.2 KB: 7 resident systems per SM
// instead of 5 (the kernel runs ONE warp per system and is latency-bound at ~1.1 warps per scheduler).
template <int N, int S, int G, bool kGJ = false>
struct RadauGroupLayout {   // offsets in doubles inside one group's shared-memory block
    static constexpr int NCS = S / 2;
    static constexpr int LD = N | 1;                       // odd leading dimension: conflict-free rows
    static constexpr int LDC = N | 1;                      // in double2 units
    static constexpr int kJ = N * LD;                      // doubles per Jacobian
    static constexpr int kJs = kGJ ? 0 : kJ;               // ... of which in shared memory
    static constexpr int kLUr0 = N * LD + (kGJ ? 0 : N);   // real factors (+ N read-ahead pad, see above)
    static constexpr int kLUr = kLUr0 + (kLUr0 & 1);
    static constexpr int kLUc = (N * LDC + N) * 2;         // one complex factor matrix, in doubles
    static constexpr int kCplx = NCS * kLUc;
    static constexpr int kFd = 3 * G * LD;                 // FD scratch: x copy, f(attempt 0), f(attempt 1) per lane
    static constexpr int kWork = (kCplx > kFd ? kCplx : kFd) + ((kCplx > kFd ? kCplx : kFd) & 1);
    static constexpr int kVec = S * LD;                    // one staged S x n block
    static constexpr int oJy = 0, oJyp = oJy + kJs, oLUr = oJyp + kJs + ((2 * kJs) & 1);
    static constexpr int oWork = oLUr + kLUr;
    static constexpr int oY = oWork + kWork, oYp = oY + kVec, oF = oYp + kVec;
    static constexpr int kIdx = (G + 1) / 2;               // G ints
    static_assert(!kGJ || kVec >= kIdx, "pivot index buffer aliases the staged residuals");
    static constexpr int oIdx = kGJ ? oF : oF + kVec + ((3 * kVec) & 1);
    static constexpr int oPar = kGJ ? oF + kVec + ((3 * kVec) & 1) : oIdx + kIdx;   // NP parameters (appended by the kernel)
    static constexpr int kScratch = kGJ ? 2 * kJ : 0;      // doubles of global scratch per instance
    __host__ __device__ static constexpr int doubles(int np) { return oPar + np + ((oPar + np) & 1); }   // keep groups 16 B aligned
};

// Residual call of the cooperative kernels (all pointers point into shared memory).  Every call
// is made by ALL lanes of a group (lanes without a stage / column of their own repeat another
// lane's evaluation and write identical values): free in SIMT, and it removes the divergent
// regions around the calls (+5..9 % on the amplifier).  Small functors are inlined at the four
// call sites; a large one (Andrews) declares itself __noinline__ and then exists once per kernel.
// A __noinline__ wrapper for every functor (-DB200DAE_NOINLINE_EVAL, 10 KB less code) is NOT
// usable: with it the amplifier ensembles fail deterministically after t = 7e-4 on sm_100a
// (nvcc 12.9; same step counts with divergent and with convergent call sites, so not a race),
// while the inlined build of the same source matches the reference; Akzo Nobel is correct
// either way and not faster.
#ifdef B200DAE_NOINLINE_EVAL
#define B200DAE_EVAL_ATTR __noinline__
#else
#define B200DAE_EVAL_ATTR __forceinline__
#endif
template <class P>
__device__ B200DAE_EVAL_ATTR void eval_shared(double t, const double* y, const double* yp, const double* p,
                                         double* out) {
    P::eval(t, y, yp, p, out);
}

// Forward-difference Jacobians, columns in parallel (lane c = column c = perturbed entry c).
// Per column exactly the arithmetic of fd_jacobian.cuh / scipy num_jac.
template <class P, int G, int LD, class Grp>
__device__ __forceinline__ void fd_jacobian_group(const Grp& g, double t, double yj, double ypj,
                                                  const double* __restrict__ p, double thr_j,
                                                  double& fac_y, double& fac_yp,
                                                  double* __restrict__ sJy, double* __restrict__ sJyp,
                                                  double* __restrict__ sy, double* __restrict__ syp,
                                                  double* __restrict__ sf, double* __restrict__ scratch) {
    constexpr int N = P::N;
    const int c = g.lane;
    const bool act = c < N;
    if (act) { sy[c] = yj; syp[c] = ypj; }
    g.sync();
    // Every residual call below is made by ALL lanes of the group (idle lanes repeat lane work
    // and discard it): in SIMT that costs nothing extra, and the call sites stay convergent.
    eval_shared<P>(t, sy, syp, p, sf);                       // identical writes from every lane
    g.sync();
    const int cc = act ? c : 0;                              // idle lanes shadow column 0
    double* xc = scratch + c * (3 * LD);
    double* fbuf0 = xc + LD;
    double* fbuf1 = fbuf0 + LD;
#pragma unroll 1
    for (int which = 0; which < 2; ++which) {
        const bool wrt_yp = (which == 1);
        const double* xs = wrt_yp ? syp : sy;
        const double xi = act ? (wrt_yp ? ypj : yj) : xs[0];
        const double f_sign = (sf[cc] >= 0.0) ? 1.0 : -1.0;
        const double y_scale = f_sign * amax_sel(thr_j, xi);
        double fac = wrt_yp ? fac_yp : fac_y;
        double xp = fma(fac, y_scale, xi);
        double h = xp - xi;
        for (int trip = 0; h == 0.0 && trip < 400; ++trip) {   // bounded, see fd_jacobian.cuh: fd_batch
            fac *= 10.0;
            xp = fma(fac, y_scale, xi);
            h = xp - xi;
        }
        for (int r = 0; r < N; ++r) xc[r] = xs[r];
        double max_diff = 0.0, scale = 0.0;
        double fac_try = fac, xp_try = xp, h_try = h;
        const double* fsel = fbuf0;
        bool want = true;                                    // this lane takes part in the attempt
#pragma unroll 1
        for (int attempt = 0; attempt < 2; ++attempt) {
            double* fnew = attempt ? fbuf1 : fbuf0;
            xc[cc] = xp_try;
            eval_shared<P>(t, wrt_yp ? sy : xc, wrt_yp ? xc : syp, p, fnew);
            if (want) {
                // running argmax on the raw differences; a NaN difference (np.argmax: the first NaN
                // wins, max_diff = NaN) is restored through the sum, see fd_probe in fd_jacobian.cuh
                double f_at = sf[0], fn_at = fnew[0], bd = fn_at - f_at, sum = 0.0;
                for (int r = 1; r < N; ++r) {
                    const double fr = sf[r], fn = fnew[r];
                    const double dr = fn - fr;
                    sum += dr;
                    if (fabs(dr) > fabs(bd)) { bd = dr; f_at = fr; fn_at = fn; }
                }
                const double md2 = fma(0.0, sum, fabs(bd));
                const double sc2 = amax_sel(f_at, fn_at);
                if (attempt == 0 || max_diff * sc2 < md2 * scale) {
                    fac = fac_try;
                    h = h_try;
                    fsel = fnew;
                    scale = sc2;
                    max_diff = md2;
                }
            }
            // second attempt only for the lanes whose difference drowned in round-off
            // (SP/common.py:350-371); the others re-evaluate their first point and ignore it
            want = want && attempt == 0 && (max_diff < kNjDiffReject * scale);
            if (attempt == 1 || !g.any(act && want)) break;
            if (want) {
                fac_try = 10.0 * fac;
                xp_try = fma(fac_try, y_scale, xi);
                h_try = xp_try - xi;
            }
        }
        if (act) {
            const double rh = fast_rcp(h);
            double* J = wrt_yp ? sJyp : sJy;
            for (int r = 0; r < N; ++r) J[r * LD + c] = (fsel[r] - sf[r]) * rh;
            if (max_diff < kNjDiffSmall * scale) fac *= 10.0;
            if (max_diff > kNjDiffBig * scale) fac *= 0.1;
            fac = max_sel(kNjMinFactor, fac);
            if (wrt_yp) fac_yp = fac; else fac_y = fac;
        }
    }
    g.sync();
}

// Analytic Jacobians (jac=callable, base.py:299-331): the state is staged in shared memory, the
// lanes zero the matrices together, lane 0 lets the functor fill the structural non-zeros.
template <class P, int LD, class Grp>
__device__ __forceinline__ void analytic_jacobian_group(const Grp& g, double t, double yj, double ypj,
                                                        const double* __restrict__ p,
                                                        double* __restrict__ sJy, double* __restrict__ sJyp,
                                                        double* __restrict__ sy, double* __restrict__ syp) {
    constexpr int N = P::N;
    const int c = g.lane;
    if (c < N) {
        sy[c] = yj;
        syp[c] = ypj;
        for (int r = 0; r < N; ++r) { sJy[r * LD + c] = 0.0; sJyp[r * LD + c] = 0.0; }
    }
    g.sync();
    if (c == 0) P::jac(t, sy, syp, p, MatRef<double*, LD>{sJy}, MatRef<double*, LD>{sJyp});
    g.sync();
}

// kLean: default error estimator, no t_eval / step trace / events (compiled out: these kernels
// are bound by instruction fetch, every KB of cold code costs throughput).
template <class P, int S, int G, int kThreads, int kMinBlocks, bool kJac = false, bool kLean = false, bool kGJ = false>
__global__ void __launch_bounds__(kThreads, kMinBlocks)
radau_group_kernel(const __grid_constant__ RadauLaunch L) {
    constexpr int N = P::N, NP = P::NP, NCS = S / 2;
    using Lay = RadauGroupLayout<N, S, G, kGJ>;
    constexpr int LD = Lay::LD, LDC = Lay::LDC;
    static_assert(G >= N && G >= S, "group must cover the components and the stages");
    extern __shared__ __align__(16) double smem[];

    const Group<G> g;
    const int lane = g.lane;
    const bool act = lane < N;
    const int64_t inst = ((int64_t)blockIdx.x * kThreads + threadIdx.x) / G;
    if (inst >= L.N) return;                              // whole groups leave together
    const int64_t ld = L.ld;
    const b200dae_radau_tables& tb = L.tb;

    double* sm = smem + (size_t)(threadIdx.x / G) * Lay::doubles(NP);
    double* sJy = kGJ ? L.jscratch + inst * Lay::kScratch : sm + Lay::oJy;
    double* sJyp = kGJ ? sJy + Lay::kJ : sm + Lay::oJyp;
    double* sWork = sm + Lay::oWork;
    double* sY = sm + Lay::oY;
    double* sYp = sm + Lay::oYp;
    double* sF = sm + Lay::oF;
    double* sLUr = sm + Lay::oLUr;
    int* sIdx = reinterpret_cast<int*>(sm + Lay::oIdx);
    double2* sCplx = reinterpret_cast<double2*>(sWork);

    // ---- per-instance state: scalars replicated in every lane, vectors one component per lane --
    double* p = sm + Lay::oPar;                            // per-instance parameters, shared by the group
    for (int j = lane; j < NP; j += G) p[j] = L.params[io_index(L.io_aos, ld, NP, j, inst)];
    g.sync();
    const int jc = act ? lane : 0;                        // clamped component index for loads
    double y = act ? L.y0[io_index(L.io_aos, ld, N, jc, inst)] : 0.0;
    double yp = act ? L.yp0[io_index(L.io_aos, ld, N, jc, inst)] : 0.0;
    const double atol_j = L.atol[jc], thr_j = L.thr[jc];
    double fac_y = kNjInitFactor, fac_yp = kNjInitFactor;

    int perm = lane;
    int cperm[NCS > 0 ? NCS : 1];
    const double2* crow[NCS > 0 ? NCS : 1];
#pragma unroll
    for (int c = 0; c < (NCS > 0 ? NCS : 1); ++c) {
        cperm[c] = lane;
        crow[c] = sCplx + c * (Lay::kLUc / 2) + jc * LDC;
    }
    const double* rrow = sLUr + jc * LD;
    double Y[S], Yp[S], W[S], Gv[S], ZP[S];
#pragma unroll
    for (int i = 0; i < S; ++i) Y[i] = Yp[i] = W[i] = Gv[i] = ZP[i] = 0.0;
    double h_inv_scale = 0.0;

    const double dir = L.dir, rtol = L.rtol;
    double t = L.t0, t_old = L.t0, h_old = 1.0;
    double h_abs;
    if (L.first_step > 0.0) {
        h_abs = L.first_step;
    } else {                                              // select_initial_step, common.py:129-183
        const double v = act ? fabs(yp * fast_rcp(fmax(fabs(y), thr_j))) : 0.0;
        double e = g.max(v);
        if (g.any(v != v)) e = nan("");
        const double rh = e * fast_rcp(fast_sqrt(rtol) * 0.8);
        h_abs = fmin(L.max_step, fabs(L.t_bound - L.t0));
        if (h_abs * rh > 1.0) h_abs = fast_rcp(rh);
        h_abs = fmax(h_abs, 16.0 * kEps * fabs(L.t0));
    }
    double h_abs_old = 0.0, err_old = 0.0;
    bool have_old = false, have_sol = false;
    bool need_jac = true, current_jac = true;
    bool lu_valid = false, step_begin = true, retry = false;
    int32_t nfev = 0, njev = 0, nlu = 0, nsolve = 0, nacc = 0, nrej = 0, nfail = 0, niter = 0;
    int32_t status = B200DAE_ST_RUNNING;
    int64_t ie = 0;
    int32_t nst = 0;
    const int newton_max_iter = L.newton_max_iter;

    if (!kLean && L.out.t_steps != nullptr && L.out.steps_cap > 0) {
        if (lane == 0) L.out.t_steps[inst] = t;
        if (act) {
            if (L.out.y_steps) L.out.y_steps[lane * ld + inst] = y;
            if (L.out.yp_steps) L.out.yp_steps[lane * ld + inst] = yp;
        }
        nst = 1;
    }
    if (L.t0 == L.t_bound) { status = B200DAE_ST_FINISHED; need_jac = false; njev = 1; }
    EventState evs;                  // g = event(t0, y0, yp0), dae.py:436 (identical in every lane)
    if (!kLean) for (int k = 0; k < B200DAE_MAX_EVENTS; ++k) {
        evs.count[k] = evs.found[k] = 0;
        evs.g[k] = (k < L.ev.spec.n_events)
            ? event_quantity_group<P, N>(g, L.ev.spec, k, t, y, yp, p) - L.ev.spec.value[k] : 0.0;
    }

    double hl = 0.0, hl_old = 0.0, errl_old = 0.0, min_step = 0.0, factor = 0.0;
    double h = 0.0, t_new = 0.0;
    bool havel_old = false, have_factor = false;

    while (status == B200DAE_ST_RUNNING) {
        // ---- phase J ----------------------------------------------------------------------
        if (need_jac) {
            if constexpr (kJac) analytic_jacobian_group<P, LD>(g, t, y, yp, p, sJy, sJyp, sY, sYp);
            else fd_jacobian_group<P, G, LD>(g, t, y, yp, p, thr_j, fac_y, fac_yp, sJy, sJyp, sY, sYp, sF, sWork);
            ++njev;
            need_jac = false;
        }
        // ---- phase B (radau.py:489-512) -----------------------------------------------------
        if (step_begin) {
            min_step = 10.0 * fabs(next_toward(t, dir) - t);
            if (h_abs > L.max_step) { hl = L.max_step; havel_old = false; }
            else if (h_abs < min_step) { hl = min_step; havel_old = false; }
            else { hl = h_abs; havel_old = have_old; }
            hl_old = h_abs_old; errl_old = err_old;
            have_factor = false;
            step_begin = false;
        }
        // ---- phase A (radau.py:514-532) -----------------------------------------------------
        if (!retry) {
            if (hl < min_step) { status = B200DAE_ST_TOO_SMALL_STEP; break; }
            h = hl * dir;
            t_new = t + h;
            if (dir * (t_new - L.t_bound) > 0.0) t_new = L.t_bound;
            h = t_new - t;
            hl = fabs(h);
            if (!have_sol) {
#pragma unroll
                for (int i = 0; i < S; ++i) Yp[i] = yp;
            } else {
                const double rho = fast_rcp(h_old);
#pragma unroll
                for (int i = 0; i < S; ++i) {
                    const double x = (fma(h, tb.C[i], t) - t_old) * rho;
                    double xk = rho, acc = 0.0;
#pragma unroll
                    for (int k = 0; k < S; ++k) { acc = fma(ZP[k], (k + 1) * xk, acc); xk *= x; }
                    Yp[i] = acc;
                }
            }
            h_inv_scale = h * fast_rcp(fma(fabs(y), rtol, atol_j));
        }
        retry = false;

        // ---- phase F (radau.py:536-538) -----------------------------------------------------
        if (!lu_valid) {
            bool fin = true;
#if B200DAE_LU_INPLACE
            // rows assembled straight into the factor storage (group.cuh: in-place variants)
#pragma unroll 1
            for (int c = 0; c < NCS; ++c) {
                const double mr = h * tb.mu_re[c], mi = h * tb.mu_im[c];
                double2* cm = sCplx + c * (Lay::kLUc / 2);
                if (act) {
#pragma unroll 2
                    for (int k = 0; k < N; ++k) {
                        const double jy = sJy[jc * LD + k], jyp = sJyp[jc * LD + k];
                        const double re = fma(mr, jy, jyp), im = mi * jy;
                        fin = fin && finite_d(re) && finite_d(im);
                        cm[jc * LDC + k] = make_double2(re, im);
                    }
                }
                g.sync();
                int pm;
                group_zlu_factor_inplace<N, LDC>(g, pm, cm);
#pragma unroll
                for (int cc = 0; cc < NCS; ++cc)
                    if (cc == c) cperm[cc] = pm;
            }
            {
                const double cr = h * tb.mu_real;
                if (act) {
#pragma unroll 2
                    for (int k = 0; k < N; ++k) {
                        const double v = fma(cr, sJy[jc * LD + k], sJyp[jc * LD + k]);
                        fin = fin && finite_d(v);
                        sLUr[jc * LD + k] = v;
                    }
                }
                g.sync();
                group_lu_factor_inplace<N, LD>(g, perm, sLUr);
            }
#else
#pragma unroll 1
            for (int c = 0; c < NCS; ++c) {
                const double mr = h * tb.mu_re[c], mi = h * tb.mu_im[c];
                double ar[N], ai[N];
#pragma unroll
                for (int k = 0; k < N; ++k) {
                    const double jy = act ? sJy[jc * LD + k] : 0.0, jyp = act ? sJyp[jc * LD + k] : 0.0;
                    ar[k] = fma(mr, jy, jyp);
                    ai[k] = mi * jy;
                    fin = fin && finite_d(ar[k]) && finite_d(ai[k]);
                }
                int pm;
                group_zlu_factor<N, LDC>(g, ar, ai, pm, sCplx + c * (Lay::kLUc / 2), sIdx);
#pragma unroll
                for (int cc = 0; cc < NCS; ++cc)
                    if (cc == c) cperm[cc] = pm;
            }
            {
                const double cr = h * tb.mu_real;
                double a[N];
#pragma unroll
                for (int k = 0; k < N; ++k) {
                    const double jy = act ? sJy[jc * LD + k] : 0.0, jyp = act ? sJyp[jc * LD + k] : 0.0;
                    a[k] = fma(cr, jy, jyp);
                    fin = fin && finite_d(a[k]);
                }
                group_lu_factor<N, LD>(g, a, perm, sLUr, sIdx);
            }
#endif
            // lu_factor(check_finite=True) raises before factoring; the factors of a non-finite
            // matrix are never used because the instance stops here
            if (!g.all(fin)) { status = B200DAE_ST_NONFINITE; break; }
            nlu += 1 + NCS;
            lu_valid = true;
        }

        // ---- phase N (radau.py:137-195) -------------------------------------------------------
        bool converged = false, have_rate = false;
        double rate = 0.0;
        int n_iter = 0;
        {
#pragma unroll
            for (int i = 0; i < S; ++i) {
                double acc = 0.0;
#pragma unroll
                for (int l = 0; l < S; ++l) acc = fma(tb.A[i * S + l], Yp[l], acc);
                Y[i] = fma(h, acc, y);
            }
            double inv_norm_old = 0.0;
            bool has_old = false;
            int k = 0;
            for (; k < newton_max_iter; ++k) {
                if (act) {
#pragma unroll
                    for (int i = 0; i < S; ++i) { sY[i * LD + lane] = Y[i]; sYp[i * LD + lane] = Yp[i]; }
                }
                g.sync();
                {
                    const int st = lane % S;             // lanes >= S repeat a stage (identical writes)
                    eval_shared<P>(fma(h, tb.C[st], t), sY + st * LD, sYp + st * LD, p, sF + st * LD);
                }
                g.sync();
                bool fin = true;
#pragma unroll
                for (int i = 0; i < S; ++i) {
                    W[i] = act ? sF[i * LD + lane] : 0.0;
                    fin = fin && finite_d(W[i]);
                }
                nfev += S;
                if (!g.all(fin)) break;                           // radau.py:157-158
#pragma unroll
                for (int i = 0; i < S; ++i) {
                    double acc = 0.0;
#pragma unroll
                    for (int l = 0; l < S; ++l) acc = fma(tb.TI[i * S + l], W[l], acc);
                    Gv[i] = -acc;
                }
                {
                    double br[NCS > 0 ? NCS : 1], bi[NCS > 0 ? NCS : 1];
#pragma unroll
                    for (int c = 0; c < NCS; ++c) { br[c] = Gv[2 * c + 1]; bi[c] = Gv[2 * c + 2]; }
                    group_lu_solve<N, NCS, true>(g, rrow, perm, Gv[0], crow, cperm, br, bi);
#pragma unroll
                    for (int c = 0; c < NCS; ++c) { Gv[2 * c + 1] = br[c]; Gv[2 * c + 2] = bi[c]; }
                }
                nsolve += 1 + NCS;
#pragma unroll
                for (int i = 0; i < S; ++i) {
                    double acc = 0.0;
#pragma unroll
                    for (int l = 0; l < S; ++l) acc = fma(tb.T[i * S + l], Gv[l], acc);
                    W[i] = acc;
                    Yp[i] += acc;
                }
                double ss = 0.0;
#pragma unroll
                for (int i = 0; i < S; ++i) {
                    double acc = 0.0;
#pragma unroll
                    for (int l = 0; l < S; ++l) acc = fma(tb.A[i * S + l], W[l], acc);
                    Y[i] = fma(h, acc, Y[i]);
                    const double q = acc * h_inv_scale;
                    ss = fma(q, q, ss);
                }
                ss = g.sum(act ? ss : 0.0);
                double inv_norm;
                const double dY_norm = fast_sqrt_rsqrt(ss * (1.0 / (S * N)), inv_norm);
                if (has_old) {
                    rate = dY_norm * inv_norm_old;
                    have_rate = true;
                    const double c = dY_norm * fast_rcp(1.0 - rate);
                    if (rate >= 1.0 || powi(rate, newton_max_iter - k) * c > L.newton_tol) break;
                    if (dY_norm == 0.0 || rate * c < L.newton_tol) { converged = true; break; }
                } else if (dY_norm == 0.0) {
                    converged = true;
                    break;
                }
                inv_norm_old = inv_norm;
                has_old = true;
            }
            n_iter = (k == newton_max_iter) ? newton_max_iter : k + 1;
            niter += n_iter;
        }
        if (!converged) {
            ++nfail;
            if (!current_jac) {
                need_jac = true;
                current_jac = true;
                lu_valid = false;
                retry = true;
                continue;
            }
            hl *= 0.5;
            lu_valid = false;
            continue;
        }

        // ---- phase E (radau.py:560-596) -------------------------------------------------------
        double err;
        {
            const int emb = kLean ? 1 : L.newton_iter_embedded;
            if (emb == 0) {
                double acc = 0.0;
#pragma unroll
                for (int l = 0; l < S; ++l) acc = fma(tb.v_explicit[l], Yp[l], acc);
                err = h * fma(tb.mu_real, yp, acc);
            } else {
                // implicit embedded method, `emb` Newton iterations (radau.py:568-587); the
                // first iterate of the general loop is the single-iteration formula
                double yph0, yh = Y[S - 1];
                bool fin_all = true;
                if (emb == 1) {
                    double acc = 0.0;
#pragma unroll
                    for (int l = 0; l < S; ++l) acc = fma(tb.v_implicit[l], Yp[l], acc);
                    yph0 = (acc - tb.b_hat1_implicit * yp) * L.inv_mu_real;
                } else {
                    double acc = 0.0;
#pragma unroll
                    for (int l = 0; l < S; ++l) acc = fma(tb.b_hat_implicit[l], Yp[l], acc);
                    yph0 = -(y / h + tb.b_hat1_implicit * yp + acc) / tb.mu_real;
                }
                const double rhm = 1.0 / (h * tb.mu_real);
                double e = 0.0;
                for (int it = 0; it < emb; ++it) {
                    const double yph = (emb == 1) ? yph0 : fma(yh, rhm, yph0);
                    if (act) { sY[lane] = yh; sYp[lane] = yph; }
                    g.sync();
                    eval_shared<P>(t_new, sY, sYp, p, sF);       // every lane, identical writes
                    g.sync();
                    e = act ? sF[lane] : 0.0;
                    ++nfev;
                    if (!g.all(finite_d(e))) { fin_all = false; break; }
                    group_lu_solve_real<N>(g, rrow, perm, e);
                    ++nsolve;
                    yh = fma(-h * tb.mu_real, e, yh);
                }
                if (!fin_all) { status = B200DAE_ST_NONFINITE; break; }
                err = (emb == 1) ? e * (-h * tb.mu_real) : yh - Y[S - 1];
            }
        }
        double error_norm;
        {
            const double w = kLean ? 0.0 : L.cew;
            double e = (1.0 - w) * fabs(err);
            if (w != 0.0) {
                double acc = 0.0;
#pragma unroll
                for (int l = 0; l < S; ++l) acc = fma(tb.P2row0[l], Y[l], acc);
                e = fma(w, pow(fabs(y - acc), (S + 1.0) / S), e);
            }
            const double sc = fma(amax_sel(y, Y[S - 1]), rtol, atol_j);
            const double q = e * fast_rcp(sc);
            const double ss = g.sum(act ? q * q : 0.0);
            error_norm = fast_sqrt(ss * (1.0 / N));
        }
        const double safety = 0.9 * (2 * newton_max_iter + 1) * fast_rcp((double)(2 * newton_max_iter + n_iter));
        const double pf = predict_factor<S>(hl, hl_old, error_norm, errl_old, havel_old);
        if (error_norm > 1.0) {
            factor = pf;
            have_factor = true;
            hl *= safety * factor;
            lu_valid = false;
            ++nrej;
            continue;
        }

        // ---- phase C (radau.py:607-657) -------------------------------------------------------
        const bool recompute_jac = (n_iter > 2) && have_rate && (rate > L.jac_rate);
        if (!have_factor) factor = safety * pf;
        if (!recompute_jac && L.db_lo <= factor && factor <= L.db_hi) factor = 1.0;
        else lu_valid = false;
        need_jac = recompute_jac;
        current_jac = recompute_jac;
        h_abs_old = h_abs;
        err_old = error_norm;
        have_old = true;
        h_abs = hl * factor;

        {
            double z[S];
#pragma unroll
            for (int l = 0; l < S; ++l) z[l] = Y[l] - y;
#pragma unroll
            for (int k = 0; k < S; ++k) {
                double acc = 0.0;
#pragma unroll
                for (int l = 0; l < S; ++l) acc = fma(z[l], tb.P[l * S + k], acc);
                ZP[k] = acc;
            }
        }
        t_old = t;
        h_old = h;
        have_sol = true;

        // my component of RadauDenseOutput at te (radau.py:681-699)
        const double rho = fast_rcp(h);
        auto dense_lane = [&](double te, double& av, double& bv) {
            const double x = (te - t) * rho;
            double xk = 1.0;
            av = 0.0; bv = 0.0;
#pragma unroll
            for (int k = 0; k < S; ++k) {
                bv = fma(ZP[k], (k + 1) * rho * xk, bv);
                xk *= x;
                av = fma(ZP[k], xk, av);
            }
            av += y;
        };
        // events on this step's dense output (dae.py:466-490); a terminal one moves the step end
        double t_lim = t_new, y_stop = 0.0, yp_stop = 0.0;
        bool ev_stop = false;
        if (!kLean && L.ev.spec.n_events > 0) {
            auto dense = [&](double te, int comp, int kind) -> double {
                double av, bv;
                dense_lane(te, av, bv);
                if constexpr (EventOf<P>::kNum > 0) {
                    if (kind == 2) {
                        double ya[N], ypa[N];
#pragma unroll
                        for (int j = 0; j < N; ++j) { ya[j] = g.bcast(av, j); ypa[j] = g.bcast(bv, j); }
                        return EventOf<P>::eval(comp, te, ya, ypa, p);
                    }
                }
                return g.bcast(kind == 1 ? bv : av, comp);
            };
            auto record = [&](int k, int idx, double tr) {
                double av, bv;
                dense_lane(tr, av, bv);
                const int64_t row = (int64_t)k * L.ev.cap + idx;
                if (lane == 0) L.ev.t_events[row * ld + inst] = tr;
                if (act) {
                    if (L.ev.y_events) L.ev.y_events[(row * N + lane) * ld + inst] = av;
                    if (L.ev.yp_events) L.ev.yp_events[(row * N + lane) * ld + inst] = bv;
                }
            };
            double gnew[B200DAE_MAX_EVENTS];
            for (int k = 0; k < B200DAE_MAX_EVENTS; ++k)
                gnew[k] = event_quantity_group<P, N>(g, L.ev.spec, k, t_new, Y[S - 1], Yp[S - 1], p) - L.ev.spec.value[k];
            double ts = t_new;
            if (handle_events(L.ev, evs, t, t_new, dir, gnew, dense, record, ts)) {
                ev_stop = true;
                t_lim = ts;
                dense_lane(ts, y_stop, yp_stop);
            }
        }
        if (!kLean && L.out.t_eval != nullptr) {
            while (ie < L.out.m) {
                const double te = __ldg(L.out.t_eval + ie);
                if (!(dir > 0.0 ? te <= t_lim : te >= t_lim)) break;
                double av, bv;
                dense_lane(te, av, bv);
                if (act) {
                    L.out.y_eval[ie * L.out.eval_stride_row + lane * L.out.eval_stride_comp + inst * L.out.eval_stride_inst] = av;
                    if (L.out.yp_eval) L.out.yp_eval[ie * L.out.eval_stride_row + lane * L.out.eval_stride_comp + inst * L.out.eval_stride_inst] = bv;
                }
                ++ie;
            }
        }
        t = t_lim;
        y = ev_stop ? y_stop : Y[S - 1];
        yp = ev_stop ? yp_stop : Yp[S - 1];
        ++nacc;
        if (!kLean && L.out.t_steps != nullptr && nst < L.out.steps_cap) {
            if (lane == 0) L.out.t_steps[(int64_t)nst * ld + inst] = t;
            if (act) {
                if (L.out.y_steps) L.out.y_steps[((int64_t)nst * N + lane) * ld + inst] = y;
                if (L.out.yp_steps) L.out.yp_steps[((int64_t)nst * N + lane) * ld + inst] = yp;
            }
            if (L.out.dense_coef != nullptr) {     // this step's RadauDenseOutput (radau.py:661-679)
                if (lane == 0) {
                    L.out.h_steps[((int64_t)nst * 2) * ld + inst] = h;
                    L.out.h_steps[((int64_t)nst * 2 + 1) * ld + inst] = t_new;
                    if (L.out.order_steps) L.out.order_steps[(int64_t)nst * ld + inst] = S;
                }
                if (act) {
#pragma unroll
                    for (int k = 0; k < S; ++k)
                        L.out.dense_coef[((int64_t)nst * (N * S) + lane * S + k) * ld + inst] = ZP[k];
                }
            }
            ++nst;
        }
        step_begin = true;
        if (ev_stop) status = B200DAE_ST_EVENT;
        else if (dir * (t - L.t_bound) >= 0.0) status = B200DAE_ST_FINISHED;
        else if (L.max_steps > 0 && nacc >= L.max_steps) status = B200DAE_ST_MAX_STEPS;
        if (status != B200DAE_ST_RUNNING && need_jac) ++njev;
    }

    // ---- results ----------------------------------------------------------------------------
    if (act) {
        L.out.y_final[io_index(L.io_aos, ld, N, lane, inst)] = y;
        if (L.out.yp_final) L.out.yp_final[io_index(L.io_aos, ld, N, lane, inst)] = yp;
    }
    if (lane == 0) {
        L.out.t_final[inst] = t;
        L.out.status[inst] = status;
        if (L.out.bad_input != nullptr && status == B200DAE_ST_NONFINITE && nacc == 0) *L.out.bad_input = 1;
        if (L.out.n_eval_valid) L.out.n_eval_valid[inst] = (int32_t)ie;
        if (L.out.n_steps) L.out.n_steps[inst] = nst;
        if (!kLean) for (int k = 0; k < L.ev.spec.n_events; ++k) L.ev.n_found[(int64_t)k * ld + inst] = evs.found[k];
        if (L.out.counters) {
            int32_t* c = L.out.counters + (L.io_aos ? inst * 8 : inst);
            const int64_t cs_ = L.io_aos ? 1 : ld;
            c[B200DAE_CNT_NFEV * cs_] = nfev;
            c[B200DAE_CNT_NJEV * cs_] = njev;
            c[B200DAE_CNT_NLU * cs_] = nlu;
            c[B200DAE_CNT_NSOLVE * cs_] = nsolve;
            c[B200DAE_CNT_ACCEPTED * cs_] = nacc;
            c[B200DAE_CNT_ERR_REJECTED * cs_] = nrej;
            c[B200DAE_CNT_NEWTON_FAIL * cs_] = nfail;
            c[B200DAE_CNT_NEWTON_ITER * cs_] = niter;
        }
    }
}

}  // namespace b200dae
