// bdf.cuh — the whole variable-order BDF/NDF integration of one system per thread.
//
// One launch replaces, for N independent instances, BDFDAE.__init__ (bdf.py:224-261), the
// step loop + t_eval slicing of solve_dae (dae.py:446-518), BDFDAE._step_impl (bdf.py:263-415),
// solve_bdf_system (bdf.py:32-64), change_D / compute_R (bdf.py:14-29) and BdfDenseOutput
// (bdf.py:422-548).  Same flattened control flow as radau.cuh: one loop trip = one Newton solve,
// each heavy phase at a single code site.
//
// The quasi-constant-step difference array D[(max_order+3)][n] is indexed by the run-time order,
// so it lives in per-thread local memory (L1-resident: 9*n doubles per instance).
#pragma once
#include "rt_compat.cuh"
#include "../../include/b200dae.h"
#include "events.cuh"
#include "fd_jacobian.cuh"
#include "radau.cuh"   // finite_d, powi

// 1 (default): rolled FD-Jacobian column loop in the thread BDF kernel beyond n = 4 (fd_jacobian.cuh)
#ifndef B200DAE_BDF_FD_ROLL
#define B200DAE_BDF_FD_ROLL 1
#endif

namespace b200dae {

constexpr int kBdfNewtonMaxIter = 4;     // bdf.py:9
constexpr int kBdfRows = B200DAE_MAX_ORDER + 3;
constexpr int kBdfQ = B200DAE_MAX_ORDER + 1;

struct BdfLaunch {
    double t0, t_bound, dir;
    double rtol, max_step, first_step, newton_tol;
    double atol[B200DAE_MAX_N];
    double thr[B200DAE_MAX_N];
    double gamma[kBdfQ], alpha[kBdfQ], error_const[kBdfQ + 1];
    double U[kBdfQ * kBdfQ];   // compute_R(6, 1): its leading blocks are compute_R(order, 1)
    double neg_inv[kBdfQ + 2]; // neg_inv[k] = -1.0 / k, k = 1 .. max_order + 2 (exponents of the controller), from the host
    int32_t max_order;
    int64_t max_steps;
    int32_t analytic_jac;       // host-side dispatch only (kJac template parameter)
    int64_t N, ld;
    const double *y0, *yp0, *params;
    b200dae_outputs out;
    EventLaunch ev;             // ev.spec.n_events == 0: no events
    int32_t io_aos;             // instance-major main inputs / outputs, see RadauLaunch
};

// norm ** (-1 / k) of the step-size / order controller (bdf.py:353,398-404), one shared copy of
// libdevice's pow for all call sites (these kernels are instruction-fetch bound).  The exponent
// comes from a host-built table, and an infinite norm (order at its bound) short-cuts to the
// limit 0 that pow would return through its special-case path.
static __device__ __noinline__ double pow_ctrl(double x, double e) { return pow(x, e); }
__device__ __forceinline__ double norm_root(double norm, double e) {
    return (norm == INFINITY) ? 0.0 : pow_ctrl(norm, e);
}

// log2(x) to ~1e-13 for the ORDER SELECTION (bdf.py:398-404: factors = error_norms ** (-1 / k),
// np.argmax(factors), np.max(factors)).  The three candidate factors are compared through
// key_i = (-1 / k_i) log2(norm_i) (factor_i = 2 ** key_i, same order) and only the winner goes
// through pow: one libdevice pow (~200 instructions, a call) instead of three per selection.  The
// argmax can differ from the reference's only when two factors agree to ~1e-13 relative.
// 0 -> -inf (and subnormals with it: the factor is beyond any bound either way), inf -> inf, NaN -> NaN.
__device__ __forceinline__ double log2_key(double x) {
    int hi = __double2hiint(x);
    const int lo = __double2loint(x);
    int e = (hi >> 20) - 1023;
    hi = (hi & 0x000fffff) | 0x3ff00000;                    // mantissa in [1, 2)
    double m = __hiloint2double(hi, lo);
    if (m > 1.4142135623730951) { m *= 0.5; ++e; }         // [sqrt(1/2), sqrt(2))
    const double s = (m - 1.0) * fast_rcp_nz(m + 1.0);     // log m = 2 atanh(s), |s| < 0.1716
    const double z = s * s;
    double q = 1.0 / 15.0;
    q = fma(q, z, 1.0 / 13.0);
    q = fma(q, z, 1.0 / 11.0);
    q = fma(q, z, 1.0 / 9.0);
    q = fma(q, z, 1.0 / 7.0);
    q = fma(q, z, 1.0 / 5.0);
    q = fma(q, z, 1.0 / 3.0);
    q = fma(q, z, 1.0);
    const double r = fma(2.0 * 1.4426950408889634 * s, q, (double)e);
    const bool normal = (x >= 2.2250738585072014e-308) && (x < INFINITY);
    return normal ? r : ((x < 2.2250738585072014e-308) ? -INFINITY : x);    // x < tiny is false for NaN: NaN stays
}

// argmax and max of norms[i] ** neg_inv_i over the three candidate orders, np.argmax semantics
// (first maximum, a NaN wins): returns the index, `best` = the winning factor from ONE pow.
__device__ __forceinline__ int select_order(const double (&norms)[3], double e0, double e1, double e2, double& best) {
    const double k0 = e0 * log2_key(norms[0]), k1 = e1 * log2_key(norms[1]), k2 = e2 * log2_key(norms[2]);
    int arg = 0;
    double kb = k0;
    if (kb == kb && (k1 > kb || k1 != k1)) { kb = k1; arg = 1; }
    if (kb == kb && (k2 > kb || k2 != k2)) { kb = k2; arg = 2; }
    const double nb = (arg == 0) ? norms[0] : ((arg == 1) ? norms[1] : norms[2]);
    const double eb = (arg == 0) ? e0 : ((arg == 1) ? e1 : e2);
    best = norm_root(nb, eb);
    return arg;
}

// change_D (bdf.py:24-29): D[:order+1] = (R @ U).T @ D[:order+1], evaluated as U.T @ (R.T @ D)
// with R = cumprod_i (i - 1 - factor*j) / i generated column by column (bdf.py:14-21).
//
// Cost matters more than it looks: with one system per lane, SOME lane of a warp needs a rescale in
// nearly every trip of the flattened step loop, so the warp runs this function almost every trip
// (ncu, Robertson ensemble: 60 % of all instructions, 10 of 32 lanes active).  The inner loops are
// therefore unrolled over compile-time l (masked by l < q) and the division by l is a
// multiplication by the constant 1 / l as in the group kernel (<= 1 ulp from the reference's true
// division for l = 3, 5; exact for l = 1, 2, 4): the rolled form with q^2 IEEE divisions cost ~4x
// as many instructions.
template <int N>
__device__ __noinline__ void change_D(double* __restrict__ D, int order, double factor,
                                      const double* __restrict__ Umat) {
    double E[kBdfQ * N];
    const int q = order + 1;
#pragma unroll 1
    for (int j = 0; j < q; ++j) {
        double acc[N];
#pragma unroll
        for (int c = 0; c < N; ++c) acc[c] = D[c];          // R[0][j] = 1
        double r = 1.0;
#pragma unroll
        for (int l = 1; l < kBdfQ; ++l) {
            r *= (j == 0) ? 0.0 : ((l - 1) - factor * j) * (1.0 / l);
            if (l < q) {
#pragma unroll
                for (int c = 0; c < N; ++c) acc[c] = fma(r, D[l * N + c], acc[c]);
            }
        }
#pragma unroll
        for (int c = 0; c < N; ++c) E[j * N + c] = acc[c];
    }
#pragma unroll 1
    for (int i = 0; i < q; ++i) {
        double acc[N];
#pragma unroll
        for (int c = 0; c < N; ++c) acc[c] = 0.0;
#pragma unroll
        for (int l = 0; l < kBdfQ; ++l) {
            // U = compute_R(order, 1) is upper triangular (its entry (l, i) carries the factor l - 1 - i + ... = 0
            // for l > i, exactly), so rows l > i add exact zeros in the reference's product: skipped
            if (l <= i) {
                const double u = Umat[l * kBdfQ + i];
#pragma unroll
                for (int c = 0; c < N; ++c) acc[c] = fma(u, E[l * N + c], acc[c]);
            }
        }
#pragma unroll
        for (int c = 0; c < N; ++c) D[i * N + c] = acc[c];
    }
}

// kLean: no t_eval output, no step trace, no events -- those paths are compiled out (the kernel is
// instruction-fetch bound: ncu stall_no_inst 38 % with all options in, 7200 instructions).
// kSharedCold: Jy, Jyp and the FD step factors (touched once per Jacobian refresh / factorisation)
// live in shared memory [element][thread] instead of registers / local memory (as in radau.cuh).
template <class P, int kThreads, bool kJac = false, bool kLean = false, bool kSharedCold = false>
__global__ void __launch_bounds__(kThreads)
bdf_kernel(const __grid_constant__ BdfLaunch L) {
    constexpr int N = P::N, NP = P::NP;
    constexpr int U = Regime<N>::kUnroll;
    const int64_t inst = (int64_t)blockIdx.x * kThreads + threadIdx.x;
    if (inst >= L.N) return;
    const int64_t ld = L.ld;

    double p[NP > 0 ? NP : 1], y[N], yp[N];
#pragma unroll U
    for (int j = 0; j < NP; ++j) p[j] = L.params[io_index(L.io_aos, ld, NP, j, inst)];
#pragma unroll U
    for (int j = 0; j < N; ++j) { y[j] = L.y0[io_index(L.io_aos, ld, N, j, inst)]; yp[j] = L.yp0[io_index(L.io_aos, ld, N, j, inst)]; }

    constexpr int kCold = 2 * N * N + 2 * N;
    constexpr bool kDynCold = kSharedCold && ((size_t)kCold * kThreads * sizeof(double) > 48 * 1024);
    extern __shared__ __align__(16) double s_cold_dyn[];
    __shared__ double s_cold_static[(kSharedCold && !kDynCold) ? kCold * kThreads : 1];
    double r_cold[kSharedCold ? 1 : kCold];
    using V = Vec<kSharedCold, kThreads>;
    const V Jy{kSharedCold ? ((kDynCold ? s_cold_dyn : s_cold_static) + threadIdx.x) : r_cold};
    const V Jyp = Jy + N * N, factor_y = Jyp + N * N, factor_yp = factor_y + N;
    double LU[N * N];
    int piv[N];
    double D[kBdfRows * N];
    double y_predict[N], psi[N], inv_scale[N], y_new[N], yp_new[N], d[N], f[N];
#pragma unroll U
    for (int j = 0; j < N; ++j) factor_y[j] = factor_yp[j] = kNjInitFactor;

    const double dir = L.dir, rtol = L.rtol;
    double t = L.t0;
    double h_abs;
    if (L.first_step > 0.0) {
        h_abs = L.first_step;
    } else {                                               // select_initial_step, common.py:129-183
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
    for (int k = 0; k < kBdfRows * N; ++k) D[k] = 0.0;     // bdf.py:254-257
    for (int j = 0; j < N; ++j) { D[j] = y[j]; D[N + j] = yp[j] * h_abs * dir; }
    int order = 1, n_equal_steps = 0;
    const int max_order = L.max_order;

    bool need_jac = true, lu_valid = false, step_begin = true, retry = false, current_jac = false;
    int32_t nfev = 0, njev = 0, nlu = 0, nsolve = 0, nacc = 0, nrej = 0, nfail = 0, niter = 0;
    int32_t status = B200DAE_ST_RUNNING;
    int64_t ie = 0;
    int32_t nst = 0;

    if (!kLean && L.out.t_steps != nullptr && L.out.steps_cap > 0) {
        L.out.t_steps[inst] = t;
        for (int j = 0; j < N; ++j) {
            if (L.out.y_steps) L.out.y_steps[j * ld + inst] = y[j];
            if (L.out.yp_steps) L.out.yp_steps[j * ld + inst] = yp[j];
        }
        nst = 1;
    }
    if (L.t0 == L.t_bound) { status = B200DAE_ST_FINISHED; njev = 1; }
    EventState evs;                  // g = event(t0, y0, yp0), dae.py:436
    if (!kLean) for (int k = 0; k < B200DAE_MAX_EVENTS; ++k) {
        evs.count[k] = evs.found[k] = 0;
        evs.g[k] = (k < L.ev.spec.n_events) ? event_quantity<P, N>(L.ev.spec, k, t, y, yp, p) - L.ev.spec.value[k] : 0.0;
    }

    double hl = 0.0, min_step = 0.0, h = 0.0, t_new = 0.0, c = 0.0;

    while (status == B200DAE_ST_RUNNING) {
        // ---- phase J: FD Jacobian at the committed (t, y, yp) (base.py:297, bdf.py:330) ------
        if (need_jac) {
            if constexpr (kJac) analytic_jacobian<P, N>(t, y, yp, p, Jy, Jyp);
            else fd_jacobian<P, V, V, (B200DAE_BDF_FD_ROLL != 0)>(t, y, yp, p, L.thr, factor_y, factor_yp, Jy, Jyp);
            ++njev;
            need_jac = false;
        }
        // ---- phase B: a new step begins (bdf.py:269-293) --------------------------------------
        if (step_begin) {
            min_step = 10.0 * fabs(next_toward(t, dir) - t);
            if (h_abs > L.max_step) {
                hl = L.max_step;
                change_D<N>(D, order, L.max_step / h_abs, L.U);
                n_equal_steps = 0;
            } else if (h_abs < min_step) {
                hl = min_step;
                change_D<N>(D, order, min_step / h_abs, L.U);
                n_equal_steps = 0;
            } else {
                hl = h_abs;
            }
            current_jac = false;      // `self.jac is None` is False for the FD wrapper, bdf.py:293
            step_begin = false;
        }
        // ---- phase A: a new attempt (bdf.py:296-318) ---------------------------------------------
        if (!retry) {
            if (hl < min_step) { status = B200DAE_ST_TOO_SMALL_STEP; break; }
            h = hl * dir;
            t_new = t + h;
            if (dir * (t_new - L.t_bound) > 0.0) {
                t_new = L.t_bound;
                change_D<N>(D, order, fabs(t_new - t) / hl, L.U);
                n_equal_steps = 0;
                lu_valid = false;
            }
            h = t_new - t;
            hl = fabs(h);
            const double rh = fast_rcp(h);
            // loops over the order unrolled and masked (lanes of a warp run different orders, so a
            // rolled loop runs to the largest one anyway and pays the loop overhead on top)
#pragma unroll U
            for (int j = 0; j < N; ++j) {
                double acc = 0.0, ps = 0.0;
#pragma unroll
                for (int i = 0; i < kBdfQ; ++i)
                    if (i <= order) acc += D[i * N + j];
#pragma unroll
                for (int i = 1; i < kBdfQ; ++i)
                    if (i <= order) ps = fma(D[i * N + j], L.gamma[i], ps);
                y_predict[j] = acc;
                inv_scale[j] = fast_rcp(fma(rtol, fabs(acc), L.atol[j]));
                psi[j] = ps * rh;
            }
            c = L.alpha[order] * rh;
        }
        retry = false;

        // ---- phase F: LU of Jy + c*Jyp (bdf.py:320-321) -------------------------------------------
        if (!lu_valid) {
            FiniteAcc fin;
#pragma unroll U
            for (int k = 0; k < N * N; ++k) {
                const double v = fma(c, Jyp[k], Jy[k]);
                fin.add(v);
                LU[k] = v;
            }
            if (!fin.all_finite()) { status = B200DAE_ST_NONFINITE; break; }
            lu_factor<N>(LU, piv);
            ++nlu;
            lu_valid = true;
        }

        // ---- phase N: simplified Newton (solve_bdf_system, bdf.py:32-64) ---------------------------
        bool converged = false;
        int n_iter;
        {
#pragma unroll U
            for (int j = 0; j < N; ++j) { d[j] = 0.0; y_new[j] = y_predict[j]; }
            double inv_norm_old = 0.0;
            bool has_old = false;
            int k = 0;
            for (; k < kBdfNewtonMaxIter; ++k) {
#pragma unroll U
                for (int j = 0; j < N; ++j) yp_new[j] = fma(c, d[j], psi[j]);
                P::eval(t_new, y_new, yp_new, p, f);
                ++nfev;
                FiniteAcc fin;
#pragma unroll U
                for (int j = 0; j < N; ++j) { fin.add(f[j]); f[j] = -f[j]; }
                if (!fin.all_finite()) break;
                lu_solve<N>(LU, piv, f);                          // f <- dy
                ++nsolve;
                double ss = 0.0;
#pragma unroll U
                for (int j = 0; j < N; ++j) { const double q = f[j] * inv_scale[j]; ss = fma(q, q, ss); }
                const double dy_norm = fast_sqrt(ss * (1.0 / N));
                double rate = 0.0;
                if (has_old) rate = dy_norm * inv_norm_old;
                const double cc = dy_norm * fast_rcp(1.0 - rate);
                if (has_old && (rate >= 1.0 || powi<3>(rate, kBdfNewtonMaxIter - k) * cc > L.newton_tol))
                    break;                                        // BEFORE the update (bdf.py:52-53)
#pragma unroll U
                for (int j = 0; j < N; ++j) { y_new[j] += f[j]; d[j] += f[j]; }
                if (dy_norm == 0.0 || (has_old && rate * cc < L.newton_tol)) {
                    converged = true;
                    break;
                }
                inv_norm_old = fast_rcp(dy_norm);
                has_old = true;
            }
            n_iter = (k == kBdfNewtonMaxIter) ? kBdfNewtonMaxIter : k + 1;
            niter += n_iter;
        }
        if (!converged) {
            ++nfail;
            if (!current_jac) {                    // bdf.py:327-332: refresh J at the OLD point
                need_jac = true;
                current_jac = true;
                lu_valid = false;
                retry = true;
                continue;
            }
            hl *= 0.5;                             // bdf.py:334-340
            change_D<N>(D, order, 0.5, L.U);
            n_equal_steps = 0;
            lu_valid = false;
            continue;
        }

        // ---- phase E: local error (bdf.py:342-355) -------------------------------------------------
        const double safety = 0.9 * (2 * kBdfNewtonMaxIter + 1) * fast_rcp((double)(2 * kBdfNewtonMaxIter + n_iter));
        double error_norm;
        {
            const double ec = L.error_const[order];
            double ss = 0.0;
#pragma unroll U
            for (int j = 0; j < N; ++j) {
                inv_scale[j] = fast_rcp(fma(rtol, fabs(y_new[j]), L.atol[j]));
                const double q = ec * d[j] * inv_scale[j];
                ss = fma(q, q, ss);
            }
            error_norm = fast_sqrt(ss * (1.0 / N));
        }
        if (error_norm > 1.0) {
            const double factor = fmax(0.2, safety * norm_root(error_norm, L.neg_inv[order + 1]));
            hl *= factor;
            change_D<N>(D, order, factor, L.U);
            n_equal_steps = 0;                     // LU is kept (bdf.py:353-354)
            ++nrej;
            continue;
        }

        // ---- phase C: accepted (bdf.py:359-415) ----------------------------------------------------
        ++n_equal_steps;
        const double t_prev = t;
        t = t_new;
#pragma unroll U
        for (int j = 0; j < N; ++j) { y[j] = y_new[j]; yp[j] = yp_new[j]; }
        h_abs = hl;
        ++nacc;
        for (int j = 0; j < N; ++j) {              // bdf.py:374-377
            D[(order + 2) * N + j] = d[j] - D[(order + 1) * N + j];
            D[(order + 1) * N + j] = d[j];
        }
#pragma unroll
        for (int i = kBdfQ - 1; i >= 0; --i)
            if (i <= order) {
#pragma unroll U
                for (int j = 0; j < N; ++j) D[i * N + j] += D[(i + 1) * N + j];
            }

        if (n_equal_steps >= order + 1) {          // order / step selection, bdf.py:379-413
            double norms[3];
            norms[1] = error_norm;
            if (order > 1) {
                const double ec = L.error_const[order - 1];
                double ss = 0.0;
                for (int j = 0; j < N; ++j) { const double q = ec * D[order * N + j] * inv_scale[j]; ss = fma(q, q, ss); }
                norms[0] = sqrt(ss * (1.0 / N));
            } else {
                norms[0] = INFINITY;
            }
            if (order < max_order) {
                const double ec = L.error_const[order + 1];
                double ss = 0.0;
                for (int j = 0; j < N; ++j) { const double q = ec * D[(order + 2) * N + j] * inv_scale[j]; ss = fma(q, q, ss); }
                norms[2] = sqrt(ss * (1.0 / N));
            } else {
                norms[2] = INFINITY;
            }
            double best;
            const int arg = select_order(norms, L.neg_inv[order], L.neg_inv[order + 1], L.neg_inv[order + 2], best);
            order += arg - 1;
            const double factor = fmin(10.0, safety * best);
            h_abs *= factor;
            change_D<N>(D, order, factor, L.U);
            n_equal_steps = 0;
            lu_valid = false;
        }

        // BdfDenseOutput is built AFTER the rescale with h = h_abs*direction and the NEW order
        // (bdf.py:417-419, 525-548): value and derivative of the Newton interpolant at te
        const double hh = h_abs * dir;
        const double t_end = t;                  // the interpolant is anchored at the step end
        auto dense_all = [&](double te, double* ya, double* ypa) {
            for (int j = 0; j < N; ++j) { ya[j] = D[j]; ypa[j] = 0.0; }
            double fact = 1.0, hj = 1.0;
            for (int jj = 1; jj <= order; ++jj) {
                fact *= jj;
                hj *= hh;
                double fac = 1.0, dfac = 0.0;
                for (int mm = 0; mm < jj; ++mm) {
                    fac *= te - (t_end - mm * hh);
                    double dprod = 1.0;
                    for (int i = 0; i < jj; ++i)
                        if (i != mm) dprod *= te - (t_end - i * hh);
                    dfac += dprod;
                }
                const double rden = 1.0 / (fact * hj);
                for (int j = 0; j < N; ++j) {
                    ya[j] = fma(D[jj * N + j], fac * rden, ya[j]);
                    ypa[j] = fma(D[jj * N + j], dfac * rden, ypa[j]);
                }
            }
        };
        // events on this step's dense output (dae.py:466-490); a terminal one moves the step end
        bool ev_stop = false;
        if (!kLean && L.ev.spec.n_events > 0) {
            auto dense = [&](double te, int comp, int kind) -> double {
                double ya[N], ypa[N];
                dense_all(te, ya, ypa);
                if constexpr (EventOf<P>::kNum > 0) {
                    if (kind == 2) return EventOf<P>::eval(comp, te, ya, ypa, p);
                }
                return kind == 1 ? ypa[comp] : ya[comp];
            };
            auto record = [&](int k, int idx, double tr) {
                double ya[N], ypa[N];
                dense_all(tr, ya, ypa);
                const int64_t row = (int64_t)k * L.ev.cap + idx;
                L.ev.t_events[row * ld + inst] = tr;
                for (int j = 0; j < N; ++j) {
                    if (L.ev.y_events) L.ev.y_events[(row * N + j) * ld + inst] = ya[j];
                    if (L.ev.yp_events) L.ev.yp_events[(row * N + j) * ld + inst] = ypa[j];
                }
            };
            double gnew[B200DAE_MAX_EVENTS];
            for (int k = 0; k < B200DAE_MAX_EVENTS; ++k)
                gnew[k] = event_quantity<P, N>(L.ev.spec, k, t, y, yp, p) - L.ev.spec.value[k];
            double ts = t;
            if (handle_events(L.ev, evs, t_prev, t, dir, gnew, dense, record, ts)) {
                ev_stop = true;
                double ya[N], ypa[N];
                dense_all(ts, ya, ypa);
                t = ts;
                for (int j = 0; j < N; ++j) { y[j] = ya[j]; yp[j] = ypa[j]; }
            }
        }
        if (!kLean && L.out.t_eval != nullptr) {
            while (ie < L.out.m) {
                const double te = __ldg(L.out.t_eval + ie);
                if (!(dir > 0.0 ? te <= t : te >= t)) break;
                double* ye = L.out.y_eval + ie * L.out.eval_stride_row + inst * L.out.eval_stride_inst;
                double* ype = L.out.yp_eval ? L.out.yp_eval + ie * L.out.eval_stride_row + inst * L.out.eval_stride_inst : nullptr;
                double ya[N], ypa[N];
                dense_all(te, ya, ypa);
                for (int j = 0; j < N; ++j) {
                    ye[j * L.out.eval_stride_comp] = ya[j];
                    if (ype) ype[j * L.out.eval_stride_comp] = ypa[j];
                }
                ++ie;
            }
        }
        if (!kLean && L.out.t_steps != nullptr && nst < L.out.steps_cap) {
            L.out.t_steps[(int64_t)nst * ld + inst] = t;
            for (int j = 0; j < N; ++j) {
                if (L.out.y_steps) L.out.y_steps[((int64_t)nst * N + j) * ld + inst] = y[j];
                if (L.out.yp_steps) L.out.yp_steps[((int64_t)nst * N + j) * ld + inst] = yp[j];
            }
            if (L.out.dense_coef != nullptr) {     // this step's BdfDenseOutput (bdf.py:417-419)
                L.out.h_steps[((int64_t)nst * 2) * ld + inst] = hh;
                L.out.h_steps[((int64_t)nst * 2 + 1) * ld + inst] = t_end;
                if (L.out.order_steps) L.out.order_steps[(int64_t)nst * ld + inst] = order;
                for (int i = 0; i <= order; ++i)
                    for (int j = 0; j < N; ++j)
                        L.out.dense_coef[((int64_t)nst * (kBdfQ * N) + i * N + j) * ld + inst] = D[i * N + j];
            }
            ++nst;
        }
        step_begin = true;
        if (ev_stop) status = B200DAE_ST_EVENT;
        else if (dir * (t - L.t_bound) >= 0.0) status = B200DAE_ST_FINISHED;
        else if (L.max_steps > 0 && nacc >= L.max_steps) status = B200DAE_ST_MAX_STEPS;
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
    if (!kLean) for (int k = 0; k < L.ev.spec.n_events; ++k) L.ev.n_found[(int64_t)k * ld + inst] = evs.found[k];
    if (L.out.counters) {
        int32_t* cn = L.out.counters + (L.io_aos ? inst * 8 : inst);
            const int64_t cs_ = L.io_aos ? 1 : ld;
        cn[B200DAE_CNT_NFEV * cs_] = nfev;
        cn[B200DAE_CNT_NJEV * cs_] = njev;
        cn[B200DAE_CNT_NLU * cs_] = nlu;
        cn[B200DAE_CNT_NSOLVE * cs_] = nsolve;
        cn[B200DAE_CNT_ACCEPTED * cs_] = nacc;
        cn[B200DAE_CNT_ERR_REJECTED * cs_] = nrej;
        cn[B200DAE_CNT_NEWTON_FAIL * cs_] = nfail;
        cn[B200DAE_CNT_NEWTON_ITER * cs_] = niter;
    }
}

}  // namespace b200dae
