// fd_jacobian.cuh — forward-difference Jacobians Jy = dF/dy, Jyp = dF/dy' with per-column
// adaptive step factors carried across the whole solve.
//
// Replaces jac_wrapped (base.py:264-295), which calls scipy's num_jac / _dense_num_jac twice
// (scipy/integrate/_ivp/common.py:268-382) with threshold = atol / rtol (base.py:277):
//   f_sign = +1 where f >= 0 else -1 (indexed like the perturbed variable);
//   y_scale = f_sign * max(threshold, |x|);  h = (x + factor*y_scale) - x;  factor *= 10 while h == 0;
//   column: diff = F(x + h e_i) - f;  max_ind = argmax|diff|;  scale = max(|f[max_ind]|, |f_new[max_ind]|);
//   if max_diff < EPS^0.875 * scale: retry with 10*factor, keep it if max_diff*scale_new < max_diff_new*scale;
//   J[:, i] = diff / h;  factor *= 10 if max_diff < EPS^0.75*scale;  factor *= 0.1 if max_diff > EPS^0.25*scale;
//   factor = max(factor, 1e3*EPS).
// The FD residual evaluations are not counted in nfev (base.py:83-87,267).
#pragma once
#include "dense_lu.cuh"

namespace b200dae {

constexpr double kEps = 2.220446049250313e-16;
constexpr double kNjDiffReject = 2.0194839173657902e-14;   // EPS ** 0.875
constexpr double kNjDiffSmall = 1.8189894035458565e-12;    // EPS ** 0.75
constexpr double kNjDiffBig = 1.220703125e-04;             // EPS ** 0.25
constexpr double kNjMinFactor = 1e3 * kEps;
constexpr double kNjInitFactor = 1.4901161193847656e-08;   // EPS ** 0.5

// One column of the difference quotient.  `x` is y or yp (the perturbed vector), `col` the
// column, `xp` the perturbed entry.  Returns max_diff and scale (NaN-propagating like
// np.argmax / np.maximum).
template <class P>
__device__ __forceinline__ void fd_column(double t, double* __restrict__ y, double* __restrict__ yp,
                                          const double* __restrict__ p, const double* __restrict__ f,
                                          bool wrt_yp, int col, double xp,
                                          double* __restrict__ diff, double& max_diff, double& scale) {
    constexpr int N = P::N;
    constexpr int U = Regime<N>::kUnroll;
    double fnew[N];
    if constexpr (Regime<N>::kReg) {
        double yy[N], yyp[N];
#pragma unroll
        for (int j = 0; j < N; ++j) {
            yy[j] = (!wrt_yp && j == col) ? xp : y[j];
            yyp[j] = (wrt_yp && j == col) ? xp : yp[j];
        }
        P::eval(t, yy, yyp, p, fnew);
    } else {
        double* x = wrt_yp ? yp : y;
        const double saved = x[col];
        x[col] = xp;
        P::eval(t, y, yp, p, fnew);
        x[col] = saved;
    }
    double best = -1.0, f_at = 0.0, fn_at = 0.0, d_at = 0.0;
#pragma unroll U
    for (int r = 0; r < N; ++r) {
        const double dr = fnew[r] - f[r];
        diff[r] = dr;
        const double a = fabs(dr);
        if (a > best || (a != a && best == best)) { best = a; f_at = f[r]; fn_at = fnew[r]; d_at = dr; }
    }
    max_diff = fabs(d_at);
    scale = amax_sel(f_at, fn_at);
    if (f_at != f_at || fn_at != fn_at) scale = nan("");
}

// Analytic Jacobians from the functor (the reference's `jac=` callable path, base.py:299-331):
// zero both matrices, let the functor fill its structural non-zeros.  `V` is any indexable view
// with row-major N x N entries at stride `LD`.
template <class V, int LD>
struct MatRef {
    V v;
    __device__ __forceinline__ double& operator()(int r, int c) const { return v[r * LD + c]; }
};

template <class P, int LD, class V>
__device__ __forceinline__ void analytic_jacobian(double t, const double* __restrict__ y,
                                                  const double* __restrict__ yp,
                                                  const double* __restrict__ p, V Jy, V Jyp) {
    constexpr int N = P::N;
    constexpr int U = Regime<N>::kUnroll;
#pragma unroll U
    for (int r = 0; r < N; ++r)
#pragma unroll U
        for (int c = 0; c < N; ++c) { Jy[r * LD + c] = 0.0; Jyp[r * LD + c] = 0.0; }
    P::jac(t, y, yp, p, MatRef<V, LD>{Jy}, MatRef<V, LD>{Jyp});
}

// ---------------------------------------------------------------------------------------------
// Batched form for the register regime (n <= 4), B200DAE_FD_BATCH = 1 (the y columns, then the y'
// columns) or 2 (all 2n columns at once).
//
// The column-at-a-time form below is a chain of ~25 dependent FP64 operations per column
// (perturb -> residual -> argmax -> scale test -> reciprocal -> quotient), seven times per
// Jacobian, with nothing else for the warp to issue: 29 % of the Robertson solve time for 15 % of
// its instructions.  Here the FIRST attempts of all columns of a batch form one basic block (n or
// 2n independent chains), the second attempts (a column of Jyp that belongs to an algebraic
// variable always takes one: about one column per Robertson Jacobian) follow as predicated
// blocks, and the reciprocals / quotients / factor updates of all columns close the batch.
// Same arithmetic per column as the reference's vectorised num_jac, which also evaluates all
// first attempts before any second one (SP/common.py:330-371).
//
// Reduction of a column: only max_diff and scale leave it.  np.argmax returns the first NaN if
// there is one (max_diff = NaN: no second attempt, no factor update); that case is restored by
// max_diff = |best| + 0 * (sum of the other differences), NaN exactly when a difference is
// non-finite (then J is non-finite and the solve ends with B200DAE_ST_NONFINITE either way).
template <class P>
__device__ __forceinline__ void fd_probe(double t, const double* __restrict__ y, const double* __restrict__ yp,
                                         const double* __restrict__ p, const double* __restrict__ f,
                                         int col, double xp,
                                         double* __restrict__ diff, double& max_diff, double& scale) {
    constexpr int N = P::N;
    double yy[N], yyp[N], fnew[N];
#pragma unroll
    for (int j = 0; j < N; ++j) {
        yy[j] = (col == j) ? xp : y[j];
        yyp[j] = (col == N + j) ? xp : yp[j];
    }
    P::eval(t, yy, yyp, p, fnew);
    // running argmax on the raw differences (|.| is a free operand modifier of the compare; a
    // materialised fabs() is one more FP64 instruction per row), carrying f and f_new of that row
    double bd = 0.0, f_at = 0.0, fn_at = 0.0, sum = 0.0;
#pragma unroll
    for (int r = 0; r < N; ++r) {
        const double dr = fnew[r] - f[r];
        diff[r] = dr;
        if (r == 0) { bd = dr; f_at = f[r]; fn_at = fnew[r]; }
        else {
            sum += dr;                       // a NaN in row 0 already sticks to bd
            if (fabs(dr) > fabs(bd)) { bd = dr; f_at = f[r]; fn_at = fnew[r]; }
        }
    }
    max_diff = fma(0.0, sum, fabs(bd));
    scale = amax_sel(f_at, fn_at);
}

template <class P, class V, class VJ, int C0, int C1>
__device__ __forceinline__ void fd_batch(double t, const double* __restrict__ y, const double* __restrict__ yp,
                                         const double* __restrict__ p, const double* __restrict__ thr,
                                         const double* __restrict__ f,
                                         V factor_y, V factor_yp, VJ Jy, VJ Jyp) {
    constexpr int N = P::N, NC = C1 - C0;
    double fac[NC], ysc[NC], h[NC], xp[NC], diff[NC][N], md[NC], sc[NC];
    bool any_zero = false;
#pragma unroll
    for (int c = 0; c < NC; ++c) {
        const int col = C0 + c, i = col % N;
        const double xi = (col >= N) ? yp[i] : y[i];
        ysc[c] = amax_signed(thr[i], xi, f[i]);            // f_sign * max(thr, |x|), thr > 0
        fac[c] = (col >= N) ? factor_yp[i] : factor_y[i];
        xp[c] = fma(fac[c], ysc[c], xi);
        h[c] = xp[c] - xi;
        any_zero = any_zero || (h[c] == 0.0);
    }
    if (any_zero) {                                   // rare: the step vanished in round-off
#pragma unroll
        for (int c = 0; c < NC; ++c) {
            const int col = C0 + c, i = col % N;
            const double xi = (col >= N) ? yp[i] : y[i];
            // bounded: with atol = 0 and x = 0 the scale is 0 and scipy's `while h == 0` never ends; here
            // h stays 0 after 400 trips (factor has overflowed), J becomes non-finite -> status -2
            for (int trip = 0; h[c] == 0.0 && trip < 400; ++trip) {
                fac[c] *= 10.0;
                xp[c] = fma(fac[c], ysc[c], xi);
                h[c] = xp[c] - xi;
            }
        }
    }
#pragma unroll
    for (int c = 0; c < NC; ++c) fd_probe<P>(t, y, yp, p, f, C0 + c, xp[c], diff[c], md[c], sc[c]);
    // second attempt with 10 * factor where the difference drowns in round-off (SP/common.py:350-371)
#pragma unroll
    for (int c = 0; c < NC; ++c) {
        if (md[c] < kNjDiffReject * sc[c]) {
            const int col = C0 + c, i = col % N;
            const double xi = (col >= N) ? yp[i] : y[i];
            const double fac2 = 10.0 * fac[c];
            const double xp2 = fma(fac2, ysc[c], xi);
            const double h2 = xp2 - xi;
            double d2[N], md2, sc2;
            fd_probe<P>(t, y, yp, p, f, col, xp2, d2, md2, sc2);
            if (md[c] * sc2 < md2 * sc[c]) {
                fac[c] = fac2;
                h[c] = h2;
#pragma unroll
                for (int r = 0; r < N; ++r) diff[c][r] = d2[r];
                sc[c] = sc2;
                md[c] = md2;
            }
        }
    }
#pragma unroll
    for (int c = 0; c < NC; ++c) {
        const int col = C0 + c, i = col % N;
        const double rh = fast_rcp(h[c]);     // h = inf (a factor that grew for 300+ refreshes on an algebraic column) -> 0
#pragma unroll
        for (int r = 0; r < N; ++r) {
            if (col >= N) Jyp[r * N + i] = diff[c][r] * rh;
            else Jy[r * N + i] = diff[c][r] * rh;
        }
        double fc = fac[c];
        if (md[c] < kNjDiffSmall * sc[c]) fc *= 10.0;
        if (md[c] > kNjDiffBig * sc[c]) fc *= 0.1;
        fc = max_sel(kNjMinFactor, fc);
        if (col >= N) factor_yp[i] = fc;
        else factor_y[i] = fc;
    }
}

#ifndef B200DAE_FD_ROLL_MINN
#define B200DAE_FD_ROLL_MINN 4      // rolled FD columns (thread BDF kernel) only beyond this system size
                                     // (measured with 0: Robertson BDF 16.13 -> 16.44 M solves/s, +2 %; left at 4)
#endif
#ifndef B200DAE_FD_BATCH
#define B200DAE_FD_BATCH 1
#endif

// Jy and Jyp (row-major N*N) at (t, y, yp).  thr[i] = atol[i] / rtol.
//
// Tried and rejected (round 1, B200): a branch-free form for n <= 4 that evaluates both attempts
// of every column and picks by selects (one basic block of 4n residual evaluations).  It is
// algorithmically identical (checked against this form on 20 000 random states), but it spilled
// at the 168-register cap and ran 16.6 ms instead of 15.3 ms per 2^20 Robertson solves, and the
// different FMA contraction of the inlined residuals moved the step sequences (config 1: nfev
// 914 instead of the reference's 911; +6 % accepted steps on the vector-atol golden).  The FD
// phase is latency-bound (29 % of the solve time for 15 % of the instructions: with the analytic
// Jacobian the same launch takes 10.9 ms); the second attempt is not rare either: a column of
// Jyp that belongs to an algebraic variable has diff = 0 and always takes it (16.4 % of the
// Robertson columns, counted with the CPU oracle).
template <class P, class V, class VJ = V, bool kRollColumns = false>
__device__ __forceinline__ void fd_jacobian(double t, double* __restrict__ y, double* __restrict__ yp,
                                            const double* __restrict__ p, const double* __restrict__ thr,
                                            V factor_y, V factor_yp, VJ Jy, VJ Jyp) {
    constexpr int N = P::N;
    constexpr int U = Regime<N>::kUnroll;
    double f[N];
    P::eval(t, y, yp, p, f);
    // kRollColumns (the thread BDF kernel, n > 4): the columns run through ONE rolled loop body instead of 2n
    // inlined residual evaluations.  The unrolled n = 6 BDF kernel is 104 KB of code with 8 resident warps in
    // different phases and stalls on instruction fetch (ncu: stall_no_inst 34 %); rolled it is 67 KB and Akzo
    // Nobel BDF goes from 6.2 M to 7.3 M solves/s with identical work counters.  The Radau kernel of the same
    // regime runs 2 warps per SM and is latency-bound instead: rolled it LOSES 19 % (9.1 -> 7.4 M), so it
    // keeps the batched form.
    constexpr bool kRollBig = kRollColumns && N > B200DAE_FD_ROLL_MINN;
    constexpr int UF = kRollBig ? 1 : U;
    if constexpr (kRollBig) {
        // falls through to the column loop below
    } else
    if constexpr (Regime<N>::kReg && B200DAE_FD_BATCH == 2) {
        fd_batch<P, V, VJ, 0, 2 * N>(t, y, yp, p, thr, f, factor_y, factor_yp, Jy, Jyp);
        return;
    } else if constexpr (Regime<N>::kReg && B200DAE_FD_BATCH == 1) {
        fd_batch<P, V, VJ, 0, N>(t, y, yp, p, thr, f, factor_y, factor_yp, Jy, Jyp);
        fd_batch<P, V, VJ, N, 2 * N>(t, y, yp, p, thr, f, factor_y, factor_yp, Jy, Jyp);
        return;
    }
#pragma unroll UF
    for (int which = 0; which < 2; ++which) {
        const bool wrt_yp = (which == 1);
        const double* x = wrt_yp ? yp : y;
        V factor = wrt_yp ? factor_yp : factor_y;
        VJ J = wrt_yp ? Jyp : Jy;
#pragma unroll UF
        for (int i = 0; i < N; ++i) {
            const double xi = x[i];
            const double f_sign = (f[i] >= 0.0) ? 1.0 : -1.0;
            const double y_scale = f_sign * amax_sel(thr[i], xi);   // thr > 0
            double fac = factor[i];
            double xp = fma(fac, y_scale, xi);
            double h = xp - xi;
            for (int trip = 0; h == 0.0 && trip < 400; ++trip) {   // bounded, see fd_batch
                fac *= 10.0;
                xp = fma(fac, y_scale, xi);
                h = xp - xi;
            }
            // first try with `fac`; if the difference drowns in round-off (SP/common.py:350-371)
            // a second try with 10*fac replaces it when it resolves the column better.  Written
            // as a 2-trip rolled loop so the residual evaluation is emitted once per column.
            double diff[N], max_diff = 0.0, scale = 0.0;
            double fac_try = fac, xp_try = xp, h_try = h;
#pragma unroll 1
            for (int attempt = 0; attempt < 2; ++attempt) {
                double d2[N], md2, sc2;
                fd_column<P>(t, y, yp, p, f, wrt_yp, i, xp_try, d2, md2, sc2);
                if (attempt == 0 || max_diff * sc2 < md2 * scale) {
                    fac = fac_try;
                    h = h_try;
#pragma unroll U
                    for (int r = 0; r < N; ++r) diff[r] = d2[r];
                    scale = sc2;
                    max_diff = md2;
                }
                if (attempt == 1 || !(max_diff < kNjDiffReject * scale)) break;
                fac_try = 10.0 * fac;
                xp_try = fma(fac_try, y_scale, xi);
                h_try = xp_try - xi;
            }
            const double rh = fast_rcp(h);
#pragma unroll U
            for (int r = 0; r < N; ++r) J[r * N + i] = diff[r] * rh;
            if (max_diff < kNjDiffSmall * scale) fac *= 10.0;
            if (max_diff > kNjDiffBig * scale) fac *= 0.1;
            factor[i] = max_sel(kNjMinFactor, fac);
        }
    }
}

}  // namespace b200dae
