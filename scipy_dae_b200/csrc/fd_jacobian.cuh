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
    scale = fmax(fabs(f_at), fabs(fn_at));
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
template <class P, class V>
__device__ __forceinline__ void fd_jacobian(double t, double* __restrict__ y, double* __restrict__ yp,
                                            const double* __restrict__ p, const double* __restrict__ thr,
                                            V factor_y, V factor_yp, V Jy, V Jyp) {
    constexpr int N = P::N;
    constexpr int U = Regime<N>::kUnroll;
    double f[N];
    P::eval(t, y, yp, p, f);
#pragma unroll U
    for (int which = 0; which < 2; ++which) {
        const bool wrt_yp = (which == 1);
        const double* x = wrt_yp ? yp : y;
        V factor = wrt_yp ? factor_yp : factor_y;
        V J = wrt_yp ? Jyp : Jy;
#pragma unroll U
        for (int i = 0; i < N; ++i) {
            const double xi = x[i];
            const double f_sign = (f[i] >= 0.0) ? 1.0 : -1.0;
            const double y_scale = f_sign * fmax(thr[i], fabs(xi));
            double fac = factor[i];
            double xp = fma(fac, y_scale, xi);
            double h = xp - xi;
            while (h == 0.0) {
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
            factor[i] = fmax(fac, kNjMinFactor);
        }
    }
}

}  // namespace b200dae
