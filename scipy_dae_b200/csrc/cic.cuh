// cic.cuh — batched consistent initial conditions, one system per thread.
//
// Replaces, for N independent instances, consistent_initial_conditions with its helpers
// solve_underdetermined_system and qrank (scipy_dae/integrate/_dae/common.py:186-372, after
// Shampine 2002): a Newton / chord iteration on 0 = F(t0, y0 + dy, yp0 + dyp) whose
// under-determined linear systems 0 = f + Jy dy + Jyp dyp are solved with as many components of
// (dy, dyp) as possible left at zero, using two QR factorisations with column pivoting
// (scipy.linalg.qr(pivoting=True) = LAPACK dgeqp3 / dlaqp2: partial column norms with the
// downdate-or-recompute rule, Householder vectors from dlarfg).  The default Jacobian is the
// reference's approx_derivative(method="2-point") on z = (y, yp)
// (scipy/optimize/_numdiff.py:147-195, 683-712); with kJac the functor's analytic Jacobians.
//
// This runs once per instance before the integration, so it is written for clarity: rolled
// loops over run-time sizes, matrices in per-thread local memory (5 n^2 doubles).
#pragma once
#include "rt_compat.cuh"
#include "../../include/b200dae.h"
#include "fd_jacobian.cuh"

namespace b200dae {

struct CicLaunch {
    double t0, rtol, safety;
    double atol[B200DAE_MAX_N];
    int32_t fixed_y[B200DAE_MAX_N], fixed_yp[B200DAE_MAX_N];
    int32_t newton_maxiter, chord_iter, analytic_jac;
    int64_t N, ld;
    double *y0, *yp0;            // [n][ld], updated in place
    const double* params;        // [n_params][ld]
    double* f_out;               // [n][ld] or NULL
    int32_t* status;             // [ld]
    int32_t* n_iter;             // [ld] or NULL
};

// A (m x n, row-major, leading dimension lda): R in the upper triangle, Householder vectors below
// the diagonal, tau[min(m,n)], jpvt[n].  vn1 / vn2: scratch of n doubles each.
__device__ inline void qr_colpiv(int m, int n, double* A, int lda, double* tau, int* jpvt,
                                 double* vn1, double* vn2) {
    const double tol3z = 1.4901161193847656e-08;            // sqrt(EPS)
    for (int j = 0; j < n; ++j) {
        double s = 0.0;
        for (int i = 0; i < m; ++i) s = fma(A[i * lda + j], A[i * lda + j], s);
        vn1[j] = vn2[j] = sqrt(s);
        jpvt[j] = j;
    }
    const int mn = m < n ? m : n;
    for (int k = 0; k < mn; ++k) {
        int pvt = k;                                         // idamax: first maximum
        for (int j = k + 1; j < n; ++j)
            if (vn1[j] > vn1[pvt]) pvt = j;
        if (pvt != k) {
            for (int i = 0; i < m; ++i) {
                const double x = A[i * lda + pvt];
                A[i * lda + pvt] = A[i * lda + k];
                A[i * lda + k] = x;
            }
            const int ti = jpvt[pvt]; jpvt[pvt] = jpvt[k]; jpvt[k] = ti;
            vn1[pvt] = vn1[k]; vn2[pvt] = vn2[k];
        }
        // dlarfg on A[k:m, k]
        const double alpha = A[k * lda + k];
        double xn = 0.0;
        for (int i = k + 1; i < m; ++i) xn = fma(A[i * lda + k], A[i * lda + k], xn);
        xn = sqrt(xn);
        double tk = 0.0;
        if (xn != 0.0) {
            const double beta = -copysign(hypot(alpha, xn), alpha);
            tk = (beta - alpha) / beta;
            const double sc = 1.0 / (alpha - beta);
            for (int i = k + 1; i < m; ++i) A[i * lda + k] *= sc;
            A[k * lda + k] = beta;
        }
        tau[k] = tk;
        if (tk != 0.0)
            for (int j = k + 1; j < n; ++j) {               // a_j -= tau v (v^T a_j)
                double w = A[k * lda + j];
                for (int i = k + 1; i < m; ++i) w = fma(A[i * lda + k], A[i * lda + j], w);
                w *= tk;
                A[k * lda + j] -= w;
                for (int i = k + 1; i < m; ++i) A[i * lda + j] = fma(-A[i * lda + k], w, A[i * lda + j]);
            }
        for (int j = k + 1; j < n; ++j)                     // dlaqp2 norm downdate
            if (vn1[j] != 0.0) {
                const double r = fabs(A[k * lda + j]) / vn1[j];
                double temp = 1.0 - r * r;
                if (temp < 0.0) temp = 0.0;
                const double q = vn1[j] / vn2[j];
                if (temp * q * q <= tol3z) {
                    double s = 0.0;
                    for (int i = k + 1; i < m; ++i) s = fma(A[i * lda + j], A[i * lda + j], s);
                    vn1[j] = vn2[j] = sqrt(s);
                } else {
                    vn1[j] *= sqrt(temp);
                }
            }
    }
}

__device__ inline void apply_qt(int m, int k, const double* A, int lda, const double* tau, double* b) {
    for (int r = 0; r < k; ++r) {
        if (tau[r] == 0.0) continue;
        double w = b[r];
        for (int i = r + 1; i < m; ++i) w = fma(A[i * lda + r], b[i], w);
        w *= tau[r];
        b[r] -= w;
        for (int i = r + 1; i < m; ++i) b[i] = fma(-A[i * lda + r], w, b[i]);
    }
}

__device__ inline int qrank_of(int m, int n, const double* A, int lda) {   // common.py:261-266
    const int mn = m < n ? m : n;
    if (mn == 0) return 0;
    const double tol = (m > n ? m : n) * kEps * fabs(A[0]);
    int rank = 0;
    for (int k = 0; k < mn; ++k) rank += fabs(A[k * lda + k]) > tol;
    return rank;
}

__device__ inline void back_subst(int r, const double* R, int lda, double* x) {
    for (int i = r - 1; i >= 0; --i) {
        double s = x[i];
        for (int j = i + 1; j < r; ++j) s = fma(-R[i * lda + j], x[j], s);
        x[i] = s / R[i * lda + i];
    }
}

template <class P, bool kJac>
__global__ void __launch_bounds__(32)
cic_kernel(const __grid_constant__ CicLaunch L) {
    constexpr int N = P::N, NP = P::NP;
    const int64_t inst = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (inst >= L.N) return;
    const int64_t ld = L.ld;

    double p[NP > 0 ? NP : 1], y[N], yp[N], f[N], f1[N], scale[N], dy[N], dyp[N];
    double Jy[N * N], Jyp[N * N], A[N * N], S[N * N], S2[N * N];
    double tau[N], tau2[N], d[N], w[N], x[N], c[N], vn1[N], vn2[N];
    int pv[N], pS[N], free_y[N], free_yp[N];
    for (int j = 0; j < NP; ++j) p[j] = L.params[j * ld + inst];
    for (int j = 0; j < N; ++j) { y[j] = L.y0[j * ld + inst]; yp[j] = L.yp0[j * ld + inst]; }
    int nfy = 0, nfyp = 0;
    for (int j = 0; j < N; ++j) {
        if (!L.fixed_y[j]) free_y[nfy++] = j;
        if (!L.fixed_yp[j]) free_yp[nfyp++] = j;
    }
    const int fixed = (N - nfy) + (N - nfyp);
    const double rtol = fmax(L.rtol, 100.0 * kEps);
    int status = B200DAE_CIC_NOCONV, iters = 0;

    auto jacobian = [&]() {
        if constexpr (kJac) {
            analytic_jacobian<P, N>(L.t0, y, yp, p, (double*)Jy, (double*)Jyp);
        } else {
            // approx_derivative(method="2-point") wrt z = (y, yp); it evaluates its own f0
            double f0[N];
            P::eval(L.t0, y, yp, p, f0);
            const double rstep = 1.4901161193847656e-08;     // EPS ** 0.5
            for (int which = 0; which < 2; ++which) {
                double* xv = which ? yp : y;
                double* J = which ? Jyp : Jy;
                for (int i = 0; i < N; ++i) {
                    const double x0 = xv[i];
                    const double h = rstep * (x0 >= 0.0 ? 1.0 : -1.0) * fmax(1.0, fabs(x0));
                    const double x1 = x0 + h;
                    const double dx = x1 - x0;
                    xv[i] = x1;
                    P::eval(L.t0, y, yp, p, f1);
                    xv[i] = x0;
                    for (int r = 0; r < N; ++r) J[r * N + i] = (f1[r] - f0[r]) / dx;
                }
            }
        }
    };

    if (nfy + nfyp < N) {
        status = B200DAE_CIC_TOO_MANY_FIXED;
    } else {
        P::eval(L.t0, y, yp, p, f);
        jacobian();
        for (int i = 0; i < N; ++i) scale[i] = L.atol[i] + fabs(f[i]) * rtol;
        for (int nit = 0; nit < L.newton_maxiter && status == B200DAE_CIC_NOCONV; ++nit) {
            for (int ch = 0; ch < L.chord_iter; ++ch) {
                // ---- solve_underdetermined_system (common.py:269-372) ---------------------------
                int rc = 0;
                for (int i = 0; i < N; ++i) dy[i] = dyp[i] = 0.0;
                if (nfy == 0 || nfyp == 0) {
                    const double* J = (nfy == 0) ? Jyp : Jy;
                    for (int k = 0; k < N * N; ++k) A[k] = J[k];
                    qr_colpiv(N, N, A, N, tau, pv, vn1, vn2);
                    const int rankdef = N - qrank_of(N, N, A, N);
                    if (rankdef > 0) {
                        rc = rankdef <= fixed ? B200DAE_CIC_TOO_MANY_FIXED : B200DAE_CIC_INDEX_GT_ONE;
                    } else {
                        for (int i = 0; i < N; ++i) d[i] = -f[i];
                        apply_qt(N, N, A, N, tau, d);
                        back_subst(N, A, N, d);
                        for (int k = 0; k < N; ++k) x[pv[k]] = d[k];
                        for (int k = 0; k < N; ++k) {
                            if (nfy == 0) dyp[free_yp[k]] = x[k]; else dy[free_y[k]] = x[k];
                        }
                    }
                } else {
                    for (int i = 0; i < N; ++i)
                        for (int k = 0; k < nfyp; ++k) A[i * N + k] = Jyp[i * N + free_yp[k]];
                    qr_colpiv(N, nfyp, A, N, tau, pv, vn1, vn2);
                    const int rank = qrank_of(N, nfyp, A, N);
                    const int kq = N < nfyp ? N : nfyp;
                    for (int i = 0; i < N; ++i) d[i] = -f[i];
                    apply_qt(N, kq, A, N, tau, d);
                    if (rank == N) {
                        back_subst(N, A, N, d);
                        for (int k = 0; k < nfyp; ++k) x[k] = 0.0;
                        for (int k = 0; k < N; ++k) x[pv[k]] = d[k];
                        for (int k = 0; k < nfyp; ++k) dyp[free_yp[k]] = x[k];
                    } else {
                        for (int k = 0; k < nfy; ++k) {              // S = Q^T Jy[:, free_y]
                            for (int i = 0; i < N; ++i) c[i] = Jy[i * N + free_y[k]];
                            apply_qt(N, kq, A, N, tau, c);
                            for (int i = 0; i < N; ++i) S[i * N + k] = c[i];
                        }
                        const int m2 = N - rank;
                        for (int i = 0; i < m2; ++i)
                            for (int k = 0; k < nfy; ++k) S2[i * N + k] = S[(rank + i) * N + k];
                        qr_colpiv(m2, nfy, S2, N, tau2, pS, vn1, vn2);
                        const int rankS = qrank_of(m2, nfy, S2, N);
                        const int rankdef = N - (rank + rankS);
                        if (rankdef > 0) {
                            rc = rankdef <= fixed ? B200DAE_CIC_TOO_MANY_FIXED : B200DAE_CIC_INDEX_GT_ONE;
                        } else {
                            for (int i = 0; i < m2; ++i) c[i] = d[rank + i];
                            apply_qt(m2, m2 < nfy ? m2 : nfy, S2, N, tau2, c);
                            back_subst(rankS, S2, N, c);
                            for (int k = 0; k < nfy; ++k) w[k] = 0.0;
                            for (int k = 0; k < rankS; ++k) w[pS[k]] = c[k];
                            for (int k = 0; k < nfyp; ++k) x[k] = 0.0;
                            if (rank > 0) {
                                for (int i = 0; i < rank; ++i) {
                                    double s = d[i];
                                    for (int k = 0; k < nfy; ++k) s = fma(-S[i * N + k], w[k], s);
                                    d[i] = s;
                                }
                                back_subst(rank, A, N, d);
                                for (int k = 0; k < rank; ++k) x[pv[k]] = d[k];
                            }
                            for (int k = 0; k < nfy; ++k) dy[free_y[k]] = w[k];
                            for (int k = 0; k < nfyp; ++k) dyp[free_yp[k]] = x[k];
                        }
                    }
                }
                if (rc != 0) { status = rc; break; }
                // ---- update and test (common.py:246-253) ----------------------------------------
                for (int i = 0; i < N; ++i) { y[i] += dy[i]; yp[i] += dyp[i]; }
                P::eval(L.t0, y, yp, p, f);
                ++iters;
                double ss = 0.0;
                for (int i = 0; i < N; ++i) { const double q = f[i] / scale[i]; ss = fma(q, q, ss); }
                const double error = sqrt(ss) / sqrt((double)N);
                if (error < L.safety) { status = B200DAE_CIC_OK; break; }
            }
            if (status != B200DAE_CIC_NOCONV) break;
            jacobian();
        }
    }
    for (int j = 0; j < N; ++j) {
        L.y0[j * ld + inst] = y[j];
        L.yp0[j * ld + inst] = yp[j];
        if (L.f_out) L.f_out[j * ld + inst] = f[j];
    }
    L.status[inst] = status;
    if (L.n_iter) L.n_iter[inst] = iters;
}

}  // namespace b200dae
