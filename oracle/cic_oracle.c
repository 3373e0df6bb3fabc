/*
 * cic_oracle.c — CPU ORACLE (TEST INFRASTRUCTURE) for consistent initial conditions.
 *
 * Plain-C restatement of consistent_initial_conditions / solve_underdetermined_system / qrank
 * (scipy_dae/integrate/_dae/common.py:186-372) with the third-party pieces it calls:
 *   - scipy.optimize.approx_derivative(method="2-point") on z = (y, yp)
 *     (scipy 1.18.1 optimize/_numdiff.py: _compute_absolute_step :147-195, _dense_difference
 *     :683-712): h_i = sqrt(EPS) * sign(x_i) * max(1, |x_i|) with sign(0) = +1,
 *     dx = (x_i + h_i) - x_i, J[:, i] = (F(x + h_i e_i) - F(x)) / dx;
 *   - scipy.linalg.qr(A, pivoting=True) = LAPACK dgeqp3, which for these sizes runs the
 *     unblocked dlaqp2: column pivoting on partial column norms with the downdate
 *     vn1 *= sqrt(max(0, 1 - (|a_ij| / vn1)^2)) and the recomputation test
 *     temp * (vn1 / vn2)^2 <= sqrt(EPS); Householder vectors from dlarfg
 *     (beta = -sign(alpha) * hypot(alpha, ||x||), tau = (beta - alpha) / beta, v = x / (alpha - beta));
 *   - scipy.linalg.solve_triangular (back substitution), norm = RMS (scipy/integrate/_ivp/common.py:63-65).
 * Q is never formed: Q^T b is the product of the reflectors applied in order.
 */
#include "dae_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

#define EPS 2.220446049250313e-16

/* A is m x n row-major with leading dimension lda.  On return: R in the upper triangle, the
 * Householder vectors below the diagonal (v_k[k] = 1 implicit), tau[min(m,n)], jpvt[n]. */
static void qr_colpiv(int m, int n, double *A, int lda, double *tau, int *jpvt)
{
    double *vn1 = malloc(sizeof(double) * (n > 0 ? n : 1)), *vn2 = malloc(sizeof(double) * (n > 0 ? n : 1));
    const double tol3z = sqrt(EPS);
    for (int j = 0; j < n; ++j) {
        double s = 0.0;
        for (int i = 0; i < m; ++i) s += A[i * lda + j] * A[i * lda + j];
        vn1[j] = vn2[j] = sqrt(s);
        jpvt[j] = j;
    }
    int mn = m < n ? m : n;
    for (int k = 0; k < mn; ++k) {
        int pvt = k;                                           /* idamax: first maximum */
        for (int j = k + 1; j < n; ++j)
            if (vn1[j] > vn1[pvt]) pvt = j;
        if (pvt != k) {
            for (int i = 0; i < m; ++i) {
                double x = A[i * lda + pvt]; A[i * lda + pvt] = A[i * lda + k]; A[i * lda + k] = x;
            }
            int ti = jpvt[pvt]; jpvt[pvt] = jpvt[k]; jpvt[k] = ti;
            vn1[pvt] = vn1[k]; vn2[pvt] = vn2[k];
        }
        /* dlarfg on A[k:m, k] */
        double alpha = A[k * lda + k], xn = 0.0;
        for (int i = k + 1; i < m; ++i) xn += A[i * lda + k] * A[i * lda + k];
        xn = sqrt(xn);
        if (xn == 0.0) {
            tau[k] = 0.0;
        } else {
            double beta = -copysign(hypot(alpha, xn), alpha);
            tau[k] = (beta - alpha) / beta;
            double sc = 1.0 / (alpha - beta);
            for (int i = k + 1; i < m; ++i) A[i * lda + k] *= sc;
            A[k * lda + k] = beta;
        }
        /* apply H(k) to A[k:m, k+1:n] from the left: a_j -= tau * v * (v^T a_j) */
        if (tau[k] != 0.0)
            for (int j = k + 1; j < n; ++j) {
                double w = A[k * lda + j];
                for (int i = k + 1; i < m; ++i) w += A[i * lda + k] * A[i * lda + j];
                w *= tau[k];
                A[k * lda + j] -= w;
                for (int i = k + 1; i < m; ++i) A[i * lda + j] -= A[i * lda + k] * w;
            }
        /* partial column norm downdate (dlaqp2) */
        for (int j = k + 1; j < n; ++j)
            if (vn1[j] != 0.0) {
                double r = fabs(A[k * lda + j]) / vn1[j];
                double temp = 1.0 - r * r;
                if (temp < 0.0) temp = 0.0;
                double q = vn1[j] / vn2[j];
                double temp2 = temp * q * q;
                if (temp2 <= tol3z) {
                    double s = 0.0;
                    for (int i = k + 1; i < m; ++i) s += A[i * lda + j] * A[i * lda + j];
                    vn1[j] = vn2[j] = sqrt(s);
                } else {
                    vn1[j] *= sqrt(temp);
                }
            }
    }
    free(vn1); free(vn2);
}

/* b <- Q^T b for the reflectors stored in A (m x n, k reflectors) */
static void apply_qt(int m, int k, const double *A, int lda, const double *tau, double *b)
{
    for (int r = 0; r < k; ++r) {
        if (tau[r] == 0.0) continue;
        double w = b[r];
        for (int i = r + 1; i < m; ++i) w += A[i * lda + r] * b[i];
        w *= tau[r];
        b[r] -= w;
        for (int i = r + 1; i < m; ++i) b[i] -= A[i * lda + r] * w;
    }
}

/* qrank (common.py:261-266): rank = #{ |R_kk| > max(shape) * EPS * |R_00| } */
static int qrank_of(int m, int n, const double *A, int lda)
{
    int mn = m < n ? m : n;
    if (mn == 0) return 0;
    double tol = (m > n ? m : n) * EPS * fabs(A[0]);
    int rank = 0;
    for (int k = 0; k < mn; ++k) rank += fabs(A[k * lda + k]) > tol;
    return rank;
}

/* x[0:r] <- R[0:r,0:r]^{-1} x[0:r] (upper triangular, row-major) */
static void back_subst(int r, const double *R, int lda, double *x)
{
    for (int i = r - 1; i >= 0; --i) {
        double s = x[i];
        for (int j = i + 1; j < r; ++j) s -= R[i * lda + j] * x[j];
        x[i] = s / R[i * lda + i];
    }
}

/* status codes of orc_consistent_ic */
enum { CIC_OK = 0, CIC_NOCONV = -1, CIC_TOO_MANY_FIXED = -2, CIC_INDEX_GT_ONE = -3 };

/* solve_underdetermined_system (common.py:269-372).  free_y[nfy], free_yp[nfyp] are index lists. */
static int solve_underdetermined(int n, const double *f, const double *Jy, const double *Jyp,
                                 const int *free_y, int nfy, const int *free_yp, int nfyp,
                                 double *dy, double *dyp)
{
    for (int i = 0; i < n; ++i) dy[i] = dyp[i] = 0.0;
    int fixed = (n - nfy) + (n - nfyp);
    double *A = malloc(sizeof(double) * n * n), *S = malloc(sizeof(double) * n * n);
    double *tau = malloc(sizeof(double) * n), *tau2 = malloc(sizeof(double) * n);
    double *d = malloc(sizeof(double) * n), *w = malloc(sizeof(double) * n), *x = malloc(sizeof(double) * n);
    int *p = malloc(sizeof(int) * n), *pS = malloc(sizeof(int) * n);
    int status = CIC_OK;

    if (nfy == 0 || nfyp == 0) {
        /* ODE case (solve for dyp) or purely algebraic case (solve for dy): square n x n systems */
        const double *J = (nfy == 0) ? Jyp : Jy;
        memcpy(A, J, sizeof(double) * n * n);
        qr_colpiv(n, n, A, n, tau, p);
        int rank = qrank_of(n, n, A, n);
        int rankdef = n - rank;
        if (rankdef > 0) { status = rankdef <= fixed ? CIC_TOO_MANY_FIXED : CIC_INDEX_GT_ONE; goto done; }
        for (int i = 0; i < n; ++i) d[i] = -f[i];
        apply_qt(n, n, A, n, tau, d);
        back_subst(n, A, n, d);
        /* Delta_[p] = solve(R, d); Delta[free] = Delta_  (free == arange(n) here, common.py:289,304) */
        double *out = (nfy == 0) ? dyp : dy;
        const int *fr = (nfy == 0) ? free_yp : free_y;
        for (int k = 0; k < n; ++k) x[p[k]] = d[k];
        for (int k = 0; k < n; ++k) out[fr[k]] = x[k];
        goto done;
    }
    {
        /* eliminate the fixed variables: Jy[:, free_y] (n x nfy), Jyp[:, free_yp] (n x nfyp) */
        for (int i = 0; i < n; ++i)
            for (int k = 0; k < nfyp; ++k) A[i * n + k] = Jyp[i * n + free_yp[k]];
        qr_colpiv(n, nfyp, A, n, tau, p);
        int rank = qrank_of(n, nfyp, A, n);
        int kq = n < nfyp ? n : nfyp;
        for (int i = 0; i < n; ++i) d[i] = -f[i];
        apply_qt(n, kq, A, n, tau, d);
        if (rank == n) {
            /* full rank: dy = 0, R wp_ = d (R is n x nfyp upper trapezoidal; scipy solves the
             * square system, so nfyp == n here) */
            back_subst(n, A, n, d);
            for (int k = 0; k < nfyp; ++k) x[k] = 0.0;
            for (int k = 0; k < n; ++k) x[p[k]] = d[k];
            for (int k = 0; k < nfyp; ++k) dyp[free_yp[k]] = x[k];
        } else {
            /* S = Q^T Jy[:, free_y]  (n x nfy) */
            double *col = malloc(sizeof(double) * n);
            for (int k = 0; k < nfy; ++k) {
                for (int i = 0; i < n; ++i) col[i] = Jy[i * n + free_y[k]];
                apply_qt(n, kq, A, n, tau, col);
                for (int i = 0; i < n; ++i) S[i * n + k] = col[i];
            }
            free(col);
            int m2 = n - rank;
            /* QR with pivoting of S[rank:, :] in a copy (S[:rank] is still needed) */
            double *S2 = malloc(sizeof(double) * (m2 > 0 ? m2 : 1) * n);
            for (int i = 0; i < m2; ++i)
                for (int k = 0; k < nfy; ++k) S2[i * n + k] = S[(rank + i) * n + k];
            qr_colpiv(m2, nfy, S2, n, tau2, pS);
            int rankS = qrank_of(m2, nfy, S2, n);
            int rankdef = n - (rank + rankS);
            if (rankdef > 0) {
                status = rankdef <= fixed ? CIC_TOO_MANY_FIXED : CIC_INDEX_GT_ONE;
                free(S2);
                goto done;
            }
            /* c = QS^T d2; RS11 v1 = c1; w[pS] = v */
            double *c = malloc(sizeof(double) * (m2 > 0 ? m2 : 1));
            for (int i = 0; i < m2; ++i) c[i] = d[rank + i];
            apply_qt(m2, m2 < nfy ? m2 : nfy, S2, n, tau2, c);
            back_subst(rankS, S2, n, c);
            for (int k = 0; k < nfy; ++k) w[k] = 0.0;
            for (int k = 0; k < rankS; ++k) w[pS[k]] = c[k];
            free(c); free(S2);
            /* R11 wp_ = d1 - S[:rank] w; wp[p] = wp_ */
            for (int k = 0; k < nfyp; ++k) x[k] = 0.0;
            if (rank > 0) {
                for (int i = 0; i < rank; ++i) {
                    double s = d[i];
                    for (int k = 0; k < nfy; ++k) s -= S[i * n + k] * w[k];
                    d[i] = s;
                }
                back_subst(rank, A, n, d);
                for (int k = 0; k < rank; ++k) x[p[k]] = d[k];
            }
            for (int k = 0; k < nfy; ++k) dy[free_y[k]] = w[k];
            for (int k = 0; k < nfyp; ++k) dyp[free_yp[k]] = x[k];
        }
    }
done:
    free(A); free(S); free(tau); free(tau2); free(d); free(w); free(x); free(p); free(pS);
    return status;
}

/* approx_derivative(method="2-point") of the composite residual wrt z = (y, yp) */
static void fd_composite(orc_residual_fn fn, int n, const double *p, double t, const double *y,
                         const double *yp, double *Jy, double *Jyp)
{
    double *f0 = malloc(sizeof(double) * n), *f1 = malloc(sizeof(double) * n);
    double *yy = malloc(sizeof(double) * n), *yyp = malloc(sizeof(double) * n);
    memcpy(yy, y, sizeof(double) * n); memcpy(yyp, yp, sizeof(double) * n);
    fn(t, y, yp, p, f0);
    const double rstep = sqrt(EPS);
    for (int which = 0; which < 2; ++which) {
        double *x = which ? yyp : yy;
        double *J = which ? Jyp : Jy;
        for (int i = 0; i < n; ++i) {
            double x0 = x[i];
            double h = rstep * (x0 >= 0 ? 1.0 : -1.0) * fmax(1.0, fabs(x0));
            double x1 = x0 + h;
            double dx = x1 - x0;
            x[i] = x1;
            fn(t, yy, yyp, p, f1);
            x[i] = x0;
            for (int r = 0; r < n; ++r) J[r * n + i] = (f1[r] - f0[r]) / dx;
        }
    }
    free(f0); free(f1); free(yy); free(yyp);
}

/*
 * consistent_initial_conditions (common.py:186-258) for one instance.  y0 / yp0 are updated in
 * place; f_out[n] receives the final residual.  fixed_y / fixed_yp: 0/1 masks of length n.
 * jac_fn == NULL -> finite differences (jac=None).  Returns CIC_*.
 */
int orc_consistent_ic(orc_residual_fn fn, orc_jacobian_fn jac_fn, int n, const double *p, double t0,
                      double *y0, double *yp0, const int *fixed_y, const int *fixed_yp,
                      double rtol, const double *atol, int newton_maxiter, int chord_iter,
                      double safety, double *f_out, int *n_iter_out)
{
    int *free_y = malloc(sizeof(int) * n), *free_yp = malloc(sizeof(int) * n);
    int nfy = 0, nfyp = 0;
    for (int i = 0; i < n; ++i) {
        if (!fixed_y || !fixed_y[i]) free_y[nfy++] = i;
        if (!fixed_yp || !fixed_yp[i]) free_yp[nfyp++] = i;
    }
    if (rtol < 100 * EPS) rtol = 100 * EPS;
    double *f = malloc(sizeof(double) * n), *Jy = malloc(sizeof(double) * n * n), *Jyp = malloc(sizeof(double) * n * n);
    double *scale = malloc(sizeof(double) * n), *dy = malloc(sizeof(double) * n), *dyp = malloc(sizeof(double) * n);
    int status = CIC_NOCONV, iters = 0;
    if (nfy + nfyp < n) { status = CIC_TOO_MANY_FIXED; goto out; }
    fn(t0, y0, yp0, p, f);
    if (jac_fn) jac_fn(t0, y0, yp0, p, Jy, Jyp); else fd_composite(fn, n, p, t0, y0, yp0, Jy, Jyp);
    for (int i = 0; i < n; ++i) scale[i] = atol[i] + fabs(f[i]) * rtol;
    for (int nit = 0; nit < newton_maxiter && status == CIC_NOCONV; ++nit) {
        for (int c = 0; c < chord_iter; ++c) {
            int rc = solve_underdetermined(n, f, Jy, Jyp, free_y, nfy, free_yp, nfyp, dy, dyp);
            if (rc != CIC_OK) { status = rc; goto out; }
            for (int i = 0; i < n; ++i) { y0[i] += dy[i]; yp0[i] += dyp[i]; }
            fn(t0, y0, yp0, p, f);
            ++iters;
            double ss = 0.0;
            for (int i = 0; i < n; ++i) { double q = f[i] / scale[i]; ss += q * q; }
            double error = sqrt(ss) / sqrt((double)n);
            if (error < safety) { status = CIC_OK; break; }
        }
        if (status == CIC_OK) break;
        if (jac_fn) jac_fn(t0, y0, yp0, p, Jy, Jyp); else fd_composite(fn, n, p, t0, y0, yp0, Jy, Jyp);
    }
out:
    if (f_out) memcpy(f_out, f, sizeof(double) * n);
    if (n_iter_out) *n_iter_out = iters;
    free(free_y); free(free_yp); free(f); free(Jy); free(Jyp); free(scale); free(dy); free(dyp);
    return status;
}

/* ensemble driver (AoS): y0[N][n], yp0[N][n] in place, f_out[N][n], status[N], n_iter[N] */
int orc_consistent_ic_ensemble(int problem_id, int use_analytic_jac, int64_t N, const double *params,
                               double t0, double *y0, double *yp0, const int *fixed_y, const int *fixed_yp,
                               double rtol, const double *atol, int newton_maxiter, int chord_iter,
                               double safety, double *f_out, int32_t *status, int32_t *n_iter)
{
    int n, np;
    if (orc_problem_dims(problem_id, &n, &np)) return -1;
    orc_residual_fn fn = orc_problem_residual(problem_id);
    orc_jacobian_fn jf = use_analytic_jac ? orc_problem_jacobian(problem_id) : 0;
    if (use_analytic_jac && !jf) return -2;
    for (int64_t i = 0; i < N; ++i) {
        int it = 0;
        status[i] = orc_consistent_ic(fn, jf, n, params + i * np, t0, y0 + i * n, yp0 + i * n, fixed_y, fixed_yp,
                                      rtol, atol, newton_maxiter, chord_iter, safety, f_out + i * n, &it);
        if (n_iter) n_iter[i] = it;
    }
    return 0;
}
