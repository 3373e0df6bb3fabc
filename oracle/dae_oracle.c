/*
 * dae_oracle.c — CPU ORACLE (TEST INFRASTRUCTURE, NOT PRODUCT CODE).  See dae_oracle.h.
 *
 * Scalar, runtime-sized (n, s) restatement of the reference's hot path.  Citations are
 * to /root/reference/scipy_dae/integrate/_dae/<file>:<line> (short: radau.py, bdf.py,
 * base.py, common.py, dae.py) and to scipy 1.18.1's scipy/integrate/_ivp/common.py
 * (short: SP/common.py) for num_jac / norm, which the reference imports.
 *
 * Pinned by tests/test_oracle_golden.py against tests/golden/ (.npz) (minted from the
 * unmodified Python reference by oracle/gen_golden.py).
 */
#include "dae_oracle.h"

#include <float.h>
#include <math.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>

#define EPS DBL_EPSILON

/* ------------------------------------------------------------------------- */
/* small helpers                                                             */
/* ------------------------------------------------------------------------- */

/* SP/common.py:63-65  norm(x) = ||x||_2 / sqrt(size) */
static double rms_norm(const double *x, int len)
{
    double ss = 0.0;
    for (int i = 0; i < len; ++i) ss += x[i] * x[i];
    return sqrt(ss) / sqrt((double)len);
}

static int all_finite(const double *x, int len)
{
    for (int i = 0; i < len; ++i)
        if (!isfinite(x[i])) return 0;
    return 1;
}

/* ------------------------------------------------------------------------- */
/* LAPACK-style LU with row partial pivoting (base.py:232-238 -> d/zgetrf, d/zgetrs) */
/* ------------------------------------------------------------------------- */

/* real: a row-major n*n, overwritten by L\U; piv[k] = row swapped with k.  Returns 0,
 * or k+1 of the first exactly-zero pivot (LAPACK info; factorisation still completes). */
int orc_lu_factor(int n, double *a, int *piv)
{
    int info = 0;
    for (int k = 0; k < n; ++k) {
        int p = k;
        double best = fabs(a[k * n + k]);
        for (int i = k + 1; i < n; ++i) {          /* idamax: first maximal element */
            double v = fabs(a[i * n + k]);
            if (v > best) { best = v; p = i; }
        }
        piv[k] = p;
        if (p != k)
            for (int j = 0; j < n; ++j) {
                double tmp = a[k * n + j]; a[k * n + j] = a[p * n + j]; a[p * n + j] = tmp;
            }
        double pv = a[k * n + k];
        if (pv != 0.0) {
            double r = 1.0 / pv;                   /* dgetf2 scales by the reciprocal */
            for (int i = k + 1; i < n; ++i) a[i * n + k] *= r;
        } else if (!info) {
            info = k + 1;
        }
        for (int i = k + 1; i < n; ++i) {
            double l = a[i * n + k];
            for (int j = k + 1; j < n; ++j) a[i * n + j] -= l * a[k * n + j];
        }
    }
    return info;
}

void orc_lu_solve(int n, const double *lu, const int *piv, double *b)
{
    for (int k = 0; k < n; ++k) {
        int p = piv[k];
        if (p != k) { double tmp = b[k]; b[k] = b[p]; b[p] = tmp; }
    }
    for (int i = 1; i < n; ++i) {
        double acc = b[i];
        for (int j = 0; j < i; ++j) acc -= lu[i * n + j] * b[j];
        b[i] = acc;
    }
    for (int i = n - 1; i >= 0; --i) {
        double acc = b[i];
        for (int j = i + 1; j < n; ++j) acc -= lu[i * n + j] * b[j];
        b[i] = acc / lu[i * n + i];
    }
}

/* complex: split storage ar/ai row-major; pivot on |re|+|im| (izamax / dcabs1) */
static int zlu_factor(int n, double *ar, double *ai, int *piv)
{
    int info = 0;
    for (int k = 0; k < n; ++k) {
        int p = k;
        double best = fabs(ar[k * n + k]) + fabs(ai[k * n + k]);
        for (int i = k + 1; i < n; ++i) {
            double v = fabs(ar[i * n + k]) + fabs(ai[i * n + k]);
            if (v > best) { best = v; p = i; }
        }
        piv[k] = p;
        if (p != k)
            for (int j = 0; j < n; ++j) {
                double tmp = ar[k * n + j]; ar[k * n + j] = ar[p * n + j]; ar[p * n + j] = tmp;
                tmp = ai[k * n + j]; ai[k * n + j] = ai[p * n + j]; ai[p * n + j] = tmp;
            }
        double pr = ar[k * n + k], pi = ai[k * n + k];
        if (pr != 0.0 || pi != 0.0) {
            /* reciprocal of the pivot (Smith), then scale the column */
            double rr, ri;
            if (fabs(pr) >= fabs(pi)) {
                double q = pi / pr, d = pr + pi * q;
                rr = 1.0 / d; ri = -q / d;
            } else {
                double q = pr / pi, d = pr * q + pi;
                rr = q / d; ri = -1.0 / d;
            }
            for (int i = k + 1; i < n; ++i) {
                double xr = ar[i * n + k], xi = ai[i * n + k];
                ar[i * n + k] = xr * rr - xi * ri;
                ai[i * n + k] = xr * ri + xi * rr;
            }
        } else if (!info) {
            info = k + 1;
        }
        for (int i = k + 1; i < n; ++i) {
            double lr = ar[i * n + k], li = ai[i * n + k];
            for (int j = k + 1; j < n; ++j) {
                double ur = ar[k * n + j], ui = ai[k * n + j];
                ar[i * n + j] -= lr * ur - li * ui;
                ai[i * n + j] -= lr * ui + li * ur;
            }
        }
    }
    return info;
}

static void zlu_solve(int n, const double *ar, const double *ai, const int *piv,
                      double *br, double *bi)
{
    for (int k = 0; k < n; ++k) {
        int p = piv[k];
        if (p != k) {
            double tmp = br[k]; br[k] = br[p]; br[p] = tmp;
            tmp = bi[k]; bi[k] = bi[p]; bi[p] = tmp;
        }
    }
    for (int i = 1; i < n; ++i) {
        double sr = br[i], si = bi[i];
        for (int j = 0; j < i; ++j) {
            double lr = ar[i * n + j], li = ai[i * n + j];
            sr -= lr * br[j] - li * bi[j];
            si -= lr * bi[j] + li * br[j];
        }
        br[i] = sr; bi[i] = si;
    }
    for (int i = n - 1; i >= 0; --i) {
        double sr = br[i], si = bi[i];
        for (int j = i + 1; j < n; ++j) {
            double ur = ar[i * n + j], ui = ai[i * n + j];
            sr -= ur * br[j] - ui * bi[j];
            si -= ur * bi[j] + ui * br[j];
        }
        double dr = ar[i * n + i], di = ai[i * n + i];
        if (fabs(dr) >= fabs(di)) {
            double q = di / dr, d = dr + di * q;
            br[i] = (sr + si * q) / d; bi[i] = (si - sr * q) / d;
        } else {
            double q = dr / di, d = dr * q + di;
            br[i] = (sr * q + si) / d; bi[i] = (si * q - sr) / d;
        }
    }
}

/* ------------------------------------------------------------------------- */
/* select_initial_step (common.py:129-183)                                   */
/* ------------------------------------------------------------------------- */
double orc_select_initial_step(int n, double t0, const double *y0, const double *yp0,
                               double t_bound, double rtol, const double *atol, double max_step)
{
    const double safety = 0.8;
    double min_step = 16 * EPS * fabs(t0);
    double hspan = fabs(t_bound - t0);
    double e = 0.0;
    for (int i = 0; i < n; ++i) {
        double wt = fmax(fabs(y0[i]), atol[i] / rtol);
        double v = fabs(yp0[i] / wt);
        if (v > e || isnan(v)) e = v;
    }
    double rh = e / sqrt(rtol) / safety;
    double h_abs = fmin(max_step, hspan);
    if (h_abs * rh > 1) h_abs = 1 / rh;
    return fmax(h_abs, min_step);
}

/* ------------------------------------------------------------------------- */
/* finite-difference Jacobian: scipy num_jac / _dense_num_jac (SP/common.py:268-382), */
/* called from jac_wrapped (base.py:264-295) with threshold = atol / rtol             */
/* ------------------------------------------------------------------------- */
#define NJ_DIFF_REJECT pow(EPS, 0.875)
#define NJ_DIFF_SMALL pow(EPS, 0.75)
#define NJ_DIFF_BIG pow(EPS, 0.25)
#define NJ_MIN_FACTOR (1e3 * EPS)

/* one column: residual at x + h e_i (x = y or yp), diff = f_new - f; returns max_diff, scale */
static void nj_column(orc_residual_fn fn, int n, const double *p, double t, double *y,
                      double *yp, int wrt_yp, int col, double h, const double *f,
                      double *f_new, double *diff, double *max_diff, double *scale)
{
    double *x = wrt_yp ? yp : y;
    double saved = x[col];
    x[col] = saved + h;
    fn(t, y, yp, p, f_new);
    x[col] = saved;
    int max_ind = 0;
    double best = -1.0;
    for (int r = 0; r < n; ++r) {
        diff[r] = f_new[r] - f[r];
        double a = fabs(diff[r]);
        if (a > best || (isnan(a) && !isnan(best))) { best = a; max_ind = r; }  /* np.argmax: first max, NaN wins */
    }
    *max_diff = fabs(diff[max_ind]);
    *scale = fmax(fabs(f[max_ind]), fabs(f_new[max_ind]));
    if (isnan(f[max_ind]) || isnan(f_new[max_ind])) *scale = NAN;  /* np.maximum propagates NaN */
}

void orc_num_jac(orc_residual_fn fn, int n, const double *p, double t, const double *y_in,
                 const double *yp_in, const double *f, int wrt_yp, double rtol, const double *atol,
                 double *factor, double *J)
{
    double *y = malloc(sizeof(double) * n), *yp = malloc(sizeof(double) * n);
    double *f_new = malloc(sizeof(double) * n), *diff = malloc(sizeof(double) * n);
    double *diff2 = malloc(sizeof(double) * n);
    memcpy(y, y_in, sizeof(double) * n);
    memcpy(yp, yp_in, sizeof(double) * n);
    const double *x = wrt_yp ? yp : y;
    for (int i = 0; i < n; ++i) {
        double thr = atol[i] / rtol;                       /* base.py:277 */
        double f_sign = (f[i] >= 0) ? 1.0 : -1.0;          /* SP/common.py:323 (indexed like y) */
        double y_scale = f_sign * fmax(thr, fabs(x[i]));
        double h = (x[i] + factor[i] * y_scale) - x[i];
        while (h == 0) {                                   /* SP/common.py:329-332 */
            factor[i] *= 10;
            h = (x[i] + factor[i] * y_scale) - x[i];
        }
        double max_diff, scale;
        nj_column(fn, n, p, t, y, yp, wrt_yp, i, h, f, f_new, diff, &max_diff, &scale);
        if (max_diff < NJ_DIFF_REJECT * scale) {           /* SP/common.py:350-371 */
            double new_factor = 10 * factor[i];
            double h_new = (x[i] + new_factor * y_scale) - x[i];
            double max_diff_new, scale_new;
            nj_column(fn, n, p, t, y, yp, wrt_yp, i, h_new, f, f_new, diff2, &max_diff_new, &scale_new);
            if (max_diff * scale_new < max_diff_new * scale) {
                factor[i] = new_factor;
                h = h_new;
                memcpy(diff, diff2, sizeof(double) * n);
                scale = scale_new;
                max_diff = max_diff_new;
            }
        }
        for (int r = 0; r < n; ++r) J[r * n + i] = diff[r] / h;   /* SP/common.py:373 */
        if (max_diff < NJ_DIFF_SMALL * scale) factor[i] *= 10;   /* SP/common.py:375-377 */
        if (max_diff > NJ_DIFF_BIG * scale) factor[i] *= 0.1;
        factor[i] = fmax(factor[i], NJ_MIN_FACTOR);
    }
    free(y); free(yp); free(f_new); free(diff); free(diff2);
}

/* ------------------------------------------------------------------------- */
/* shared per-instance solver state (base.py:160-242)                        */
/* ------------------------------------------------------------------------- */
typedef struct {
    orc_residual_fn fn;
    orc_jacobian_fn jacfn;   /* NULL: finite differences (jac=None) */
    int n;
    const double *p;
    double rtol;
    const double *atol;
    double t, t_bound, dir, max_step;
    double *y, *yp;          /* current state */
    double *Jy, *Jyp;        /* row-major n*n */
    double *factor_y, *factor_yp;
    double *ftmp;
    int64_t *cnt;
} base_state;

static void count_fun(base_state *b, double t, const double *y, const double *yp, double *out)
{
    b->cnt[ORC_CNT_NFEV] += 1;                              /* base.py:206-208 */
    b->fn(t, y, yp, b->p, out);
}

/* jac_wrapped, base.py:264-295 (FD residual calls are not counted in nfev) */
static int g_analytic_jac = 0;
void orc_set_analytic_jac(int on) { g_analytic_jac = on; }

static void fd_jacobian(base_state *b, double t, const double *y, const double *yp)
{
    b->cnt[ORC_CNT_NJEV] += 1;
    if (b->jacfn) {                                         /* callable jac, base.py:299-331 */
        b->jacfn(t, y, yp, b->p, b->Jy, b->Jyp);
        return;
    }
    b->fn(t, y, yp, b->p, b->ftmp);
    orc_num_jac(b->fn, b->n, b->p, t, y, yp, b->ftmp, 0, b->rtol, b->atol, b->factor_y, b->Jy);
    orc_num_jac(b->fn, b->n, b->p, t, y, yp, b->ftmp, 1, b->rtol, b->atol, b->factor_yp, b->Jyp);
}

static double min_step_at(double t, double dir)
{
    /* radau.py:489, bdf.py:270 */
    return 10 * fabs(nextafter(t, dir * INFINITY) - t);
}

/* ------------------------------------------------------------------------- */
/* Radau IIA                                                                 */
/* ------------------------------------------------------------------------- */
typedef struct {
    base_state b;
    const orc_radau_tables *tb;
    const orc_radau_opts *o;
    int s, ncs, newton_max_iter;
    double newton_tol;
    double h_abs, h_abs_old, error_norm_old;
    int have_old;                       /* h_abs_old / error_norm_old are not None */
    int current_jac, lu_valid, have_sol;
    double *LUr; int *pivr;             /* n*n */
    double *LUcr, *LUci; int *pivc;     /* ncs * n*n */
    double *Y, *Yp, *F, *dW, *dYp;      /* s*n */
    double *scale, *err, *tmpn, *tmpn2, *yhat;
    /* dense output of the last step (radau.py:661-699) */
    double t_old, h_old, *y_old, *ZP;   /* ZP: n*s */
    int status;
} radau_state;

/* RadauDenseOutput._call_impl, radau.py:681-699 */
static void radau_dense_eval(const radau_state *st, double t, double *y, double *yp)
{
    int n = st->b.n, s = st->s;
    double x = (t - st->t_old) / st->h_old;
    for (int i = 0; i < n; ++i) {
        double ay = 0.0, ayp = 0.0;
        for (int k = 1; k <= s; ++k) {
            double zp = st->ZP[i * s + (k - 1)];
            ay += zp * pow(x, k);
            ayp += zp * ((k / st->h_old) * pow(x, k - 1));
        }
        if (y) y[i] = ay + st->y_old[i];
        if (yp) yp[i] = ayp;
    }
}

/* solve_collocation_system, radau.py:137-195.  Yp is updated in place (quirk A.3-1). */
static int radau_newton(radau_state *st, double t, const double *y, double h,
                        int *n_iter_out, double *rate_out, int *have_rate)
{
    base_state *b = &st->b;
    const orc_radau_tables *tb = st->tb;
    int n = b->n, s = st->s, ncs = st->ncs;
    double *Y = st->Y, *Yp = st->Yp, *F = st->F, *dW = st->dW, *dYp = st->dYp;

    for (int i = 0; i < s; ++i)                              /* radau.py:145 */
        for (int j = 0; j < n; ++j) {
            double acc = 0.0;
            for (int l = 0; l < s; ++l) acc += tb->A[i * s + l] * Yp[l * n + j];
            Y[i * n + j] = y[j] + h * acc;
        }

    double dY_norm_old = 0.0, rate = 0.0;
    int has_old = 0, has_rate = 0, converged = 0, k;
    for (k = 0; k < st->newton_max_iter; ++k) {
        for (int i = 0; i < s; ++i)
            count_fun(b, t + h * tb->C[i], Y + i * n, Yp + i * n, F + i * n);
        if (!all_finite(F, s * n)) break;                    /* radau.py:157-158 */

        /* G = TI @ F; real block and complex blocks, radau.py:160-174 */
        for (int j = 0; j < n; ++j) {
            double acc = 0.0;
            for (int l = 0; l < s; ++l) acc += tb->TI[0 * s + l] * F[l * n + j];
            dW[j] = -acc;
        }
        orc_lu_solve(n, st->LUr, st->pivr, dW);
        b->cnt[ORC_CNT_NSOLVE] += 1;
        for (int c = 0; c < ncs; ++c) {
            double *gr = dW + (2 * c + 1) * n, *gi = dW + (2 * c + 2) * n;
            for (int j = 0; j < n; ++j) {
                double a1 = 0.0, a2 = 0.0;
                for (int l = 0; l < s; ++l) {
                    a1 += tb->TI[(2 * c + 1) * s + l] * F[l * n + j];
                    a2 += tb->TI[(2 * c + 2) * s + l] * F[l * n + j];
                }
                gr[j] = -a1; gi[j] = -a2;
            }
            zlu_solve(n, st->LUcr + c * n * n, st->LUci + c * n * n, st->pivc + c * n, gr, gi);
            b->cnt[ORC_CNT_NSOLVE] += 1;
        }
        /* dYp = T @ dW ; dY = h * A @ dYp ; radau.py:176-180 */
        for (int i = 0; i < s; ++i)
            for (int j = 0; j < n; ++j) {
                double acc = 0.0;
                for (int l = 0; l < s; ++l) acc += tb->T[i * s + l] * dW[l * n + j];
                dYp[i * n + j] = acc;
            }
        double ss = 0.0;
        for (int i = 0; i < s; ++i)
            for (int j = 0; j < n; ++j) {
                double acc = 0.0;
                for (int l = 0; l < s; ++l) acc += tb->A[i * s + l] * dYp[l * n + j];
                double dY = h * acc;
                Y[i * n + j] += dY;
                double q = dY / st->scale[j];
                ss += q * q;
            }
        for (int i = 0; i < s * n; ++i) Yp[i] += dYp[i];
        double dY_norm = sqrt(ss) / sqrt((double)(s * n));   /* radau.py:182 */
        if (has_old) { rate = dY_norm / dY_norm_old; has_rate = 1; }

        if (has_rate && (rate >= 1 ||
                         pow(rate, st->newton_max_iter - k) / (1 - rate) * dY_norm > st->newton_tol))
            break;                                           /* radau.py:186-187 */
        if (dY_norm == 0 || (has_rate && rate / (1 - rate) * dY_norm < st->newton_tol)) {
            converged = 1;                                   /* radau.py:189-191 */
            break;
        }
        dY_norm_old = dY_norm;
        has_old = 1;
    }
    if (k == st->newton_max_iter) k = st->newton_max_iter - 1;  /* python's k after exhausting range */
    *n_iter_out = k + 1;
    *rate_out = rate;
    *have_rate = has_rate;
    return converged;
}

/* predict_factor, radau.py:198-240 */
static double predict_factor(double h_abs, double h_abs_old, double error_norm,
                             double error_norm_old, int have_old, int s)
{
    double multiplier;
    if (!have_old || error_norm == 0)
        multiplier = 1;
    else
        multiplier = h_abs / h_abs_old * pow(error_norm_old / error_norm, 1.0 / (s + 1));
    double factor = fmin(1, multiplier) * pow(error_norm, -1.0 / (s + 1));
    return 1 + atan(factor - 1);                             /* KAPPA = 1 */
}

static int radau_factor_matrices(radau_state *st, double h)
{
    /* radau.py:536-538: LU_real = lu(Jyp + h*MU_REAL*Jy); LU_complex = lu(Jyp + h*MU*Jy) */
    base_state *b = &st->b;
    int n = b->n;
    double cr = h * st->tb->mu_real;
    for (int i = 0; i < n * n; ++i) st->LUr[i] = b->Jyp[i] + cr * b->Jy[i];
    if (!all_finite(st->LUr, n * n)) return -2;             /* lu_factor(check_finite=True) raises */
    orc_lu_factor(n, st->LUr, st->pivr);
    b->cnt[ORC_CNT_NLU] += 1;
    for (int c = 0; c < st->ncs; ++c) {
        double mr = h * st->tb->mu_re[c], mi = h * st->tb->mu_im[c];
        double *ar = st->LUcr + c * n * n, *ai = st->LUci + c * n * n;
        for (int i = 0; i < n * n; ++i) {
            ar[i] = b->Jyp[i] + mr * b->Jy[i];
            ai[i] = mi * b->Jy[i];
        }
        if (!all_finite(ar, n * n) || !all_finite(ai, n * n)) return -2;
        zlu_factor(n, ar, ai, st->pivc + c * n);
        b->cnt[ORC_CNT_NLU] += 1;
    }
    st->lu_valid = 1;
    return 0;
}

/* RadauDAE._step_impl, radau.py:465-659.  Returns 1 on an accepted step, 0 on failure
 * (st->status = -1 too small step, -2 non-finite matrix). */
static int radau_step(radau_state *st)
{
    base_state *b = &st->b;
    const orc_radau_tables *tb = st->tb;
    const orc_radau_opts *o = st->o;
    int n = b->n, s = st->s;
    double t = b->t;
    double *y = b->y, *yp = b->yp;

    double min_step = min_step_at(t, b->dir);
    double h_abs, h_abs_old = st->h_abs_old, error_norm_old = st->error_norm_old;
    int have_old = st->have_old;
    if (st->h_abs > b->max_step) { h_abs = b->max_step; have_old = 0; }       /* radau.py:490-493 */
    else if (st->h_abs < min_step) { h_abs = min_step; have_old = 0; }        /* radau.py:494-497 */
    else h_abs = st->h_abs;

    int have_factor = 0, n_iter = 0, have_rate = 0;
    double factor = 0.0, rate = 0.0, error_norm = 0.0, safety = 0.0, h = 0.0, t_new = t;
    for (;;) {
        if (h_abs < min_step) { st->status = -1; return 0; }                 /* radau.py:515-516 */
        h = h_abs * b->dir;
        t_new = t + h;
        if (b->dir * (t_new - b->t_bound) > 0) t_new = b->t_bound;
        h = t_new - t;
        h_abs = fabs(h);

        /* initial guess, radau.py:527-531 */
        for (int i = 0; i < s; ++i) {
            if (!st->have_sol) memcpy(st->Yp + i * n, yp, sizeof(double) * n);
            else radau_dense_eval(st, t + h * tb->C[i], NULL, st->Yp + i * n);
        }
        for (int j = 0; j < n; ++j) st->scale[j] = b->atol[j] + fabs(y[j]) * b->rtol;

        int converged = 0;
        for (;;) {
            if (!st->lu_valid) {
                int rc = radau_factor_matrices(st, h);
                if (rc) { st->status = rc; return 0; }
            }
            converged = radau_newton(st, t, y, h, &n_iter, &rate, &have_rate);
            b->cnt[ORC_CNT_NEWTON_ITER] += n_iter;
            if (converged) break;
            b->cnt[ORC_CNT_NEWTON_FAIL] += 1;
            if (st->current_jac) break;
            fd_jacobian(b, t, y, yp);                                        /* radau.py:549-552 */
            st->current_jac = 1;
            st->lu_valid = 0;
        }
        if (!converged) {                                                    /* radau.py:554-558 */
            h_abs *= 0.5;
            st->lu_valid = 0;
            continue;
        }

        const double *y_new = st->Y + (s - 1) * n;
        for (int j = 0; j < n; ++j)
            st->scale[j] = b->atol[j] + fmax(fabs(y[j]), fabs(y_new[j])) * b->rtol;

        /* embedded error, radau.py:568-587 */
        double *err = st->err;
        if (o->newton_iter_embedded == 0) {
            for (int j = 0; j < n; ++j) {
                double acc = 0.0;
                for (int l = 0; l < s; ++l) acc += tb->v_explicit[l] * st->Yp[l * n + j];
                err[j] = h * (tb->mu_real * yp[j] + acc);
            }
        } else if (o->newton_iter_embedded == 1) {
            for (int j = 0; j < n; ++j) {
                double acc = 0.0;
                for (int l = 0; l < s; ++l) acc += tb->v_implicit[l] * st->Yp[l * n + j];
                st->tmpn[j] = (acc - tb->b_hat1_implicit * yp[j]) / tb->mu_real;
            }
            count_fun(b, t_new, y_new, st->tmpn, err);
            if (!all_finite(err, n)) { st->status = -2; return 0; }          /* lu_solve(check_finite) raises */
            orc_lu_solve(n, st->LUr, st->pivr, err);
            b->cnt[ORC_CNT_NSOLVE] += 1;
            for (int j = 0; j < n; ++j) err[j] = -h * tb->mu_real * err[j];
        } else {
            double *yh = st->yhat, *yph0 = st->tmpn2, *yph = st->tmpn;
            for (int j = 0; j < n; ++j) {
                double acc = 0.0;
                for (int l = 0; l < s; ++l) acc += tb->b_hat_implicit[l] * st->Yp[l * n + j];
                yph0[j] = -(y[j] / h + tb->b_hat1_implicit * yp[j] + acc) / tb->mu_real;
                yh[j] = y_new[j];
            }
            for (int it = 0; it < o->newton_iter_embedded; ++it) {
                for (int j = 0; j < n; ++j) yph[j] = yph0[j] + yh[j] / (h * tb->mu_real);
                count_fun(b, t_new, yh, yph, err);
                if (!all_finite(err, n)) { st->status = -2; return 0; }
                orc_lu_solve(n, st->LUr, st->pivr, err);
                b->cnt[ORC_CNT_NSOLVE] += 1;
                for (int j = 0; j < n; ++j) yh[j] -= h * tb->mu_real * err[j];
            }
            for (int j = 0; j < n; ++j) err[j] = yh[j] - y_new[j];
        }
        /* mix with the collocation error, radau.py:566,590-594 */
        {
            double w = o->continuous_error_weight;
            for (int j = 0; j < n; ++j) {
                double e = (1 - w) * fabs(err[j]);
                if (w != 0) {
                    double acc = 0.0;
                    for (int l = 0; l < s; ++l) acc += tb->P2row0[l] * st->Y[l * n + j];
                    e = w * pow(fabs(y[j] - acc), (s + 1.0) / s) + e;
                }
                st->tmpn[j] = e / st->scale[j];
            }
            error_norm = rms_norm(st->tmpn, n);
        }
        safety = 0.9 * (2 * st->newton_max_iter + 1) / (2 * st->newton_max_iter + n_iter);

        if (error_norm > 1) {                                                /* radau.py:598-603 */
            factor = predict_factor(h_abs, h_abs_old, error_norm, error_norm_old, have_old, s);
            have_factor = 1;
            h_abs *= safety * factor;
            st->lu_valid = 0;
            b->cnt[ORC_CNT_ERR_REJECTED] += 1;
        } else {
            break;
        }
    }

    /* accepted, radau.py:607-657 */
    int recompute_jac = (n_iter > 2) && have_rate && (rate > o->jac_recompute_rate);
    if (!have_factor)
        factor = safety * predict_factor(h_abs, h_abs_old, error_norm, error_norm_old, have_old, s);
    if (!recompute_jac && o->deadband_lo <= factor && factor <= o->deadband_hi)
        factor = 1;
    else
        st->lu_valid = 0;

    const double *y_new = st->Y + (s - 1) * n, *yp_new = st->Yp + (s - 1) * n;
    if (recompute_jac) {
        fd_jacobian(b, t_new, y_new, yp_new);
        st->current_jac = 1;
    } else {
        st->current_jac = 0;
    }

    st->h_abs_old = st->h_abs;            /* the PRE-step proposal (quirk A.3-4) */
    st->error_norm_old = error_norm;
    st->have_old = 1;
    st->h_abs = h_abs * factor;

    /* dense output: ZP = (Y - y_old).T @ P, radau.py:661-666 */
    for (int i = 0; i < n; ++i)
        for (int k = 0; k < s; ++k) {
            double acc = 0.0;
            for (int l = 0; l < s; ++l) acc += (st->Y[l * n + i] - y[i]) * tb->P[l * s + k];
            st->ZP[i * s + k] = acc;
        }
    memcpy(st->y_old, y, sizeof(double) * n);
    st->t_old = t;
    st->h_old = t_new - t;
    st->have_sol = 1;

    memcpy(b->y, y_new, sizeof(double) * n);
    memcpy(b->yp, yp_new, sizeof(double) * n);
    b->t = t_new;
    b->cnt[ORC_CNT_ACCEPTED] += 1;
    return 1;
}

static void base_init(base_state *b, orc_residual_fn fn, int n, const double *p, double rtol,
                      const double *atol, double t0, double t_bound, double max_step,
                      const double *y0, const double *yp0, int64_t *counters)
{
    memset(b, 0, sizeof(*b));
    b->fn = fn; b->n = n; b->p = p; b->rtol = rtol; b->atol = atol;
    b->jacfn = g_analytic_jac ? orc_jacobian_of(fn) : 0;
    b->t = t0; b->t_bound = t_bound; b->max_step = max_step;
    b->dir = (t_bound != t0) ? ((t_bound > t0) ? 1.0 : -1.0) : 1.0;         /* base.py:214 */
    b->y = malloc(sizeof(double) * n); b->yp = malloc(sizeof(double) * n);
    memcpy(b->y, y0, sizeof(double) * n); memcpy(b->yp, yp0, sizeof(double) * n);
    b->Jy = malloc(sizeof(double) * n * n); b->Jyp = malloc(sizeof(double) * n * n);
    b->factor_y = malloc(sizeof(double) * n); b->factor_yp = malloc(sizeof(double) * n);
    b->ftmp = malloc(sizeof(double) * n);
    for (int i = 0; i < n; ++i) b->factor_y[i] = b->factor_yp[i] = sqrt(EPS); /* SP/common.py:316 */
    b->cnt = counters;
    for (int i = 0; i < ORC_NUM_COUNTERS; ++i) counters[i] = 0;
}

static void base_free(base_state *b)
{
    free(b->y); free(b->yp); free(b->Jy); free(b->Jyp);
    free(b->factor_y); free(b->factor_yp); free(b->ftmp);
}

/* direction-aware "te has been passed by t": solve_dae's searchsorted slicing, dae.py:496-515 */
static int eval_reached(double te, double t, double dir)
{
    return dir > 0 ? (te <= t) : (te >= t);
}


/* ------------------------------------------------------------------------- */
/* events (dae.py:30-109, 426-490): threshold events g(t, y, yp) = x[component] - value on   */
/* x = y or yp; sign-change detection = scipy find_active_events, root = scipy brentq with    */
/* xtol = rtol = 4 EPS on the step's dense output, terminal handling = handle_events          */
/* ------------------------------------------------------------------------- */
typedef void (*dense_fn)(const void *st, double t, double *y, double *yp);

typedef struct {
    const orc_events *spec;
    int64_t cap;
    double *t_events, *y_events, *yp_events;     /* [ne][cap], [ne][cap][n] */
    int32_t *n_found;                            /* [ne] */
    double g[ORC_MAX_EVENTS];
    int64_t count[ORC_MAX_EVENTS];
    orc_event_fn fn;                             /* the problem's event functions (of_yp = 2) */
    const double *p;
} event_ctx;

static __thread event_ctx *tl_ev = NULL;

static double event_value(const event_ctx *c, int k, double t, const double *y, const double *yp)
{
    const orc_events *e = c->spec;
    if (e->of_yp[k] == 2) return c->fn(e->component[k], t, y, yp, c->p) - e->value[k];
    return (e->of_yp[k] ? yp[e->component[k]] : y[e->component[k]]) - e->value[k];
}

/* scipy/optimize/Zeros/brentq.c (maxiter = 100); f(t) = event k on the dense output */
static double brentq_event(const event_ctx *e, int k, dense_fn dense, const void *st, int n,
                           double xa, double xb, double *ytmp, double *yptmp)
{
    const double xtol = 4 * EPS, rtol = 4 * EPS;
    double xpre = xa, xcur = xb, xblk = 0., fpre, fcur, fblk = 0., spre = 0., scur = 0., sbis, delta, stry, dpre, dblk;
    (void)n;
    dense(st, xpre, ytmp, yptmp); fpre = event_value(e, k, xpre, ytmp, yptmp);
    dense(st, xcur, ytmp, yptmp); fcur = event_value(e, k, xcur, ytmp, yptmp);
    if (fpre == 0) return xpre;
    if (fcur == 0) return xcur;
    if (signbit(fpre) == signbit(fcur)) return NAN;                          /* ValueError in scipy */
    for (int i = 0; i < 100; ++i) {
        if (fpre != 0 && fcur != 0 && (signbit(fpre) != signbit(fcur))) {
            xblk = xpre; fblk = fpre; spre = scur = xcur - xpre;
        }
        if (fabs(fblk) < fabs(fcur)) {
            xpre = xcur; xcur = xblk; xblk = xpre;
            fpre = fcur; fcur = fblk; fblk = fpre;
        }
        delta = (xtol + rtol * fabs(xcur)) / 2;
        sbis = (xblk - xcur) / 2;
        if (fcur == 0 || fabs(sbis) < delta) return xcur;
        if (fabs(spre) > delta && fabs(fcur) < fabs(fpre)) {
            if (xpre == xblk) {
                stry = -fcur * (xcur - xpre) / (fcur - fpre);                /* secant */
            } else {
                dpre = (fpre - fcur) / (xpre - xcur);                        /* inverse quadratic */
                dblk = (fblk - fcur) / (xblk - xcur);
                stry = -fcur * (fblk * dblk - fpre * dpre) / (dblk * dpre * (fblk - fpre));
            }
            double lim = fmin(fabs(spre), 3 * fabs(sbis) - delta);
            if (2 * fabs(stry) < lim) { spre = scur; scur = stry; }
            else { spre = sbis; scur = sbis; }
        } else {
            spre = sbis; scur = sbis;
        }
        xpre = xcur; fpre = fcur;
        if (fabs(scur) > delta) xcur += scur;
        else xcur += (sbis > 0 ? delta : -delta);
        dense(st, xcur, ytmp, yptmp); fcur = event_value(e, k, xcur, ytmp, yptmp);
    }
    return xcur;
}

/* After an accepted step [t_old, t]: dae.py:466-490.  Returns 1 when a terminal event fired;
 * then *t_stop, y_stop, yp_stop hold the event point that replaces the step end. */
static int handle_events_step(event_ctx *c, dense_fn dense, const void *st, int n, double t_old,
                              double t, const double *y, const double *yp, double dir,
                              double *t_stop, double *y_stop, double *yp_stop)
{
    const orc_events *e = c->spec;
    int ne = e->n_events, act[ORC_MAX_EVENTS], na = 0;
    double gnew[ORC_MAX_EVENTS], roots[ORC_MAX_EVENTS];
    double *ytmp = malloc(sizeof(double) * n), *yptmp = malloc(sizeof(double) * n);
    for (int k = 0; k < ne; ++k) {
        gnew[k] = event_value(c, k, t, y, yp);
        int up = (c->g[k] <= 0) && (gnew[k] >= 0), down = (c->g[k] >= 0) && (gnew[k] <= 0);
        int d = e->direction[k];
        if ((up && d > 0) || (down && d < 0) || ((up || down) && d == 0)) act[na++] = k;
    }
    int terminate = 0;
    if (na > 0) {
        int any_term = 0;
        for (int a = 0; a < na; ++a) {
            int k = act[a];
            c->count[k] += 1;
            roots[a] = brentq_event(c, k, dense, st, n, t_old, t, ytmp, yptmp);
        }
        for (int a = 0; a < na; ++a)
            if (e->max_events[act[a]] > 0 && c->count[act[a]] >= e->max_events[act[a]]) any_term = 1;
        int nrec = na;
        if (any_term) {
            /* stable sort by root in the direction of integration, cut after the first terminal one */
            for (int i = 1; i < na; ++i)
                for (int j = i; j > 0; --j) {
                    int swap = dir > 0 ? (roots[j] < roots[j - 1]) : (-roots[j] < -roots[j - 1]);
                    if (!swap) break;
                    double tr = roots[j]; roots[j] = roots[j - 1]; roots[j - 1] = tr;
                    int ti = act[j]; act[j] = act[j - 1]; act[j - 1] = ti;
                }
            for (int a = 0; a < na; ++a)
                if (e->max_events[act[a]] > 0 && c->count[act[a]] >= e->max_events[act[a]]) { nrec = a + 1; break; }
            terminate = 1;
        }
        for (int a = 0; a < nrec; ++a) {
            int k = act[a];
            dense(st, roots[a], ytmp, yptmp);
            int32_t idx = c->n_found[k];
            if (idx < c->cap) {
                c->t_events[k * c->cap + idx] = roots[a];
                memcpy(c->y_events + (k * c->cap + idx) * n, ytmp, sizeof(double) * n);
                memcpy(c->yp_events + (k * c->cap + idx) * n, yptmp, sizeof(double) * n);
            }
            c->n_found[k] = idx + 1;
        }
        if (terminate) {
            *t_stop = roots[nrec - 1];
            dense(st, *t_stop, y_stop, yp_stop);
        }
    }
    for (int k = 0; k < ne; ++k) c->g[k] = gnew[k];
    free(ytmp); free(yptmp);
    return terminate;
}

static void radau_dense_void(const void *st, double t, double *y, double *yp) { radau_dense_eval((const radau_state *)st, t, y, yp); }

int orc_solve_radau(orc_residual_fn fn, int n, const double *p,
                    double t0, double t_bound, const double *y0, const double *yp0,
                    const double *atol, const orc_radau_opts *o, const orc_radau_tables *tb,
                    const double *t_eval, int64_t m, double *y_eval, double *yp_eval,
                    int64_t *n_eval_valid,
                    double *y_final, double *yp_final, double *t_final, int64_t *counters,
                    double *t_steps, double *y_steps, int64_t steps_cap, int64_t *n_steps)
{
    radau_state st;
    memset(&st, 0, sizeof(st));
    base_state *b = &st.b;
    double rtol = o->rtol;
    if (rtol < 100 * EPS) rtol = 100 * EPS;                                 /* validate_tol */
    base_init(b, fn, n, p, rtol, atol, t0, t_bound, o->max_step, y0, yp0, counters);
    int s = tb->s, ncs = s / 2;
    st.tb = tb; st.o = o; st.s = s; st.ncs = ncs;
    st.newton_max_iter = o->newton_max_iter > 0 ? o->newton_max_iter
                                                : 7 + (int)((s - 3) * 2.5);  /* radau.py:440-441 */
    {
        double expmns = (s + 1.0) / (2.0 * s);                              /* radau.py:422-433 */
        st.newton_tol = fmax(10 * EPS / rtol, fmin(0.03, pow(rtol, 1.0 / expmns - 1)));
    }
    st.h_abs = o->first_step > 0 ? o->first_step
             : orc_select_initial_step(n, t0, y0, yp0, t_bound, rtol, atol, o->max_step);
    st.LUr = malloc(sizeof(double) * n * n); st.pivr = malloc(sizeof(int) * n);
    st.LUcr = malloc(sizeof(double) * n * n * (ncs + 1)); st.LUci = malloc(sizeof(double) * n * n * (ncs + 1));
    st.pivc = malloc(sizeof(int) * n * (ncs + 1));
    st.Y = malloc(sizeof(double) * s * n); st.Yp = malloc(sizeof(double) * s * n);
    st.F = malloc(sizeof(double) * s * n); st.dW = malloc(sizeof(double) * s * n);
    st.dYp = malloc(sizeof(double) * s * n);
    st.scale = malloc(sizeof(double) * n); st.err = malloc(sizeof(double) * n);
    st.tmpn = malloc(sizeof(double) * n); st.tmpn2 = malloc(sizeof(double) * n);
    st.yhat = malloc(sizeof(double) * n);
    st.y_old = malloc(sizeof(double) * n); st.ZP = malloc(sizeof(double) * n * s);

    /* first Jacobian in the constructor (base.py:297); current_jac starts True (radau.py:460) */
    fd_jacobian(b, t0, b->y, b->yp);
    st.current_jac = 1;

    int64_t ie = 0, nsteps = 0;
    int status = 0;
    if (t_steps && steps_cap > 0) {
        t_steps[0] = t0; memcpy(y_steps, y0, sizeof(double) * n); nsteps = 1;
    }
    if (n == 0 || t0 == t_bound) {
        /* base.py:376-381: degenerate, finishes immediately */
    } else {
        event_ctx *ec = tl_ev;
        if (ec) for (int k = 0; k < ec->spec->n_events; ++k) ec->g[k] = event_value(ec, k, t0, y0, yp0);
        for (;;) {
            double t_prev = b->t;
            if (!radau_step(&st)) { status = st.status; break; }
            if (ec) {
                double ts;
                if (handle_events_step(ec, radau_dense_void, &st, n, t_prev, b->t, b->y, b->yp, b->dir,
                                       &ts, st.tmpn, st.tmpn2)) {
                    b->t = ts; memcpy(b->y, st.tmpn, sizeof(double) * n); memcpy(b->yp, st.tmpn2, sizeof(double) * n);
                    status = 1;                                             /* dae.py:484-487 */
                }
            }
            if (t_steps && nsteps < steps_cap) {
                t_steps[nsteps] = b->t; memcpy(y_steps + nsteps * n, b->y, sizeof(double) * n); ++nsteps;
            }
            while (ie < m && eval_reached(t_eval[ie], b->t, b->dir)) {
                radau_dense_eval(&st, t_eval[ie], y_eval + ie * n, yp_eval + ie * n);
                ++ie;
            }
            if (status == 1) break;
            if (b->dir * (b->t - b->t_bound) >= 0) break;                   /* base.py:389-390 */
            if (o->max_steps > 0 && counters[ORC_CNT_ACCEPTED] >= o->max_steps) { status = -3; break; }
        }
    }
    memcpy(y_final, b->y, sizeof(double) * n);
    memcpy(yp_final, b->yp, sizeof(double) * n);
    *t_final = b->t;
    if (n_eval_valid) *n_eval_valid = ie;
    if (n_steps) *n_steps = nsteps;

    free(st.LUr); free(st.pivr); free(st.LUcr); free(st.LUci); free(st.pivc);
    free(st.Y); free(st.Yp); free(st.F); free(st.dW); free(st.dYp);
    free(st.scale); free(st.err); free(st.tmpn); free(st.tmpn2); free(st.yhat);
    free(st.y_old); free(st.ZP);
    base_free(b);
    return status;
}

/* ------------------------------------------------------------------------- */
/* BDF / NDF                                                                 */
/* ------------------------------------------------------------------------- */
#define BDF_NEWTON_MAXITER 4
#define BDF_MIN_FACTOR 0.2
#define BDF_MAX_FACTOR 10.0

typedef struct {
    base_state b;
    const orc_bdf_opts *o;
    int max_order, order, n_equal_steps, lu_valid;
    double newton_tol, h_abs;
    double gamma[ORC_MAX_ORDER + 1], alpha[ORC_MAX_ORDER + 1], error_const[ORC_MAX_ORDER + 2];
    double *D;                           /* (max_order+3) * n */
    double *LU; int *piv;
    double *y_predict, *psi, *scale, *y_new, *yp_new, *d, *f, *tmp;
    double t_old;
    int status;
} bdf_state;

/* compute_R, bdf.py:14-21:  cumprod over rows of M[i,j] = (i-1-factor*j)/i, M[0,:] = 1 */
static void compute_R(int order, double factor, double *R /* (order+1)^2 */)
{
    int q = order + 1;
    for (int j = 0; j < q; ++j) R[j] = 1.0;
    for (int i = 1; i < q; ++i)
        for (int j = 0; j < q; ++j) {
            double mij = (j == 0) ? 0.0 : ((i - 1 - factor * j) / i);
            R[i * q + j] = R[(i - 1) * q + j] * mij;
        }
}

/* change_D, bdf.py:24-29: D[:order+1] = (R @ U).T @ D[:order+1] */
static void change_D(double *D, int n, int order, double factor)
{
    int q = order + 1;
    double R[(ORC_MAX_ORDER + 1) * (ORC_MAX_ORDER + 1)], U[(ORC_MAX_ORDER + 1) * (ORC_MAX_ORDER + 1)];
    double RU[(ORC_MAX_ORDER + 1) * (ORC_MAX_ORDER + 1)], col[ORC_MAX_ORDER + 1];
    compute_R(order, factor, R);
    compute_R(order, 1.0, U);
    for (int i = 0; i < q; ++i)
        for (int j = 0; j < q; ++j) {
            double acc = 0.0;
            for (int l = 0; l < q; ++l) acc += R[i * q + l] * U[l * q + j];
            RU[i * q + j] = acc;
        }
    for (int c = 0; c < n; ++c) {
        for (int i = 0; i < q; ++i) {
            double acc = 0.0;
            for (int l = 0; l < q; ++l) acc += RU[l * q + i] * D[l * n + c];
            col[i] = acc;
        }
        for (int i = 0; i < q; ++i) D[i * n + c] = col[i];
    }
}

/* solve_bdf_system, bdf.py:32-64 */
static int bdf_newton(bdf_state *st, double t_new, double c, int *n_iter_out)
{
    base_state *b = &st->b;
    int n = b->n;
    double *y = st->y_new, *yp = st->yp_new, *d = st->d, *f = st->f;
    for (int j = 0; j < n; ++j) { d[j] = 0.0; y[j] = st->y_predict[j]; }
    double dy_norm_old = 0.0;
    int has_old = 0, converged = 0, k;
    for (k = 0; k < BDF_NEWTON_MAXITER; ++k) {
        for (int j = 0; j < n; ++j) yp[j] = c * d[j] + st->psi[j];
        count_fun(b, t_new, y, yp, f);
        if (!all_finite(f, n)) break;
        for (int j = 0; j < n; ++j) f[j] = -f[j];
        orc_lu_solve(n, st->LU, st->piv, f);                /* f now holds dy */
        b->cnt[ORC_CNT_NSOLVE] += 1;
        for (int j = 0; j < n; ++j) st->tmp[j] = f[j] / st->scale[j];
        double dy_norm = rms_norm(st->tmp, n);
        double rate = 0.0;
        if (has_old) rate = dy_norm / dy_norm_old;
        if (has_old && (rate >= 1 ||
                        pow(rate, BDF_NEWTON_MAXITER - k) / (1 - rate) * dy_norm > st->newton_tol))
            break;                                           /* before the update, bdf.py:52-53 */
        for (int j = 0; j < n; ++j) { y[j] += f[j]; d[j] += f[j]; }
        if (dy_norm == 0 || (has_old && rate / (1 - rate) * dy_norm < st->newton_tol)) {
            converged = 1;
            break;
        }
        dy_norm_old = dy_norm;
        has_old = 1;
    }
    if (k == BDF_NEWTON_MAXITER) k = BDF_NEWTON_MAXITER - 1;
    *n_iter_out = k + 1;
    return converged;
}

/* BDFDAE._step_impl, bdf.py:263-415 */
static int bdf_step(bdf_state *st)
{
    base_state *b = &st->b;
    int n = b->n;
    double t = b->t;
    double *D = st->D;

    double min_step = min_step_at(t, b->dir);
    double h_abs;
    if (st->h_abs > b->max_step) {                                           /* bdf.py:271-274 */
        h_abs = b->max_step;
        change_D(D, n, st->order, b->max_step / st->h_abs);
        st->n_equal_steps = 0;
    } else if (st->h_abs < min_step) {                                       /* bdf.py:275-278 */
        h_abs = min_step;
        change_D(D, n, st->order, min_step / st->h_abs);
        st->n_equal_steps = 0;
    } else {
        h_abs = st->h_abs;
    }
    int order = st->order;
    int current_jac = 0;                     /* self.jac is None is False for the FD wrapper, bdf.py:293 */
    int n_iter = 0;
    double t_new = t, safety = 0.0, error_norm = 0.0;

    for (;;) {
        if (h_abs < min_step) { st->status = -1; return 0; }
        double h = h_abs * b->dir;
        t_new = t + h;
        if (b->dir * (t_new - b->t_bound) > 0) {                             /* bdf.py:303-307 */
            t_new = b->t_bound;
            change_D(D, n, order, fabs(t_new - t) / h_abs);
            st->n_equal_steps = 0;
            st->lu_valid = 0;
        }
        h = t_new - t;
        h_abs = fabs(h);

        for (int j = 0; j < n; ++j) {                                        /* bdf.py:312-316 */
            double acc = 0.0;
            for (int i = 0; i <= order; ++i) acc += D[i * n + j];
            st->y_predict[j] = acc;
            st->scale[j] = b->atol[j] + b->rtol * fabs(acc);
            double ps = 0.0;
            for (int i = 1; i <= order; ++i) ps += D[i * n + j] * st->gamma[i];
            st->psi[j] = ps / h;
        }
        double c = st->alpha[order] / h;
        int converged = 0;
        for (;;) {
            if (!st->lu_valid) {                                             /* bdf.py:320-321 */
                for (int i = 0; i < n * n; ++i) st->LU[i] = b->Jy[i] + c * b->Jyp[i];
                if (!all_finite(st->LU, n * n)) { st->status = -2; return 0; }
                orc_lu_factor(n, st->LU, st->piv);
                b->cnt[ORC_CNT_NLU] += 1;
                st->lu_valid = 1;
            }
            converged = bdf_newton(st, t_new, c, &n_iter);
            b->cnt[ORC_CNT_NEWTON_ITER] += n_iter;
            if (converged) break;
            b->cnt[ORC_CNT_NEWTON_FAIL] += 1;
            if (current_jac) break;
            fd_jacobian(b, t, b->y, b->yp);                                  /* bdf.py:330 */
            st->lu_valid = 0;
            current_jac = 1;
        }
        if (!converged) {                                                    /* bdf.py:334-340 */
            h_abs *= 0.5;
            change_D(D, n, order, 0.5);
            st->n_equal_steps = 0;
            st->lu_valid = 0;
            continue;
        }
        safety = 0.9 * (2 * BDF_NEWTON_MAXITER + 1) / (2 * BDF_NEWTON_MAXITER + n_iter);
        for (int j = 0; j < n; ++j) {
            st->scale[j] = b->atol[j] + b->rtol * fabs(st->y_new[j]);
            st->tmp[j] = st->error_const[order] * st->d[j] / st->scale[j];
        }
        error_norm = rms_norm(st->tmp, n);
        if (error_norm > 1) {                                                /* bdf.py:348-355 */
            double factor = fmax(BDF_MIN_FACTOR, safety * pow(error_norm, -1.0 / (order + 1)));
            h_abs *= factor;
            change_D(D, n, order, factor);
            st->n_equal_steps = 0;
            b->cnt[ORC_CNT_ERR_REJECTED] += 1;
        } else {
            break;
        }
    }

    st->n_equal_steps += 1;
    st->t_old = t;
    b->t = t_new;
    memcpy(b->y, st->y_new, sizeof(double) * n);
    memcpy(b->yp, st->yp_new, sizeof(double) * n);
    st->h_abs = h_abs;
    b->cnt[ORC_CNT_ACCEPTED] += 1;

    /* update differences, bdf.py:374-377 */
    for (int j = 0; j < n; ++j) {
        D[(order + 2) * n + j] = st->d[j] - D[(order + 1) * n + j];
        D[(order + 1) * n + j] = st->d[j];
    }
    for (int i = order; i >= 0; --i)
        for (int j = 0; j < n; ++j) D[i * n + j] += D[(i + 1) * n + j];

    if (st->n_equal_steps < order + 1) return 1;

    double norms[3];
    if (order > 1) {                                                         /* bdf.py:382-386 */
        for (int j = 0; j < n; ++j) st->tmp[j] = st->error_const[order - 1] * D[order * n + j] / st->scale[j];
        norms[0] = rms_norm(st->tmp, n);
    } else norms[0] = INFINITY;
    norms[1] = error_norm;
    if (order < st->max_order) {                                             /* bdf.py:388-392 */
        for (int j = 0; j < n; ++j) st->tmp[j] = st->error_const[order + 1] * D[(order + 2) * n + j] / st->scale[j];
        norms[2] = rms_norm(st->tmp, n);
    } else norms[2] = INFINITY;

    double fcts[3];                                                          /* bdf.py:394-399 */
    for (int i = 0; i < 3; ++i) fcts[i] = pow(norms[i], -1.0 / (order + i));
    double best = fcts[0]; int arg = 0;                                      /* np.argmax: first max, NaN wins */
    for (int i = 1; i < 3; ++i)
        if (!isnan(best) && (fcts[i] > best || isnan(fcts[i]))) { best = fcts[i]; arg = i; }
    order += arg - 1;
    st->order = order;
    double factor = fmin(BDF_MAX_FACTOR, safety * best);
    st->h_abs *= factor;
    change_D(D, n, order, factor);
    st->n_equal_steps = 0;
    st->lu_valid = 0;
    return 1;
}

/* BdfDenseOutput._call_impl ("naive implementation"), bdf.py:525-548, built from
 * D[:order+1] with h = h_abs*direction AFTER the step's rescale (bdf.py:417-419) */
static void bdf_dense_eval(const bdf_state *st, double t, double *y, double *yp)
{
    int n = st->b.n, order = st->order;
    double h = st->h_abs * st->b.dir, tn = st->b.t;
    for (int c = 0; c < n; ++c) { y[c] = st->D[c]; yp[c] = 0.0; }
    double fact = 1.0;
    for (int j = 1; j <= order; ++j) {
        fact *= j;
        double fac = 1.0, dfac = 0.0;
        for (int mm = 0; mm < j; ++mm) {
            fac *= t - (tn - mm * h);
            double dprod = 1.0;
            for (int i = 0; i < j; ++i)
                if (i != mm) dprod *= t - (tn - i * h);
            dfac += dprod;
        }
        double denom = fact * pow(h, j);
        for (int c = 0; c < n; ++c) {
            y[c] += st->D[j * n + c] * fac / denom;
            yp[c] += st->D[j * n + c] * dfac / denom;
        }
    }
}

static void bdf_dense_void(const void *st, double t, double *y, double *yp) { bdf_dense_eval((const bdf_state *)st, t, y, yp); }

int orc_solve_bdf(orc_residual_fn fn, int n, const double *p,
                  double t0, double t_bound, const double *y0, const double *yp0,
                  const double *atol, const orc_bdf_opts *o,
                  const double *t_eval, int64_t m, double *y_eval, double *yp_eval,
                  int64_t *n_eval_valid,
                  double *y_final, double *yp_final, double *t_final, int64_t *counters,
                  double *t_steps, double *y_steps, int64_t steps_cap, int64_t *n_steps)
{
    bdf_state st;
    memset(&st, 0, sizeof(st));
    base_state *b = &st.b;
    double rtol = o->rtol;
    if (rtol < 100 * EPS) rtol = 100 * EPS;
    base_init(b, fn, n, p, rtol, atol, t0, t_bound, o->max_step, y0, yp0, counters);
    st.o = o;
    st.max_order = o->max_order > 0 ? o->max_order : 5;
    st.newton_tol = fmax(10 * EPS / rtol, fmin(0.03, sqrt(rtol)));          /* bdf.py:234 */
    {
        static const double k_stab[7] = {0, -37.0 / 200, -1.0 / 9, 0.0834, 0.0665, 0.0551, 0.0464};
        static const double k_eff[7] = {0, -37.0 / 200, -1.0 / 9, -0.0823, -0.0415, 0, 0};
        static const double k_bdf[7] = {0, -37.0 / 200, -1.0 / 9, 0, 0, 0, 0};
        const double *kappa = o->ndf_strategy == 0 ? k_stab : (o->ndf_strategy == 1 ? k_eff : k_bdf);
        double g = 0.0;
        st.gamma[0] = 0.0;
        for (int k = 1; k <= st.max_order; ++k) { g += 1.0 / k; st.gamma[k] = g; }
        for (int k = 0; k <= st.max_order; ++k) {
            st.alpha[k] = (1 - kappa[k]) * st.gamma[k];
            st.error_const[k] = kappa[k] * st.gamma[k] + 1.0 / (k + 1);
        }
    }
    st.h_abs = o->first_step > 0 ? o->first_step
             : orc_select_initial_step(n, t0, y0, yp0, t_bound, rtol, atol, o->max_step);
    st.D = calloc((size_t)(st.max_order + 3) * n, sizeof(double));
    for (int j = 0; j < n; ++j) {                                           /* bdf.py:254-257 */
        st.D[j] = y0[j];
        st.D[n + j] = yp0[j] * st.h_abs * b->dir;
    }
    st.order = 1; st.n_equal_steps = 0; st.lu_valid = 0;
    st.LU = malloc(sizeof(double) * n * n); st.piv = malloc(sizeof(int) * n);
    st.y_predict = malloc(sizeof(double) * n); st.psi = malloc(sizeof(double) * n);
    st.scale = malloc(sizeof(double) * n); st.y_new = malloc(sizeof(double) * n);
    st.yp_new = malloc(sizeof(double) * n); st.d = malloc(sizeof(double) * n);
    st.f = malloc(sizeof(double) * n); st.tmp = malloc(sizeof(double) * n);

    fd_jacobian(b, t0, b->y, b->yp);                                        /* base.py:297 */

    int64_t ie = 0, nsteps = 0;
    int status = 0;
    if (t_steps && steps_cap > 0) {
        t_steps[0] = t0; memcpy(y_steps, y0, sizeof(double) * n); nsteps = 1;
    }
    if (!(n == 0 || t0 == t_bound)) {
        event_ctx *ec = tl_ev;
        if (ec) for (int k = 0; k < ec->spec->n_events; ++k) ec->g[k] = event_value(ec, k, t0, y0, yp0);
        for (;;) {
            double t_prev = b->t;
            if (!bdf_step(&st)) { status = st.status; break; }
            if (ec) {
                double ts;
                if (handle_events_step(ec, bdf_dense_void, &st, n, t_prev, b->t, b->y, b->yp, b->dir,
                                       &ts, st.tmp, st.f)) {
                    b->t = ts; memcpy(b->y, st.tmp, sizeof(double) * n); memcpy(b->yp, st.f, sizeof(double) * n);
                    status = 1;
                }
            }
            if (t_steps && nsteps < steps_cap) {
                t_steps[nsteps] = b->t; memcpy(y_steps + nsteps * n, b->y, sizeof(double) * n); ++nsteps;
            }
            while (ie < m && eval_reached(t_eval[ie], b->t, b->dir)) {
                bdf_dense_eval(&st, t_eval[ie], y_eval + ie * n, yp_eval + ie * n);
                ++ie;
            }
            if (status == 1) break;
            if (b->dir * (b->t - b->t_bound) >= 0) break;
            if (o->max_steps > 0 && counters[ORC_CNT_ACCEPTED] >= o->max_steps) { status = -3; break; }
        }
    }
    memcpy(y_final, b->y, sizeof(double) * n);
    memcpy(yp_final, b->yp, sizeof(double) * n);
    *t_final = b->t;
    if (n_eval_valid) *n_eval_valid = ie;
    if (n_steps) *n_steps = nsteps;

    free(st.D); free(st.LU); free(st.piv); free(st.y_predict); free(st.psi); free(st.scale);
    free(st.y_new); free(st.yp_new); free(st.d); free(st.f); free(st.tmp);
    base_free(b);
    return status;
}

/* ------------------------------------------------------------------------- */
/* ensemble driver over POSIX threads (the CPU baseline of bench.py)         */
/* ------------------------------------------------------------------------- */
typedef struct {
    int problem_id, method, n, np;
    int64_t N;
    const double *params, *y0, *yp0, *atol, *t_eval;
    double t0, t_bound;
    const orc_radau_opts *ropts; const orc_radau_tables *tb; const orc_bdf_opts *bopts;
    int64_t m;
    double *y_eval, *yp_eval, *y_final, *yp_final, *t_final;
    int64_t *n_eval_valid, *counters;
    int32_t *status;
    const orc_events *ev; int64_t ev_cap;
    double *t_events, *y_events, *yp_events; int32_t *n_found;     /* [N][ne][cap]([n]), [N][ne] */
    int64_t next;
    pthread_mutex_t lock;
} ens_job;

static void *ens_worker(void *arg)
{
    ens_job *J = (ens_job *)arg;
    orc_residual_fn fn = orc_problem_residual(J->problem_id);
    const int64_t chunk = 16;
    for (;;) {
        pthread_mutex_lock(&J->lock);
        int64_t lo = J->next;
        J->next += chunk;
        pthread_mutex_unlock(&J->lock);
        if (lo >= J->N) break;
        int64_t hi = lo + chunk < J->N ? lo + chunk : J->N;
        for (int64_t i = lo; i < hi; ++i) {
            int n = J->n;
            double *ye = J->y_eval ? J->y_eval + i * J->m * n : NULL;
            double *ype = J->yp_eval ? J->yp_eval + i * J->m * n : NULL;
            int64_t nev = 0;
            int64_t mm = (ye && ype) ? J->m : 0;
            int st;
            event_ctx ec;
            if (J->ev && J->ev->n_events > 0) {
                int ne = J->ev->n_events;
                memset(&ec, 0, sizeof(ec));
                ec.spec = J->ev; ec.cap = J->ev_cap;
                ec.t_events = J->t_events + i * ne * J->ev_cap;
                ec.y_events = J->y_events + i * ne * J->ev_cap * n;
                ec.yp_events = J->yp_events + i * ne * J->ev_cap * n;
                ec.n_found = J->n_found + i * ne;
                ec.fn = orc_event_of(fn); ec.p = J->params + i * J->np;
                for (int k = 0; k < ne; ++k) ec.n_found[k] = 0;
                tl_ev = &ec;
            } else {
                tl_ev = NULL;
            }
            if (J->method == 0)
                st = orc_solve_radau(fn, n, J->params + i * J->np, J->t0, J->t_bound,
                                     J->y0 + i * n, J->yp0 + i * n, J->atol, J->ropts, J->tb,
                                     J->t_eval, mm, ye, ype, &nev,
                                     J->y_final + i * n, J->yp_final + i * n, J->t_final + i,
                                     J->counters + i * ORC_NUM_COUNTERS, NULL, NULL, 0, NULL);
            else
                st = orc_solve_bdf(fn, n, J->params + i * J->np, J->t0, J->t_bound,
                                   J->y0 + i * n, J->yp0 + i * n, J->atol, J->bopts,
                                   J->t_eval, mm, ye, ype, &nev,
                                   J->y_final + i * n, J->yp_final + i * n, J->t_final + i,
                                   J->counters + i * ORC_NUM_COUNTERS, NULL, NULL, 0, NULL);
            tl_ev = NULL;
            J->status[i] = st;
            if (J->n_eval_valid) J->n_eval_valid[i] = nev;
        }
    }
    return NULL;
}

static const orc_events *g_ev = NULL;
static int64_t g_ev_cap = 0;
static double *g_t_events = NULL, *g_y_events = NULL, *g_yp_events = NULL;
static int32_t *g_n_found = NULL;

/* events for the NEXT orc_solve_ensemble call (NULL spec switches them off again) */
void orc_set_events(const orc_events *spec, int64_t cap, double *t_events, double *y_events,
                    double *yp_events, int32_t *n_found)
{
    g_ev = spec; g_ev_cap = cap; g_t_events = t_events; g_y_events = y_events;
    g_yp_events = yp_events; g_n_found = n_found;
}

int orc_solve_ensemble(int problem_id, int method, int64_t N,
                       const double *params, const double *y0, const double *yp0,
                       double t0, double t_bound, const double *atol,
                       const orc_radau_opts *ropts, const orc_radau_tables *tb,
                       const orc_bdf_opts *bopts,
                       const double *t_eval, int64_t m, double *y_eval, double *yp_eval,
                       int64_t *n_eval_valid,
                       double *y_final, double *yp_final, double *t_final,
                       int32_t *status, int64_t *counters, int nthreads)
{
    ens_job J;
    memset(&J, 0, sizeof(J));
    if (orc_problem_dims(problem_id, &J.n, &J.np)) return -1;
    J.problem_id = problem_id; J.method = method; J.N = N;
    J.params = params; J.y0 = y0; J.yp0 = yp0; J.atol = atol; J.t_eval = t_eval;
    J.t0 = t0; J.t_bound = t_bound; J.ropts = ropts; J.tb = tb; J.bopts = bopts; J.m = m;
    J.y_eval = y_eval; J.yp_eval = yp_eval; J.y_final = y_final; J.yp_final = yp_final;
    J.t_final = t_final; J.n_eval_valid = n_eval_valid; J.counters = counters; J.status = status;
    J.ev = g_ev; J.ev_cap = g_ev_cap; J.t_events = g_t_events; J.y_events = g_y_events;
    J.yp_events = g_yp_events; J.n_found = g_n_found;
    J.next = 0;
    pthread_mutex_init(&J.lock, NULL);
    if (nthreads < 1) nthreads = 1;
    pthread_t *th = malloc(sizeof(pthread_t) * nthreads);
    for (int i = 0; i < nthreads; ++i) pthread_create(&th[i], NULL, ens_worker, &J);
    for (int i = 0; i < nthreads; ++i) pthread_join(th[i], NULL);
    free(th);
    pthread_mutex_destroy(&J.lock);
    return 0;
}
