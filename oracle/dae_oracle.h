/*
 * dae_oracle.h — CPU ORACLE (TEST INFRASTRUCTURE, NOT PRODUCT CODE).
 *
 * A plain-C, scalar restatement of the reference's hot path (JonasBreuling/scipy_dae
 * v0.1.0): RadauDAE / BDFDAE step loops, simplified Newton, scipy's num_jac finite
 * difference Jacobian, LAPACK-style partial-pivot LU (real and complex), embedded
 * error estimate, step-size / order control, dense output and solve_dae's t_eval
 * slicing.  Each function in dae_oracle.c cites the reference file:line it follows.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load this library.  The product (scipy_dae_b200 + libb200dae.so) never does.
 *
 * PARITY PINNING: the reference has no golden vectors; this oracle is
 * pinned against tests/golden/ (.npz), which oracle/gen_golden.py minted by running the
 * unmodified Python reference in the build container (tests/test_oracle_golden.py).
 */
#ifndef DAE_ORACLE_H
#define DAE_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ORC_MAX_STAGES 7
#define ORC_MAX_ORDER 6

/* residual F(t, y, yp; p) -> out[n]  (reference convention fun(t, y, yp, *args), dae.py:365-383) */
typedef void (*orc_residual_fn)(double t, const double *y, const double *yp,
                                const double *p, double *out);

/* Same layout as struct b200dae_radau_tables (include/b200dae.h) and
 * scipy_dae_b200.constants.CRadauTables; matrices row-major with stride s. */
typedef struct orc_radau_tables {
    int s;
    int order;
    double A[ORC_MAX_STAGES * ORC_MAX_STAGES];
    double C[ORC_MAX_STAGES];
    double T[ORC_MAX_STAGES * ORC_MAX_STAGES];
    double TI[ORC_MAX_STAGES * ORC_MAX_STAGES];
    double P[ORC_MAX_STAGES * ORC_MAX_STAGES];
    double P2row0[ORC_MAX_STAGES];
    double v_implicit[ORC_MAX_STAGES];
    double v_explicit[ORC_MAX_STAGES];
    double b_hat_implicit[ORC_MAX_STAGES];
    double mu_real;
    double mu_re[ORC_MAX_STAGES / 2];
    double mu_im[ORC_MAX_STAGES / 2];
    double b_hat1_implicit;
} orc_radau_tables;

typedef struct orc_radau_opts {
    double rtol;
    double max_step;                 /* +inf: unbounded */
    double first_step;               /* <= 0: select_initial_step */
    double continuous_error_weight;  /* 0 */
    double jac_recompute_rate;       /* 1e-3 */
    double deadband_lo, deadband_hi; /* 1.0, 1.2 */
    int newton_max_iter;             /* <= 0: 7 + int((s-3)*2.5) */
    int newton_iter_embedded;        /* 1 */
    int64_t max_steps;               /* safety cap on accepted steps, <= 0: none */
} orc_radau_opts;

typedef struct orc_bdf_opts {
    double rtol;
    double max_step;
    double first_step;
    int max_order;                   /* 5 */
    int ndf_strategy;                /* 0 stability, 1 efficiency, 2 plain BDF */
    int64_t max_steps;
} orc_bdf_opts;

/* per-instance counters, same order as the rows of the product's counters[8][N] */
enum {
    ORC_CNT_NFEV = 0, ORC_CNT_NJEV, ORC_CNT_NLU, ORC_CNT_NSOLVE,
    ORC_CNT_ACCEPTED, ORC_CNT_ERR_REJECTED, ORC_CNT_NEWTON_FAIL, ORC_CNT_NEWTON_ITER,
    ORC_NUM_COUNTERS
};

/* status: 0 finished, -1 step size too small, -2 non-finite iteration matrix, -3 max_steps hit */

/*
 * One instance.  Outputs:
 *   y_final[n], yp_final[n], *t_final, counters[8]
 *   t_eval[m] (sorted in integration direction) -> y_eval[m*n], yp_eval[m*n] (row k = point k),
 *   *n_eval_valid points were produced (solve_dae emits only points reached before a failure).
 *   optional step trace: t_steps[cap], y_steps[cap*n] (accepted step ends), *n_steps.
 * Returns the status.
 */
int orc_solve_radau(orc_residual_fn fn, int n, const double *p,
                    double t0, double t_bound, const double *y0, const double *yp0,
                    const double *atol /*[n]*/, const orc_radau_opts *opts,
                    const orc_radau_tables *tb,
                    const double *t_eval, int64_t m, double *y_eval, double *yp_eval,
                    int64_t *n_eval_valid,
                    double *y_final, double *yp_final, double *t_final, int64_t *counters,
                    double *t_steps, double *y_steps, int64_t steps_cap, int64_t *n_steps);

int orc_solve_bdf(orc_residual_fn fn, int n, const double *p,
                  double t0, double t_bound, const double *y0, const double *yp0,
                  const double *atol /*[n]*/, const orc_bdf_opts *opts,
                  const double *t_eval, int64_t m, double *y_eval, double *yp_eval,
                  int64_t *n_eval_valid,
                  double *y_final, double *yp_final, double *t_final, int64_t *counters,
                  double *t_steps, double *y_steps, int64_t steps_cap, int64_t *n_steps);

/* event functions of a problem (twins of Problem.events / the device functor's event()) */
typedef double (*orc_event_fn)(int which, double t, const double *y, const double *yp, const double *p);
orc_event_fn orc_problem_event(int problem_id);
orc_event_fn orc_event_of(orc_residual_fn fn);

/* Events g_k(t, y, yp) = x[component_k] - value_k, x = y (of_yp = 0) or yp (of_yp = 1), or the
 * problem's event function number component_k minus value_k (of_yp = 2);
 * direction / terminal as the attributes of a solve_dae event function (dae.py:226-247):
 * max_events <= 0 = never terminal, 1 = `terminal = True`, k = `terminal = k`. */
#define ORC_MAX_EVENTS 4
typedef struct orc_events {
    int32_t n_events;
    int32_t of_yp[ORC_MAX_EVENTS];
    int32_t component[ORC_MAX_EVENTS];
    double value[ORC_MAX_EVENTS];
    int32_t direction[ORC_MAX_EVENTS];
    int32_t max_events[ORC_MAX_EVENTS];
} orc_events;
/* outputs: t_events[N][ne][cap], y_events / yp_events[N][ne][cap][n], n_found[N][ne];
 * instance status 1 = a terminal event stopped the integration (solve_dae status 1) */
void orc_set_events(const orc_events *spec, int64_t cap, double *t_events, double *y_events,
                    double *yp_events, int32_t *n_found);

/* analytic Jacobian (the reference's `jac=` callable, base.py:299-331): Jy, Jyp row-major n*n */
typedef void (*orc_jacobian_fn)(double t, const double *y, const double *yp, const double *p,
                                double *Jy, double *Jyp);
/* 1: the solvers call the problem's analytic Jacobian instead of num_jac (process-wide switch,
 * read when an instance starts; problems without one keep the FD Jacobian) */
void orc_set_analytic_jac(int on);
orc_jacobian_fn orc_problem_jacobian(int problem_id);
orc_jacobian_fn orc_jacobian_of(orc_residual_fn fn);
int orc_problem_jac_eval(int problem_id, double t, const double *y, const double *yp,
                         const double *p, double *Jy, double *Jyp);

/* problem registry (oracle/problems.c) */
int orc_problem_lookup(const char *name, int *n, int *n_params);
orc_residual_fn orc_problem_residual(int problem_id);
int orc_problem_dims(int problem_id, int *n, int *n_params);

/*
 * Ensemble driver: N independent instances over `nthreads` POSIX threads (dynamic chunks).
 * AoS host layout: params[N][n_params], y0[N][n], yp0[N][n]; outputs y_final[N][n],
 * yp_final[N][n], t_final[N], status[N], counters[N][8], y_eval[N][m][n], yp_eval[N][m][n],
 * n_eval_valid[N].  method: 0 Radau (radau_opts, tb), 1 BDF (bdf_opts).
 */
int orc_solve_ensemble(int problem_id, int method, int64_t N,
                       const double *params, const double *y0, const double *yp0,
                       double t0, double t_bound, const double *atol,
                       const orc_radau_opts *ropts, const orc_radau_tables *tb,
                       const orc_bdf_opts *bopts,
                       const double *t_eval, int64_t m, double *y_eval, double *yp_eval,
                       int64_t *n_eval_valid,
                       double *y_final, double *yp_final, double *t_final,
                       int32_t *status, int64_t *counters, int nthreads);

/* consistent_initial_conditions (common.py:186-372), oracle/cic_oracle.c.
 * status: 0 converged, -1 "Convergence failed." (RuntimeError), -2 "Too many fixed components",
 * -3 "Index greater than one." (ValueErrors of solve_underdetermined_system). */
int orc_consistent_ic(orc_residual_fn fn, orc_jacobian_fn jac_fn, int n, const double *p, double t0,
                      double *y0, double *yp0, const int *fixed_y, const int *fixed_yp,
                      double rtol, const double *atol, int newton_maxiter, int chord_iter,
                      double safety, double *f_out, int *n_iter_out);
int orc_consistent_ic_ensemble(int problem_id, int use_analytic_jac, int64_t N, const double *params,
                               double t0, double *y0, double *yp0, const int *fixed_y, const int *fixed_yp,
                               double rtol, const double *atol, int newton_maxiter, int chord_iter,
                               double safety, double *f_out, int32_t *status, int32_t *n_iter);

/* Building blocks exported for unit tests */
double orc_select_initial_step(int n, double t0, const double *y0, const double *yp0,
                               double t_bound, double rtol, const double *atol, double max_step);
void orc_num_jac(orc_residual_fn fn, int n, const double *p, double t, const double *y,
                 const double *yp, const double *f, int wrt_yp, double rtol, const double *atol,
                 double *factor, double *J /* row-major n*n */);
int orc_lu_factor(int n, double *a /* row-major, overwritten */, int *piv);
void orc_lu_solve(int n, const double *lu, const int *piv, double *b);

#ifdef __cplusplus
}
#endif
#endif /* DAE_ORACLE_H */
