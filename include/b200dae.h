/*
 * b200dae.h — C-ABI of libb200dae.so, the B200 (sm_100a) ensemble DAE/IDE integrator.
 *
 * The reference (JonasBreuling/scipy_dae v0.1.0) is pure Python and has no FFI; its plug-in
 * boundary for the hot path is the DaeSolver protocol consumed by solve_dae
 * (scipy_dae/integrate/_dae/dae.py:359-361,405-408,446-518; base.py:50-416).  This header is
 * what a binding of that path binds instead.  Each entry point names the reference interface
 * it replaces.  INTEGRATION.md shows the ctypes stub a maintainer of the reference would add.
 *
 * Conventions
 *   - Plain C types only.  All bulk arrays are DEVICE pointers owned by the caller (the Python
 *     host side passes torch.Tensor.data_ptr()).  Small descriptors (options, tables, atol[n])
 *     are HOST pointers, copied into the launch.  The library allocates nothing persistent.
 *   - Layout is structure-of-arrays with the instance index innermost: component j of instance
 *     i lives at  a[j * ld + i]  (ld >= N, the leading dimension in instances), so that the 32
 *     lanes of a warp (32 neighbouring instances) load and store 256 contiguous bytes.
 *   - Work is queued on `stream` (a cudaStream_t passed as void*, NULL = legacy default
 *     stream); the call returns without synchronising.
 *   - Return value: 0 on success, otherwise a negative B200DAE_E* code; the message is
 *     available from b200dae_last_error().  Per-instance solver outcomes are NOT errors: they
 *     are written to status[] (the reference likewise reports them as DaeResult.status,
 *     dae.py:451-453, and not as exceptions).
 */
#ifndef B200DAE_H
#define B200DAE_H

#ifdef __CUDACC_RTC__            /* run-time compiled kernels: NVRTC has no <stdint.h> */
typedef int int32_t;
typedef long long int64_t;
#else
#include <stdint.h>
#endif

#ifdef __cplusplus
extern "C" {
#endif

#define B200DAE_VERSION 106
#define B200DAE_MAX_STAGES 7
#define B200DAE_MAX_ORDER 6
#define B200DAE_MAX_N 32
#define B200DAE_MAX_EVENTS 4

/* library-level error codes (return values) */
#define B200DAE_OK 0
#define B200DAE_EINVAL (-1)      /* bad argument (message says which) */
#define B200DAE_ENOPROBLEM (-2)  /* unknown problem id / (problem, stages) not compiled */
#define B200DAE_ECUDA (-3)       /* CUDA runtime error at launch */
#define B200DAE_ERTC (-4)        /* run-time compilation of a registered problem failed (message = NVRTC log) */

/* per-instance status[] values */
#define B200DAE_ST_FINISHED 0         /* reached t_bound           (solve_dae status 0)  */
#define B200DAE_ST_TOO_SMALL_STEP (-1)/* DaeSolver.TOO_SMALL_STEP  (base.py:158; solve_dae status -1) */
#define B200DAE_ST_NONFINITE (-2)     /* non-finite iteration matrix / rhs: the reference raises
                                         ValueError from scipy lu_factor/lu_solve(check_finite=True),
                                         base.py:234,238 */
#define B200DAE_ST_MAX_STEPS (-3)     /* opts.max_steps accepted steps done, t_bound not reached
                                         (resumable; no reference equivalent) */
#define B200DAE_ST_RUNNING 1          /* only ever seen in a state block between resumed launches */
#define B200DAE_ST_EVENT 2            /* a terminal event stopped the integration (solve_dae status 1,
                                         "A termination event occurred.", dae.py:484-487) */

/* rows of counters[8][ld] (int32).  nfev/njev/nlu are DaeSolver.nfev/njev/nlu (base.py:180-182),
 * nsolve is DaeSolver._nlusove (base.py:183); the rest are derived in the reference by wrapping
 * solve_collocation_system / solve_bdf_system. */
enum {
    B200DAE_CNT_NFEV = 0,
    B200DAE_CNT_NJEV = 1,
    B200DAE_CNT_NLU = 2,
    B200DAE_CNT_NSOLVE = 3,
    B200DAE_CNT_ACCEPTED = 4,
    B200DAE_CNT_ERR_REJECTED = 5,
    B200DAE_CNT_NEWTON_FAIL = 6,
    B200DAE_CNT_NEWTON_ITER = 7,
    B200DAE_NUM_COUNTERS = 8
};

/* Radau IIA method tables = the return value of radau_constants(s) (radau.py:51-134), computed
 * on the host in float64 with the same NumPy/SciPy primitives (scipy_dae_b200/constants.py).
 * Matrices are row-major with stride s.  The kernels rely on the normalisation of the reference's
 * eigen-transformation (radau.py:72-84): T[s-1] = (1, 0, 1, 0, ..., 1) and no zero in TI[0];
 * b200dae_solve_radau returns B200DAE_EINVAL for tables that do not have it. */
typedef struct b200dae_radau_tables {
    int32_t s;
    int32_t order;
    double A[B200DAE_MAX_STAGES * B200DAE_MAX_STAGES];
    double C[B200DAE_MAX_STAGES];
    double T[B200DAE_MAX_STAGES * B200DAE_MAX_STAGES];
    double TI[B200DAE_MAX_STAGES * B200DAE_MAX_STAGES];
    double P[B200DAE_MAX_STAGES * B200DAE_MAX_STAGES];
    double P2row0[B200DAE_MAX_STAGES];
    double v_implicit[B200DAE_MAX_STAGES];
    double v_explicit[B200DAE_MAX_STAGES];
    double b_hat_implicit[B200DAE_MAX_STAGES];
    double mu_real;
    double mu_re[B200DAE_MAX_STAGES / 2];
    double mu_im[B200DAE_MAX_STAGES / 2];
    double b_hat1_implicit;
} b200dae_radau_tables;

/* keyword arguments of RadauDAE.__init__ (radau.py:393-399) */
typedef struct b200dae_radau_opts {
    double rtol;
    double max_step;                 /* +inf: unbounded */
    double first_step;               /* <= 0: select_initial_step (common.py:129-183) */
    double continuous_error_weight;  /* 0 */
    double jac_recompute_rate;       /* 1e-3 */
    double deadband_lo, deadband_hi; /* controller_deadband = (1.0, 1.2) */
    int32_t newton_max_iter;         /* <= 0: 7 + int((s-3)*2.5) */
    int32_t newton_iter_embedded;    /* 1 */
    int64_t max_steps;               /* <= 0: no cap; otherwise stop after this many accepted steps */
    int32_t analytic_jac;            /* 0: jac=None, finite differences (base.py:250-297);
                                        1: the functor's analytic Jacobians = a `jac` callable
                                        (base.py:299-331); compiled for stages = 3 */
    int32_t io_layout;               /* 0: y0 / yp0 [n][ld], params [p][ld], y_final / yp_final [n][ld],
                                        counters [8][ld] (instance index innermost);
                                        1: instance-major, the reference's own array shapes: y0 / yp0 /
                                        y_final / yp_final [ld][n], params [ld][p], counters [ld][8].
                                        With 1 the pointers may be page-locked HOST arrays (cudaHostAlloc /
                                        cudaHostRegister, see b200dae_host_device_pointer): the launch then
                                        reads its inputs and writes its results over PCIe in place, no
                                        staging copies.  t_final, status, t_eval rows and traces are unaffected */
} b200dae_radau_opts;

/* keyword arguments of BDFDAE.__init__ (bdf.py:224-227) */
typedef struct b200dae_bdf_opts {
    double rtol;
    double max_step;
    double first_step;
    int32_t max_order;               /* 1..6, default 5 */
    int32_t ndf_strategy;            /* 0 "stability", 1 "efficiency", 2 plain BDF (bdf.py:242-247) */
    int64_t max_steps;
    int32_t analytic_jac;            /* as in b200dae_radau_opts */
    int32_t io_layout;               /* as in b200dae_radau_opts */
} b200dae_bdf_opts;

/* Events of solve_dae (dae.py:30-109, 426-490).  A Python event callable cannot run inside the
 * device step loop; the device offers THRESHOLD events g_k(t, y, y') = x[component_k] - value_k,
 * x = y (of_yp = 0) or y' (of_yp = 1), and the event functions registered with the problem's device
 * functor, g_k = event_{component_k}(t, y, y'; p) - value_k (of_yp = 2; b200dae_problem_num_events
 * says how many a problem has), with the attributes of a solve_dae event function:
 * direction (> 0 only - to +, < 0 only + to -, 0 both) and terminal (max_events <= 0: never
 * terminal, 1: `terminal = True`, k: `terminal = k`, i.e. stop at the k-th occurrence).  Roots are
 * located on the step's dense output with Brent's method, xtol = rtol = 4 EPS (dae.py:30-55).
 * HOST struct; the output pointers inside are DEVICE:
 *   t_events [n_events][cap][ld], y_events / yp_events [n_events][cap][n][ld],
 *   n_found [n_events][ld] (occurrences seen; only the first `cap` are stored). */
typedef struct b200dae_event_spec {
    int32_t n_events;
    int32_t of_yp[B200DAE_MAX_EVENTS];
    int32_t component[B200DAE_MAX_EVENTS];
    double value[B200DAE_MAX_EVENTS];
    int32_t direction[B200DAE_MAX_EVENTS];
    int32_t max_events[B200DAE_MAX_EVENTS];
} b200dae_event_spec;

typedef struct b200dae_events {
    b200dae_event_spec spec;
    int64_t cap;
    double *t_events, *y_events, *yp_events;
    int32_t *n_found;
} b200dae_events;

/* Output block shared by both integrators.  Any pointer may be NULL (that output is skipped),
 * except y_final/t_final/status.
 *   y_final, yp_final [n][ld]   state at t_final            (solver.y / solver.yp after the loop)
 *   t_final           [ld]      solver.t
 *   status            [ld]      B200DAE_ST_*
 *   counters          [8][ld]   see enum above
 *   t_eval            [m]       DEVICE, sorted in the direction of integration (dae.py:384-403)
 *   y_eval, yp_eval   [m][n][ld] dense output at t_eval (dae.py:496-515; Radau: radau.py:681-699,
 *                                BDF: bdf.py:525-548); rows >= n_eval_valid[i] are left untouched.
 *                                Entry (row r, component j, instance i) lives at
 *                                r * eval_stride_row + j * eval_stride_comp + i * eval_stride_inst
 *                                (in doubles); all three 0 = the default [m][n][ld] (n * ld, ld, 1).
 *                                The instance-major form [ld][m][n] is (n, 1, m * n): every
 *                                instance's (m, n) block contiguous; the default is the faster
 *                                one on the device (a warp writes a row as 256-byte segments)
 *   n_eval_valid      [ld]      number of t_eval points produced (all m unless the instance failed:
 *                               solve_dae breaks before recording a failed step, dae.py:451-453)
 *   t_steps           [steps_cap][ld]     accepted step ends incl. t0 (what solve_dae returns as
 *   y_steps, yp_steps [steps_cap][n][ld]   DaeResult.t/y/yp when t_eval is None, dae.py:492-495)
 *   n_steps           [ld]      number of rows written to the step trace
 *
 * Dense-output trace (solve_dae(dense_output=True) -> DaeSolution, common.py:9-126, and the
 * DaeSolver.dense_output() of every step, base.py:394-411): with dense_coef != NULL row k >= 1 of
 * the step trace also carries the data of the reference's per-step interpolant object,
 *   h_steps     [steps_cap][2][ld]  [0] the interpolant's h: Radau t - t_old (RadauDenseOutput,
 *                                   radau.py:672-679), BDF h_abs * direction AFTER the step's
 *                                   rescale (bdf.py:417-419); [1] the step end the interpolant is
 *                                   anchored at (differs from t_steps[k] only when a terminal event
 *                                   cut the step short, dae.py:484-487)
 *   order_steps [steps_cap][ld]     BDF: order of the interpolant (bdf.py:418); Radau: s
 *   dense_coef  [steps_cap][coef_rows][ld]
 *               Radau: ZP = (Y - y_old).T @ P, entry (j, k) in row j * s + k, coef_rows = n * s
 *                      (radau.py:661-666; y(t) = y_old + ZP @ x^(1..s), x = (t - t_old) / h)
 *               BDF:   D[:order + 1], entry (i, j) in row i * n + j, coef_rows = 7 * n (rows of
 *                      orders above order_steps[k] are left untouched) (bdf.py:525-548)
 */
typedef struct b200dae_outputs {
    double *y_final, *yp_final, *t_final;
    int32_t *status;
    int32_t *counters;
    const double *t_eval;
    int64_t m;
    double *y_eval, *yp_eval;
    int32_t *n_eval_valid;
    int64_t steps_cap;
    double *t_steps, *y_steps, *yp_steps;
    int32_t *n_steps;
    double *h_steps;
    int32_t *order_steps;
    double *dense_coef;
    int64_t coef_rows;
    const b200dae_events *events;   /* HOST pointer, NULL = no events */
    int64_t eval_stride_row, eval_stride_comp, eval_stride_inst;   /* layout of y_eval / yp_eval, see above */
    int32_t *bad_input;             /* optional, ONE int (device or page-locked host): set to 1 by the lean Radau
                                       kernel in io_layout = 1 launches when an instance stops with NONFINITE before
                                       its first accepted step (non-finite y0 / yp0 / parameters: what the reference
                                       rejects in check_arguments, base.py:33-37); the host then looks at that one
                                       word instead of scanning status[] */
} b200dae_outputs;

int b200dae_version(void);
const char *b200dae_last_error(void);

/* Problem registry: residuals F(t, y, y'; p) compiled into the library as device functors
 * (replaces the Python callable `fun(t, y, yp, *args)`, base.py:44-45, dae.py:365-383).
 * Returns the problem id (>= 0) or B200DAE_ENOPROBLEM. */
int b200dae_num_problems(void);
int b200dae_problem_lookup(const char *name, int32_t *n, int32_t *n_params);
int b200dae_problem_num_events(int problem_id);   /* event functions g(t, y, y'; p) of the functor */
const char *b200dae_problem_name(int problem_id);

/* Register a residual at RUN time: the reference takes any Python callable fun(t, y, yp, *args)
 * (base.py:44-45, dae.py:365-383); here `cuda_src` is CUDA C++ source text that defines
 *
 *     struct Problem {
 *         static constexpr int N = <n>, NP = <n_params>;
 *         static constexpr bool kHasJac = false;            // true: also provide jac(), see problems.cuh
 *         __device__ static void eval(double t, const double* y, const double* yp,
 *                                     const double* p, double* out);       // out[N] = F(t, y, y'; p)
 *         // optional: kNumEvents + static __device__ double event(int which, t, y, yp, p)
 *     };
 *
 * It is compiled with NVRTC against the library's own kernel templates (csrc/ next to the
 * library), one kernel variant at a time on first use, cubins cached under B200DAE_RT_CACHE
 * (default ~/.cache/scipy_dae_b200/rt).  Returns the problem id (usable with every entry point
 * below), B200DAE_ERTC with the compiler's message in b200dae_last_error(), or B200DAE_EINVAL.
 * Registering the same name with the same source again returns the same id. */
int b200dae_register_problem(const char *name, const char *cuda_src, int32_t n, int32_t n_params);
/* Compile (no device needed) the kernel a solve with these settings would launch:
 * method 0 = Radau (stages), 1 = BDF; with_t_eval / analytic_jac select the variant. */
int b200dae_precompile(int problem_id, int32_t method, int32_t stages, int32_t with_t_eval, int32_t analytic_jac);

/* Evaluate the registered residual on the device for N instances (SoA, ld): out[n][ld].
 * Lets the host side verify that a device functor matches its NumPy twin. */
int b200dae_residual(int problem_id, int64_t N, int64_t ld, double t,
                     const double *y, const double *yp, const double *params,
                     double *out, void *stream);

/* solve_dae(fun, (t0, t_bound), y0, yp0, method="Radau", rtol, atol, t_eval, **opts) for N
 * instances at once: replaces RadauDAE.__init__ + the `while status is None: solver.step()` loop
 * of solve_dae (dae.py:405-518) + RadauDAE._step_impl (radau.py:465-659) and everything below it
 * (solve_collocation_system radau.py:137-195, num_jac, LU, predict_factor, dense output).
 *   y0, yp0 [n][ld], params [n_params][ld] : DEVICE;  atol[n], opts, tb : HOST. */
int b200dae_solve_radau(int problem_id, int64_t N, int64_t ld,
                        const double *y0, const double *yp0, const double *params,
                        double t0, double t_bound, const double *atol,
                        const b200dae_radau_opts *opts, const b200dae_radau_tables *tb,
                        const b200dae_outputs *out, void *stream);

/* Same for method="BDF": BDFDAE.__init__ (bdf.py:224-261), BDFDAE._step_impl (bdf.py:263-415),
 * solve_bdf_system (bdf.py:32-64), change_D (bdf.py:14-29), BdfDenseOutput (bdf.py:422-548). */
int b200dae_solve_bdf(int problem_id, int64_t N, int64_t ld,
                      const double *y0, const double *yp0, const double *params,
                      double t0, double t_bound, const double *atol,
                      const b200dae_bdf_opts *opts,
                      const b200dae_outputs *out, void *stream);

/* consistent_initial_conditions(fun, t0, y0, yp0, jac, fixed_y0, fixed_yp0, rtol, atol,
 * newton_maxiter, chord_iter, safety) for N instances at once (common.py:186-258 with
 * solve_underdetermined_system :269-372 and qrank :261-266).  y0 / yp0 [n][ld] are DEVICE arrays
 * updated in place; atol[n], fixed_y[n], fixed_yp[n] (0 / 1 masks, NULL = nothing fixed) and opts
 * are HOST.  Per-instance outcome in status[]: the reference raises RuntimeError("Convergence
 * failed.") / ValueError("Too many fixed components, ...") / ValueError("Index greater than
 * one.") where this writes NOCONV / TOO_MANY_FIXED / INDEX_GT_ONE.  f_out [n][ld] (may be NULL)
 * receives the final residual, n_iter [ld] (may be NULL) the number of chord iterations. */
#define B200DAE_CIC_OK 0
#define B200DAE_CIC_NOCONV (-1)
#define B200DAE_CIC_TOO_MANY_FIXED (-2)
#define B200DAE_CIC_INDEX_GT_ONE (-3)
typedef struct b200dae_cic_opts {
    double rtol;                 /* 1e-8 */
    double safety;               /* 0.5, in (0, 1] */
    int32_t newton_maxiter;      /* 10 */
    int32_t chord_iter;          /* 3 */
    int32_t analytic_jac;        /* 0: approx_derivative("2-point") (jac=None); 1: functor Jacobians */
    int32_t reserved;
} b200dae_cic_opts;
int b200dae_consistent_ic(int problem_id, int64_t N, int64_t ld, double t0,
                          double *y0, double *yp0, const double *params, const double *atol,
                          const int32_t *fixed_y, const int32_t *fixed_yp,
                          const b200dae_cic_opts *opts,
                          double *f_out, int32_t *status, int32_t *n_iter, void *stream);

/* Device address of a page-locked host allocation (cudaHostGetDevicePointer), for io_layout = 1 launches
 * that work on host arrays in place.  On x86-64 Linux it equals the host address. */
int b200dae_host_device_pointer(void *host_ptr, void **device_ptr);

/* Number of kernels launched by this library since load (bench.py's "gpu_launches" claim). */
int64_t b200dae_launch_count(void);

/* FP64 FMA-pipe peak microbenchmark used as the roofline denominator for the FP64 bound:
 * every thread runs `iters` rounds of 8 independent dependent-chain DFMAs.  `sink` is a device
 * buffer of at least `blocks*threads` doubles.  flops = blocks*threads*iters*8*2. */
int b200dae_fp64_peak_kernel(int blocks, int threads, int64_t iters, double *sink, void *stream);

/* Self-test hook: rcp_out[i] = 1/x[i], sqrt_out[i] = sqrt(x[i]) computed with the library's
 * branch-free FP64 reciprocal / square root (csrc/fastmath.cuh), which stand in for the IEEE
 * divisions and sqrt of the reference's NumPy arithmetic.  All pointers DEVICE. */
int b200dae_selftest_fastmath(int64_t n, const double *x, double *rcp_out, double *sqrt_out, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* B200DAE_H */
