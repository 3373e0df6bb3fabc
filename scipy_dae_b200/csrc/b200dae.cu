// b200dae.cu — C-ABI of libb200dae.so (include/b200dae.h): argument checking, launch
// descriptors and the (problem, stages) -> kernel dispatch table.  No torch types, no allocation.
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <atomic>

#include "../../include/b200dae.h"
#include "launchers.cuh"
#include "rt_registry.h"

namespace b200dae {

static thread_local char g_err[512] = "";
std::atomic<int64_t> g_launches{0};

static int fail(int code, const char* fmt, const char* a = "", long long b = 0) {
    snprintf(g_err, sizeof(g_err), fmt, a, b);
    return code;
}

static const ProblemEntry* registry(int* count) {
    static const ProblemEntry table[] = {entry_robertson(), entry_akzo_nobel(), entry_transamp(), entry_andrews(),
                                         entry_rational(), entry_linear(), entry_cosine(), entry_weissinger(),
                                         entry_brenan(), entry_transamp1(), entry_blowup()};
    *count = (int)(sizeof(table) / sizeof(table[0]));
    return table;
}

// Problems registered at run time (rt_registry.cu) follow the built-in table: id = count + index.
// Their entries carry no launchers (rt >= 0 routes the call through NVRTC-compiled kernels).
static const ProblemEntry* find_problem(int id, int* rt = nullptr) {
    int cnt;
    const ProblemEntry* t = registry(&cnt);
    if (rt) *rt = -1;
    if (id >= 0 && id < cnt) return &t[id];
    static thread_local ProblemEntry dyn;
    const char* name;
    int n, np, nev;
    if (id >= cnt && rt_info(id - cnt, &name, &n, &np, &nev, nullptr)) {
        memset(&dyn, 0, sizeof(dyn));
        dyn.name = name;
        dyn.n = n;
        dyn.n_params = np;
        dyn.n_events = nev;
        if (rt) *rt = id - cnt;
        return &dyn;
    }
    return nullptr;
}

static int rt_result(int rc, const std::string& err, const ProblemEntry* e) {
    if (rc == 0) return 0;
    if (rc == kRtErrNoJac) return fail(B200DAE_ENOPROBLEM, "problem '%s' provides no analytic Jacobian (kHasJac)", e->name);
    snprintf(g_err, sizeof(g_err), "%s", err.c_str());
    return rc;
}

static int check_common(const ProblemEntry* e, int64_t N, int64_t ld, const double* y0,
                        const double* yp0, const double* params, const double* atol,
                        const b200dae_outputs* out) {
    if (!e) return fail(B200DAE_ENOPROBLEM, "unknown problem id%s (%lld)", "", 0);
    if (N < 0 || ld < N) return fail(B200DAE_EINVAL, "need 0 <= N <= ld%s (N=%lld)", "", (long long)N);
    if (N > 0 && (!y0 || !yp0 || (e->n_params > 0 && !params)))
        return fail(B200DAE_EINVAL, "y0 / yp0 / params must not be NULL%s", "");
    if (!atol) return fail(B200DAE_EINVAL, "atol[n] must not be NULL%s", "");
    for (int j = 0; j < e->n; ++j)
        if (!(atol[j] >= 0.0)) return fail(B200DAE_EINVAL, "`atol` must be positive.%s", "");
    if (!out || !out->y_final || !out->t_final || !out->status)
        return fail(B200DAE_EINVAL, "out->y_final / t_final / status must not be NULL%s", "");
    if (out->m < 0 || (out->m > 0 && (!out->t_eval || !out->y_eval)))
        return fail(B200DAE_EINVAL, "m > 0 needs t_eval and y_eval%s", "");
    if (out->eval_stride_row < 0 || out->eval_stride_comp < 0 || out->eval_stride_inst < 0)
        return fail(B200DAE_EINVAL, "y_eval strides must not be negative%s", "");
    if (out->steps_cap < 0 || (out->steps_cap > 0 && !out->t_steps))
        return fail(B200DAE_EINVAL, "steps_cap > 0 needs t_steps%s", "");
    if (out->dense_coef && (out->steps_cap <= 0 || !out->h_steps || !out->y_steps))
        return fail(B200DAE_EINVAL, "dense_coef needs a step trace (steps_cap > 0, t_steps, y_steps) and h_steps%s", "");
    return 0;
}

static int check_events(const ProblemEntry* e, const b200dae_outputs* out, EventLaunch* ev) {
    memset(ev, 0, sizeof(*ev));
    const b200dae_events* s = out->events;
    if (!s || s->spec.n_events == 0) return 0;
    if (s->spec.n_events < 0 || s->spec.n_events > B200DAE_MAX_EVENTS)
        return fail(B200DAE_EINVAL, "need 0 <= n_events <= 4%s", "");
    for (int k = 0; k < s->spec.n_events; ++k) {
        const int kind = s->spec.of_yp[k];
        if (kind < 0 || kind > 2) return fail(B200DAE_EINVAL, "event kind must be 0 (y), 1 (y') or 2 (functor event)%s", "");
        if (s->spec.component[k] < 0 || s->spec.component[k] >= (kind == 2 ? e->n_events : e->n))
            return fail(B200DAE_EINVAL, kind == 2 ? "problem '%s' has no such event function" : "event component out of range%s", e->name);
    }
    if (s->cap <= 0 || !s->t_events || !s->n_found)
        return fail(B200DAE_EINVAL, "events need cap > 0, t_events and n_found%s", "");
    ev->spec = s->spec;
    ev->cap = s->cap;
    ev->t_events = s->t_events;
    ev->y_events = s->y_events;
    ev->yp_events = s->yp_events;
    ev->n_found = s->n_found;
    return 0;
}

static b200dae_outputs normalise(const b200dae_outputs* out, int n, int64_t ld) {
    b200dae_outputs o = *out;
    o.events = nullptr;
    if (o.m == 0) { o.t_eval = nullptr; o.y_eval = nullptr; o.yp_eval = nullptr; }
    if (o.eval_stride_row == 0 && o.eval_stride_comp == 0 && o.eval_stride_inst == 0) {
        o.eval_stride_row = (int64_t)n * ld;              // the default [m][n][ld]
        o.eval_stride_comp = ld;
        o.eval_stride_inst = 1;
    }
    if (o.steps_cap == 0) { o.t_steps = nullptr; o.y_steps = nullptr; o.yp_steps = nullptr; o.dense_coef = nullptr; }
    if (o.dense_coef == nullptr) { o.h_steps = nullptr; o.order_steps = nullptr; }
    return o;
}

// compute_R(order = 6, factor = 1) (bdf.py:14-21); leading blocks serve every lower order
static void fill_U(double* Um) {
    const int q = kBdfQ;
    for (int j = 0; j < q; ++j) Um[j] = 1.0;
    for (int i = 1; i < q; ++i)
        for (int j = 0; j < q; ++j) {
            const double mij = (j == 0) ? 0.0 : ((double)(i - 1) - 1.0 * j) / i;
            Um[i * q + j] = Um[(i - 1) * q + j] * mij;
        }
}

__global__ void fp64_peak_kernel(int64_t iters, double* __restrict__ sink) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    double a0 = 1.0 + 1e-9 * threadIdx.x, a1 = a0 + 0.1, a2 = a0 + 0.2, a3 = a0 + 0.3, a4 = a0 + 0.4,
           a5 = a0 + 0.5, a6 = a0 + 0.6, a7 = a0 + 0.7;
    const double m = 0.9999999, c = 1e-7;
    for (int64_t k = 0; k < iters; ++k) {
        a0 = fma(a0, m, c); a1 = fma(a1, m, c); a2 = fma(a2, m, c); a3 = fma(a3, m, c);
        a4 = fma(a4, m, c); a5 = fma(a5, m, c); a6 = fma(a6, m, c); a7 = fma(a7, m, c);
    }
    sink[i] = ((a0 + a1) + (a2 + a3)) + ((a4 + a5) + (a6 + a7));
}

__global__ void fastmath_kernel(int64_t n, const double* __restrict__ x, double* __restrict__ rcp,
                                double* __restrict__ sq) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    rcp[i] = fast_rcp(x[i]);
    sq[i] = fast_sqrt(x[i]);
}

}  // namespace b200dae

using namespace b200dae;

extern "C" {

int b200dae_version(void) { return B200DAE_VERSION; }

const char* b200dae_last_error(void) { return g_err; }

int b200dae_num_problems(void) {
    int cnt;
    registry(&cnt);
    return cnt + rt_count();
}

int b200dae_problem_num_events(int problem_id) {
    const ProblemEntry* e = find_problem(problem_id);
    return e ? e->n_events : B200DAE_ENOPROBLEM;
}

int b200dae_problem_lookup(const char* name, int32_t* n, int32_t* n_params) {
    int cnt;
    const ProblemEntry* t = registry(&cnt);
    if (name)
        for (int i = 0; i < cnt; ++i)
            if (strcmp(t[i].name, name) == 0) {
                if (n) *n = t[i].n;
                if (n_params) *n_params = t[i].n_params;
                return i;
            }
    if (name)
        for (int i = 0; i < rt_count(); ++i) {
            const char* nm;
            int rn, rnp;
            if (rt_info(i, &nm, &rn, &rnp, nullptr, nullptr) && strcmp(nm, name) == 0) {
                if (n) *n = rn;
                if (n_params) *n_params = rnp;
                return cnt + i;
            }
        }
    return fail(B200DAE_ENOPROBLEM, "unknown problem '%s'", name ? name : "(null)");
}

const char* b200dae_problem_name(int problem_id) {
    const ProblemEntry* e = find_problem(problem_id);
    return e ? e->name : nullptr;
}

int64_t b200dae_launch_count(void) { return g_launches.load(); }

int b200dae_host_device_pointer(void* host_ptr, void** device_ptr) {
    if (!host_ptr || !device_ptr) return fail(B200DAE_EINVAL, "host_device_pointer needs two pointers%s", "");
    const cudaError_t rc = cudaHostGetDevicePointer(device_ptr, host_ptr, 0);
    if (rc != cudaSuccess) {
        (void)cudaGetLastError();
        return fail(B200DAE_ECUDA, "CUDA: %s (is the host allocation page-locked?)", cudaGetErrorString(rc));
    }
    return 0;
}

int b200dae_register_problem(const char* name, const char* cuda_src, int32_t n, int32_t n_params) {
    if (!name || !*name || !cuda_src) return fail(B200DAE_EINVAL, "register_problem needs a name and a source%s", "");
    for (const char* c = name; *c; ++c)
        if (!((*c >= 'a' && *c <= 'z') || (*c >= 'A' && *c <= 'Z') || (*c >= '0' && *c <= '9') || *c == '_'))
            return fail(B200DAE_EINVAL, "problem name '%s' must be an identifier", name);
    if (n < 1 || n > B200DAE_MAX_N || n_params < 0 || n_params > 64)
        return fail(B200DAE_EINVAL, "need 1 <= n <= 32 and 0 <= n_params <= 64%s (n=%lld)", "", (long long)n);
    int cnt;
    const ProblemEntry* t = registry(&cnt);
    for (int i = 0; i < cnt; ++i)
        if (strcmp(t[i].name, name) == 0) return fail(B200DAE_EINVAL, "'%s' is a built-in problem", name);
    std::string err;
    const int idx = rt_register(name, cuda_src, n, n_params, err);
    if (idx < 0) {
        snprintf(g_err, sizeof(g_err), "%s", err.c_str());
        return idx;
    }
    return cnt + idx;
}

int b200dae_precompile(int problem_id, int32_t method, int32_t stages, int32_t with_t_eval, int32_t analytic_jac) {
    int rt;
    const ProblemEntry* e = find_problem(problem_id, &rt);
    if (!e) return fail(B200DAE_ENOPROBLEM, "unknown problem id%s (%lld)", "", problem_id);
    if (rt < 0) return 0;                                 // built-in: compiled by nvcc
    static const double dummy = 0.0;
    std::string err;
    if (method == 0) {
        if (stages < 1 || stages > B200DAE_MAX_STAGES || stages % 2 != 1) return fail(B200DAE_EINVAL, "bad number of stages%s", "");
        RadauLaunch L;
        memset(&L, 0, sizeof(L));
        L.tb.s = stages;
        L.newton_iter_embedded = 1;
        L.newton_max_iter = 7 + (int)((stages - 3) * 2.5);
        L.analytic_jac = analytic_jac;
        if (with_t_eval) { L.out.t_eval = &dummy; L.out.m = 1; }
        return rt_result(rt_launch_radau(rt, L, nullptr, false, err), err, e);
    }
    BdfLaunch L;
    memset(&L, 0, sizeof(L));
    L.analytic_jac = analytic_jac;
    if (with_t_eval) { L.out.t_eval = &dummy; L.out.m = 1; }
    return rt_result(rt_launch_bdf(rt, L, nullptr, false, err), err, e);
}

int b200dae_residual(int problem_id, int64_t N, int64_t ld, double t, const double* y,
                     const double* yp, const double* params, double* out, void* stream) {
    int rt;
    const ProblemEntry* e = find_problem(problem_id, &rt);
    if (!e) return fail(B200DAE_ENOPROBLEM, "unknown problem id%s (%lld)", "", problem_id);
    if (N < 0 || ld < N || !y || !yp || !out) return fail(B200DAE_EINVAL, "bad residual arguments%s", "");
    if (N == 0) return 0;
    if (rt >= 0) {
        std::string err;
        return rt_result(rt_launch_residual(rt, N, ld, t, y, yp, params, out, (cudaStream_t)stream, err), err, e);
    }
    const int rc = e->residual(N, ld, t, y, yp, params, out, (cudaStream_t)stream);
    if (rc) return fail(B200DAE_ECUDA, "CUDA: %s", cudaGetErrorString((cudaError_t)rc));
    return 0;
}

int b200dae_solve_radau(int problem_id, int64_t N, int64_t ld, const double* y0, const double* yp0,
                        const double* params, double t0, double t_bound, const double* atol,
                        const b200dae_radau_opts* opts, const b200dae_radau_tables* tb,
                        const b200dae_outputs* out, void* stream) {
    int rt;
    const ProblemEntry* e = find_problem(problem_id, &rt);
    int rc = check_common(e, N, ld, y0, yp0, params, atol, out);
    if (rc) return rc;
    if (!opts || !tb) return fail(B200DAE_EINVAL, "opts / tables must not be NULL%s", "");
    const int s = tb->s;
    if (s < 1 || s > B200DAE_MAX_STAGES || s % 2 != 1 || (rt < 0 && !e->radau[s]))
        return fail(B200DAE_ENOPROBLEM, "Radau stages not compiled for problem '%s' (stages=%lld)", e->name, s);
    if (!(opts->max_step > 0)) return fail(B200DAE_EINVAL, "`max_step` must be positive.%s", "");
    if (opts->first_step > fabs(t_bound - t0)) return fail(B200DAE_EINVAL, "`first_step` exceeds bounds.%s", "");
    if (!(opts->continuous_error_weight >= 0 && opts->continuous_error_weight <= 1))
        return fail(B200DAE_EINVAL, "need 0 <= continuous_error_weight <= 1%s", "");
    if (!(opts->jac_recompute_rate > 0 && opts->jac_recompute_rate < 1))
        return fail(B200DAE_EINVAL, "need 0 < jac_recompute_rate < 1%s", "");
    if (!(opts->deadband_lo > 0 && opts->deadband_lo <= opts->deadband_hi))
        return fail(B200DAE_EINVAL, "need 0 < controller_deadband[0] <= controller_deadband[1]%s", "");
    if (opts->newton_iter_embedded < 0) return fail(B200DAE_EINVAL, "need newton_iter_embedded >= 0%s", "");
    if (opts->io_layout != 0 && opts->io_layout != 1) return fail(B200DAE_EINVAL, "io_layout must be 0 or 1%s", "");
    // the kernels evaluate rate ** (newton_max_iter - k) (radau.py:186) with a 5-bit square-and-multiply
    if (opts->newton_max_iter > 32) return fail(B200DAE_EINVAL, "need newton_max_iter <= 32%s (got %lld)", "", (long long)opts->newton_max_iter);
    // The thread-per-system kernel relies on the normalisation of the reference's eigen-transformation
    // (radau.py:72-84): eigenvectors scaled to a unit last row, so T[s-1] = (1, 0, 1, 0, ..., 1), and
    // on a dense first row of TI (a non-finite stage residual then shows in TI[0] @ F).
    for (int l = 0; l < s; ++l) {
        if (tb->T[(s - 1) * s + l] != ((l % 2 == 0) ? 1.0 : 0.0) || tb->TI[l] == 0.0)
            return fail(B200DAE_EINVAL, "Radau tables are not in the reference's normalisation (radau.py:72-84)%s", "");
    }
    if (out->dense_coef && out->coef_rows != (int64_t)e->n * s)
        return fail(B200DAE_EINVAL, "Radau dense_coef needs coef_rows = n * stages%s (got %lld)", "", (long long)out->coef_rows);
    if (N == 0) return 0;

    RadauLaunch L;
    memset(&L, 0, sizeof(L));
    L.tb = *tb;
    L.t0 = t0;
    L.t_bound = t_bound;
    L.dir = (t_bound != t0) ? ((t_bound > t0) ? 1.0 : -1.0) : 1.0;            // base.py:214
    L.rtol = fmax(opts->rtol, 100 * kEps);                                     // validate_tol
    L.max_step = opts->max_step;
    L.first_step = opts->first_step;
    L.cew = opts->continuous_error_weight;
    L.jac_rate = opts->jac_recompute_rate;
    L.db_lo = opts->deadband_lo;
    L.db_hi = opts->deadband_hi;
    {
        const double expmns = (s + 1.0) / (2.0 * s);                           // radau.py:422-433
        L.newton_tol = fmax(10 * kEps / L.rtol, fmin(0.03, pow(L.rtol, 1.0 / expmns - 1)));
    }
    L.inv_mu_real = 1.0 / tb->mu_real;
    for (int j = 0; j < e->n; ++j) { L.atol[j] = atol[j]; L.thr[j] = atol[j] / L.rtol; }
    L.newton_max_iter = opts->newton_max_iter > 0 ? opts->newton_max_iter
                                                  : 7 + (int)((s - 3) * 2.5);  // radau.py:440-441
    L.newton_iter_embedded = opts->newton_iter_embedded;
    L.max_steps = opts->max_steps;
    L.analytic_jac = opts->analytic_jac;
    L.io_aos = opts->io_layout == 1;
    L.N = N;
    L.ld = ld;
    L.y0 = y0;
    L.yp0 = yp0;
    L.params = params;
    L.out = normalise(out, e->n, ld);
    rc = check_events(e, out, &L.ev);
    if (rc) return rc;
    if (rt >= 0) {
        std::string err;
        return rt_result(rt_launch_radau(rt, L, (cudaStream_t)stream, true, err), err, e);
    }
    rc = e->radau[s](L, (cudaStream_t)stream);
    if (rc == kErrNoJac)
        return fail(B200DAE_ENOPROBLEM, "analytic Jacobian not compiled for problem '%s' (stages=%lld)", e->name, s);
    if (rc) return fail(B200DAE_ECUDA, "CUDA: %s", cudaGetErrorString((cudaError_t)rc));
    return 0;
}

int b200dae_solve_bdf(int problem_id, int64_t N, int64_t ld, const double* y0, const double* yp0,
                      const double* params, double t0, double t_bound, const double* atol,
                      const b200dae_bdf_opts* opts, const b200dae_outputs* out, void* stream) {
    int rt;
    const ProblemEntry* e = find_problem(problem_id, &rt);
    int rc = check_common(e, N, ld, y0, yp0, params, atol, out);
    if (rc) return rc;
    if (!opts) return fail(B200DAE_EINVAL, "opts must not be NULL%s", "");
    if (opts->max_order < 1 || opts->max_order > B200DAE_MAX_ORDER)
        return fail(B200DAE_EINVAL, "Ensure that 1 <= max_order <= 6.%s", "");
    if (!(opts->max_step > 0)) return fail(B200DAE_EINVAL, "`max_step` must be positive.%s", "");
    if (opts->first_step > fabs(t_bound - t0)) return fail(B200DAE_EINVAL, "`first_step` exceeds bounds.%s", "");
    if (opts->io_layout != 0 && opts->io_layout != 1) return fail(B200DAE_EINVAL, "io_layout must be 0 or 1%s", "");
    if (out->dense_coef && (out->coef_rows != (int64_t)e->n * kBdfQ || !out->order_steps))
        return fail(B200DAE_EINVAL, "BDF dense_coef needs order_steps and coef_rows = 7 * n%s (got %lld)", "", (long long)out->coef_rows);
    if (N == 0) return 0;

    BdfLaunch L;
    memset(&L, 0, sizeof(L));
    L.t0 = t0;
    L.t_bound = t_bound;
    L.dir = (t_bound != t0) ? ((t_bound > t0) ? 1.0 : -1.0) : 1.0;
    L.rtol = fmax(opts->rtol, 100 * kEps);
    L.max_step = opts->max_step;
    L.first_step = opts->first_step;
    L.newton_tol = fmax(10 * kEps / L.rtol, fmin(0.03, sqrt(L.rtol)));         // bdf.py:234
    for (int j = 0; j < e->n; ++j) { L.atol[j] = atol[j]; L.thr[j] = atol[j] / L.rtol; }
    {
        static const double k_stab[7] = {0, -37.0 / 200, -1.0 / 9, 0.0834, 0.0665, 0.0551, 0.0464};
        static const double k_eff[7] = {0, -37.0 / 200, -1.0 / 9, -0.0823, -0.0415, 0, 0};
        static const double k_bdf[7] = {0, -37.0 / 200, -1.0 / 9, 0, 0, 0, 0};
        const double* kappa = opts->ndf_strategy == 0 ? k_stab : (opts->ndf_strategy == 1 ? k_eff : k_bdf);
        double g = 0.0;
        for (int k = 1; k <= opts->max_order; ++k) { g += 1.0 / k; L.gamma[k] = g; }   // bdf.py:250-252
        for (int k = 0; k <= opts->max_order; ++k) {
            L.alpha[k] = (1 - kappa[k]) * L.gamma[k];
            L.error_const[k] = kappa[k] * L.gamma[k] + 1.0 / (k + 1);
        }
    }
    fill_U(L.U);
    for (int k = 1; k < kBdfQ + 2; ++k) L.neg_inv[k] = -1.0 / k;
    L.max_order = opts->max_order;
    L.max_steps = opts->max_steps;
    L.analytic_jac = opts->analytic_jac;
    L.io_aos = opts->io_layout == 1;
    L.N = N;
    L.ld = ld;
    L.y0 = y0;
    L.yp0 = yp0;
    L.params = params;
    L.out = normalise(out, e->n, ld);
    rc = check_events(e, out, &L.ev);
    if (rc) return rc;
    if (rt >= 0) {
        std::string err;
        return rt_result(rt_launch_bdf(rt, L, (cudaStream_t)stream, true, err), err, e);
    }
    rc = e->bdf(L, (cudaStream_t)stream);
    if (rc == kErrNoJac)
        return fail(B200DAE_ENOPROBLEM, "analytic Jacobian not compiled for problem '%s'", e->name);
    if (rc) return fail(B200DAE_ECUDA, "CUDA: %s", cudaGetErrorString((cudaError_t)rc));
    return 0;
}

int b200dae_consistent_ic(int problem_id, int64_t N, int64_t ld, double t0, double* y0, double* yp0,
                          const double* params, const double* atol, const int32_t* fixed_y,
                          const int32_t* fixed_yp, const b200dae_cic_opts* opts, double* f_out,
                          int32_t* status, int32_t* n_iter, void* stream) {
    int rt;
    const ProblemEntry* e = find_problem(problem_id, &rt);
    if (!e) return fail(B200DAE_ENOPROBLEM, "unknown problem id%s (%lld)", "", problem_id);
    if (N < 0 || ld < N) return fail(B200DAE_EINVAL, "need 0 <= N <= ld%s (N=%lld)", "", (long long)N);
    if (N > 0 && (!y0 || !yp0 || !status || (e->n_params > 0 && !params)))
        return fail(B200DAE_EINVAL, "y0 / yp0 / params / status must not be NULL%s", "");
    if (!atol || !opts) return fail(B200DAE_EINVAL, "atol[n] and opts must not be NULL%s", "");
    if (!(opts->rtol > 0)) return fail(B200DAE_EINVAL, "Relative tolerance must be a positive scalar.%s", "");
    if (!(opts->safety > 0 && opts->safety <= 1)) return fail(B200DAE_EINVAL, "safety factor has to be in (0, 1]%s", "");
    if (opts->newton_maxiter < 0 || opts->chord_iter < 0) return fail(B200DAE_EINVAL, "iteration counts must be >= 0%s", "");
    CicLaunch L;
    memset(&L, 0, sizeof(L));
    for (int j = 0; j < e->n; ++j) {
        if (!(atol[j] > 0)) return fail(B200DAE_EINVAL, "Absolute tolerance must be positive.%s", "");
        L.atol[j] = atol[j];
        L.fixed_y[j] = fixed_y ? (fixed_y[j] != 0) : 0;
        L.fixed_yp[j] = fixed_yp ? (fixed_yp[j] != 0) : 0;
    }
    if (N == 0) return 0;
    L.t0 = t0;
    L.rtol = opts->rtol;
    L.safety = opts->safety;
    L.newton_maxiter = opts->newton_maxiter;
    L.chord_iter = opts->chord_iter;
    L.analytic_jac = opts->analytic_jac;
    L.N = N;
    L.ld = ld;
    L.y0 = y0;
    L.yp0 = yp0;
    L.params = params;
    L.f_out = f_out;
    L.status = status;
    L.n_iter = n_iter;
    if (rt >= 0) {
        std::string err;
        return rt_result(rt_launch_cic(rt, L, (cudaStream_t)stream, true, err), err, e);
    }
    const int rc = e->cic(L, (cudaStream_t)stream);
    if (rc == kErrNoJac)
        return fail(B200DAE_ENOPROBLEM, "analytic Jacobian not compiled for problem '%s'", e->name);
    if (rc) return fail(B200DAE_ECUDA, "CUDA: %s", cudaGetErrorString((cudaError_t)rc));
    return 0;
}

int b200dae_selftest_fastmath(int64_t n, const double* x, double* rcp_out, double* sqrt_out, void* stream) {
    if (n < 0 || !x || !rcp_out || !sqrt_out) return fail(B200DAE_EINVAL, "bad selftest arguments%s", "");
    if (n == 0) return 0;
    fastmath_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(n, x, rcp_out, sqrt_out);
    g_launches.fetch_add(1);
    const cudaError_t rc = cudaGetLastError();
    if (rc != cudaSuccess) return fail(B200DAE_ECUDA, "CUDA: %s", cudaGetErrorString(rc));
    return 0;
}

int b200dae_fp64_peak_kernel(int blocks, int threads, int64_t iters, double* sink, void* stream) {
    if (blocks <= 0 || threads <= 0 || threads > 1024 || !sink)
        return fail(B200DAE_EINVAL, "bad fp64 peak arguments%s", "");
    fp64_peak_kernel<<<blocks, threads, 0, (cudaStream_t)stream>>>(iters, sink);
    g_launches.fetch_add(1);
    const cudaError_t rc = cudaGetLastError();
    if (rc != cudaSuccess) return fail(B200DAE_ECUDA, "CUDA: %s", cudaGetErrorString(rc));
    return 0;
}

}  // extern "C"
