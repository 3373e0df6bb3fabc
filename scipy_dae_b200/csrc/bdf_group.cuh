// bdf_group.cuh — variable-order BDF/NDF integration with G cooperating lanes per system.
//
// Same replacement scope and flattened control flow as bdf.cuh (BDFDAE._step_impl bdf.py:263-415,
// solve_bdf_system bdf.py:32-64, change_D / compute_R bdf.py:14-29, BdfDenseOutput
// bdf.py:422-548, solve_dae's loop dae.py:446-518) for n > 4.  Lane j owns component j: its
// column of the backward-difference array D lives in registers (rows beyond the current order
// are masked, not indexed), norms are butterfly all-reduces, the LU and the triangular solves
// are the row-per-lane routines of group.cuh, the FD Jacobian runs its columns on n lanes at
// once (radau_group.cuh: fd_jacobian_group).
#pragma once
#include "rt_compat.cuh"
#include "../../include/b200dae.h"
#include "bdf.cuh"
#include "radau_group.cuh"

namespace b200dae {

template <int N, int G>
struct BdfGroupLayout {
    static constexpr int LD = N | 1;
    static constexpr int kJ = N * LD;
    static constexpr int kLU = N * LD + N + ((N * LD + N) & 1);
    static constexpr int kFd = 3 * G * LD + ((3 * G * LD) & 1);
    static constexpr int oJy = 0, oJyp = kJ, oLU = 2 * kJ, oWork = oLU + kLU;
    static constexpr int oY = oWork + kFd, oYp = oY + LD, oF = oYp + LD;
    static constexpr int oIdx = oF + LD + ((3 * LD) & 1);
    static constexpr int oPar = oIdx + (G + 1) / 2;
    __host__ __device__ static constexpr int doubles(int np) { return oPar + np + ((oPar + np) & 1); }
};

// change_D (bdf.py:24-29) for one component: D[:order+1] <- (R U)^T D[:order+1], evaluated as
// U^T (R^T D) with R generated column by column (bdf.py:14-21).
// Not inlined: six call sites; the array travels through the local stack (9 loads + 9 stores).
// (Funnelling the five pre-attempt calls through one site, 170 instructions less, measured
// neutral: C3 +1 %, C5 -2.5 %.)
// The outer loops are ROLLED (the inner ones unrolled over registers): fully unrolled the body was
// 690 of the kernel's 4600 instructions, and this kernel is instruction-fetch bound.  Same
// operations in the same order as the unrolled form.
static __device__ __noinline__ void change_D_lane(double* __restrict__ D, int order, double factor,
                                           const double* __restrict__ Umat) {
    const int q = order + 1;
    double Dl[kBdfQ], E[kBdfQ];
#pragma unroll
    for (int l = 0; l < kBdfQ; ++l) Dl[l] = D[l];
#pragma unroll 1
    for (int j = 0; j < kBdfQ; ++j) {
        double acc = Dl[0];
        double r = 1.0;
#pragma unroll
        for (int l = 1; l < kBdfQ; ++l) {
            r *= (j == 0) ? 0.0 : ((l - 1) - factor * j) * (1.0 / l);
            if (l < q) acc = fma(r, Dl[l], acc);
        }
        E[j] = acc;
    }
    double El[kBdfQ];
#pragma unroll
    for (int l = 0; l < kBdfQ; ++l) El[l] = E[l];
#pragma unroll 1
    for (int i = 0; i < kBdfQ; ++i) {
        double acc = 0.0;
#pragma unroll
        for (int l = 0; l < kBdfQ; ++l)
            if (l <= i && l < q) acc = fma(Umat[l * kBdfQ + i], El[l], acc);   // U is upper triangular: rows l > i add exact zeros
        if (i < q) D[i] = acc;
    }
}

__device__ __forceinline__ void change_D_call(double (&D)[kBdfRows], int order, double factor,
                                              const double* __restrict__ Umat) {
    double tmp[kBdfRows];
#pragma unroll
    for (int i = 0; i < kBdfRows; ++i) tmp[i] = D[i];
    change_D_lane(tmp, order, factor, Umat);
#pragma unroll
    for (int i = 0; i < kBdfRows; ++i) D[i] = tmp[i];
}

// kOut: t_eval output and step trace compiled in (false: the lean variant of a plain solve)
template <class P, int G, int kThreads, int kMinBlocks, bool kJac = false, bool kEv = false, bool kOut = true>
__global__ void __launch_bounds__(kThreads, kMinBlocks)
bdf_group_kernel(const __grid_constant__ BdfLaunch L) {
    constexpr int N = P::N, NP = P::NP;
    using Lay = BdfGroupLayout<N, G>;
    constexpr int LD = Lay::LD;
    static_assert(G >= N, "group must cover the components");
    extern __shared__ __align__(16) double smem[];

    const Group<G> g;
    const int lane = g.lane;
    const bool act = lane < N;
    const int64_t inst = ((int64_t)blockIdx.x * kThreads + threadIdx.x) / G;
    if (inst >= L.N) return;
    const int64_t ld = L.ld;

    double* sm = smem + (size_t)(threadIdx.x / G) * Lay::doubles(NP);
    double* sJy = sm + Lay::oJy;
    double* sJyp = sm + Lay::oJyp;
    double* sWork = sm + Lay::oWork;
    double* sY = sm + Lay::oY;
    double* sYp = sm + Lay::oYp;
    double* sF = sm + Lay::oF;
    double* sLU = sm + Lay::oLU;
    int* sIdx = reinterpret_cast<int*>(sm + Lay::oIdx);

    double* p = sm + Lay::oPar;
    for (int j = lane; j < NP; j += G) p[j] = L.params[io_index(L.io_aos, ld, NP, j, inst)];
    g.sync();
    const int jc = act ? lane : 0;
    double y = act ? L.y0[io_index(L.io_aos, ld, N, jc, inst)] : 0.0;
    double yp = act ? L.yp0[io_index(L.io_aos, ld, N, jc, inst)] : 0.0;
    const double atol_j = L.atol[jc], thr_j = L.thr[jc];
    double fac_y = kNjInitFactor, fac_yp = kNjInitFactor;

    int perm = lane;
    const double* rrow = sLU + jc * LD;

    const double dir = L.dir, rtol = L.rtol;
    double t = L.t0;
    double h_abs;
    if (L.first_step > 0.0) {
        h_abs = L.first_step;
    } else {                                               // select_initial_step, common.py:129-183
        const double v = act ? fabs(yp * fast_rcp(fmax(fabs(y), thr_j))) : 0.0;
        double e = g.max(v);
        if (g.any(v != v)) e = nan("");
        const double rh = e * fast_rcp(fast_sqrt(rtol) * 0.8);
        h_abs = fmin(L.max_step, fabs(L.t_bound - L.t0));
        if (h_abs * rh > 1.0) h_abs = fast_rcp(rh);
        h_abs = fmax(h_abs, 16.0 * kEps * fabs(L.t0));
    }
    double D[kBdfRows];
#pragma unroll
    for (int i = 0; i < kBdfRows; ++i) D[i] = 0.0;         // bdf.py:254-257
    D[0] = y;
    D[1] = yp * h_abs * dir;
    int order = 1, n_equal_steps = 0;
    const int max_order = L.max_order;

    bool need_jac = true, lu_valid = false, step_begin = true, retry = false, current_jac = false;
    int32_t nfev = 0, njev = 0, nlu = 0, nsolve = 0, nacc = 0, nrej = 0, nfail = 0, niter = 0;
    int32_t status = B200DAE_ST_RUNNING;
    int64_t ie = 0;
    int32_t nst = 0;

    if (kOut && L.out.t_steps != nullptr && L.out.steps_cap > 0) {
        if (lane == 0) L.out.t_steps[inst] = t;
        if (act) {
            if (L.out.y_steps) L.out.y_steps[lane * ld + inst] = y;
            if (L.out.yp_steps) L.out.yp_steps[lane * ld + inst] = yp;
        }
        nst = 1;
    }
    if (L.t0 == L.t_bound) { status = B200DAE_ST_FINISHED; njev = 1; }
    EventState evs;                  // g = event(t0, y0, yp0), dae.py:436 (identical in every lane)
    if (kEv) for (int k = 0; k < B200DAE_MAX_EVENTS; ++k) {
        evs.count[k] = evs.found[k] = 0;
        evs.g[k] = (k < L.ev.spec.n_events)
            ? event_quantity_group<P, N>(g, L.ev.spec, k, t, y, yp, p) - L.ev.spec.value[k] : 0.0;
    }

    double hl = 0.0, min_step = 0.0, h = 0.0, t_new = 0.0, c = 0.0;
    double y_predict = 0.0, psi = 0.0, inv_scale = 0.0, y_new = 0.0, yp_new = 0.0, d = 0.0;

    while (status == B200DAE_ST_RUNNING) {
        // ---- phase J (base.py:297, bdf.py:330) ------------------------------------------------
        if (need_jac) {
            if constexpr (kJac) analytic_jacobian_group<P, LD>(g, t, y, yp, p, sJy, sJyp, sY, sYp);
            else fd_jacobian_group<P, G, LD>(g, t, y, yp, p, thr_j, fac_y, fac_yp, sJy, sJyp, sY, sYp, sF, sWork);
            ++njev;
            need_jac = false;
        }
        // ---- phase B (bdf.py:269-293) -----------------------------------------------------------
        if (step_begin) {
            min_step = 10.0 * fabs(next_toward(t, dir) - t);
            if (h_abs > L.max_step) {
                hl = L.max_step;
                change_D_call(D, order, L.max_step / h_abs, L.U);
                n_equal_steps = 0;
            } else if (h_abs < min_step) {
                hl = min_step;
                change_D_call(D, order, min_step / h_abs, L.U);
                n_equal_steps = 0;
            } else {
                hl = h_abs;
            }
            current_jac = false;
            step_begin = false;
        }
        // ---- phase A (bdf.py:296-318) -------------------------------------------------------------
        if (!retry) {
            if (hl < min_step) { status = B200DAE_ST_TOO_SMALL_STEP; break; }
            h = hl * dir;
            t_new = t + h;
            if (dir * (t_new - L.t_bound) > 0.0) {
                t_new = L.t_bound;
                change_D_call(D, order, fabs(t_new - t) / hl, L.U);
                n_equal_steps = 0;
                lu_valid = false;
            }
            h = t_new - t;
            hl = fabs(h);
            const double rh = fast_rcp(h);
            double acc = 0.0, ps = 0.0;
#pragma unroll
            for (int i = 0; i < kBdfQ; ++i)
                if (i <= order) acc += D[i];
#pragma unroll
            for (int i = 1; i < kBdfQ; ++i)
                if (i <= order) ps = fma(D[i], L.gamma[i], ps);
            y_predict = acc;
            inv_scale = fast_rcp(fma(rtol, fabs(acc), atol_j));
            psi = ps * rh;
            c = L.alpha[order] * rh;
        }
        retry = false;

        // ---- phase F: LU of Jy + c*Jyp (bdf.py:320-321) ------------------------------------------
        if (!lu_valid) {
            bool fin = true;
#if B200DAE_LU_INPLACE
            if (act) {
#pragma unroll 2
                for (int k = 0; k < N; ++k) {
                    const double v = fma(c, sJyp[jc * LD + k], sJy[jc * LD + k]);
                    fin = fin && finite_d(v);
                    sLU[jc * LD + k] = v;
                }
            }
            if (!g.all(fin)) { status = B200DAE_ST_NONFINITE; break; }
            g.sync();
            group_lu_factor_inplace<N, LD>(g, perm, sLU);
#else
            double a[N];
#pragma unroll
            for (int k = 0; k < N; ++k) {
                const double jy = act ? sJy[jc * LD + k] : 0.0, jyp = act ? sJyp[jc * LD + k] : 0.0;
                a[k] = fma(c, jyp, jy);
                fin = fin && finite_d(a[k]);
            }
            if (!g.all(fin)) { status = B200DAE_ST_NONFINITE; break; }
            group_lu_factor<N, LD>(g, a, perm, sLU, sIdx);
#endif
            ++nlu;
            lu_valid = true;
        }

        // ---- phase N (solve_bdf_system, bdf.py:32-64) ----------------------------------------------
        bool converged = false;
        int n_iter;
        {
            d = 0.0;
            y_new = y_predict;
            double inv_norm_old = 0.0;
            bool has_old = false;
            int k = 0;
            for (; k < kBdfNewtonMaxIter; ++k) {
                yp_new = fma(c, d, psi);
                if (act) { sY[lane] = y_new; sYp[lane] = yp_new; }
                g.sync();
                eval_shared<P>(t_new, sY, sYp, p, sF);           // every lane, identical writes
                g.sync();
                double f = act ? sF[lane] : 0.0;
                ++nfev;
                if (!g.all(finite_d(f))) break;
                f = -f;
                group_lu_solve_real<N>(g, rrow, perm, f);           // f <- dy
                ++nsolve;
                const double q = f * inv_scale;
                const double ss = g.sum(act ? q * q : 0.0);
                const double dy_norm = fast_sqrt(ss * (1.0 / N));
                double rate = 0.0;
                if (has_old) rate = dy_norm * inv_norm_old;
                const double cc = dy_norm * fast_rcp(1.0 - rate);
                if (has_old && (rate >= 1.0 || powi<3>(rate, kBdfNewtonMaxIter - k) * cc > L.newton_tol))
                    break;                                        // BEFORE the update (bdf.py:52-53)
                y_new += f;
                d += f;
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
            if (!current_jac) {
                need_jac = true;
                current_jac = true;
                lu_valid = false;
                retry = true;
                continue;
            }
            hl *= 0.5;
            change_D_call(D, order, 0.5, L.U);
            n_equal_steps = 0;
            lu_valid = false;
            continue;
        }

        // ---- phase E (bdf.py:342-355) ------------------------------------------------------------
        const double safety = 0.9 * (2 * kBdfNewtonMaxIter + 1) * fast_rcp((double)(2 * kBdfNewtonMaxIter + n_iter));
        double error_norm;
        {
            inv_scale = fast_rcp(fma(rtol, fabs(y_new), atol_j));
            const double q = L.error_const[order] * d * inv_scale;
            const double ss = g.sum(act ? q * q : 0.0);
            error_norm = fast_sqrt(ss * (1.0 / N));
        }
        if (error_norm > 1.0) {
            const double factor = fmax(0.2, safety * norm_root(error_norm, L.neg_inv[order + 1]));
            hl *= factor;
            change_D_call(D, order, factor, L.U);
            n_equal_steps = 0;
            ++nrej;
            continue;
        }

        // ---- phase C (bdf.py:359-415) ------------------------------------------------------------
        ++n_equal_steps;
        const double t_prev = t;
        t = t_new;
        y = y_new;
        yp = yp_new;
        h_abs = hl;
        ++nacc;
        {
            // D[order+2] = d - D[order+1]; D[order+1] = d; D[i] += D[i+1] for i = order..0
            double Dn[kBdfRows];
#pragma unroll
            for (int i = 0; i < kBdfRows; ++i) {
                double v = D[i];
                if (i == order + 2) v = d - D[i - 1 >= 0 ? i - 1 : 0];
                if (i == order + 1) v = d;
                Dn[i] = v;
            }
#pragma unroll
            for (int i = kBdfRows - 2; i >= 0; --i)
                if (i <= order) Dn[i] += Dn[i + 1];
#pragma unroll
            for (int i = 0; i < kBdfRows; ++i) D[i] = Dn[i];
        }

        if (n_equal_steps >= order + 1) {
            double norms[3];
            norms[1] = error_norm;
            double dm = 0.0, dp = 0.0;                   // D[order], D[order + 2]
#pragma unroll
            for (int i = 0; i < kBdfRows; ++i) {
                if (i == order) dm = D[i];
                if (i == order + 2) dp = D[i];
            }
            if (order > 1) {
                const double q = L.error_const[order - 1] * dm * inv_scale;
                norms[0] = sqrt(g.sum(act ? q * q : 0.0) * (1.0 / N));
            } else {
                norms[0] = INFINITY;
            }
            if (order < max_order) {
                const double q = L.error_const[order + 1] * dp * inv_scale;
                norms[2] = sqrt(g.sum(act ? q * q : 0.0) * (1.0 / N));
            } else {
                norms[2] = INFINITY;
            }
            double best;
            const int arg = select_order(norms, L.neg_inv[order], L.neg_inv[order + 1], L.neg_inv[order + 2], best);
            order += arg - 1;
            const double factor = fmin(10.0, safety * best);
            h_abs *= factor;
            change_D_call(D, order, factor, L.U);
            n_equal_steps = 0;
            lu_valid = false;
        }

        // my component of BdfDenseOutput at te: built AFTER the rescale (bdf.py:417-419, 525-548)
        const double hh = h_abs * dir;
        const double t_end = t;
        auto dense_lane = [&](double te, double& ya, double& ypa) {
            ya = D[0];
            ypa = 0.0;
            double fact = 1.0, hj = 1.0;
#pragma unroll
            for (int jj = 1; jj < kBdfQ; ++jj) {
                fact *= jj;
                hj *= hh;
                double fc = 1.0, dfac = 0.0;
#pragma unroll
                for (int mm = 0; mm < jj; ++mm) {
                    fc *= te - (t_end - mm * hh);
                    double dprod = 1.0;
#pragma unroll
                    for (int i = 0; i < jj; ++i)
                        if (i != mm) dprod *= te - (t_end - i * hh);
                    dfac += dprod;
                }
                const double rden = 1.0 / (fact * hj);
                if (jj <= order) {
                    ya = fma(D[jj], fc * rden, ya);
                    ypa = fma(D[jj], dfac * rden, ypa);
                }
            }
        };
        // events on this step's dense output (dae.py:466-490); a terminal one moves the step end
        bool ev_stop = false;
        if (kEv && L.ev.spec.n_events > 0) {
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
                gnew[k] = event_quantity_group<P, N>(g, L.ev.spec, k, t, y, yp, p) - L.ev.spec.value[k];
            double ts = t;
            if (handle_events(L.ev, evs, t_prev, t, dir, gnew, dense, record, ts)) {
                ev_stop = true;
                double av, bv;
                dense_lane(ts, av, bv);
                t = ts;
                y = av;
                yp = bv;
            }
        }
        if (kOut && L.out.t_eval != nullptr) {
            while (ie < L.out.m) {
                const double te = __ldg(L.out.t_eval + ie);
                if (!(dir > 0.0 ? te <= t : te >= t)) break;
                double ya, ypa;
                dense_lane(te, ya, ypa);
                if (act) {
                    L.out.y_eval[ie * L.out.eval_stride_row + lane * L.out.eval_stride_comp + inst * L.out.eval_stride_inst] = ya;
                    if (L.out.yp_eval) L.out.yp_eval[ie * L.out.eval_stride_row + lane * L.out.eval_stride_comp + inst * L.out.eval_stride_inst] = ypa;
                }
                ++ie;
            }
        }
        if (kOut && L.out.t_steps != nullptr && nst < L.out.steps_cap) {
            if (lane == 0) L.out.t_steps[(int64_t)nst * ld + inst] = t;
            if (act) {
                if (L.out.y_steps) L.out.y_steps[((int64_t)nst * N + lane) * ld + inst] = y;
                if (L.out.yp_steps) L.out.yp_steps[((int64_t)nst * N + lane) * ld + inst] = yp;
            }
            if (L.out.dense_coef != nullptr) {     // this step's BdfDenseOutput (bdf.py:417-419)
                if (lane == 0) {
                    L.out.h_steps[((int64_t)nst * 2) * ld + inst] = hh;
                    L.out.h_steps[((int64_t)nst * 2 + 1) * ld + inst] = t_end;
                    if (L.out.order_steps) L.out.order_steps[(int64_t)nst * ld + inst] = order;
                }
                if (act) {
#pragma unroll
                    for (int i = 0; i < kBdfQ; ++i)
                        if (i <= order) L.out.dense_coef[((int64_t)nst * (kBdfQ * N) + i * N + lane) * ld + inst] = D[i];
                }
            }
            ++nst;
        }
        step_begin = true;
        if (ev_stop) status = B200DAE_ST_EVENT;
        else if (dir * (t - L.t_bound) >= 0.0) status = B200DAE_ST_FINISHED;
        else if (L.max_steps > 0 && nacc >= L.max_steps) status = B200DAE_ST_MAX_STEPS;
    }

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
        if (kEv) for (int k = 0; k < L.ev.spec.n_events; ++k) L.ev.n_found[(int64_t)k * ld + inst] = evs.found[k];
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
}

}  // namespace b200dae
