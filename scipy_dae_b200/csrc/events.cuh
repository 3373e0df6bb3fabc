// events.cuh — event detection and location inside the device step loop.
//
// Replaces, per instance, the event part of solve_dae's loop (dae.py:426-443, 466-490) with
// handle_events (dae.py:58-109), solve_event_equation (dae.py:30-55: scipy.optimize.brentq with
// xtol = rtol = 4 EPS on the step's dense output) and scipy's find_active_events
// (scipy/integrate/_ivp/ivp.py).  A Python event callable cannot run on the device; an event is
// either a THRESHOLD  g_k(t, y, y') = x[component_k] - value_k,  x = y (kind 0) or y' (kind 1), or
// one of the problem functor's own event functions  g_k = P::event(component_k, t, y, y', p) - value_k
// (kind 2, problems.cuh: EventOf), with the reference's `direction` and `terminal` attributes
// (terminal = k stops at the k-th occurrence).  The kind is spec.of_yp[k].
#pragma once
#include "rt_compat.cuh"
#include "../../include/b200dae.h"
#include "problems.cuh"

namespace b200dae {

struct EventLaunch {
    b200dae_event_spec spec;
    int64_t cap;                         // rows per event in the outputs
    double *t_events;                    // [ne][cap][ld]
    double *y_events, *yp_events;        // [ne][cap][n][ld]
    int32_t* n_found;                    // [ne][ld]
};

struct EventState {
    double g[B200DAE_MAX_EVENTS];
    int32_t count[B200DAE_MAX_EVENTS];
    int32_t found[B200DAE_MAX_EVENTS];
};

// scipy/optimize/Zeros/brentq.c, maxiter = 100, xtol = rtol = 4 EPS.  f is evaluated with the
// same arguments by every cooperating lane, so the control flow is uniform in a group.
template <class F>
__device__ __forceinline__ double brentq_dev(F f, double xa, double xb) {
    const double xtol = 4.0 * 2.220446049250313e-16, rtol = xtol;
    double xpre = xa, xcur = xb, xblk = 0.0, fblk = 0.0, spre = 0.0, scur = 0.0;
    double fpre = f(xpre), fcur = f(xcur);
    if (fpre == 0.0) return xpre;
    if (fcur == 0.0) return xcur;
    if (signbit(fpre) == signbit(fcur)) return nan("");     // scipy raises ValueError
#pragma unroll 1
    for (int i = 0; i < 100; ++i) {
        if (fpre != 0.0 && fcur != 0.0 && (signbit(fpre) != signbit(fcur))) {
            xblk = xpre; fblk = fpre; spre = scur = xcur - xpre;
        }
        if (fabs(fblk) < fabs(fcur)) {
            xpre = xcur; xcur = xblk; xblk = xpre;
            fpre = fcur; fcur = fblk; fblk = fpre;
        }
        const double delta = (xtol + rtol * fabs(xcur)) / 2;
        const double sbis = (xblk - xcur) / 2;
        if (fcur == 0.0 || fabs(sbis) < delta) return xcur;
        if (fabs(spre) > delta && fabs(fcur) < fabs(fpre)) {
            double stry;
            if (xpre == xblk) {
                stry = -fcur * (xcur - xpre) / (fcur - fpre);
            } else {
                const double dpre = (fpre - fcur) / (xpre - xcur);
                const double dblk = (fblk - fcur) / (xblk - xcur);
                stry = -fcur * (fblk * dblk - fpre * dpre) / (dblk * dpre * (fblk - fpre));
            }
            if (2 * fabs(stry) < fmin(fabs(spre), 3 * fabs(sbis) - delta)) { spre = scur; scur = stry; }
            else { spre = sbis; scur = sbis; }
        } else {
            spre = sbis; scur = sbis;
        }
        xpre = xcur; fpre = fcur;
        if (fabs(scur) > delta) xcur += scur;
        else xcur += (sbis > 0 ? delta : -delta);
        fcur = f(xcur);
    }
    return xcur;
}

// After an accepted step [t_old, t_new].
//   gnew[k]                    event values at the step end
//   dense(t, comp, kind)       the event quantity on the step's dense output at t (uniform across a
//                              group): y[comp], y'[comp] or the functor's event function `comp`
//   record(k, idx, t_root)     store y(t_root), y'(t_root) as occurrence idx of event k
// Returns true when a terminal event fired; t_stop is then the event time that replaces t_new.
template <class Dense, class Record>
__device__ __forceinline__ bool handle_events(const EventLaunch& E, EventState& S, double t_old,
                                              double t_new, double dir, const double* gnew,
                                              Dense dense, Record record, double& t_stop) {
    const int ne = E.spec.n_events;
    int act[B200DAE_MAX_EVENTS], na = 0;
    double roots[B200DAE_MAX_EVENTS];
    for (int k = 0; k < ne; ++k) {
        const bool up = (S.g[k] <= 0.0) && (gnew[k] >= 0.0), down = (S.g[k] >= 0.0) && (gnew[k] <= 0.0);
        const int d = E.spec.direction[k];
        if ((up && d > 0) || (down && d < 0) || ((up || down) && d == 0)) act[na++] = k;
    }
    bool terminate = false;
    if (na > 0) {
        for (int a = 0; a < na; ++a) {
            const int k = act[a];
            S.count[k] += 1;
            const int comp = E.spec.component[k];
            const int kind = E.spec.of_yp[k];
            const double c = E.spec.value[k];
            roots[a] = brentq_dev([&](double tt) { return dense(tt, comp, kind) - c; }, t_old, t_new);
        }
        bool any_term = false;
        for (int a = 0; a < na; ++a)
            any_term = any_term || (E.spec.max_events[act[a]] > 0 && S.count[act[a]] >= E.spec.max_events[act[a]]);
        int nrec = na;
        if (any_term) {
            for (int i = 1; i < na; ++i)                       // stable sort in integration direction
                for (int j = i; j > 0; --j) {
                    const bool sw = dir > 0.0 ? (roots[j] < roots[j - 1]) : (-roots[j] < -roots[j - 1]);
                    if (!sw) break;
                    const double tr = roots[j]; roots[j] = roots[j - 1]; roots[j - 1] = tr;
                    const int ti = act[j]; act[j] = act[j - 1]; act[j - 1] = ti;
                }
            for (int a = 0; a < na; ++a)
                if (E.spec.max_events[act[a]] > 0 && S.count[act[a]] >= E.spec.max_events[act[a]]) { nrec = a + 1; break; }
            terminate = true;
        }
        for (int a = 0; a < nrec; ++a) {
            const int k = act[a];
            if (S.found[k] < E.cap) record(k, S.found[k], roots[a]);
            S.found[k] += 1;
        }
        if (terminate) t_stop = roots[nrec - 1];
    }
    for (int k = 0; k < ne; ++k) S.g[k] = gnew[k];
    return terminate;
}

// a[j] for a run-time j without dynamic register indexing
template <int N>
__device__ __forceinline__ double pick(const double* a, int j) {
    double r = a[0];
#pragma unroll
    for (int i = 1; i < N; ++i) r = (j == i) ? a[i] : r;
    return r;
}

// event quantity k at a point where the whole state is at hand (thread-per-system kernels)
template <class P, int N>
__device__ __forceinline__ double event_quantity(const b200dae_event_spec& sp, int k, double t, const double* y,
                                                 const double* yp, const double* p) {
    const int kind = sp.of_yp[k], comp = sp.component[k];
    if constexpr (EventOf<P>::kNum > 0) {
        if (kind == 2) return EventOf<P>::eval(comp, t, y, yp, p);
    }
    return kind == 1 ? pick<N>(yp, comp) : pick<N>(y, comp);
}

// the same in the group kernels: lane j holds component j (yv, ypv); the event functor needs the whole
// vectors, which every lane gathers by shuffles (the kind is uniform across the group)
template <class P, int N, class Grp>
__device__ __forceinline__ double event_quantity_group(const Grp& g, const b200dae_event_spec& sp, int k, double t,
                                                       double yv, double ypv, const double* p) {
    const int kind = sp.of_yp[k], comp = sp.component[k];
    if constexpr (EventOf<P>::kNum > 0) {
        if (kind == 2) {
            double ya[N], ypa[N];
#pragma unroll
            for (int j = 0; j < N; ++j) { ya[j] = g.bcast(yv, j); ypa[j] = g.bcast(ypv, j); }
            return EventOf<P>::eval(comp, t, ya, ypa, p);
        }
    }
    return g.bcast(kind == 1 ? ypv : yv, comp);
}

}  // namespace b200dae
