// problems.cuh — residuals F(t, y, y'; p) = 0 as device functors.
//
// Replaces the Python callable `fun(t, y, yp, *args)` of the reference (base.py:44-45,
// dae.py:365-383).  Twins: scipy_dae_b200/problems.py (NumPy, drives the unmodified reference
// when goldens are minted) and oracle/problems.c (CPU oracle).  Definitions: SURVEY.md App. D
// (reference README.md:37-47, examples/daes/akzo_nobel.py:18-50, examples/daes/andrews.py:20-225;
// "transamp" is the IVP-test-set two-transistor amplifier).
//
// A functor provides  N (state size), NP (parameters per instance), name, and
//   static __device__ void eval(double t, const double* y, const double* yp, const double* p, double* out)
// and optionally (kHasJac) the analytic Jacobians of the reference's `jac=` callable (base.py:299-331)
//   template <class M> static __device__ void jac(t, y, yp, p, M Jy, M Jyp)
// where M(r, c) is a reference to entry (r, c) of a matrix the caller has zeroed.
// All indexing is static so that, after inlining into a fully unrolled caller, every array
// lives in registers.
//
// Rounding.  Every residual rounds each operation on its own, in the order its twins are written
// in (R64 below: operators that compile to __dadd_rn / __dmul_rn / __ddiv_rn, which ptxas never
// contracts into FMAs).  num_jac differences F(x + h e_i) - F(x); left to the compiler's FMA
// contraction the two evaluations round differently per inlined call site, which puts noise of
// eps |F| / h into structurally zero Jacobian entries (measured in round 1: +6 % steps on a Robertson
// golden, +2.4 % LU factorisations on the Akzo Nobel ensemble in the column-parallel FD of the
// group kernel).  With R64 a residual is a pure function of its arguments, bit-identical at every
// call site and, up to the libm functions (exp, sin, cos; sqrt and division are IEEE), to the C
// oracle twin (built with -ffp-contract=off) and the NumPy twin the reference integrates.
#pragma once
#include "rt_compat.cuh"

namespace b200dae {

struct R64 {
    double v;
    __device__ __forceinline__ R64() : v(0.0) {}
    __device__ __forceinline__ R64(double x) : v(x) {}
};
__device__ __forceinline__ R64 operator+(R64 a, R64 b) { return R64(__dadd_rn(a.v, b.v)); }
__device__ __forceinline__ R64 operator-(R64 a, R64 b) { return R64(__dsub_rn(a.v, b.v)); }
__device__ __forceinline__ R64 operator*(R64 a, R64 b) { return R64(__dmul_rn(a.v, b.v)); }
__device__ __forceinline__ R64 operator/(R64 a, R64 b) { return R64(__ddiv_rn(a.v, b.v)); }
__device__ __forceinline__ R64 operator-(R64 a) { return R64(-a.v); }
__device__ __forceinline__ R64 rsin(R64 a) { return R64(sin(a.v)); }
__device__ __forceinline__ R64 rcos(R64 a) { return R64(cos(a.v)); }
__device__ __forceinline__ R64 rexp(R64 a) { return R64(exp(a.v)); }
__device__ __forceinline__ R64 rsqrt64(R64 a) { return R64(sqrt(a.v)); }

// Event functions g(t, y, y'; p) of a problem (the reference's `events=` callables, dae.py:226-247,
// 426-443): a functor may provide  kNumEvents  and
//   static __device__ double event(int which, double t, const double* y, const double* yp, const double* p).
// EventOf<P> answers for functors that provide none.
template <class P, class = void>
struct EventOf {
    static constexpr int kNum = 0;
    __device__ __forceinline__ static double eval(int, double, const double*, const double*, const double*) { return nan(""); }
};
template <class P>
struct EventOf<P, decltype((void)P::kNumEvents)> {
    static constexpr int kNum = P::kNumEvents;
    __device__ __forceinline__ static double eval(int which, double t, const double* y, const double* yp, const double* p) {
        return P::event(which, t, y, yp, p);
    }
};

struct Robertson {
    static constexpr int N = 3, NP = 3;
    static constexpr const char* name() { return "robertson"; }
    // flops of one residual as written (for the roofline model): 15
    static constexpr int kFlops = 15;
    __device__ __forceinline__ static void eval(double, const double* y, const double* yp,
                                                const double* p, double* out) {
        const double k1 = p[0], k2 = p[1], k3 = p[2];
        // Every operation rounded on its own, in the order of the reference's NumPy callable
        // (README.md:43-45): the residual is bit-identical to the reference's for the same
        // arguments and rounds the same way at every inlined call site, which the differences
        // F(x + h e_i) - F(x) of num_jac need.  Measured alternatives: the plain expression left to
        // the compiler's FMA contraction (0.25 ms faster per 2^20 solves) gave site-dependent noise
        // in structurally zero Jacobian columns (+6 % steps on the vector-atol golden), and a fixed
        // three-FMA form (0.4 ms faster) moved the s = 7 golden by 12 % in accepted steps.
        // -k1 y1 + k2 y2 y3 is formed as -(k1 y1 - k2 y2 y3): exact negation, one DADD less.
        const double r23 = __dmul_rn(__dmul_rn(k2, y[1]), y[2]);
        const double b = __dsub_rn(__dmul_rn(k1, y[0]), r23);
        out[0] = __dadd_rn(yp[0], b);
        out[1] = __dsub_rn(yp[1], __dsub_rn(b, __dmul_rn(__dmul_rn(k3, y[1]), y[1])));
        out[2] = __dsub_rn(__dadd_rn(__dadd_rn(y[0], y[1]), y[2]), 1.0);
    }
    // event 0: y1 - y3 (the reactant and the product cross), event 1: y1' + y2' + k1 y1 / 2
    static constexpr int kNumEvents = 2;
    __device__ __forceinline__ static double event(int which, double, const double* y, const double* yp, const double* p) {
        return which == 0 ? (R64(y[0]) - R64(y[2])).v : (R64(yp[0]) + R64(yp[1]) + R64(p[0]) * R64(y[0]) * 0.5).v;
    }
    static constexpr bool kHasJac = true;      // reference: examples/daes/robertson.py:38-47
    template <class M>
    __device__ __forceinline__ static void jac(double, const double* y, const double*, const double* p,
                                               M Jy, M Jyp) {
        const double k1 = p[0], k2 = p[1], k3 = p[2];
        Jy(0, 0) = k1;  Jy(0, 1) = -k2 * y[2];                     Jy(0, 2) = -k2 * y[1];
        Jy(1, 0) = -k1; Jy(1, 1) = fma(2.0 * k3, y[1], k2 * y[2]); Jy(1, 2) = k2 * y[1];
        Jy(2, 0) = 1.0; Jy(2, 1) = 1.0;                            Jy(2, 2) = 1.0;
        Jyp(0, 0) = 1.0; Jyp(1, 1) = 1.0;
    }
};

struct AkzoNobel {
    static constexpr int N = 6, NP = 7;
    static constexpr const char* name() { return "akzo_nobel"; }
    static constexpr int kFlops = 52;
    __device__ __forceinline__ static void eval(double, const double* y, const double* yp,
                                                const double* p, double* out) {
        const double po2 = 0.9, hen = 737.0;
        const R64 k1 = p[0], k2 = p[1], k3 = p[2], k4 = p[3], kbig = p[4], kla = p[5], ks = p[6];
        const R64 y0 = y[0], y1 = y[1], y2 = y[2], y3 = y[3], y4 = y[4], y5 = y[5];
        const R64 sq = rsqrt64(y1);              // NaN for negative trial iterates: Newton treats it
        const R64 y0sq = y0 * y0;                // as non-convergence (radau.py:157-158, bdf.py:41-42)
        const R64 r1 = k1 * (y0sq * y0sq) * sq;
        const R64 r2 = k2 * y2 * y3;
        const R64 r3 = k2 / kbig * y0 * y4;      // k2 / kbig is loop-invariant: hoisted by the compiler
        const R64 r4 = k3 * y0 * (y3 * y3);
        const R64 r5 = k4 * (y5 * y5) * sq;
        const R64 fin = kla * (R64(po2 / hen) - y1);
        out[0] = (R64(yp[0]) - (-2.0 * r1 + r2 - r3 - r4)).v;
        out[1] = (R64(yp[1]) - (-0.5 * r1 - r4 - 0.5 * r5 + fin)).v;
        out[2] = (R64(yp[2]) - (r1 - r2 + r3)).v;
        out[3] = (R64(yp[3]) - (-r2 + r3 - 2.0 * r4)).v;
        out[4] = (R64(yp[4]) - (r2 - r3 + r5)).v;
        out[5] = (-(ks * y0 * y3 - y5)).v;
    }
    static constexpr int kNumEvents = 1;       // event 0: y1 - y3
    __device__ __forceinline__ static double event(int, double, const double* y, const double*, const double*) {
        return (R64(y[0]) - R64(y[2])).v;
    }
    static constexpr bool kHasJac = true;      // reference: examples/daes/akzo_nobel.py:52-103
    template <class M>
    __device__ __forceinline__ static void jac(double, const double* y, const double*, const double* p,
                                               M Jy, M Jyp) {
        const double k1 = p[0], k2 = p[1], k3 = p[2], k4 = p[3], kbig = p[4], kla = p[5], ks = p[6];
        const double sq = sqrt(y[1]), hsq = 0.5 / sq;
        const double y0c = y[0] * y[0] * y[0];
        const double r1_0 = 4.0 * k1 * y0c * sq, r1_1 = k1 * (y0c * y[0]) * hsq;
        const double r2_2 = k2 * y[3], r2_3 = k2 * y[2];
        const double kk = k2 / kbig;
        const double r3_0 = kk * y[4], r3_4 = kk * y[0];
        const double r4_0 = k3 * (y[3] * y[3]), r4_3 = 2.0 * k3 * y[0] * y[3];
        const double r5_1 = k4 * (y[5] * y[5]) * hsq, r5_5 = 2.0 * k4 * y[5] * sq;
        Jy(0, 0) = -(-2.0 * r1_0 - r3_0 - r4_0); Jy(0, 1) = 2.0 * r1_1; Jy(0, 2) = -r2_2;
        Jy(0, 3) = -(r2_3 - r4_3); Jy(0, 4) = r3_4;
        Jy(1, 0) = -(-0.5 * r1_0 - r4_0); Jy(1, 1) = -(-0.5 * r1_1 - 0.5 * r5_1 - kla);
        Jy(1, 3) = r4_3; Jy(1, 5) = 0.5 * r5_5;
        Jy(2, 0) = -(r1_0 + r3_0); Jy(2, 1) = -r1_1; Jy(2, 2) = r2_2; Jy(2, 3) = r2_3; Jy(2, 4) = -r3_4;
        Jy(3, 0) = -(r3_0 - 2.0 * r4_0); Jy(3, 2) = r2_2; Jy(3, 3) = -(-r2_3 - 2.0 * r4_3); Jy(3, 4) = -r3_4;
        Jy(4, 0) = r3_0; Jy(4, 1) = -r5_1; Jy(4, 2) = -r2_2; Jy(4, 3) = -r2_3; Jy(4, 4) = r3_4; Jy(4, 5) = -r5_5;
        Jy(5, 0) = -ks * y[3]; Jy(5, 3) = -ks * y[0]; Jy(5, 5) = 1.0;
        Jyp(0, 0) = 1.0; Jyp(1, 1) = 1.0; Jyp(2, 2) = 1.0; Jyp(3, 3) = 1.0; Jyp(4, 4) = 1.0;
    }
};

struct Transamp {
    static constexpr int N = 8, NP = 3;
    static constexpr const char* name() { return "transamp"; }
    static constexpr int kFlops = 75;
    __device__ __forceinline__ static void eval(double t, const double* y, const double* yp,
                                                const double* p, double* out) {
        const double UF = 0.026, ALPHA = 0.99, R = 9000.0;
        const double C1 = 1e-6, C2 = 2e-6, C3 = 3e-6, C4 = 4e-6, C5 = 5e-6;
        const R64 Ub = p[0], R0 = p[1], beta = p[2];
        // divisions by the constants UF, R and by the parameter R0 are multiplications by their
        // reciprocals IN ALL THREE TWINS (problems.py, oracle/problems.c): the problem is restated here
        // (it is not in the reference tree), so its arithmetic is ours to fix
        const double rUF = 1.0 / UF, rR = 1.0 / R, r2R = 2.0 / R;
        const R64 rR0 = R64(1.0) / R0;           // loop-invariant
        const R64 Ue = 0.1 * rsin(R64(200.0 * 3.141592653589793) * R64(t));
        const R64 g23 = beta * (rexp((R64(y[1]) - R64(y[2])) * rUF) - 1.0);
        const R64 g56 = beta * (rexp((R64(y[4]) - R64(y[5])) * rUF) - 1.0);
        const R64 UbR = Ub * rR;
        const R64 f0 = (R64(y[0]) - Ue) * rR0;
        const R64 f1 = -UbR + R64(y[1]) * r2R - (ALPHA - 1.0) * g23;
        const R64 f2 = -g23 + R64(y[2]) * rR;
        const R64 f3 = -UbR + R64(y[3]) * rR + ALPHA * g23;
        const R64 f4 = -UbR + R64(y[4]) * r2R - (ALPHA - 1.0) * g56;
        const R64 f5 = -g56 + R64(y[5]) * rR;
        const R64 f6 = -UbR + R64(y[6]) * rR + ALPHA * g56;
        const R64 f7 = R64(y[7]) * rR;
        const R64 yp0 = yp[0], yp1 = yp[1], yp2 = yp[2], yp3 = yp[3], yp4 = yp[4], yp5 = yp[5], yp6 = yp[6], yp7 = yp[7];
        out[0] = ((-C1 * yp0 + C1 * yp1) - f0).v;
        out[1] = ((C1 * yp0 - C1 * yp1) - f1).v;
        out[2] = ((-C2 * yp2) - f2).v;
        out[3] = ((-C3 * yp3 + C3 * yp4) - f3).v;
        out[4] = ((C3 * yp3 - C3 * yp4) - f4).v;
        out[5] = ((-C4 * yp5) - f5).v;
        out[6] = ((-C5 * yp6 + C5 * yp7) - f6).v;
        out[7] = ((C5 * yp6 - C5 * yp7) - f7).v;
    }
    static constexpr bool kHasJac = true;      // F = M yp - f(t, y):  Jyp = M,  Jy = -df/dy
    template <class M>
    __device__ __forceinline__ static void jac(double, const double* y, const double*, const double* p,
                                               M Jy, M Jyp) {
        const double UF = 0.026, a = 0.99, R = 9000.0;
        const double C1 = 1e-6, C2 = 2e-6, C3 = 3e-6, C4 = 4e-6, C5 = 5e-6;
        const double R0 = p[1], beta = p[2];
        const double rUF = 1.0 / UF, rR = 1.0 / R;
        const double e23 = beta * exp((y[1] - y[2]) * rUF) * rUF;
        const double e56 = beta * exp((y[4] - y[5]) * rUF) * rUF;
        Jy(0, 0) = -1.0 / R0;
        Jy(1, 1) = -(2.0 * rR - (a - 1.0) * e23); Jy(1, 2) = -(a - 1.0) * e23;
        Jy(2, 1) = e23; Jy(2, 2) = -(e23 + rR);
        Jy(3, 1) = -a * e23; Jy(3, 2) = a * e23; Jy(3, 3) = -rR;
        Jy(4, 4) = -(2.0 * rR - (a - 1.0) * e56); Jy(4, 5) = -(a - 1.0) * e56;
        Jy(5, 4) = e56; Jy(5, 5) = -(e56 + rR);
        Jy(6, 4) = -a * e56; Jy(6, 5) = a * e56; Jy(6, 6) = -rR;
        Jy(7, 7) = -rR;
        Jyp(0, 0) = -C1; Jyp(0, 1) = C1; Jyp(1, 0) = C1; Jyp(1, 1) = -C1;
        Jyp(2, 2) = -C2;
        Jyp(3, 3) = -C3; Jyp(3, 4) = C3; Jyp(4, 3) = C3; Jyp(4, 4) = -C3;
        Jyp(5, 5) = -C4;
        Jyp(6, 6) = -C5; Jyp(6, 7) = C5; Jyp(7, 6) = C5; Jyp(7, 7) = -C5;
    }
};

// ---- the reference's own known-answer test problems ---------------------------

// 2-D rational ODE of tests/test_integration_rational.py (scipy test_ivp.fun_rational) as
// F = y' - f(t, y); exact solution (t / (t + 10), 10 t / (t + 10)^2).
struct Rational {
    static constexpr int N = 2, NP = 0;
    static constexpr const char* name() { return "rational"; }
    static constexpr int kFlops = 12;
    __device__ __forceinline__ static void eval(double t, const double* y, const double* yp, const double*,
                                                double* out) {
        const R64 T = t, y0 = y[0], y1 = y[1];
        out[0] = (R64(yp[0]) - y1 / T).v;
        out[1] = (R64(yp[1]) - y1 * (y0 + 2.0 * y1 - 1.0) / (T * (y0 - 1.0))).v;
    }
    static constexpr bool kHasJac = true;      // J_rational, tests/test_integration_rational.py:20-23
    template <class M>
    __device__ __forceinline__ static void jac(double t, const double* y, const double*, const double*,
                                               M Jy, M Jyp) {
        const double d = y[0] - 1.0;
        Jy(0, 1) = -1.0 / t;
        Jy(1, 0) = 2.0 * y[1] * y[1] / (t * d * d);
        Jy(1, 1) = -(y[0] + 4.0 * y[1] - 1.0) / (t * d);
        Jyp(0, 0) = 1.0; Jyp(1, 1) = 1.0;
    }
};

// linear 2-D ODE of tests/test_integration_linear.py (scipy test_ivp.fun_linear):
// exact solution (-5 sin 2t, 2 cos 2t + sin 2t).
struct Linear {
    static constexpr int N = 2, NP = 0;
    static constexpr const char* name() { return "linear"; }
    static constexpr int kFlops = 6;
    __device__ __forceinline__ static void eval(double, const double* y, const double* yp, const double*,
                                                double* out) {
        const R64 y0 = y[0], y1 = y[1];
        out[0] = (R64(yp[0]) - (-y0 - 5.0 * y1)).v;
        out[1] = (R64(yp[1]) - (y0 + y1)).v;
    }
    static constexpr bool kHasJac = true;      // J_linear (a constant tuple in the reference's test)
    template <class M>
    __device__ __forceinline__ static void jac(double, const double*, const double*, const double*, M Jy, M Jyp) {
        Jy(0, 0) = 1.0; Jy(0, 1) = 5.0; Jy(1, 0) = -1.0; Jy(1, 1) = -1.0;
        Jyp(0, 0) = 1.0; Jyp(1, 1) = 1.0;
    }
};

// y' - a cos t = 0 (tests/test_events.py:8-9 with a = 1): y = a sin t, the event test problem.
struct Cosine {
    static constexpr int N = 1, NP = 1;
    static constexpr const char* name() { return "cosine"; }
    static constexpr int kFlops = 3;
    __device__ __forceinline__ static void eval(double t, const double*, const double* yp, const double* p,
                                                double* out) {
        out[0] = (R64(yp[0]) - R64(p[0]) * rcos(R64(t))).v;
    }
    // events that are not thresholds on one component: 0: y - y' (a sin t = a cos t at t = pi/4 + k pi),
    // 1: y - a t / 2 (depends on t)
    static constexpr int kNumEvents = 2;
    __device__ __forceinline__ static double event(int which, double t, const double* y, const double* yp, const double* p) {
        return which == 0 ? (R64(y[0]) - R64(yp[0])).v : (R64(y[0]) - R64(p[0]) * R64(t) * 0.5).v;
    }
    static constexpr bool kHasJac = false;
};

// Weissinger's fully implicit equation (benchmarks/weissinger/weissinger.py:11-17,
// tests/test_consistent_initial_conditions.py) for z = a y:  F(t, z / a, z' / a) = 0,
// exact solution z = a sqrt(t^2 + 1/2).
struct Weissinger {
    static constexpr int N = 1, NP = 1;
    static constexpr const char* name() { return "weissinger"; }
    static constexpr int kFlops = 16;
    __device__ __forceinline__ static void eval(double t, const double* y, const double* yp, const double* p,
                                                double* out) {
        const R64 T = t, a = p[0];
        const R64 u = R64(y[0]) / a, up = R64(yp[0]) / a;
        out[0] = (T * (u * u) * (up * up * up) - (u * u * u) * (up * up) + T * (T * T + 1.0) * up - (T * T) * u).v;
    }
    static constexpr bool kHasJac = false;
};

// Brenan's index-1 problem (benchmarks/brenan/brenan.py:11-19) with amplitude A:
// y1' - t y2' + y1 - (1 + t) y2 = 0,  y2 - A sin t = 0;  y1 = C exp(-t) + A t sin t, y2 = A sin t.
struct Brenan {
    static constexpr int N = 2, NP = 1;
    static constexpr const char* name() { return "brenan"; }
    static constexpr int kFlops = 10;
    __device__ __forceinline__ static void eval(double t, const double* y, const double* yp, const double* p,
                                                double* out) {
        const R64 T = t;
        out[0] = (R64(yp[0]) - T * R64(yp[1]) + R64(y[0]) - (1.0 + T) * R64(y[1])).v;
        out[1] = (R64(y[1]) - R64(p[0]) * rsin(T)).v;
    }
    static constexpr bool kHasJac = false;
};

// One-transistor amplifier of the reference's examples/daes/stiff_transistor_amplifier.py:20-50
// (n = 5, Hairer & Wanner II p. 377), p = (Ub, R0, beta); R1..R5 = Ri = 9000, Uf = 0.026,
// alpha = 0.99, C_k = k * 1e-6, Ue = 0.4 sin(200 pi t).  At rtol = atol = 1e-6 the reference ends
// with status -1 ("Required step size is less than spacing between numbers") for Radau and BDF:
// the failure-path goldens (tests/golden/transamp1_*).  Divisions as in the example.
struct Transamp1 {
    static constexpr int N = 5, NP = 3;
    static constexpr const char* name() { return "transamp1"; }
    static constexpr int kFlops = 40;
    __device__ __forceinline__ static void eval(double t, const double* y, const double* yp,
                                                const double* p, double* out) {
        const double UF = 0.026, ALPHA = 0.99, R = 9000.0;
        const double C1 = 1e-6, C2 = 2e-6, C3 = 3e-6;
        const R64 Ub = p[0], R0 = p[1], beta = p[2];
        const R64 U1 = y[0], U2 = y[1], U3 = y[2], U4 = y[3], U5 = y[4];
        const R64 Up1 = yp[0], Up2 = yp[1], Up3 = yp[2], Up4 = yp[3], Up5 = yp[4];
        const R64 Ue = 0.4 * rsin(R64(200.0 * 3.141592653589793) * R64(t));
        const R64 f23 = beta * (rexp((U2 - U3) / UF) - 1.0);
        out[0] = ((Ue - U1) / R0 + C1 * (Up2 - Up1)).v;
        out[1] = ((Ub - U2) / R - U2 / R + C1 * (Up1 - Up2) - (1.0 - ALPHA) * f23).v;
        out[2] = (f23 - U3 / R - C2 * Up3).v;
        out[3] = ((Ub - U4) / R + C3 * (Up5 - Up4) - ALPHA * f23).v;
        out[4] = (-U5 / R + C3 * (Up4 - Up5)).v;
    }
    static constexpr bool kHasJac = false;
};

// y' = p y^2: y = y0 / (1 - p y0 (t - t0)) blows up at t* = t0 + 1 / (p y0).  Every adaptive
// solver fails just before t* with "step size too small" (DaeSolver.TOO_SMALL_STEP, base.py:158),
// at a time that rounding does not move: the deterministic failure-path case (solve_dae breaks
// before recording the failed step, dae.py:451-453).
struct Blowup {
    static constexpr int N = 1, NP = 1;
    static constexpr const char* name() { return "blowup"; }
    static constexpr int kFlops = 3;
    __device__ __forceinline__ static void eval(double, const double* y, const double* yp, const double* p,
                                                double* out) {
        out[0] = (R64(yp[0]) - R64(p[0]) * R64(y[0]) * R64(y[0])).v;
    }
    static constexpr bool kHasJac = false;
};

// Andrews' squeezer, stabilised index-1 form, n = 26: y = (q[7], u[7], la[6], mu[6]),
// yp = (q'[7], u'[7], lambda[6], mu-multipliers[6]); p = (mom, c0).
struct Andrews {
    static constexpr int N = 26, NP = 2;
    static constexpr const char* name() { return "andrews"; }
    static constexpr int kFlops = 520;
    static constexpr bool kHasJac = false;     // FD Jacobian only (as examples/daes/andrews.py)
    __device__ __noinline__ static void eval(double, const double* y, const double* yp,
                                             const double* p, double* out) {
        const double m1 = 0.04325, m2 = 0.00365, m3 = 0.02373, m4 = 0.00706, m5 = 0.07050,
                     m6 = 0.00706, m7 = 0.05498;
        const double I1 = 2.194e-6, I2 = 4.410e-7, I3 = 5.255e-6, I4 = 5.667e-7, I5 = 1.169e-5,
                     I6 = 5.667e-7, I7 = 1.912e-5;
        const double xa = -0.06934, ya = -0.00227, xb = -0.03635, yb = 0.03273, xc = 0.014,
                     yc = 0.072;
        const double d = 0.028, da = 0.0115, e = 0.02, ea = 0.01421, zf = 0.02, fa = 0.01421;
        const double rr = 0.007, ra = 0.00092, ss = 0.035, sa = 0.01874, sb = 0.01043, sc = 0.018;
        const double sd = 0.02, zt = 0.04, ta = 0.02308, tb = 0.00916, u = 0.04, ua = 0.01228,
                     ub = 0.00449, l0 = 0.07785;
        const R64 mom = p[0], c0 = p[1];
        R64 q[7], v[7], qd[7], vd[7], la[6], mu[6];
#pragma unroll
        for (int i = 0; i < 7; ++i) { q[i] = y[i]; v[i] = y[7 + i]; qd[i] = yp[i]; vd[i] = yp[7 + i]; }
#pragma unroll
        for (int i = 0; i < 6; ++i) { la[i] = yp[14 + i]; mu[i] = yp[20 + i]; }

        // ten angles, as in the twins (the angle sums are evaluated directly, not by addition theorems)
        R64 sin_be, cos_be, sin_Th, cos_Th, sin_ga, cos_ga, sin_Phi, cos_Phi, sin_de, cos_de,
            sin_Om, cos_Om, sin_ep, cos_ep, sin_be_Th, cos_be_Th, sin_Ph_de, cos_Ph_de, sin_Om_ep,
            cos_Om_ep;
        sincos(q[0].v, &sin_be.v, &cos_be.v);
        sincos(q[1].v, &sin_Th.v, &cos_Th.v);
        sincos(q[2].v, &sin_ga.v, &cos_ga.v);
        sincos(q[3].v, &sin_Phi.v, &cos_Phi.v);
        sincos(q[4].v, &sin_de.v, &cos_de.v);
        sincos(q[5].v, &sin_Om.v, &cos_Om.v);
        sincos(q[6].v, &sin_ep.v, &cos_ep.v);
        sincos((q[0] + q[1]).v, &sin_be_Th.v, &cos_be_Th.v);
        sincos((q[3] + q[4]).v, &sin_Ph_de.v, &cos_Ph_de.v);
        sincos((q[5] + q[6]).v, &sin_Om_ep.v, &cos_Om_ep.v);

        const R64 M00 = m1 * ra * ra + m2 * (rr * rr - 2 * da * rr * cos_Th + da * da) + I1 + I2;
        const R64 M01 = m2 * (da * da - da * rr * cos_Th) + I2;
        const double M11 = m2 * da * da + I2;
        const double M22 = m3 * (sa * sa + sb * sb) + I3;
        const double M33 = m4 * (e - ea) * (e - ea) + I4;
        const R64 M34 = m4 * ((e - ea) * (e - ea) + zt * (e - ea) * sin_Phi) + I4;
        const R64 M44 = m4 * (zt * zt + 2 * zt * (e - ea) * sin_Phi + (e - ea) * (e - ea)) +
                        m5 * (ta * ta + tb * tb) + I4 + I5;
        const double M55 = m6 * (zf - fa) * (zf - fa) + I6;
        const R64 M56 = m6 * ((zf - fa) * (zf - fa) - u * (zf - fa) * sin_Om) + I6;
        const R64 M66 = m6 * ((zf - fa) * (zf - fa) - 2 * u * (zf - fa) * sin_Om + u * u) +
                        m7 * (ua * ua + ub * ub) + I6 + I7;

        const R64 xd = sd * cos_ga + sc * sin_ga + xb;
        const R64 yd = sd * sin_ga - sc * cos_ga + yb;
        const R64 L = rsqrt64((xd - xc) * (xd - xc) + (yd - yc) * (yd - yc));
        const R64 Fs = -c0 * (L - l0) / L;
        const R64 Fx = Fs * (xd - xc), Fy = Fs * (yd - yc);
        const R64 h0 = mom - m2 * da * rr * v[1] * (v[1] + 2 * v[0]) * sin_Th;
        const R64 h1 = m2 * da * rr * v[0] * v[0] * sin_Th;
        const R64 h2 = Fx * (sc * cos_ga - sd * sin_ga) + Fy * (sd * cos_ga + sc * sin_ga);
        const R64 h3 = m4 * zt * (e - ea) * v[4] * v[4] * cos_Phi;
        const R64 h4 = -m4 * zt * (e - ea) * v[3] * (v[3] + 2 * v[4]) * cos_Phi;
        const R64 h5 = -m6 * u * (zf - fa) * v[6] * v[6] * cos_Om;
        const R64 h6 = m6 * u * (zf - fa) * v[5] * (v[5] + 2 * v[6]) * cos_Om;

        // constraint Jacobian G = g_q (6 x 7), only its structural non-zeros
        const R64 a0 = -rr * sin_be + d * sin_be_Th, a1 = d * sin_be_Th;
        const R64 b0 = rr * cos_be - d * cos_be_Th, b1 = -d * cos_be_Th;
        const R64 G02 = -ss * cos_ga, G12 = -ss * sin_ga;
        const R64 G23 = -e * cos_Ph_de, G24 = -e * cos_Ph_de + zt * sin_de;
        const R64 G33 = -e * sin_Ph_de, G34 = -e * sin_Ph_de - zt * cos_de;
        const R64 G45 = zf * sin_Om_ep, G46 = zf * sin_Om_ep - u * cos_ep;
        const R64 G55 = -zf * cos_Om_ep, G56 = -zf * cos_Om_ep - u * sin_ep;

        // G^T w for w = mu and w = la: the non-zero terms of every column summed in row order
        // (the oracle twin's loop adds exact zeros for the others)
        const R64 gtm0 = a0 * mu[0] + b0 * mu[1] + a0 * mu[2] + b0 * mu[3] + a0 * mu[4] + b0 * mu[5];
        const R64 gtl0 = a0 * la[0] + b0 * la[1] + a0 * la[2] + b0 * la[3] + a0 * la[4] + b0 * la[5];
        const R64 gtm1 = a1 * mu[0] + b1 * mu[1] + a1 * mu[2] + b1 * mu[3] + a1 * mu[4] + b1 * mu[5];
        const R64 gtl1 = a1 * la[0] + b1 * la[1] + a1 * la[2] + b1 * la[3] + a1 * la[4] + b1 * la[5];
        const R64 gtm2 = G02 * mu[0] + G12 * mu[1], gtl2 = G02 * la[0] + G12 * la[1];
        const R64 gtm3 = G23 * mu[2] + G33 * mu[3], gtl3 = G23 * la[2] + G33 * la[3];
        const R64 gtm4 = G24 * mu[2] + G34 * mu[3], gtl4 = G24 * la[2] + G34 * la[3];
        const R64 gtm5 = G45 * mu[4] + G55 * mu[5], gtl5 = G45 * la[4] + G55 * la[5];
        const R64 gtm6 = G46 * mu[4] + G56 * mu[5], gtl6 = G46 * la[4] + G56 * la[5];

        out[0] = (qd[0] - v[0] + gtm0).v;
        out[1] = (qd[1] - v[1] + gtm1).v;
        out[2] = (qd[2] - v[2] + gtm2).v;
        out[3] = (qd[3] - v[3] + gtm3).v;
        out[4] = (qd[4] - v[4] + gtm4).v;
        out[5] = (qd[5] - v[5] + gtm5).v;
        out[6] = (qd[6] - v[6] + gtm6).v;
        out[7] = ((M00 * vd[0] + M01 * vd[1]) - h0 - gtl0).v;
        out[8] = ((M01 * vd[0] + M11 * vd[1]) - h1 - gtl1).v;
        out[9] = ((M22 * vd[2]) - h2 - gtl2).v;
        out[10] = ((M33 * vd[3] + M34 * vd[4]) - h3 - gtl3).v;
        out[11] = ((M34 * vd[3] + M44 * vd[4]) - h4 - gtl4).v;
        out[12] = ((M55 * vd[5] + M56 * vd[6]) - h5 - gtl5).v;
        out[13] = ((M56 * vd[5] + M66 * vd[6]) - h6 - gtl6).v;
        const R64 ab0 = a0 * v[0] + a1 * v[1], ab1 = b0 * v[0] + b1 * v[1];
        out[14] = (ab0 + G02 * v[2]).v;
        out[15] = (ab1 + G12 * v[2]).v;
        out[16] = (ab0 + G23 * v[3] + G24 * v[4]).v;
        out[17] = (ab1 + G33 * v[3] + G34 * v[4]).v;
        out[18] = (ab0 + G45 * v[5] + G46 * v[6]).v;
        out[19] = (ab1 + G55 * v[5] + G56 * v[6]).v;
        out[20] = (rr * cos_be - d * cos_be_Th - ss * sin_ga - xb).v;
        out[21] = (rr * sin_be - d * sin_be_Th + ss * cos_ga - yb).v;
        out[22] = (rr * cos_be - d * cos_be_Th - e * sin_Ph_de - zt * cos_de - xa).v;
        out[23] = (rr * sin_be - d * sin_be_Th + e * cos_Ph_de - zt * sin_de - ya).v;
        out[24] = (rr * cos_be - d * cos_be_Th - zf * cos_Om_ep - u * sin_ep - xa).v;
        out[25] = (rr * sin_be - d * sin_be_Th - zf * sin_Om_ep + u * cos_ep - ya).v;
    }
};

}  // namespace b200dae
