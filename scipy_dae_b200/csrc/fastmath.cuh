// fastmath.cuh — branch-free FP64 reciprocal and square root for sm_100a.
//
// Measured on B200 (profiles/microbench/fp64_latency.cu): a dependent DFMA costs 8.7 cycles, an
// IEEE `1.0 / x` 80 cycles and `sqrt(x)` 100 cycles of warp time, and four INDEPENDENT divisions
// still cost 4 x 74 cycles because each carries a slow-path branch that the scheduler cannot
// interleave across.  The integrator makes ~3000 divisions and ~700 square roots per Robertson
// solve (LU pivots, error scales, convergence rates, FD steps), so those sequences were 20 % of
// the warp lifetime.  The versions here are straight-line code (MUFU seed + DFMA refinement,
// special cases by select), so independent ones overlap and a dependent one costs ~55 cycles.
// Accuracy: <= 1 ulp (tests/test_gpu_fastmath.py); the IEEE results are reproduced for
// 0, inf, NaN and denormal arguments where the integrator relies on them (zero pivot -> inf).
#pragma once
#include <math.h>

namespace b200dae {

__device__ __forceinline__ double rcp_seed(double x) {
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    return r;
}

__device__ __forceinline__ double rsqrt_seed(double x) {
    double r;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    return r;
}

#ifndef B200DAE_FASTMATH_LITE
#define B200DAE_FASTMATH_LITE 1
#endif

// 1 / x.  x = 0 or denormal -> +-inf, x = +-inf -> +-0, NaN -> NaN.
// The seed is good to 2^-20 or better, so r1 = r0 (1 + e + e^2) already carries a relative error of
// e^3 < 2^-60 plus one rounding (B200DAE_FASTMATH_LITE stops there: <= 1 ulp); the second
// refinement of the full version only repairs that rounding.
__device__ __forceinline__ double fast_rcp(double x) {
    const double r0 = rcp_seed(x);
    const double e0 = fma(-x, r0, 1.0);
    const double e1 = fma(e0, e0, e0);          // e + e^2
    const double r1 = fma(r0, e1, r0);          // r0 (1 + e + e^2): error e^3
#if B200DAE_FASTMATH_LITE
    return (fabs(e0) < 0.25) ? r1 : r0;         // seed already exact for 0 / inf / NaN
#else
    const double e2 = fma(-x, r1, 1.0);
    const double r2 = fma(r1, e2, r1);
    return (fabs(e0) < 0.25) ? r2 : r0;         // seed already exact for 0 / inf / NaN
#endif
}

// 1 / x for an x the caller knows to be finite, normal and non-zero (error scales, step sizes,
// small integers): no special-case select.  Anything else gives NaN instead of the IEEE limit.
__device__ __forceinline__ double fast_rcp_nz(double x) {
#if B200DAE_FASTMATH_LITE
    const double r0 = rcp_seed(x);
    const double e0 = fma(-x, r0, 1.0);
    const double e1 = fma(e0, e0, e0);
    return fma(r0, e1, r0);
#else
    return fast_rcp(x);
#endif
}

__device__ __forceinline__ double fast_div(double a, double b) { return a * fast_rcp(b); }

#ifndef B200DAE_OPT2
#define B200DAE_OPT2 0
#endif

// sqrt(x), correct for 0, inf, NaN, negative and denormal arguments.
// Arguments below 1e-200 (denormals would flush in the seed) are rescaled first: by selects in
// the baseline, by a never-taken branch with B200DAE_OPT2 (6 instructions less per call; the
// integrator's arguments are error norms, which are either 0 or far above 1e-200).
__device__ __forceinline__ double fast_sqrt(double x) {
#if B200DAE_OPT2 & 1
    if (__builtin_expect(x < 1e-200 && x > 0.0, 0)) return sqrt(x);
    const double xs = x;
    constexpr bool tiny = false;
#else
    const bool tiny = x < 1e-200;
    const double xs = tiny ? x * 0x1p+600 : x;              // also normalises denormals
#endif
    const double r0 = rsqrt_seed(xs);
    double g = xs * r0, h = 0.5 * r0;
    double e = fma(-h, g, 0.5);
    g = fma(g, e, g); h = fma(h, e, h);
#if !B200DAE_FASTMATH_LITE
    // (lite: after ONE coupled step g and h are good to ~2^-39, and the correction below is itself
    // a Newton step on g, so it lands on ~2^-78 before rounding: the second step is redundant)
    e = fma(-h, g, 0.5);
    g = fma(g, e, g); h = fma(h, e, h);
#endif
    const double d = fma(-g, g, xs);
    g = fma(d, h, g);
    g = tiny ? g * 0x1p-300 : g;
    return (xs > 0.0 && xs < INFINITY) ? g : ((x < 0.0) ? nan("") : x);
}

// sqrt(x) and 1/sqrt(x) from one refinement (the coupled iteration carries h ~ 1/(2 sqrt x)).
// For x = 0 returns (0, inf); for x = inf (inf, 0); NaN / negative -> NaN.
__device__ __forceinline__ double fast_sqrt_rsqrt(double x, double& rsq) {
#if B200DAE_OPT2 & 1
    if (__builtin_expect(x < 1e-200 && x > 0.0, 0)) { const double g0 = sqrt(x); rsq = 1.0 / g0; return g0; }
    const double xs = x;
    constexpr bool tiny = false;
#else
    const bool tiny = x < 1e-200;
    const double xs = tiny ? x * 0x1p+600 : x;
#endif
    const double r0 = rsqrt_seed(xs);
    double g = xs * r0, h = 0.5 * r0;
    double e = fma(-h, g, 0.5);
    g = fma(g, e, g); h = fma(h, e, h);
#if !B200DAE_FASTMATH_LITE
    e = fma(-h, g, 0.5);
    g = fma(g, e, g); h = fma(h, e, h);
#endif
    const double d = fma(-g, g, xs);
    g = fma(d, h, g);
    e = fma(-h, g, 0.5);
    h = fma(h, e, h);
    const bool ok = (xs > 0.0 && xs < INFINITY);
    rsq = ok ? (tiny ? 2.0 * h * 0x1p+300 : 2.0 * h) : ((x < 0.0) ? nan("") : r0);
    g = tiny ? g * 0x1p-300 : g;
    return ok ? g : ((x < 0.0) ? nan("") : x);
}

// x ** k for 0 <= k < 32 without a loop
__device__ __forceinline__ double powi32(double x, int k) {
    const double x2 = x * x, x4 = x2 * x2, x8 = x4 * x4, x16 = x8 * x8;
    double r = (k & 1) ? x : 1.0;
    r *= (k & 2) ? x2 : 1.0;
    r *= (k & 4) ? x4 : 1.0;
    r *= (k & 8) ? x8 : 1.0;
    r *= (k & 16) ? x16 : 1.0;
    return r;
}

// x ** k for 0 <= k < 8
__device__ __forceinline__ double powi8(double x, int k) {
    const double x2 = x * x, x4 = x2 * x2;
    double r = (k & 1) ? x : 1.0;
    r *= (k & 2) ? x2 : 1.0;
    r *= (k & 4) ? x4 : 1.0;
    return r;
}

// all-finite test on the integer pipe: accumulates min over (~hi & 0x7ff00000); 0 <=> some value
// has an all-ones exponent (inf or NaN).  Keeps 9..35 DSETPs per Newton iteration off the FP64 pipe.
struct FiniteAcc {
    unsigned m = 0x7ff00000u;
    __device__ __forceinline__ void add(double x) {
        const unsigned t = (~(unsigned)__double2hiint(x)) & 0x7ff00000u;
        m = min(m, t);
    }
    __device__ __forceinline__ bool all_finite() const { return m != 0u; }
};

}  // namespace b200dae
