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
#include "rt_compat.cuh"

namespace b200dae {

// max / min as one compare and one select.  fmax / fmin compile to an 8-instruction IEEE-754
// maxNum sequence on sm_100a (DSETP.MAX, moves, FSEL, NaN quieting), and the integrator calls
// them ~3000 times per Robertson solve (num_jac scales, error scales): 8 % of its instructions.
// NaN: the SECOND argument is returned when the comparison is unordered, so a NaN there
// propagates (like np.maximum) and a NaN in the first is dropped; call sites order their
// arguments accordingly.
__device__ __forceinline__ double max_sel(double a, double b) { return (a > b) ? a : b; }
__device__ __forceinline__ double min_sel(double a, double b) { return (a < b) ? a : b; }
// max(|a|, |b|)
__device__ __forceinline__ double amax_sel(double a, double b) {
    // the select works on the raw words and the sign is cleared on the selected high word, so no
    // fabs() is materialised on the FP64 pipe (|.| on the compare operands is free)
    const bool gt = fabs(a) > fabs(b);
    const int hi = (gt ? __double2hiint(a) : __double2hiint(b)) & 0x7fffffff;
    const int lo = gt ? __double2loint(a) : __double2loint(b);
    return __hiloint2double(hi, lo);
}

// max(|a|, |b|) with the sign of num_jac's f_sign = 2 (f >= 0) - 1 (negative for NaN, positive for -0)
__device__ __forceinline__ double amax_signed(double a, double b, double f) {
    const bool gt = fabs(a) > fabs(b);
    const int hi = ((gt ? __double2hiint(a) : __double2hiint(b)) & 0x7fffffff) | ((f >= 0.0) ? 0 : (int)0x80000000);
    const int lo = gt ? __double2loint(a) : __double2loint(b);
    return __hiloint2double(hi, lo);
}

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
#ifndef B200DAE_RCP_IGUARD
#define B200DAE_RCP_IGUARD 0
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
#if B200DAE_RCP_IGUARD
    // |e0| < 0.25 on the integer pipe: 0.25 = 0x3fd0000000000000 has a zero low word, so the high word
    // of |e0| decides (NaN / inf compare above it); keeps ~1500 DSETPs per Robertson solve off the FP64 pipe
    return ((unsigned)(__double2hiint(e0) & 0x7fffffff) < 0x3fd00000u) ? r1 : r0;
#else
    return (fabs(e0) < 0.25) ? r1 : r0;         // seed already exact for 0 / inf / NaN
#endif
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

// 1 / x where the IEEE limits for x = 0 / inf do not change what happens next (B200DAE_RCP_LIMITS=0):
//   * the Newton scale atol + rtol |y| (zero only for atol = 0 and y = 0, where the reference divides
//     by zero too: the increment norm is non-finite either way and the Newton iteration fails; the
//     error-estimate scale keeps the IEEE limits, because inf rejects a step where NaN would accept it),
//   * LU pivots: a zero pivot leaves inf in the reference's factors and NaN here; both make the next
//     triangular solve non-finite, which the non-finite test after the solves catches (the multipliers
//     use the separate "LAPACK skips the scaling" select, not this value).
#ifndef B200DAE_RCP_LIMITS
#define B200DAE_RCP_LIMITS 0
#endif
__device__ __forceinline__ double fast_rcp_scale(double x) {
#if B200DAE_RCP_LIMITS
    return fast_rcp(x);
#else
    return fast_rcp_nz(x);
#endif
}

// sqrt(x), correct for 0, inf, NaN, negative and denormal arguments.
// Arguments below 1e-200 (denormals would flush in the seed) are rescaled first, by selects.
// (Tried on B200: a never-taken branch for that case instead of the selects, 6 instructions
// less per call: 13.83 ms instead of 13.61 ms per 2^20 Robertson solves.  The same happened with
// a uniform branch to a 3-bit powi for newton_max_iter - k < 8 (13.71 ms) and with an integer
// exponent guard + branch in crecip (13.76 ms): every extra basic-block boundary costs the
// scheduler more than the instructions it saves, so the straight-line forms stay.)
//
// kTiny = false drops the rescaling: arguments in (0, 1e-200) may then come back as 0 (and 1/sqrt
// as inf).  For callers to which "below 1e-100" and "zero" mean the same (an error norm that only
// feeds predict_factor, where both saturate at 1 + pi/2).
//
// positive, normal and finite <=> biased exponent in [1, 0x7fe] with a clear sign bit: one
// subtraction and one unsigned compare on the high word.
__device__ __forceinline__ bool pos_normal(double xs) {
    return (unsigned)(__double2hiint(xs) - 0x00100000) < 0x7fe00000u;
}

template <bool kTiny = true>
__device__ __forceinline__ double fast_sqrt(double x) {
    const bool tiny = kTiny && x < 1e-200;
    const double xs = tiny ? x * 0x1p+600 : x;              // also normalises denormals
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
    if (!kTiny) return pos_normal(xs) ? g : ((x < 0.0) ? nan("") : ((x < 1.0) ? 0.0 : x));
    return pos_normal(xs) ? g : ((x < 0.0) ? nan("") : x);
}

// sqrt(x) and 1/sqrt(x) from one refinement (the coupled iteration carries h ~ 1/(2 sqrt x)).
// For x = 0 returns (0, inf); for x = inf (inf, 0); NaN / negative -> NaN.
template <bool kTiny = true>
__device__ __forceinline__ double fast_sqrt_rsqrt(double x, double& rsq) {
    const bool tiny = kTiny && x < 1e-200;
    const double xs = tiny ? x * 0x1p+600 : x;
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
    const bool ok = pos_normal(xs);
    if (!kTiny) {
        // denormal x: the seed flushed it to zero and returned inf, which is the x = 0 answer
        rsq = ok ? 2.0 * h : ((x < 0.0) ? nan("") : r0);
        return ok ? g : ((x < 0.0) ? nan("") : ((x < 1.0) ? 0.0 : x));
    }
    rsq = ok ? (tiny ? 2.0 * h * 0x1p+300 : 2.0 * h) : ((x < 0.0) ? nan("") : r0);
    g = tiny ? g * 0x1p-300 : g;
    return ok ? g : ((x < 0.0) ? nan("") : x);
}

// nextafter(t, dir * inf) for finite t (radau.py:489, bdf.py:269) on the integer pipe: one step of the bit
// pattern away from zero when t and dir have the same sign, towards zero otherwise.
__device__ __forceinline__ double next_toward(double t, double dir) {
#if B200DAE_FASTMATH_LITE
    const long long b = __double_as_longlong(t);
    const bool away = (t > 0.0) == (dir > 0.0);
    const double nt = __longlong_as_double(b + (away ? 1LL : -1LL));
    return (t == 0.0) ? copysign(4.9406564584124654e-324, dir) : nt;
#else
    return nextafter(t, dir * INFINITY);
#endif
}

// x ** k for 0 <= k < 2^kBits without a loop
template <int kBits>
__device__ __forceinline__ double powi_bits(double x, int k) {
    double sq = x, r = (k & 1) ? x : 1.0;
#pragma unroll
    for (int b = 1; b < kBits; ++b) {
        sq *= sq;
        r *= (k & (1 << b)) ? sq : 1.0;
    }
    return r;
}
__device__ __forceinline__ double powi32(double x, int k) { return powi_bits<5>(x, k); }

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
