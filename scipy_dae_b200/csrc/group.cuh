// group.cuh — G cooperating lanes per system (G = 8, 16 or 32): lane r owns row r of the
// iteration matrices and component r of every state vector.
//
// Used for systems that do not fit one thread's registers (n > 4).  The reference factors and
// solves with LAPACK getrf / getrs through scipy.linalg.lu_factor / lu_solve (base.py:232-238);
// here the same elimination (row partial pivoting, first maximal element, rows interchanged in
// full) is spread over the lanes of a group:
//   * factorisation: each lane keeps the not-yet-eliminated part of ITS row in registers, rotated
//     one entry per step so that the pivot column is always register 0 (static register indices
//     in a ROLLED loop over the steps: the fully unrolled variant was 240 KB of straight-line
//     code and spent 73 % of its warp time on instruction-cache misses, ncu round 1).  The pivot
//     row of step k goes to its final place in the shared-memory factor matrix, from where every
//     remaining lane reads it (broadcast) to eliminate its own row; multipliers are written to
//     the row of the lane's CURRENT position and the already finished parts of two rows are
//     exchanged when LAPACK would interchange them, so the matrix ends in getrf's layout
//     (L below, U above the diagonal, rows in pivot order); the diagonal holds 1 / U[k,k].
//   * pivot search: |x| of a double orders like its bit pattern, so argmax is two integer
//     redux.sync (high word, low word) plus a min-position redux for LAPACK's first-maximum
//     tie-break.
//   * triangular solves: column-oriented substitution; the finished entry of step k is
//     broadcast by one shuffle and every other lane applies its own L / U entry.  The real and
//     the complex systems of one Newton iteration are independent and are interleaved in one
//     loop so their shuffle latencies overlap.
// All sync primitives take the group's lane mask, so the 32 / G groups of a warp run
// independent systems with independent control flow.
#pragma once
#ifndef __CUDACC_RTC__
#include <cuda_runtime.h>
#endif
#include "dense_lu.cuh"

// Unroll factor of the substitution loops.  Swept on B200 (solves/s, amplifier Radau s=3 / Andrews
// Radau s=5): 1 -> 60.8 k / 91.6 k, 2 -> 63.4 k / 93.7 k, 4 (Andrews) -> 90.8 k, full for n <= 8 -> 56.7 k.
// 0 (default): the rotated-register LU below; 1: the in-place shared-memory variants (measured slower, see there)
#ifndef B200DAE_LU_INPLACE
#define B200DAE_LU_INPLACE 0
#endif
#ifndef B200DAE_LU_INPLACE_UNROLL
#define B200DAE_LU_INPLACE_UNROLL 2
#endif
#ifndef B200DAE_SOLVE_UNROLL
#define B200DAE_SOLVE_UNROLL(N) 2
#endif

namespace b200dae {

template <int G>
struct Group {
    static_assert(G == 4 || G == 8 || G == 16 || G == 32, "group size");
    unsigned mask;
    int lane;          // 0..G-1 inside the group
    __device__ __forceinline__ Group() {
        const unsigned l = threadIdx.x & 31u;
        lane = (int)(l & (G - 1));
        mask = (G == 32) ? 0xffffffffu : (((1u << (G & 31)) - 1u) << (l & ~(unsigned)(G - 1)));
    }
    __device__ __forceinline__ void sync() const { __syncwarp(mask); }
    __device__ __forceinline__ double bcast(double v, int src) const { return __shfl_sync(mask, v, src, G); }
    __device__ __forceinline__ int bcast(int v, int src) const { return __shfl_sync(mask, v, src, G); }
    // butterfly all-reduce: every lane ends with the bit-identical sum (x + y == y + x)
    __device__ __forceinline__ double sum(double v) const {
#pragma unroll
        for (int o = G / 2; o > 0; o >>= 1) v += __shfl_xor_sync(mask, v, o, G);
        return v;
    }
    __device__ __forceinline__ double max(double v) const {
#pragma unroll
        for (int o = G / 2; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(mask, v, o, G));
        return v;
    }
    __device__ __forceinline__ bool all(bool p) const { return __all_sync(mask, p) != 0; }
    __device__ __forceinline__ bool any(bool p) const { return __any_sync(mask, p) != 0; }
    __device__ __forceinline__ unsigned ballot(bool p) const { return __ballot_sync(mask, p); }
    __device__ __forceinline__ int rmax(int v) const { return __reduce_max_sync(mask, v); }
    __device__ __forceinline__ unsigned rmax(unsigned v) const { return __reduce_max_sync(mask, v); }
    __device__ __forceinline__ int rmin(int v) const { return __reduce_min_sync(mask, v); }
};

// One pivot search: among lanes with `cand`, the one with the largest key v >= 0 (first position
// on ties).  Returns true on the winning lane; ppos = its position.
template <class Grp>
__device__ __forceinline__ bool pivot_search(const Grp& g, bool cand, double v, int pos, int& ppos) {
    const int hi = cand ? __double2hiint(v) : -1;
    const int mhi = g.rmax(hi);
    bool match = cand && hi == mhi;
#ifdef B200DAE_SLOW_PIVOT
    {
        const unsigned lo = match ? (unsigned)__double2loint(v) : 0u;
        const unsigned mlo = g.rmax(lo);
        match = match && lo == mlo;
        ppos = g.rmin(match ? pos : (1 << 20));
        return match && pos == ppos;
    }
#endif
    unsigned b = g.ballot(match);
    if (b & (b - 1u)) {                                 // several lanes share the high word (rare)
        const unsigned lo = match ? (unsigned)__double2loint(v) : 0u;
        const unsigned mlo = g.rmax(lo);
        match = match && lo == mlo;
        b = g.ballot(match);
        if (b & (b - 1u)) {                             // exact ties: LAPACK takes the first position
            ppos = g.rmin(match ? pos : (1 << 20));
            return match && pos == ppos;
        }
    }
    ppos = __shfl_sync(g.mask, pos, __ffs(b) - 1);
    return match;
}

// (Tried, B200 round 1: splitting the step loop into four ranges with narrower unrolled bodies,
// which cuts the FMA count from N^2 to 0.63 N^2.  The extra code cost more than the FMAs saved:
// Andrews 91 -> 85 k solves/s, amplifier 57 -> 44 k solves/s.  These kernels are bound by
// instruction fetch and latency, not by the FP64 pipe, so the single rolled loop stays.  A two-range
// split for n >= 16 only (-25 % FMAs, +50 % code in the elimination body) was neutral for Andrews:
// 93.8 k vs 94.3 k solves/s.)
//
// Real LU.  a[0..N-1] = row `lane` of the matrix (lanes >= N: ignored).  On return
// mat[k * LD + j] holds the factors in getrf layout with the reciprocal pivots on the diagonal and
// perm = the original row that ended at position `lane`.  `mat` needs N * LD + N doubles (the
// rotated elimination reads up to N entries past a row's end; those values are never used).
template <int N, int LD, class Grp>
__device__ __forceinline__ void group_lu_factor(const Grp& g, double (&a)[N], int& perm,
                                                double* __restrict__ mat, int* __restrict__ sidx) {
    const int lane = g.lane;
    const bool act = lane < N;
    int pos = lane;
    double my_rd = 0.0;
#pragma unroll 1
    for (int k = 0; k < N; ++k) {
        const bool cand = act && pos >= k;
        int ppos;
        const bool is_piv = pivot_search(g, cand, fabs(a[0]), pos, ppos);
        g.sync();                                       // multipliers of step k-1 are visible
        if (ppos != k && lane < k) {                    // interchange the finished parts (columns < k)
            const double x = mat[k * LD + lane];
            mat[k * LD + lane] = mat[ppos * LD + lane];
            mat[ppos * LD + lane] = x;
        }
        if (pos == k) pos = ppos;
        if (is_piv) {
            pos = k;
            double* row = mat + k * LD + k;
#pragma unroll
            for (int j = 0; j < N; ++j)
                if (j < N - k) row[j] = a[j];
        }
        g.sync();
        const double* row = mat + k * LD + k;
        const double pv = row[0];
        const double r = fast_rcp(pv);
        const double rs = (pv != 0.0) ? r : 1.0;       // LAPACK skips the scaling for a zero pivot
        if (is_piv) my_rd = r;
        if (cand && !is_piv) {
            const double l = a[0] * rs;
            mat[pos * LD + k] = l;
#pragma unroll
            for (int j = 0; j < N - 1; ++j) a[j] = fma(-l, row[j + 1], a[j + 1]);
        }
    }
    g.sync();
    if (act) { mat[pos * LD + pos] = my_rd; sidx[pos] = lane; }
    g.sync();
    perm = act ? sidx[lane] : lane;
    g.sync();                                           // sidx may be reused by the caller
}

// Complex LU (split re / im row in registers on entry).  Factors in shared memory, interleaved:
// cmat[k * LDC + j] = (re, im); same layout and padding rule as the real routine.
template <int N, int LDC, class Grp>
__device__ __forceinline__ void group_zlu_factor(const Grp& g, double (&ar)[N], double (&ai)[N], int& perm,
                                                 double2* __restrict__ cmat, int* __restrict__ sidx) {
    const int lane = g.lane;
    const bool act = lane < N;
    int pos = lane;
    double my_r = 0.0, my_i = 0.0;
#pragma unroll 1
    for (int k = 0; k < N; ++k) {
        const bool cand = act && pos >= k;
        int ppos;
        const bool is_piv = pivot_search(g, cand, fabs(ar[0]) + fabs(ai[0]), pos, ppos);   // izamax: |re| + |im|
        g.sync();
        if (ppos != k && lane < k) {
            const double2 x = cmat[k * LDC + lane];
            cmat[k * LDC + lane] = cmat[ppos * LDC + lane];
            cmat[ppos * LDC + lane] = x;
        }
        if (pos == k) pos = ppos;
        if (is_piv) {
            pos = k;
            double2* row = cmat + k * LDC + k;
#pragma unroll
            for (int j = 0; j < N; ++j)
                if (j < N - k) row[j] = make_double2(ar[j], ai[j]);
        }
        g.sync();
        const double2* row = cmat + k * LDC + k;
        const double2 pv = row[0];
        double rr, ri;
        crecip(pv.x, pv.y, rr, ri);
        const bool nz = (pv.x != 0.0) || (pv.y != 0.0);
        const double sr = nz ? rr : 1.0, si = nz ? ri : 0.0;
        if (is_piv) { my_r = rr; my_i = ri; }
        if (cand && !is_piv) {
            const double xr = ar[0], xi = ai[0];
            const double lr = fma(xr, sr, -xi * si), li = fma(xr, si, xi * sr);
            cmat[pos * LDC + k] = make_double2(lr, li);
#pragma unroll
            for (int j = 0; j < N - 1; ++j) {
                const double2 u = row[j + 1];
                ar[j] = fma(li, u.y, fma(-lr, u.x, ar[j + 1]));
                ai[j] = fma(-li, u.x, fma(-lr, u.y, ai[j + 1]));
            }
        }
    }
    g.sync();
    if (act) { cmat[pos * LDC + pos] = make_double2(my_r, my_i); sidx[pos] = lane; }
    g.sync();
    perm = act ? sidx[lane] : lane;
    g.sync();
}

// ---------------------------------------------------------------------------------------------
// In-place variants (-DB200DAE_LU_INPLACE=1; compiled out by default): the matrix is assembled straight
// into the shared-memory factor storage and eliminated there.  Lane r owns the row at POSITION r for
// the whole factorisation: an interchange swaps the two rows physically (every lane moves one
// column), the elimination of step k touches columns k+1 .. N-1 only.  Against the rotated-register
// form above this halves the multiply-adds (sum (N-k-1) instead of N (N-1) per lane), drops the
// pivot-row write (N predicated stores by ONE lane per step: 10.7 % of all instructions of the
// Andrews kernel, ncu round 2) and the 2N..4N registers that held the rows; same operations on the
// same values in the same order per entry, so the factors are bit-identical (verified: identical work
// counters on every ensemble).  MEASURED SLOWER on B200 all the same: Andrews Radau s = 5 103.3 k
// (rotated) against 94.8 / 94.4 / 97.5 / 96.8 k solves/s (in place, inner loop unrolled 1 / 2 / 4 / 8),
// amplifier 80.8 -> 80.4 k (Radau), 58.6 -> 57.3 k (BDF).  These kernels are latency-bound: the in-place
// update puts a shared-memory load AND store into every entry's dependency chain, the rotated form
// only a broadcast load, and its extra multiply-adds issue into idle slots.
template <int N, int LD, class Grp>
__device__ __forceinline__ void group_lu_factor_inplace(const Grp& g, int& perm, double* __restrict__ mat) {
    constexpr int kLuUnroll = B200DAE_LU_INPLACE_UNROLL;
    const int lane = g.lane;
    const bool act = lane < N;
    int orig = lane;
    double* myrow = mat + (act ? lane : 0) * LD;
#pragma unroll 1
    for (int k = 0; k < N; ++k) {
        const bool cand = act && lane >= k;
        double v = cand ? myrow[k] : 0.0;
        int ppos;
        pivot_search(g, cand, fabs(v), lane, ppos);
        if (ppos != k) {                                 // same decision in every lane of the group
            if (act) {                                   // LAPACK interchanges the full rows
                const double x = mat[k * LD + lane], z = mat[ppos * LD + lane];
                mat[k * LD + lane] = z;
                mat[ppos * LD + lane] = x;
            }
            const int ok = g.bcast(orig, k), op = g.bcast(orig, ppos);
            if (lane == k) orig = op;
            else if (lane == ppos) orig = ok;
            g.sync();
            if (cand) v = myrow[k];
        }
        const double* prow = mat + k * LD;
        const double pv = prow[k];
        const double r = fast_rcp(pv);
        const double rs = (pv != 0.0) ? r : 1.0;       // LAPACK skips the scaling for a zero pivot
        if (act && lane > k) {
            const double l = v * rs;
            myrow[k] = l;
#pragma unroll kLuUnroll
            for (int j = k + 1; j < N; ++j) myrow[j] = fma(-l, prow[j], myrow[j]);
        }
        g.sync();
        if (lane == k) myrow[k] = r;                    // reciprocal pivot (not read again before the solves)
    }
    g.sync();
    perm = orig;
}

template <int N, int LDC, class Grp>
__device__ __forceinline__ void group_zlu_factor_inplace(const Grp& g, int& perm, double2* __restrict__ cmat) {
    constexpr int kLuUnroll = B200DAE_LU_INPLACE_UNROLL;
    const int lane = g.lane;
    const bool act = lane < N;
    int orig = lane;
    double2* myrow = cmat + (act ? lane : 0) * LDC;
#pragma unroll 1
    for (int k = 0; k < N; ++k) {
        const bool cand = act && lane >= k;
        double2 v = cand ? myrow[k] : make_double2(0.0, 0.0);
        int ppos;
        pivot_search(g, cand, fabs(v.x) + fabs(v.y), lane, ppos);   // izamax: |re| + |im|
        if (ppos != k) {
            if (act) {
                const double2 x = cmat[k * LDC + lane], z = cmat[ppos * LDC + lane];
                cmat[k * LDC + lane] = z;
                cmat[ppos * LDC + lane] = x;
            }
            const int ok = g.bcast(orig, k), op = g.bcast(orig, ppos);
            if (lane == k) orig = op;
            else if (lane == ppos) orig = ok;
            g.sync();
            if (cand) v = myrow[k];
        }
        const double2* prow = cmat + k * LDC;
        const double2 pv = prow[k];
        double rr, ri;
        crecip(pv.x, pv.y, rr, ri);
        const bool nz = (pv.x != 0.0) || (pv.y != 0.0);
        const double sr = nz ? rr : 1.0, si = nz ? ri : 0.0;
        if (act && lane > k) {
            const double lr = fma(v.x, sr, -v.y * si), li = fma(v.x, si, v.y * sr);
            myrow[k] = make_double2(lr, li);
#pragma unroll kLuUnroll
            for (int j = k + 1; j < N; ++j) {
                const double2 u = prow[j];
                double2 a = myrow[j];
                a.x = fma(li, u.y, fma(-lr, u.x, a.x));
                a.y = fma(-li, u.x, fma(-lr, u.y, a.y));
                myrow[j] = a;
            }
        }
        g.sync();
        if (lane == k) myrow[k] = make_double2(rr, ri);
    }
    g.sync();
    perm = orig;
}

// Solves LU x = b for one real system and NC complex systems at once (same step index k, so
// the NC + 1 independent shuffle / FMA chains overlap).  Lane `lane` holds component `lane` of
// every right-hand side on entry and of every solution on return.  rrow / crow[c] = this lane's
// row of the factor matrices in shared memory; perm / cperm[c] from the factorisations.
// kReal = false skips the real system; NC may be 0.
template <int N, int NC, bool kReal, class Grp>
__device__ __forceinline__ void group_lu_solve(const Grp& g, const double* __restrict__ rrow, int perm,
                                               double& b,
                                               const double2* (&crow)[NC > 0 ? NC : 1],
                                               const int (&cperm)[NC > 0 ? NC : 1],
                                               double (&br)[NC > 0 ? NC : 1], double (&bi)[NC > 0 ? NC : 1]) {
    const int lane = g.lane;
    if (kReal) b = g.bcast(b, perm);
#pragma unroll
    for (int c = 0; c < NC; ++c) {
        br[c] = g.bcast(br[c], cperm[c]);
        bi[c] = g.bcast(bi[c], cperm[c]);
    }
    // forward substitution with the unit lower factor
    constexpr int kSolveUnroll = B200DAE_SOLVE_UNROLL(N);
#pragma unroll kSolveUnroll
    for (int k = 0; k < N - 1; ++k) {
        const bool below = lane > k && lane < N;
        if (kReal) {
            const double bk = g.bcast(b, k);
            if (below) b = fma(-rrow[k], bk, b);
        }
#pragma unroll
        for (int c = 0; c < NC; ++c) {
            const double zr = g.bcast(br[c], k), zi = g.bcast(bi[c], k);
            if (below) {
                const double2 l = crow[c][k];
                br[c] = fma(l.y, zi, fma(-l.x, zr, br[c]));
                bi[c] = fma(-l.y, zr, fma(-l.x, zi, bi[c]));
            }
        }
    }
    // back substitution with U (reciprocal pivots on the diagonal)
#pragma unroll kSolveUnroll
    for (int k = N - 1; k >= 0; --k) {
        const bool above = lane < k;
        if (kReal) {
            if (lane == k) b *= rrow[k];
            const double bk = g.bcast(b, k);
            if (above) b = fma(-rrow[k], bk, b);
        }
#pragma unroll
        for (int c = 0; c < NC; ++c) {
            if (lane == k) {
                const double2 dg = crow[c][k];
                const double sr = br[c], si = bi[c];
                br[c] = fma(sr, dg.x, -si * dg.y);
                bi[c] = fma(sr, dg.y, si * dg.x);
            }
            const double zr = g.bcast(br[c], k), zi = g.bcast(bi[c], k);
            if (above) {
                const double2 u = crow[c][k];
                br[c] = fma(u.y, zi, fma(-u.x, zr, br[c]));
                bi[c] = fma(-u.y, zr, fma(-u.x, zi, bi[c]));
            }
        }
    }
}

// real system only
template <int N, class Grp>
__device__ __forceinline__ void group_lu_solve_real(const Grp& g, const double* __restrict__ rrow, int perm,
                                                    double& b) {
    const double2* nocrow[1] = {nullptr};
    const int zi1[1] = {0};
    double zb[1] = {0.0}, zc[1] = {0.0};
    group_lu_solve<N, 0, true>(g, rrow, perm, b, nocrow, zi1, zb, zc);
}

}  // namespace b200dae
