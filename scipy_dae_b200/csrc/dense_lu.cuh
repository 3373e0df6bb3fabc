// dense_lu.cuh — per-instance dense LU with row partial pivoting and triangular solves, real and
// complex, one system per thread.
//
// Replaces scipy.linalg.lu_factor / lu_solve as called by the reference (base.py:232-238), i.e.
// LAPACK dgetrf/dgetrs for the real blocks and zgetrf/zgetrs for the complex Radau blocks:
// first-maximal-element pivoting (idamax on |x|, izamax on |re|+|im|), rows swapped in full,
// column scaled by the reciprocal of the pivot, a zero pivot is left in place (info > 0 in
// LAPACK; the solve then produces inf/nan exactly like getrs).
//
// Two storage regimes, chosen at compile time:
//   REG  (N <= 4): the matrix is a fully unrolled local array -> registers.  Row swaps and the
//                  right-hand-side permutation are predicated selects, never dynamic indexing
//                  (which would demote the arrays to local memory).
//   !REG (N  > 4): the matrix lives in (L1-resident) local memory, loops are rolled, swaps index
//                  dynamically.
// The diagonal of U is stored as its reciprocal so the solves multiply instead of divide.
#pragma once
#include "rt_compat.cuh"
#include "fastmath.cuh"

namespace b200dae {

// View of a per-thread vector that lives either in a local array (stride 1 -> registers after
// unrolling) or in shared memory laid out [element][thread] (stride = threads per block, so the
// 32 lanes of a warp hit 32 consecutive 8-byte words: conflict-free).
template <bool SH, int STRIDE>
struct Vec {
    double* p;
    __device__ __forceinline__ double& operator[](int k) const { return p[SH ? k * STRIDE : k]; }
    __device__ __forceinline__ Vec operator+(int off) const { return Vec{p + (SH ? off * STRIDE : off)}; }
};

// 1 (default): LINPACK-style partial interchanges in the unrolled regime beyond n = 4, see lu_factor.  Measured
// on B200: Akzo Nobel (n = 6) BDF 6.12 -> 6.35 M, Radau 8.62 -> 9.13 M solves/s; for n = 3 the interleaved
// interchange / sweep chain of the solve lengthens the critical path of the headline kernel (10.70 -> 11.82 ms),
// so n <= 4 keeps LAPACK's full-row interchanges.
#ifndef B200DAE_LINPACK_SWAPS
#define B200DAE_LINPACK_SWAPS 1
#endif
template <int N> struct LinpackSwaps { static constexpr bool value = (B200DAE_LINPACK_SWAPS != 0) && N > 4; };
#ifndef B200DAE_REG_MAXN
#define B200DAE_REG_MAXN 6
#endif
template <int N>
struct Regime {
    static constexpr bool kReg = (N <= B200DAE_REG_MAXN);
    static constexpr int kUnroll = kReg ? 64 : 1;   // "#pragma unroll kUnroll": full vs none
};

// ---------------------------------------------------------------------------------------------
// real
// ---------------------------------------------------------------------------------------------
// kFirstDone: the rows already stand in the order LAPACK's first interchange produces and piv[0]
// is set (pivoted assembly, see first_pivot_order): the search and the swap of step 0 are skipped.
template <int N, bool kFirstDone = false>
__device__ __forceinline__ void lu_factor(double* __restrict__ a, int* __restrict__ piv) {
    constexpr bool REG = Regime<N>::kReg;
    constexpr int U = Regime<N>::kUnroll;
#pragma unroll U
    for (int k = 0; k < N; ++k) {
        int p = k;
        if (!(kFirstDone && k == 0)) {
            double best = fabs(a[k * N + k]);
#pragma unroll U
            for (int i = k + 1; i < N; ++i) {
                const double v = fabs(a[i * N + k]);
                if (v > best) { best = v; p = i; }
            }
            piv[k] = p;
        }
        if (kFirstDone && k == 0) {
            // nothing to move
        } else if (REG) {
            // register regime: an interchange is a select per entry, so only the part of the rows that is
            // still to be eliminated moves (columns >= k, LINPACK's dgefa layout: the multipliers of the
            // earlier columns stay where they were computed and lu_solve interleaves the interchanges of
            // the right-hand side with the column sweeps, dgesl).  Same multipliers applied to the same
            // values in the same order as with LAPACK's full-row interchanges: bit-identical solutions,
            // 22 % fewer selects per factorisation at n = 6.
#pragma unroll U
            for (int i = k + 1; i < N; ++i) {
                const bool sw = (p == i);
#pragma unroll U
                for (int j = LinpackSwaps<N>::value ? k : 0; j < N; ++j) {
                    const double x = a[k * N + j], z = a[i * N + j];
                    a[k * N + j] = sw ? z : x;
                    a[i * N + j] = sw ? x : z;
                }
            }
        } else if (p != k) {
            for (int j = 0; j < N; ++j) {
                const double x = a[k * N + j];
                a[k * N + j] = a[p * N + j];
                a[p * N + j] = x;
            }
        }
        const double pv = a[k * N + k];
        const double r = fast_rcp_scale(pv);
        a[k * N + k] = r;                              // reciprocal pivot
        const double rs = (pv != 0.0) ? r : 1.0;       // LAPACK skips the scaling for a zero pivot
#pragma unroll U
        for (int i = k + 1; i < N; ++i) {
            const double l = a[i * N + k] * rs;
            a[i * N + k] = l;
#pragma unroll U
            for (int j = k + 1; j < N; ++j) a[i * N + j] = fma(-l, a[k * N + j], a[i * N + j]);
        }
    }
}

template <int N>
__device__ __forceinline__ void lu_solve(const double* __restrict__ lu, const int* __restrict__ piv,
                                         double* __restrict__ b) {
    constexpr bool REG = Regime<N>::kReg;
    if constexpr (REG) {
        // (`#pragma unroll` without a count: with the count form ptxas once kept the inner loop of the
        // complex twin rolled, which put its right-hand side into local memory)
        if constexpr (LinpackSwaps<N>::value) {
#pragma unroll
        for (int k = 0; k < N; ++k) {                  // dgesl: interchange k, then the column sweep of step k
            const int p = piv[k];
#pragma unroll
            for (int i = k + 1; i < N; ++i) {
                const bool sw = (p == i);
                const double x = b[k], z = b[i];
                b[k] = sw ? z : x;
                b[i] = sw ? x : z;
            }
#pragma unroll
            for (int i = k + 1; i < N; ++i) b[i] = fma(-lu[i * N + k], b[k], b[i]);
        }
        } else {
#pragma unroll
        for (int k = 0; k < N; ++k) {
            const int p = piv[k];
#pragma unroll
            for (int i = k + 1; i < N; ++i) {
                const bool sw = (p == i);
                const double x = b[k], z = b[i];
                b[k] = sw ? z : x;
                b[i] = sw ? x : z;
            }
        }
#pragma unroll
        for (int i = 1; i < N; ++i) {
            double acc = b[i];
#pragma unroll
            for (int j = 0; j < i; ++j) acc = fma(-lu[i * N + j], b[j], acc);
            b[i] = acc;
        }
        }
#pragma unroll
        for (int i = N - 1; i >= 0; --i) {
            double acc = b[i];
#pragma unroll
            for (int j = i + 1; j < N; ++j) acc = fma(-lu[i * N + j], b[j], acc);
            b[i] = acc * lu[i * N + i];
        }
    } else {
        for (int k = 0; k < N; ++k) {
            const int p = piv[k];
            if (p != k) { const double x = b[k]; b[k] = b[p]; b[p] = x; }
        }
        for (int i = 1; i < N; ++i) {
            double acc = b[i];
            for (int j = 0; j < i; ++j) acc = fma(-lu[i * N + j], b[j], acc);
            b[i] = acc;
        }
        for (int i = N - 1; i >= 0; --i) {
            double acc = b[i];
            for (int j = i + 1; j < N; ++j) acc = fma(-lu[i * N + j], b[j], acc);
            b[i] = acc * lu[i * N + i];
        }
    }
}

// Pivoted assembly.  The first pivot of getrf depends only on column 0 of the matrix, which is known
// before the matrix is assembled: p0 = first maximal |a[i][0]|.  LAPACK then interchanges rows 0 and
// p0, i.e. the rows stand in the order  r[0] = p0,  r[p0] = 0,  r[i] = i otherwise.  Jy and Jyp live
// in shared memory, where a run-time row index costs an address, not a select: the matrix is assembled
// with its rows already in that order and the N^2 (real) / 2 N^2 (complex) selects of the first row
// interchange disappear (they were 10 % of the select instructions of the Robertson kernel).
template <int N>
__device__ __forceinline__ void first_pivot_order(const double* __restrict__ mag, int* __restrict__ r, int& p0) {
    int p = 0;
    double best = mag[0];
#pragma unroll
    for (int i = 1; i < N; ++i)
        if (mag[i] > best) { best = mag[i]; p = i; }
    p0 = p;
    r[0] = p;
#pragma unroll
    for (int i = 1; i < N; ++i) r[i] = (i == p) ? 0 : i;
}

// ---------------------------------------------------------------------------------------------
// complex, split re/im storage
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void crecip(double pr, double pi, double& rr, double& ri) {
    // 1 / (pr + i pi) = conj / |.|^2 with one reciprocal.  When |.|^2 would leave [1e-280, 1e280]
    // (pivots beyond 1e+-140: rare branch) the operands are first rescaled by a power of two,
    // which is what Smith's formula guards against in LAPACK's zladiv.
    double sc = 1.0;
    const double m = amax_sel(pr, pi);
    if (!(m >= 1e-140 && m <= 1e140)) sc = (m > 1.0) ? 0x1p-600 : ((m > 0.0) ? 0x1p+600 : 1.0);
    const double a = pr * sc, b = pi * sc;
    const double r = fast_rcp(fma(a, a, b * b)) * sc;
    rr = a * r; ri = -b * r;
}

template <int N, class MR, class MI, bool kFirstDone = false>
__device__ __forceinline__ void zlu_factor(MR ar, MI ai, int* __restrict__ piv) {
    constexpr bool REG = Regime<N>::kReg;
    constexpr int U = Regime<N>::kUnroll;
#pragma unroll U
    for (int k = 0; k < N; ++k) {
        int p = k;
        if (!(kFirstDone && k == 0)) {
            double best = fabs(ar[k * N + k]) + fabs(ai[k * N + k]);
#pragma unroll U
            for (int i = k + 1; i < N; ++i) {
                const double v = fabs(ar[i * N + k]) + fabs(ai[i * N + k]);
                if (v > best) { best = v; p = i; }
            }
            piv[k] = p;
        }
        if (kFirstDone && k == 0) {
            // nothing to move
        } else if (REG) {
#pragma unroll U
            for (int i = k + 1; i < N; ++i) {
                const bool sw = (p == i);
#pragma unroll U
                for (int j = LinpackSwaps<N>::value ? k : 0; j < N; ++j) {      // see lu_factor
                    double x = ar[k * N + j], z = ar[i * N + j];
                    ar[k * N + j] = sw ? z : x;
                    ar[i * N + j] = sw ? x : z;
                    x = ai[k * N + j]; z = ai[i * N + j];
                    ai[k * N + j] = sw ? z : x;
                    ai[i * N + j] = sw ? x : z;
                }
            }
        } else if (p != k) {
            for (int j = 0; j < N; ++j) {
                double x = ar[k * N + j]; ar[k * N + j] = ar[p * N + j]; ar[p * N + j] = x;
                x = ai[k * N + j]; ai[k * N + j] = ai[p * N + j]; ai[p * N + j] = x;
            }
        }
        const double pr = ar[k * N + k], pi = ai[k * N + k];
        double rr, ri;
        crecip(pr, pi, rr, ri);
        ar[k * N + k] = rr;                            // reciprocal pivot
        ai[k * N + k] = ri;
        const bool nz = (pr != 0.0) || (pi != 0.0);
        const double sr = nz ? rr : 1.0, si = nz ? ri : 0.0;
#pragma unroll U
        for (int i = k + 1; i < N; ++i) {
            const double xr = ar[i * N + k], xi = ai[i * N + k];
            const double lr = fma(xr, sr, -xi * si), li = fma(xr, si, xi * sr);
            ar[i * N + k] = lr;
            ai[i * N + k] = li;
#pragma unroll U
            for (int j = k + 1; j < N; ++j) {
                const double ur = ar[k * N + j], ui = ai[k * N + j];
                ar[i * N + j] = fma(li, ui, fma(-lr, ur, ar[i * N + j]));
                ai[i * N + j] = fma(-li, ur, fma(-lr, ui, ai[i * N + j]));
            }
        }
    }
}

template <int N, class MR, class MI>
__device__ __forceinline__ void zlu_solve(MR ar, MI ai, const int* __restrict__ piv,
                                          double* __restrict__ br, double* __restrict__ bi) {
    constexpr bool REG = Regime<N>::kReg;
    if constexpr (REG) {
        if constexpr (LinpackSwaps<N>::value) {
#pragma unroll
        for (int k = 0; k < N; ++k) {
            const int p = piv[k];
#pragma unroll
            for (int i = k + 1; i < N; ++i) {
                const bool sw = (p == i);
                double x = br[k], z = br[i];
                br[k] = sw ? z : x; br[i] = sw ? x : z;
                x = bi[k]; z = bi[i];
                bi[k] = sw ? z : x; bi[i] = sw ? x : z;
            }
#pragma unroll
            for (int i = k + 1; i < N; ++i) {
                const double lr = ar[i * N + k], li = ai[i * N + k];
                const double sr = br[i], si = bi[i];
                br[i] = fma(li, bi[k], fma(-lr, br[k], sr));
                bi[i] = fma(-li, br[k], fma(-lr, bi[k], si));
            }
        }
        } else {
#pragma unroll
        for (int k = 0; k < N; ++k) {
            const int p = piv[k];
#pragma unroll
            for (int i = k + 1; i < N; ++i) {
                const bool sw = (p == i);
                double x = br[k], z = br[i];
                br[k] = sw ? z : x; br[i] = sw ? x : z;
                x = bi[k]; z = bi[i];
                bi[k] = sw ? z : x; bi[i] = sw ? x : z;
            }
        }
#pragma unroll
        for (int i = 1; i < N; ++i) {
            double sr = br[i], si = bi[i];
#pragma unroll
            for (int j = 0; j < i; ++j) {
                const double lr = ar[i * N + j], li = ai[i * N + j];
                sr = fma(li, bi[j], fma(-lr, br[j], sr));
                si = fma(-li, br[j], fma(-lr, bi[j], si));
            }
            br[i] = sr; bi[i] = si;
        }
        }
#pragma unroll
        for (int i = N - 1; i >= 0; --i) {
            double sr = br[i], si = bi[i];
#pragma unroll
            for (int j = i + 1; j < N; ++j) {
                const double ur = ar[i * N + j], ui = ai[i * N + j];
                sr = fma(ui, bi[j], fma(-ur, br[j], sr));
                si = fma(-ui, br[j], fma(-ur, bi[j], si));
            }
            const double dr = ar[i * N + i], di = ai[i * N + i];   // reciprocal pivot
            br[i] = fma(sr, dr, -si * di);
            bi[i] = fma(sr, di, si * dr);
        }
    } else {
        for (int k = 0; k < N; ++k) {
            const int p = piv[k];
            if (p != k) {
                double x = br[k]; br[k] = br[p]; br[p] = x;
                x = bi[k]; bi[k] = bi[p]; bi[p] = x;
            }
        }
        for (int i = 1; i < N; ++i) {
            double sr = br[i], si = bi[i];
            for (int j = 0; j < i; ++j) {
                const double lr = ar[i * N + j], li = ai[i * N + j];
                sr = fma(li, bi[j], fma(-lr, br[j], sr));
                si = fma(-li, br[j], fma(-lr, bi[j], si));
            }
            br[i] = sr; bi[i] = si;
        }
        for (int i = N - 1; i >= 0; --i) {
            double sr = br[i], si = bi[i];
            for (int j = i + 1; j < N; ++j) {
                const double ur = ar[i * N + j], ui = ai[i * N + j];
                sr = fma(ui, bi[j], fma(-ur, br[j], sr));
                si = fma(-ui, br[j], fma(-ur, bi[j], si));
            }
            const double dr = ar[i * N + i], di = ai[i * N + i];
            br[i] = fma(sr, dr, -si * di);
            bi[i] = fma(sr, di, si * dr);
        }
    }
}

}  // namespace b200dae
