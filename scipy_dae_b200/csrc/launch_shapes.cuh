// launch_shapes.cuh — compile-time launch shapes of every kernel family, per (problem, stages):
// block size, blocks per SM, shared-memory bytes.  Included by the nvcc launchers (launchers.cuh)
// and by the run-time compiled modules of registered problems (rt_kernels.cuh, NVRTC), so that both
// instantiate exactly the same kernels.
#pragma once
#include "bdf.cuh"
#include "bdf_group.cuh"
#include "cic.cuh"
#include "problems.cuh"
#include "radau.cuh"
#include "radau_group.cuh"

#ifndef B200DAE_REG_MINBLOCKS
#define B200DAE_REG_MINBLOCKS 6
#endif
#ifndef B200DAE_REG_THREADS
#define B200DAE_REG_THREADS 64
#endif
#ifndef B200DAE_BIG_THREADS
#define B200DAE_BIG_THREADS 64
#endif
#ifndef B200DAE_GROUP_MINBLOCKS
#define B200DAE_GROUP_MINBLOCKS(G) ((G) == 32 ? 1 : 6)
#endif

namespace b200dae {

// threads per block: register-regime kernels keep ~255 registers per thread live, so 128
// threads x 2 blocks fill an SM's register file; the local-memory regime prefers fewer
// threads per SM so that its per-thread arrays stay L1-resident.
#ifndef B200DAE_BDF_BIG_THREADS
#define B200DAE_BDF_BIG_THREADS 128
#endif
template <class P> struct BlockSize { static constexpr int value = Regime<P::N>::kReg ? (P::N > 4 ? B200DAE_BDF_BIG_THREADS : 128) : 64; };

// Launch shape of the Radau kernel per (problem size, stages).
//   register regime (n <= 4): 64-thread blocks, the cold per-instance state in shared memory
//   (when it fits the 48 KB static limit); for s = 3 the register cap is 168 so that 6 blocks =
//   3 warps per scheduler are resident (measured on B200, 2^20 Robertson instances:
//   255 regs/2 warps 18.9 ms, 168 regs/3 warps 15.6 ms, 128 regs/4 warps (spills) 17.7 ms).
//   local-memory regime (n > 4): 64-thread blocks so the per-thread arrays stay L1-resident.
template <class P, int S>
struct RadauTune {
    static constexpr int N = P::N, NCS = S / 2;
    static constexpr bool kReg = Regime<N>::kReg;
    static constexpr bool kBig = kReg && N > 4;         // unrolled thread regime beyond n = 4 (B200DAE_REG_MAXN)
    static constexpr int kThreads = kReg ? (kBig ? B200DAE_BIG_THREADS : B200DAE_REG_THREADS) : 64;
    static constexpr int kCold = 2 * N * N + 2 * N + N * S + 2 * NCS * N * N;
    static constexpr bool kSharedCold = kReg && (kBig || kCold * kThreads * 8 <= 48 * 1024);
    static constexpr size_t kDynSmem = (kSharedCold && (size_t)kCold * kThreads * 8 > 48 * 1024) ? (size_t)kCold * kThreads * 8 : 0;
    static constexpr int kMinBlocks = kBig ? 1 : ((kSharedCold && S <= 3) ? B200DAE_REG_MINBLOCKS : 1);
};

// Launch shape of the thread-per-system BDF kernel.  kSharedCold (Jacobians and FD step factors in
// shared memory [element][thread], as in the Radau kernel) is compiled but OFF: measured on B200 it
// costs the L1 capacity the spilled difference array D lives in (Akzo Nobel: 86 KB per 128-thread
// block, 5.95 M -> 5.01 M solves/s; Robertson: 16.0 M -> 15.8 M).  -DB200DAE_BDF_SHARED_COLD=1 enables it.
#ifndef B200DAE_BDF_SHARED_COLD
#define B200DAE_BDF_SHARED_COLD 0
#endif
template <class P>
struct BdfTune {
    static constexpr int N = P::N;
    static constexpr int kThreads = BlockSize<P>::value;
    static constexpr bool kSharedCold = Regime<N>::kReg && (B200DAE_BDF_SHARED_COLD != 0);
    static constexpr size_t kColdBytes = (size_t)(2 * N * N + 2 * N) * kThreads * sizeof(double);
    static constexpr size_t kDynSmem = (kSharedCold && kColdBytes > 48 * 1024) ? kColdBytes : 0;
};

// Group size of the cooperative kernels (group.cuh): the smallest power of two >= n (>= 8).
template <int N> struct GroupSize { static constexpr int value = N <= 8 ? 8 : (N <= 16 ? 16 : 32); };

// Shape of the cooperative Radau kernel.  Sub-warp groups: blocks per SM = what shared memory admits,
// capped at 8 (128 registers, 16 warps); measured on B200 for Akzo Nobel Radau s = 3: 4 blocks 3.64 M,
// 6 blocks 4.8 M, 8 blocks 5.26 M solves/s.  The amplifier's shared memory admits 6, so it stays at
// 168 registers.
template <class P, int S, bool GJ = false>
struct RadauGroupShape {
    static constexpr int G = GroupSize<P::N>::value;
    static constexpr int T = (G == 32) ? 32 : 64;
    // one warp per system (G = 32): the variant with the Jacobians in global memory (radau_group.cuh)
    // is worth its scratch block; the plain solve (lean kernel, FD Jacobian) uses it
    static constexpr bool kCanGJ = (G == 32);
    using Lay = RadauGroupLayout<P::N, S, G, GJ>;
    static constexpr size_t smem = (size_t)(T / G) * Lay::doubles(P::NP) * sizeof(double);
    static constexpr size_t scratch = (size_t)Lay::kScratch * sizeof(double);    // global bytes per instance
    static_assert(smem <= 227 * 1024, "per-block shared memory");
    static constexpr int kFit = (int)((228 * 1024) / (smem + 1024));            // 1 KB per block is reserved
    static constexpr int MB = (G == 32) ? (GJ ? (kFit > 8 ? 8 : kFit) : 1)
                                        : (kFit >= 8 ? 8 : (kFit >= B200DAE_GROUP_MINBLOCKS(G) ? B200DAE_GROUP_MINBLOCKS(G) : (kFit > 0 ? kFit : 1)));
};

template <class P>
struct BdfGroupShape {
    static constexpr int G = GroupSize<P::N>::value;
    static constexpr int T = (G == 32) ? 32 : 64;
    using Lay = BdfGroupLayout<P::N, G>;
    static constexpr size_t smem = (size_t)(T / G) * Lay::doubles(P::NP) * sizeof(double);
    static_assert(smem <= 227 * 1024, "per-block shared memory");
    static constexpr int MB = B200DAE_GROUP_MINBLOCKS(G);
};

// residual check kernel: out = F(t, y, yp; p) for N instances (SoA)
template <class P>
__global__ void residual_kernel(int64_t N, int64_t ld, double t, const double* __restrict__ y,
                                const double* __restrict__ yp, const double* __restrict__ params,
                                double* __restrict__ out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    double yy[P::N], yyp[P::N], pp[P::NP > 0 ? P::NP : 1], f[P::N];
    for (int j = 0; j < P::N; ++j) { yy[j] = y[j * ld + i]; yyp[j] = yp[j * ld + i]; }
    for (int j = 0; j < P::NP; ++j) pp[j] = params[j * ld + i];
    P::eval(t, yy, yyp, pp, f);
    for (int j = 0; j < P::N; ++j) out[j * ld + i] = f[j];
}

}  // namespace b200dae
