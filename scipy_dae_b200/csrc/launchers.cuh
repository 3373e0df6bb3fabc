// launchers.cuh — launch shapes and the per-problem dispatch entry (problem, stages) -> kernel.
// Each problem is instantiated in its own translation unit (problem_*.cu) so the library builds
// in parallel; b200dae.cu only holds the C-ABI and the registry.
#pragma once
#include <cuda_runtime.h>
#include <stdlib.h>
#include <string.h>

#include <atomic>

#include "../../include/b200dae.h"
#include "launch_shapes.cuh"

#ifndef B200DAE_BIG_CARVEOUT_DEFAULT
#define B200DAE_BIG_CARVEOUT_DEFAULT 40     // percent of the SM's 228 KB; 0: leave the carve-out to the driver
#endif

namespace b200dae {

extern std::atomic<int64_t> g_launches;

typedef int (*radau_launcher)(const RadauLaunch&, cudaStream_t);
typedef int (*bdf_launcher)(const BdfLaunch&, cudaStream_t);
typedef int (*cic_launcher)(const CicLaunch&, cudaStream_t);
typedef int (*residual_launcher)(int64_t, int64_t, double, const double*, const double*,
                                 const double*, double*, cudaStream_t);

struct ProblemEntry {
    const char* name;
    int n, n_params, n_events;                      // n_events: event functions of the functor (EventOf)
    radau_launcher radau[B200DAE_MAX_STAGES + 1];   // indexed by the number of stages
    bdf_launcher bdf;
    residual_launcher residual;
    cic_launcher cic;
};

ProblemEntry entry_robertson();
ProblemEntry entry_akzo_nobel();
ProblemEntry entry_transamp();
ProblemEntry entry_andrews();
ProblemEntry entry_rational();
ProblemEntry entry_linear();
ProblemEntry entry_cosine();
ProblemEntry entry_weissinger();
ProblemEntry entry_brenan();
ProblemEntry entry_transamp1();
ProblemEntry entry_blowup();

// B200DAE_REGIME=thread forces the one-thread-per-system kernels for n > 4 (A/B measurements).
static bool force_thread_regime() {
    static const bool v = getenv("B200DAE_REGIME") && strcmp(getenv("B200DAE_REGIME"), "thread") == 0;
    return v;
}

constexpr int kErrNoJac = -77;    // analytic Jacobian requested but not compiled for (problem, stages)

template <class P, int S, bool J, bool LEAN, bool GJ>
static int launch_radau_group_shape(const RadauLaunch& L, cudaStream_t st) {
    using Sh = RadauGroupShape<P, S, GJ>;
    constexpr int G = Sh::G, T = Sh::T, MB = Sh::MB;
    constexpr size_t smem = Sh::smem;
    auto kern = radau_group_kernel<P, S, G, T, MB, J, LEAN, GJ>;
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t rc = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (rc != cudaSuccess) return (int)rc;
        attr_set = true;
    }
    const int64_t blocks = (L.N * G + T - 1) / T;
    kern<<<(unsigned)blocks, T, smem, st>>>(L);
    g_launches.fetch_add(1);
    return (int)cudaGetLastError();
}

// B200DAE_GLOBAL_J=0 keeps the Jacobians of the one-warp-per-system kernels in shared memory (A/B)
static bool global_j_enabled() {
    static const bool v = !(getenv("B200DAE_GLOBAL_J") && strcmp(getenv("B200DAE_GLOBAL_J"), "0") == 0);
    return v;
}

template <class P, int S, bool J = false, bool LEAN = false>
static int launch_radau_group(const RadauLaunch& L, cudaStream_t st) {
    if constexpr (LEAN && !J && RadauGroupShape<P, S>::kCanGJ) {
        // Jacobians in a stream-ordered global scratch block (2 n (n|1) doubles per instance, L2-resident
        // for the resident systems); beyond 8 GB, or if the pool cannot serve it, the shared-memory variant runs
        const size_t bytes = (size_t)L.N * RadauGroupShape<P, S, true>::scratch;
        void* scr = nullptr;
        if (global_j_enabled() && bytes <= ((size_t)8 << 30) && cudaMallocAsync(&scr, bytes, st) == cudaSuccess) {
            RadauLaunch Lg = L;
            Lg.jscratch = (double*)scr;
            const int rc = launch_radau_group_shape<P, S, J, LEAN, true>(Lg, st);
            cudaFreeAsync(scr, st);
            return rc;
        }
        (void)cudaGetLastError();
    }
    return launch_radau_group_shape<P, S, J, LEAN, false>(L, st);
}

template <class P, int S>
static int launch_radau(const RadauLaunch& L, cudaStream_t st) {
    // analytic-Jacobian kernels are compiled for s = 3 of the problems that provide jac()
    constexpr bool kJacOk = P::kHasJac && S == 3;
    if (L.analytic_jac && !kJacOk) return kErrNoJac;
    if constexpr (!Regime<P::N>::kReg) {
        if (!force_thread_regime()) {
            const bool glean = L.newton_iter_embedded == 1 && L.cew == 0.0 && L.out.t_eval == nullptr &&
                               L.out.t_steps == nullptr && L.ev.spec.n_events == 0;
            if constexpr (kJacOk) {
                if (L.analytic_jac) return launch_radau_group<P, S, true, false>(L, st);
            }
            return glean ? launch_radau_group<P, S, false, true>(L, st) : launch_radau_group<P, S, false, false>(L, st);
        }
        if (L.analytic_jac) return kErrNoJac;
    }
    using Tn = RadauTune<P, S>;
    constexpr int T = Tn::kThreads;
    const int64_t blocks = (L.N + T - 1) / T;
    // lean: default options, no step trace, no events; with or without t_eval output (FD Jacobian,
    // register regime, s = 3: the headline kernel and its dense-output twin)
    const bool lean_opts = L.newton_iter_embedded == 1 && L.cew == 0.0 && L.out.t_steps == nullptr &&
                           L.ev.spec.n_events == 0;
    const bool lean = lean_opts && L.out.t_eval == nullptr;
    bool done = false;
    if (Tn::kDynSmem > 0) {
        // cold state of the unrolled thread regime beyond n = 4: dynamic shared memory above 48 KB
        static bool attr_set = false;
        if (!attr_set) {
            cudaFuncSetAttribute(radau_kernel<P, S, T, Tn::kMinBlocks, Tn::kSharedCold, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Tn::kDynSmem);
            cudaFuncSetAttribute(radau_kernel<P, S, T, Tn::kMinBlocks, Tn::kSharedCold, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Tn::kDynSmem);
            if constexpr (kJacOk && Regime<P::N>::kReg) {
                cudaFuncSetAttribute(radau_kernel<P, S, T, Tn::kMinBlocks, Tn::kSharedCold, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Tn::kDynSmem);
                cudaFuncSetAttribute(radau_kernel<P, S, T, Tn::kMinBlocks, Tn::kSharedCold, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Tn::kDynSmem);
            }
            if constexpr (Regime<P::N>::kReg && S == 3) {
                cudaFuncSetAttribute(radau_kernel<P, S, T, Tn::kMinBlocks, Tn::kSharedCold, true, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Tn::kDynSmem);
                cudaFuncSetAttribute(radau_kernel<P, S, T, Tn::kMinBlocks, Tn::kSharedCold, true, false, false, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Tn::kDynSmem);
            }
            attr_set = true;
        }
    }
    if constexpr (kJacOk && Regime<P::N>::kReg) {
        if (L.analytic_jac) {
            if (lean) radau_kernel<P, S, T, Tn::kMinBlocks, Tn::kSharedCold, true, true><<<(unsigned)blocks, T, Tn::kDynSmem, st>>>(L);
            else radau_kernel<P, S, T, Tn::kMinBlocks, Tn::kSharedCold, false, true><<<(unsigned)blocks, T, Tn::kDynSmem, st>>>(L);
            done = true;
        }
    }
    if constexpr (Regime<P::N>::kReg && S == 3) {
        if (!done && lean_opts && !lean) {
            radau_kernel<P, S, T, Tn::kMinBlocks, Tn::kSharedCold, true, false, true><<<(unsigned)blocks, T, Tn::kDynSmem, st>>>(L);
            done = true;
        }
    }
    if constexpr (Regime<P::N>::kReg && S == 3 && Tn::kSharedCold) {
        // instance-major I/O (io_layout = 1, zero-copy launches on pinned host arrays): the lean kernel with
        // its inputs and results staged through shared memory, so that PCIe sees contiguous runs
        if (!done && lean && L.io_aos && L.newton_max_iter <= 8) {
            auto kern = radau_kernel<P, S, T, Tn::kMinBlocks, Tn::kSharedCold, true, false, false, 3, true>;
            if (Tn::kDynSmem > 0) {
                static bool attr_io = false;
                if (!attr_io) { cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Tn::kDynSmem); attr_io = true; }
            }
            kern<<<(unsigned)blocks, T, Tn::kDynSmem, st>>>(L);
            done = true;
        }
    }
    if constexpr (Regime<P::N>::kReg && S == 3 && Tn::kBig) {
        // unrolled thread regime beyond n = 4: the kernel lives on its spills (1.5 KB per thread), which hit L1
        // only if the shared-memory carve-out leaves room for them; fewer resident blocks with a larger L1 win
        // (B200DAE_BIG_CARVEOUT = percent of the 228 KB given to shared memory; measured default below)
        static bool carve_set = false;
        if (!carve_set) {
            const char* env = getenv("B200DAE_BIG_CARVEOUT");
            const int pct = env ? atoi(env) : B200DAE_BIG_CARVEOUT_DEFAULT;
            if (pct > 0)
                cudaFuncSetAttribute(radau_kernel<P, S, T, Tn::kMinBlocks, Tn::kSharedCold, true, false, false, 3>,
                                     cudaFuncAttributePreferredSharedMemoryCarveout, pct);
            carve_set = true;
        }
    }
    if constexpr (Regime<P::N>::kReg && S == 3) {
        // the headline kernel with a 3-bit exponent in the Newton rate test (newton_max_iter <= 8: the default 7)
        if (!done && lean && L.newton_max_iter <= 8) {
            radau_kernel<P, S, T, Tn::kMinBlocks, Tn::kSharedCold, true, false, false, 3><<<(unsigned)blocks, T, Tn::kDynSmem, st>>>(L);
            done = true;
        }
    }
    if (!done) {
        if (lean) radau_kernel<P, S, T, Tn::kMinBlocks, Tn::kSharedCold, true><<<(unsigned)blocks, T, Tn::kDynSmem, st>>>(L);
        else radau_kernel<P, S, T, Tn::kMinBlocks, Tn::kSharedCold, false><<<(unsigned)blocks, T, Tn::kDynSmem, st>>>(L);
    }
    g_launches.fetch_add(1);
    return (int)cudaGetLastError();
}

template <class P, bool J = false, bool E = false, bool O = true>
static int launch_bdf_group(const BdfLaunch& L, cudaStream_t st) {
    using Sh = BdfGroupShape<P>;
    constexpr int G = Sh::G, T = Sh::T, MB = Sh::MB;
    constexpr size_t smem = Sh::smem;
    auto kern = bdf_group_kernel<P, G, T, MB, J, E, O>;
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t rc = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (rc != cudaSuccess) return (int)rc;
        attr_set = true;
    }
    const int64_t blocks = (L.N * G + T - 1) / T;
    kern<<<(unsigned)blocks, T, smem, st>>>(L);
    g_launches.fetch_add(1);
    return (int)cudaGetLastError();
}

template <class P>
static int launch_bdf(const BdfLaunch& L, cudaStream_t st) {
    if (L.analytic_jac && !P::kHasJac) return kErrNoJac;
    if constexpr (!Regime<P::N>::kReg) {
        if (!force_thread_regime()) {
            const bool ev = L.ev.spec.n_events > 0;      // the event code is compiled out otherwise
            const bool plain = !ev && L.out.t_eval == nullptr && L.out.t_steps == nullptr;   // ... and t_eval / trace
            if constexpr (P::kHasJac) {
                if (L.analytic_jac) return ev ? launch_bdf_group<P, true, true>(L, st) : launch_bdf_group<P, true, false>(L, st);
            }
            if (plain) return launch_bdf_group<P, false, false, false>(L, st);
            return ev ? launch_bdf_group<P, false, true>(L, st) : launch_bdf_group<P, false, false>(L, st);
        }
        if (L.analytic_jac) return kErrNoJac;
    }
    using Tn = BdfTune<P>;
    constexpr int T = Tn::kThreads;
    constexpr bool SC = Tn::kSharedCold;
    const int64_t blocks = (L.N + T - 1) / T;
    const bool lean = L.out.t_eval == nullptr && L.out.t_steps == nullptr && L.ev.spec.n_events == 0;
    if (Tn::kDynSmem > 0) {
        static bool attr_set = false;
        if (!attr_set) {
            cudaFuncSetAttribute(bdf_kernel<P, T, false, false, SC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Tn::kDynSmem);
            if constexpr (Regime<P::N>::kReg) {
                cudaFuncSetAttribute(bdf_kernel<P, T, false, true, SC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Tn::kDynSmem);
                if constexpr (P::kHasJac)
                    cudaFuncSetAttribute(bdf_kernel<P, T, true, false, SC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Tn::kDynSmem);
            }
            attr_set = true;
        }
    }
    bool done = false;
    if constexpr (P::kHasJac && Regime<P::N>::kReg) {
        if (L.analytic_jac) { bdf_kernel<P, T, true, false, SC><<<(unsigned)blocks, T, Tn::kDynSmem, st>>>(L); done = true; }
    }
    if constexpr (Regime<P::N>::kReg) {
        if (!done && lean) {
            // B200DAE_BDF_PAD_SMEM (experiments): unused dynamic shared memory per block, which caps the resident
            // blocks per SM so that the spilled difference arrays of the unrolled n > 4 kernel fit L1
            static const size_t pad = getenv("B200DAE_BDF_PAD_SMEM") ? (size_t)atol(getenv("B200DAE_BDF_PAD_SMEM")) : 0;
            size_t dyn = Tn::kDynSmem;
            if (P::N > 4 && getenv("B200DAE_BDF_CARVEOUT")) {
                static bool carve_set = false;
                if (!carve_set) {
                    cudaFuncSetAttribute(bdf_kernel<P, T, false, true, SC>, cudaFuncAttributePreferredSharedMemoryCarveout, atoi(getenv("B200DAE_BDF_CARVEOUT")));
                    carve_set = true;
                }
            }
            if (pad > dyn && P::N > 4) {
                static bool pad_set = false;
                if (!pad_set) {
                    cudaFuncSetAttribute(bdf_kernel<P, T, false, true, SC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pad);
                    if (getenv("B200DAE_BDF_CARVEOUT"))
                        cudaFuncSetAttribute(bdf_kernel<P, T, false, true, SC>, cudaFuncAttributePreferredSharedMemoryCarveout, atoi(getenv("B200DAE_BDF_CARVEOUT")));
                    pad_set = true;
                }
                dyn = pad;
            }
            bdf_kernel<P, T, false, true, SC><<<(unsigned)blocks, T, dyn, st>>>(L);
            done = true;
        }
    }
    if (!done) bdf_kernel<P, T, false, false, SC><<<(unsigned)blocks, T, Tn::kDynSmem, st>>>(L);
    g_launches.fetch_add(1);
    return (int)cudaGetLastError();
}

template <class P>
static int launch_residual(int64_t N, int64_t ld, double t, const double* y, const double* yp,
                           const double* params, double* out, cudaStream_t st) {
    residual_kernel<P><<<(unsigned)((N + 127) / 128), 128, 0, st>>>(N, ld, t, y, yp, params, out);
    g_launches.fetch_add(1);
    return (int)cudaGetLastError();
}

template <class P>
static int launch_cic(const CicLaunch& L, cudaStream_t st) {
    if (L.analytic_jac && !P::kHasJac) return kErrNoJac;
    const unsigned blocks = (unsigned)((L.N + 31) / 32);
    bool done = false;
    if constexpr (P::kHasJac) {
        if (L.analytic_jac) { cic_kernel<P, true><<<blocks, 32, 0, st>>>(L); done = true; }
    }
    if (!done) cic_kernel<P, false><<<blocks, 32, 0, st>>>(L);
    g_launches.fetch_add(1);
    return (int)cudaGetLastError();
}

template <class P, bool S1, bool S5, bool S7>
static ProblemEntry make_entry() {
    ProblemEntry e;
    memset(&e, 0, sizeof(e));
    e.name = P::name();
    e.n = P::N;
    e.n_params = P::NP;
    e.n_events = EventOf<P>::kNum;
    e.radau[3] = launch_radau<P, 3>;
    if constexpr (S1) e.radau[1] = launch_radau<P, 1>;
    if constexpr (S5) e.radau[5] = launch_radau<P, 5>;
    if constexpr (S7) e.radau[7] = launch_radau<P, 7>;
    e.bdf = launch_bdf<P>;
    e.residual = launch_residual<P>;
    e.cic = launch_cic<P>;
    return e;
}

}  // namespace b200dae
