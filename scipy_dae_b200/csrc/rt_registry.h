// rt_registry.h — run-time registered problems (rt_registry.cu), as seen by the C-ABI unit.
#pragma once
#include <cuda_runtime.h>

#include <string>

#include "launch_shapes.cuh"     // RadauLaunch, BdfLaunch, CicLaunch

namespace b200dae {

constexpr int kRtErrNoJac = -77;  // same meaning as launchers.cuh: kErrNoJac

// Registers (or finds, when name and source are identical) a problem; returns its run-time index
// (>= 0) or a negative B200DAE_E* code with `err` set.  Compiles the residual-check kernel: a
// syntax and interface check of the source that needs NVRTC but no device.
int rt_register(const char* name, const char* src, int n, int n_params, std::string& err);
int rt_count();
bool rt_info(int idx, const char** name, int* n, int* n_params, int* n_events, int* has_jac);

// launch = false: only select and compile the variant (no device needed)
int rt_launch_radau(int idx, const RadauLaunch& L, cudaStream_t st, bool launch, std::string& err);
int rt_launch_bdf(int idx, const BdfLaunch& L, cudaStream_t st, bool launch, std::string& err);
int rt_launch_cic(int idx, const CicLaunch& L, cudaStream_t st, bool launch, std::string& err);
int rt_launch_residual(int idx, int64_t N, int64_t ld, double t, const double* y, const double* yp,
                       const double* params, double* out, cudaStream_t st, std::string& err);

}  // namespace b200dae
