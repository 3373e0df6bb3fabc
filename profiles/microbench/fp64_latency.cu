// fp64_latency.cu — dependent-issue latency and per-warp throughput of the FP64 pipe on B200,
// measured with clock64() around serial chains in ONE warp per SM sub-partition.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_latency fp64_latency.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int CHAINS, int OP>
__global__ void chain_kernel(int iters, double seed, double* out, long long* cycles) {
    double a[CHAINS];
#pragma unroll
    for (int c = 0; c < CHAINS; ++c) a[c] = seed + 0.001 * c + 1e-9 * threadIdx.x;
    const double m = 0.999999, k = 1e-6;
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int c = 0; c < CHAINS; ++c) {
            if (OP == 0) a[c] = fma(a[c], m, k);
            else if (OP == 1) a[c] = a[c] + k;
            else if (OP == 2) a[c] = a[c] * m;
            else if (OP == 3) a[c] = 1.0 / a[c] + 0.5;       // IEEE division
            else if (OP == 4) a[c] = sqrt(a[c]) + 0.5;
            else if (OP == 5) a[c] = a[c] > 1.0 ? a[c] * m : a[c] + k;   // DSETP + select
            else if (OP == 6) a[c] = rsqrt(a[c]) + 0.5;
            else if (OP == 7) a[c] = __drcp_rn(a[c]) + 0.5;
        }
    }
    long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int c = 0; c < CHAINS; ++c) s += a[c];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cycles = t1 - t0;
}

template <int CHAINS, int OP>
void run(const char* name, int threads) {
    double* out; long long* cyc;
    cudaMalloc(&out, sizeof(double) * 1024); cudaMalloc(&cyc, sizeof(long long));
    const int iters = 4096;
    chain_kernel<CHAINS, OP><<<1, threads>>>(iters, 1.5, out, cyc);
    chain_kernel<CHAINS, OP><<<1, threads>>>(iters, 1.5, out, cyc);
    long long h; cudaMemcpy(&h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    printf("%-10s chains=%d threads=%4d : %7.2f cycles per op-round, %6.2f cycles/op/warp-slot\n", name, CHAINS,
           threads, (double)h / iters, (double)h / iters / CHAINS);
    cudaFree(out); cudaFree(cyc);
}

int main() {
    run<1, 0>("DFMA", 32); run<2, 0>("DFMA", 32); run<4, 0>("DFMA", 32); run<8, 0>("DFMA", 32);
    run<8, 0>("DFMA", 128); run<8, 0>("DFMA", 256); run<4, 0>("DFMA", 256); run<2, 0>("DFMA", 256); run<1, 0>("DFMA", 256);
    run<1, 1>("DADD", 32); run<1, 2>("DMUL", 32);
    run<1, 3>("DIV", 32); run<4, 3>("DIV", 32);
    run<1, 7>("DRCP", 32); run<4, 7>("DRCP", 32);
    run<1, 4>("DSQRT", 32); run<4, 4>("DSQRT", 32);
    run<1, 6>("DRSQRT", 32); run<4, 6>("DRSQRT", 32);
    run<1, 5>("DSETP+SEL", 32);
    return 0;
}
