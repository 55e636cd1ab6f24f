// Microbenchmark: dependent-chain latency of FP64 ops and how many warps per SMSP saturate the pipe.
#include <cstdio>
#include <cuda_runtime.h>
template <int CHAINS>
__global__ void k(double *sink, int iters, double seed, long long *cycles) {
    double a[CHAINS];
    for (int i = 0; i < CHAINS; ++i) a[i] = seed + threadIdx.x + i;
    const double m = 0.999999999 + seed * 1e-30, c = 1e-9 + seed * 1e-30;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 16; ++u)
#pragma unroll
            for (int i = 0; i < CHAINS; ++i) a[i] = fma(a[i], m, c);
    }
    long long t1 = clock64();
    double s = 0; for (int i = 0; i < CHAINS; ++i) s += a[i];
    if (s == 12345.678) sink[0] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) cycles[0] = t1 - t0;
}
template <int CHAINS> void run(int warps_per_sm) {
    double *sink; long long *cyc, h; cudaMalloc(&sink, 8); cudaMalloc(&cyc, 8);
    const int iters = 2000;
    k<CHAINS><<<148, warps_per_sm * 32>>>(sink, iters, 1.0, cyc);
    cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    const double per_inst = (double)h / (iters * 16.0 * CHAINS);
    printf("chains/thread %d, warps/SM %2d (%.1f per SMSP): %.2f cycles per DFMA per warp -> %.3f DFMA/clk/SMSP\n", CHAINS,
           warps_per_sm, warps_per_sm / 4.0, per_inst, (warps_per_sm / 4.0) / per_inst);
}
int main() {
    run<1>(1); run<2>(1); run<4>(1);
    run<1>(4); run<1>(8); run<1>(16); run<1>(24); run<1>(32);
    run<2>(8); run<2>(16); run<2>(24);
    return 0;
}
