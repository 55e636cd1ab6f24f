// Microbenchmark: does an FP64 instruction block the SMSP's issue port for both of its pipe cycles?
// One dependent DFMA chain per thread (two register operands + constant: 2 pipe cycles each) interleaved with
// K independent integer (IMAD) / FP32 (FFMA) / shared-memory (LDS) instructions per DFMA, 6 warps per SMSP.
// If the issue port were free in the DFMA's second cycle, K = 1 would leave the DFMA rate at ~0.5 / clk.
#include <cstdio>
#include <cuda_runtime.h>
__constant__ double c_k[4] = {0.999999999, 1e-9, 1.0000001, 0.5};
template <int K, int KIND>
__global__ void k(double *sink, int iters, double seed, long long *cycles) {
    __shared__ int sm[1024];
    sm[threadIdx.x] = threadIdx.x;
    __syncthreads();
    double a = seed + threadIdx.x;
    const double c = 1e-9 + seed * 1e-30;
    int i0 = threadIdx.x, i1 = i0 + 1, i2 = i0 + 2, i3 = i0 + 3;
    float f0 = (float)seed, f1 = f0 + 1.f, f2 = f0 + 2.f, f3 = f0 + 3.f;
    const int mi = (int)seed + 2;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 32; ++u) {
            a = fma(a, c_k[0], c);
            if (KIND == 0) {   // IMAD
                if (K >= 1) i0 = i0 * mi + 1;
                if (K >= 2) i1 = i1 * mi + 1;
                if (K >= 3) i2 = i2 * mi + 1;
                if (K >= 4) i3 = i3 * mi + 1;
            } else if (KIND == 1) {  // FFMA
                if (K >= 1) f0 = fmaf(f0, 0.999f, 1e-3f);
                if (K >= 2) f1 = fmaf(f1, 0.999f, 1e-3f);
                if (K >= 3) f2 = fmaf(f2, 0.999f, 1e-3f);
                if (K >= 4) f3 = fmaf(f3, 0.999f, 1e-3f);
            } else {                 // LDS (dependent address chain per stream)
                if (K >= 1) i0 = sm[i0 & 1023];
                if (K >= 2) i1 = sm[i1 & 1023];
                if (K >= 3) i2 = sm[i2 & 1023];
                if (K >= 4) i3 = sm[i3 & 1023];
            }
        }
    }
    long long t1 = clock64();
    if (a + i0 + i1 + i2 + i3 + f0 + f1 + f2 + f3 == 12345.678) sink[0] = a;
    if (threadIdx.x == 0 && blockIdx.x == 0) cycles[0] = t1 - t0;
}
template <int K, int KIND> void run(const char *name) {
    double *sink; long long *cyc, h; cudaMalloc(&sink, 8); cudaMalloc(&cyc, 8);
    const int iters = 1000, warps_per_sm = 24;
    k<K, KIND><<<148, warps_per_sm * 32>>>(sink, iters, 1.0, cyc);
    cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    const double n = iters * 32.0;
    const double r = (warps_per_sm / 4.0) * n / (double)h;
    printf("%-28s K=%d: DFMA %.3f /clk/SMSP, all instr %.3f /clk/SMSP, issue-port model (2 per DFMA) %.3f\n", name, K, r,
           r * (1 + K), r * (2 + K));
}
int main() {
    run<0, 0>("DFMA alone");
    run<1, 0>("DFMA + IMAD"); run<2, 0>("DFMA + IMAD"); run<3, 0>("DFMA + IMAD"); run<4, 0>("DFMA + IMAD");
    run<1, 1>("DFMA + FFMA"); run<2, 1>("DFMA + FFMA"); run<4, 1>("DFMA + FFMA");
    run<1, 2>("DFMA + LDS");  run<2, 2>("DFMA + LDS");
    return 0;
}
