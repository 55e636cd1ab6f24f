// Microbenchmark: is the FP64 issue rate bounded by register-operand reads?  One dependent chain per thread,
// 6 warps per SMSP (enough to cover the 8-cycle latency), different operand sources.
#include <cstdio>
#include <cuda_runtime.h>
__constant__ double c_k[4] = {0.999999999, 1e-9, 1.0000001, 0.5};
template <int MODE>
__global__ void k(double *sink, int iters, double seed, long long *cycles) {
    double a = seed + threadIdx.x, b = a + 1.0;
    const double m = 0.999999999 + seed * 1e-30, c = 1e-9 + seed * 1e-30;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 32; ++u) {
            if (MODE == 0) a = fma(a, m, c);            // 3 register operands
            if (MODE == 1) a = fma(a, c_k[0], c);       // 2 registers + constant/uniform
            if (MODE == 2) a = fma(a, c_k[0], c_k[1]);  // 1 register + 2 constants -> (1 const + 1 UR?)
            if (MODE == 3) a = a * m;                   // DMUL 2 registers
            if (MODE == 4) a = a * c_k[0];              // DMUL 1 register + constant
            if (MODE == 5) a = fma(a, a, c);            // 2 distinct registers (a twice)
            if (MODE == 6) a = fma(a, m, 0.5);          // 2 registers + immediate
            if (MODE == 7) { a = fma(a, m, c); b = fma(b, m, c); }  // 2 chains, shared m, c (reuse cache)
        }
    }
    long long t1 = clock64();
    if (a + b == 12345.678) sink[0] = a;
    if (threadIdx.x == 0 && blockIdx.x == 0) cycles[0] = t1 - t0;
}
template <int MODE> void run(const char *name) {
    double *sink; long long *cyc, h; cudaMalloc(&sink, 8); cudaMalloc(&cyc, 8);
    const int iters = 1000, warps_per_sm = 24;
    k<MODE><<<148, warps_per_sm * 32>>>(sink, iters, 1.0, cyc);
    cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    const double n = iters * 32.0 * (MODE == 7 ? 2 : 1);
    printf("%-44s %.3f FP64 instr/clk/SMSP\n", name, (warps_per_sm / 4.0) * n / (double)h);
}
int main() {
    run<0>("DFMA r,r,r (1 chain)"); run<1>("DFMA r,const,r"); run<2>("DFMA r,const,const"); run<3>("DMUL r,r");
    run<4>("DMUL r,const"); run<5>("DFMA r,r(same),r"); run<6>("DFMA r,r,imm"); run<7>("DFMA r,r,r 2 chains sharing m,c");
    return 0;
}
