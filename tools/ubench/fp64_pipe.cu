// Microbenchmark: issue rate of FP64 instruction flavours on B200 (warp-instructions / cycle / SM).
#include <cstdio>
#include <cuda_runtime.h>
__constant__ double c_k[8] = {0.999999999, 1e-9, 1.0000001, 0.5, 0.25, 3.0, 7.0, 9.0};

template <int MODE>
__global__ void __launch_bounds__(256) k(double *sink, int iters, double seed) {
    double a[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = seed + threadIdx.x + i;
    const double m = c_k[0] + seed * 1e-30, c = c_k[1] + seed * 1e-30;  // registers
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                if (MODE == 0) a[i] = fma(a[i], m, c);                 // DFMA reg,reg,reg
                if (MODE == 1) a[i] = a[i] * m;                        // DMUL reg
                if (MODE == 2) a[i] = a[i] + c;                        // DADD reg
                if (MODE == 3) a[i] = fma(a[i], c_k[2], c_k[1]);       // DFMA with constant-bank operands
                if (MODE == 4) a[i] = (a[i] > c) ? a[i] - c : a[i] + m; // DSETP + select + DADD mix
                if (MODE == 5) a[i] = fma(a[i], m, 0.5);               // DFMA with immediate addend
            }
        }
    }
    double s = 0; for (int i = 0; i < 8; ++i) s += a[i];
    if (s == 12345.678) sink[0] = s;
}

template <int MODE> void run(const char *name, int sms) {
    double *sink; cudaMalloc(&sink, 8);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int grid = sms * 8, block = 256, iters = 4000;
    k<MODE><<<grid, block>>>(sink, 100, 1.0);
    cudaEventRecord(e0); k<MODE><<<grid, block>>>(sink, iters, 1.0); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double winst = (double)iters * 64 * grid * block / 32;   // FP64-ish warp instructions (MODE 4: per element op)
    int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    printf("%-34s %8.3f ms  %.3f warp-instr/clk/SM (at %d MHz nominal)  %.2f T thread-op/s\n", name, ms,
           winst / (ms * 1e-3) / (clk * 1e3) / sms, clk / 1000, winst * 32 / (ms * 1e-3) / 1e12);
    cudaFree(sink);
}
int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    run<0>("DFMA reg", p.multiProcessorCount); run<1>("DMUL reg", p.multiProcessorCount);
    run<2>("DADD reg", p.multiProcessorCount); run<3>("DFMA const-bank operands", p.multiProcessorCount);
    run<4>("DSETP+FSEL+DADD", p.multiProcessorCount); run<5>("DFMA imm addend", p.multiProcessorCount);
    return 0;
}
