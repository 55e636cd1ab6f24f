// D2H bandwidth of the copy engine: one 2-D pitched copy (16 rows of 4 MB out of 8 MB-pitch rows) vs 16 1-D copies of
// 4 MB vs one 64 MB 1-D copy, pinned host memory.
#include <cstdio>
#include <cuda_runtime.h>
int main() {
    const size_t pitch = 8u << 20, width = 4u << 20, rows = 16;
    char *d, *h;
    cudaMalloc(&d, pitch * rows);
    cudaHostAlloc(&h, pitch * rows, cudaHostAllocDefault);
    cudaStream_t s; cudaStreamCreate(&s);
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    float ms;
    for (int mode = 0; mode < 3; ++mode) {
        for (int rep = 0; rep < 3; ++rep) {
            cudaEventRecord(a, s);
            for (int it = 0; it < 20; ++it) {
                if (mode == 0) cudaMemcpy2DAsync(h, pitch, d, pitch, width, rows, cudaMemcpyDeviceToHost, s);
                else if (mode == 1) for (size_t r = 0; r < rows; ++r) cudaMemcpyAsync(h + r * pitch, d + r * pitch, width, cudaMemcpyDeviceToHost, s);
                else cudaMemcpyAsync(h, d, width * rows, cudaMemcpyDeviceToHost, s);
            }
            cudaEventRecord(b, s); cudaEventSynchronize(b); cudaEventElapsedTime(&ms, a, b);
        }
        printf("%-28s %.1f GB/s\n", mode == 0 ? "2-D pitched (16 x 4 MB)" : mode == 1 ? "16 x 1-D 4 MB" : "1-D 64 MB", 20.0 * width * rows / ms / 1e6);
    }
    return 0;
}
