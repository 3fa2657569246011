// Microbenchmark: scalar FFMA vs packed FFMA2 (fma.rn.f32x2) issue/throughput on sm_100a,
// alone and mixed with ALU-pipe integer ops.  Build: nvcc -arch=sm_100a -O3 -o ffma2 ffma2.cu
#include <cstdio>
#include <cuda_runtime.h>
#define ITERS 4096
template <int MODE>
__global__ void k(float* out, float a, float b) {
    float2 x[8];
    unsigned u[4] = {threadIdx.x, threadIdx.x * 3u, 7u, 11u};
#pragma unroll
    for (int i = 0; i < 8; ++i) x[i] = make_float2(threadIdx.x * 0.001f + i, i * 0.5f);
    const float2 aa = make_float2(a, a), bb = make_float2(b, b);
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (MODE == 0 || MODE == 2) {  // scalar: 2 FFMA per pair
                x[i].x = fmaf(x[i].x, a, b);
                x[i].y = fmaf(x[i].y, a, b);
            } else {  // packed: 1 FFMA2 per pair
                x[i] = __ffma2_rn(x[i], aa, bb);
            }
        }
        if (MODE >= 2) {  // plus 4 ALU-pipe ops per 8 pairs
#pragma unroll
            for (int j = 0; j < 4; ++j) u[j] = (u[j] ^ (u[(j + 1) & 3] >> 3)) + 0x9e3779b9u;
        }
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += x[i].x + x[i].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s + (float)(u[0] ^ u[1] ^ u[2] ^ u[3]);
}
template <int MODE>
void run(const char* name, float* d) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<MODE><<<148 * 4, 512>>>(d, 1.0001f, 0.5f);
    cudaEventRecord(e0);
    k<MODE><<<148 * 4, 512>>>(d, 1.0001f, 0.5f);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double fma = 148.0 * 4 * 512 * (double)ITERS * 16;
    printf("%-28s %8.3f ms  %7.2f TFMA/s (%.1f TFLOP/s)\n", name, ms, fma / ms * 1e-9, 2 * fma / ms * 1e-9);
}
int main() {
    float* d; cudaMalloc(&d, 148 * 4 * 512 * 4);
    run<0>("scalar FFMA", d);
    run<1>("packed FFMA2", d);
    run<2>("scalar FFMA + ALU", d);
    run<3>("packed FFMA2 + ALU", d);
    cudaError_t e = cudaDeviceSynchronize();
    printf("status %s\n", cudaGetErrorString(e));
    return 0;
}
