// Microbenchmark for the round-2 K1 restructuring: the cost of K1's OUTPUT stream alone (no transform), as a function of
// the tile width F (frames per row segment a warp instruction writes) on the reference layout [clip][3][1024][T] fp32.
// With T = 862 rows are only 8-byte aligned, so the first / last 32-byte sector of most row segments is written partially
// and L2 fills it from DRAM (round 1: 2.0 GB of fill reads per 2.7 GB written at F = 16).  F = 32 / 64 halve / quarter
// the number of seams per byte.  148 persistent CTAs x 512 threads, tiles dealt round-robin exactly like K1.
//   F = 16: lanes = 2 rows x 16 frames (two 64-byte segments per instruction)       -- today's K1
//   F = 32: lanes along 32 frames (one 128-byte segment per instruction)
//   F = 64: lanes along 64 frames, 8-byte stores (one 256-byte segment per instruction)
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o store_pattern store_pattern.cu
// Run:   ./store_pattern [F=16] [T=862] [prefetch 0/1] [plain stores 0/1] [shift 0/1: per-row sector-aligned windows (F = 16 / 32)]
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e_)); exit(1); } } while (0)

constexpr int kRows = 1024, kClips = 256;

template <int PLAIN>
__device__ __forceinline__ void st1(float* p, float v) {
    if (PLAIN) *p = v; else __stcs(p, v);
}
template <int PLAIN>
__device__ __forceinline__ void st2(float* p, float2 v) {
    if (PLAIN) *reinterpret_cast<float2*>(p) = v; else __stcs(reinterpret_cast<float2*>(p), v);
}

template <int F, int PLAIN>
__global__ void __launch_bounds__(512, 1) store_kernel(float* __restrict__ out, int n_tiles, int T, int pf, int shift) {
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int tiles_per_clip = (T + F - 1) / F;
    const long long rowB = 4ll * T, planeB = rowB * kRows;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int b = tile / tiles_per_clip, t0 = (tile % tiles_per_clip) * F;
        char* clip = reinterpret_cast<char*>(out) + (long long)b * 3 * planeB;
        if (pf) {
            const int nt = tile + gridDim.x;
            if (nt < n_tiles) {
                const int nb = nt / tiles_per_clip, nt0 = (nt % tiles_per_clip) * F;
                char* nclip = reinterpret_cast<char*>(out) + (long long)nb * 3 * planeB + 4ll * nt0;
                // 64 rows x 3 planes per warp: 6 per lane
#pragma unroll
                for (int i = 0; i < 6; ++i) {
                    const int idx = lane + 32 * i;          // 0..191
                    const int pl = idx / 64, rr = idx % 64;
                    const int row = ((rr & 1) ? 31 - warp : warp) + 32 * (rr >> 1);
                    const unsigned long long a = reinterpret_cast<unsigned long long>(nclip + pl * planeB + row * rowB);
                    if (a & 31u) asm volatile("prefetch.global.L2 [%0];" ::"l"(a));
                }
            }
        }
        const float v = (float)(tile + tid);
        if (F == 16) {
            const int h = lane >> 4, t = lane & 15;
            const int rho = h ? 31 - warp : warp;
            if (t0 + t < T) {
                char* p = clip + (long long)rho * rowB + 4ll * (t0 + t);
                if (shift) {
                    const int sft = (int)((reinterpret_cast<unsigned long long>(clip + (long long)rho * rowB + 4ll * t0) & 31u) >> 2);
                    p -= 4 * sft;
                    if (t0 + t - sft < 0) p += 4 * sft;   // first tile of a row: keep in bounds (duplicate stores)
                }
#pragma unroll 8
                for (int q = 0; q < 32; ++q) {
                    st1<PLAIN>(reinterpret_cast<float*>(p), v + q);
                    st1<PLAIN>(reinterpret_cast<float*>(p + planeB), v - q);
                    st1<PLAIN>(reinterpret_cast<float*>(p + 2 * planeB), v * q);
                    p += 32 * rowB;
                }
            }
        } else if (F == 32) {
            if (t0 + lane < T) {
#pragma unroll
                for (int hh = 0; hh < 2; ++hh) {
                    const int rho = hh ? 31 - warp : warp;
                    char* p = clip + (long long)rho * rowB + 4ll * (t0 + lane);
                    if (shift) {
                        const int sft = (int)((reinterpret_cast<unsigned long long>(clip + (long long)rho * rowB + 4ll * t0) & 31u) >> 2);
                        p -= 4 * sft;
                        if (t0 + lane - sft < 0) p += 4 * sft;
                    }
#pragma unroll 8
                    for (int q = 0; q < 32; ++q) {
                        st1<PLAIN>(reinterpret_cast<float*>(p), v + q);
                        st1<PLAIN>(reinterpret_cast<float*>(p + planeB), v - q);
                        st1<PLAIN>(reinterpret_cast<float*>(p + 2 * planeB), v * q);
                        p += 32 * rowB;
                    }
                }
            }
        } else {
            if (t0 + 2 * lane < T) {   // T even
#pragma unroll
                for (int hh = 0; hh < 2; ++hh) {
                    const int rho = hh ? 31 - warp : warp;
                    char* p = clip + (long long)rho * rowB + 4ll * (t0 + 2 * lane);
#pragma unroll 8
                    for (int q = 0; q < 32; ++q) {
                        st2<PLAIN>(reinterpret_cast<float*>(p), make_float2(v + q, v));
                        st2<PLAIN>(reinterpret_cast<float*>(p + planeB), make_float2(v - q, v));
                        st2<PLAIN>(reinterpret_cast<float*>(p + 2 * planeB), make_float2(v * q, v));
                        p += 32 * rowB;
                    }
                }
            }
        }
    }
}

int main(int argc, char** argv) {
    const int F = argc > 1 ? atoi(argv[1]) : 16;
    const int T = argc > 2 ? atoi(argv[2]) : 862;
    const int pf = argc > 3 ? atoi(argv[3]) : 0;
    const int plain = argc > 4 ? atoi(argv[4]) : 0;
    const int shift = argc > 5 ? atoi(argv[5]) : 0;
    const size_t elems = (size_t)kClips * 3 * kRows * T;
    float* out;
    CK(cudaMalloc(&out, elems * 4));
    const int n_tiles = kClips * ((T + F - 1) / F);
    auto run = [&](auto kern) {
        cudaEvent_t e0, e1;
        CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
        float best = 1e9f, sum = 0.f;
        const int reps = 12;
        for (int i = 0; i < reps + 3; ++i) {
            CK(cudaEventRecord(e0));
            kern<<<148, 512>>>(out, n_tiles, T, pf, shift);
            CK(cudaEventRecord(e1));
            CK(cudaEventSynchronize(e1));
            float ms;
            CK(cudaEventElapsedTime(&ms, e0, e1));
            if (i >= 3) { best = ms < best ? ms : best; sum += ms; }
        }
        CK(cudaGetLastError());
        const double gb = (double)elems * 4 / 1e9;
        printf("store F=%d T=%d pf=%d plain=%d shift=%d: best %.3f ms (%.0f GB/s)  mean %.3f ms\n", F, T, pf, plain, shift, best, gb / best * 1e3, sum / reps);
    };
    if (F == 16) { if (plain) run(store_kernel<16, 1>); else run(store_kernel<16, 0>); }
    else if (F == 32) { if (plain) run(store_kernel<32, 1>); else run(store_kernel<32, 0>); }
    else { if (plain) run(store_kernel<64, 1>); else run(store_kernel<64, 0>); }
    return 0;
}
