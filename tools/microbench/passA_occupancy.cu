// Microbenchmark for the round-2 plan: does the HBM-facing pass of the inverse kernel get faster
// with twice the resident warps?  Two mock kernels do pass A of K2 -- 96 / 48 spectrogram loads per thread (lanes along
// 16 frames, two 8-byte-aligned 64-byte row segments per warp instruction), the packed expansion, the partner-lane
// pairing shuffles, a register-resident radix-Q DFT and the exchange stores -- on the config-2 spectrogram
// (256 x [3][1024][862] fp32):
//     Q = 32: 512 threads x 128 registers  (today's structure, 16 warps per SM)
//     Q = 16: 1024 threads x 64 registers  (the planned structure, 32 warps per SM; each thread owns half the bins of
//             a residue class -- the cross-lane radix-2 that would recombine the halves is modelled by 32 extra shuffles)
// The numbers produced are not a transform (no pass B); only the time matters.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I ../../audio_intelligence_b200/csrc -o passA_occupancy passA_occupancy.cu
#include <cstdio>
#include <type_traits>
#include <cuda_runtime.h>

#include "fftx2.cuh"

using namespace a2sb;

template <int I, int N, class Fn>
__device__ __forceinline__ void static_for(Fn&& f) {
    if constexpr (I < N) {
        f(std::integral_constant<int, I>{});
        static_for<I + 1, N>(f);
    }
}

constexpr int kM = 1024, kRB = 32, kF = 16, kT = 862, kRows = 1024, kClips = 256;

__device__ __forceinline__ float ld(const float* p) {
    float v;
    asm volatile("ld.global.nc.L2::256B.f32 %0, [%1];" : "=f"(v) : "l"(p));
    return v;
}

template <int Q>
__global__ void __launch_bounds__(kF * kRB * (32 / Q), 1) passA(const float* __restrict__ spec, float* __restrict__ sink, int n_tiles) {
    extern __shared__ float s_x[];                       // exchange: 16 frames x (2 * 1024 + 34) floats
    constexpr int PARTS = 32 / Q, FS = 2 * kM + 34;
    const int tid = threadIdx.x, part = tid / (kF * kRB), lt = tid % (kF * kRB);
    const int warp = lt >> 5, lane = lt & 31, h = lane >> 4, t = lane & 15;
    const int c = warp;                                  // residue class, ja = c (h = 0) / 32 - c (h = 1)
    const int ja = (c == 0) ? (h ? 16 : 1) : (h ? kRB - c : c);
    const unsigned long long rowB = 4ull * kT, stepB = (unsigned long long)kRB * rowB * PARTS, planeB = rowB * kRows,
                             plane2B = 2 * planeB;
    const int tiles_per_clip = (kT + kF - 1) / kF;
    float acc = 0.f;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int b = tile / tiles_per_clip, t0 = (tile % tiles_per_clip) * kF;
        const int tg = t0 + t;
        float xr[Q], xi[Q];
        if (tg < kT) {
            unsigned long long a = reinterpret_cast<unsigned long long>(spec) + (unsigned long long)b * 3 * planeB +
                                   (unsigned long long)(ja - 1 + kRB * part) * rowB + 4ull * tg;
#pragma unroll
            for (int j = 0; j < Q / 2; ++j) {
                const unsigned long long a1 = a + stepB;
                float2 m, cc, ss;
                m.x = ld(reinterpret_cast<const float*>(a)); cc.x = ld(reinterpret_cast<const float*>(a + planeB));
                ss.x = ld(reinterpret_cast<const float*>(a + plane2B));
                m.y = ld(reinterpret_cast<const float*>(a1)); cc.y = ld(reinterpret_cast<const float*>(a1 + planeB));
                ss.y = ld(reinterpret_cast<const float*>(a1 + plane2B));
                a = a1 + stepB;
                // expansion m*|m|^4/(|m|+eps) * (c,s)/sqrt(c^2+s^2), packed over the two bins
                const float2 a2 = p2_mul(m, m), a4 = p2_mul(a2, a2);
                float2 r, rn;
                r.x = rcp_approx(fabsf(m.x) + 1e-9f); r.y = rcp_approx(fabsf(m.y) + 1e-9f);
                const float2 n2 = p2_fma(cc, cc, p2_mul(ss, ss));
                rn.x = rsqrt_approx(n2.x); rn.y = rsqrt_approx(n2.y);
                const float2 g = p2_mul(p2_mul(m, p2_mul(a4, r)), rn);
                const float2 vr = p2_mul(g, cc), vi = p2_mul(g, ss);
                xr[2 * j] = vr.x; xi[2 * j] = vi.x; xr[2 * j + 1] = vr.y; xi[2 * j + 1] = vi.y;
            }
        } else {
#pragma unroll
            for (int q = 0; q < Q; ++q) { xr[q] = 0.f; xi[q] = 0.f; }
        }
        // pairing with the partner lane group (k <-> M - k), as in K2
#pragma unroll
        for (int q = 0; q < Q / 2; ++q) {
            const float xmr = __shfl_xor_sync(0xffffffffu, xr[Q - 1 - q], kF), xmi = __shfl_xor_sync(0xffffffffu, xi[Q - 1 - q], kF);
            const float er = xr[q] + xmr, ei = xi[q] - xmi, dr = xr[q] - xmr, di = xi[q] + xmi;
            const float w = 0.7071f, pr = -(w * di + w * dr), pi = w * dr - w * di;
            const float zmr = er - pr, zmi = pi - ei;
            xr[q] = er + pr; xi[q] = ei + pi;
            xr[Q - 1 - q] = __shfl_xor_sync(0xffffffffu, zmr, kF);
            xi[Q - 1 - q] = __shfl_xor_sync(0xffffffffu, zmi, kF);
        }
        // radix-Q DFT in registers (scalar DIF stage + packed radix-Q/2 x 2)
        float2 pre[Q / 2], pim[Q / 2];
        static_for<0, Q / 2>([&](auto QQ) {
            constexpr int q = decltype(QQ)::value;
            dif_first<Q, +1, q>(xr[q], xi[q], xr[q + Q / 2], xi[q + Q / 2], pre[q], pim[q]);
        });
        fft_v<Q / 2, +1, float2>(pre, pim);
        if (PARTS > 1) {   // model of the cross-lane radix-2 that recombines the two halves: one shuffle per value
#pragma unroll
            for (int k = 0; k < Q / 2; ++k) {
                pre[k].x += __shfl_xor_sync(0xffffffffu, pre[k].y, 1); pim[k].x += __shfl_xor_sync(0xffffffffu, pim[k].y, 1);
            }
        }
        float* dst = s_x + t * FS + (h ? 16 * 32 + 16 : 0) + (c % 16) * 32 + part * Q;
#pragma unroll
        for (int k = 0; k < Q / 2; ++k) {
            *reinterpret_cast<float2*>(dst + 2 * k) = pre[k];
            *reinterpret_cast<float2*>(dst + kM + 16 + 2 * k) = pim[k];
        }
        __syncthreads();
        acc += s_x[(tid * 7) % (kF * FS)];
        __syncthreads();
    }
    sink[blockIdx.x * blockDim.x + tid] = acc;
}


// Variant for the cluster-pair plan: lanes run along 32 frames, so a warp instruction reads ONE row segment of 128 bytes
// (a full L1 line when the rows are aligned, two lines otherwise) instead of two segments of 64 bytes; a CTA covers 16 of
// the 32 residue classes of a 32-frame tile (its cluster partner would cover the other 16 and the exchange would be split
// by frames over the two CTAs' shared memories).  Same bytes, same arithmetic per thread as passA<32>.
__global__ void __launch_bounds__(512, 1) passA_wide(const float* __restrict__ spec, float* __restrict__ sink, int n_items, int T_pitch) {
    extern __shared__ float s_x[];
    constexpr int Q = 32, FS = 2 * kM + 34;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const unsigned long long rowB = 4ull * T_pitch, stepB = (unsigned long long)kRB * rowB, planeB = rowB * kRows, plane2B = 2 * planeB;
    const int tiles_per_clip = (kT + 31) / 32;
    float acc = 0.f;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
        const int half = item & 1, tile = item >> 1;
        const int b = tile / tiles_per_clip, t0 = (tile % tiles_per_clip) * 32;
        const int tg = t0 + lane;
        const int c = warp + 16 * half;                  // residue class
        const int ja = (c == 0) ? 1 : c;
        float xr[Q], xi[Q];
        if (tg < kT) {
            unsigned long long a = reinterpret_cast<unsigned long long>(spec) + (unsigned long long)b * 3 * planeB +
                                   (unsigned long long)(ja - 1) * rowB + 4ull * tg;
#pragma unroll
            for (int j = 0; j < Q / 2; ++j) {
                const unsigned long long a1 = a + stepB;
                float2 m, cc, ss;
                m.x = ld(reinterpret_cast<const float*>(a)); cc.x = ld(reinterpret_cast<const float*>(a + planeB));
                ss.x = ld(reinterpret_cast<const float*>(a + plane2B));
                m.y = ld(reinterpret_cast<const float*>(a1)); cc.y = ld(reinterpret_cast<const float*>(a1 + planeB));
                ss.y = ld(reinterpret_cast<const float*>(a1 + plane2B));
                a = a1 + stepB;
                const float2 a2 = p2_mul(m, m), a4 = p2_mul(a2, a2);
                float2 r, rn;
                r.x = rcp_approx(fabsf(m.x) + 1e-9f); r.y = rcp_approx(fabsf(m.y) + 1e-9f);
                const float2 n2 = p2_fma(cc, cc, p2_mul(ss, ss));
                rn.x = rsqrt_approx(n2.x); rn.y = rsqrt_approx(n2.y);
                const float2 g = p2_mul(p2_mul(m, p2_mul(a4, r)), rn);
                const float2 vr = p2_mul(g, cc), vi = p2_mul(g, ss);
                xr[2 * j] = vr.x; xi[2 * j] = vi.x; xr[2 * j + 1] = vr.y; xi[2 * j + 1] = vi.y;
            }
        } else {
#pragma unroll
            for (int q = 0; q < Q; ++q) { xr[q] = 0.f; xi[q] = 0.f; }
        }
#pragma unroll
        for (int q = 0; q < Q / 2; ++q) {
            const float xmr = __shfl_xor_sync(0xffffffffu, xr[Q - 1 - q], kF), xmi = __shfl_xor_sync(0xffffffffu, xi[Q - 1 - q], kF);
            const float er = xr[q] + xmr, ei = xi[q] - xmi, dr = xr[q] - xmr, di = xi[q] + xmi;
            const float w = 0.7071f, pr = -(w * di + w * dr), pi = w * dr - w * di;
            const float zmr = er - pr, zmi = pi - ei;
            xr[q] = er + pr; xi[q] = ei + pi;
            xr[Q - 1 - q] = __shfl_xor_sync(0xffffffffu, zmr, kF);
            xi[Q - 1 - q] = __shfl_xor_sync(0xffffffffu, zmi, kF);
        }
        float2 pre[Q / 2], pim[Q / 2];
        static_for<0, Q / 2>([&](auto QQ) {
            constexpr int q = decltype(QQ)::value;
            dif_first<Q, +1, q>(xr[q], xi[q], xr[q + Q / 2], xi[q + Q / 2], pre[q], pim[q]);
        });
        fft_v<Q / 2, +1, float2>(pre, pim);
        float* dst = s_x + (lane & 15) * FS + ((lane >> 4) ? 16 * 32 + 16 : 0) + warp * 32;
#pragma unroll
        for (int k = 0; k < Q / 2; ++k) {
            *reinterpret_cast<float2*>(dst + 2 * k) = pre[k];
            *reinterpret_cast<float2*>(dst + kM + 16 + 2 * k) = pim[k];
        }
        __syncthreads();
        acc += s_x[(tid * 7) % (kF * FS)];
        __syncthreads();
    }
    sink[blockIdx.x * blockDim.x + tid] = acc;
}

template <int Q>
static void run(const float* spec, float* sink, const char* name) {
    constexpr int NT = kF * kRB * (32 / Q);
    const size_t smem = sizeof(float) * kF * (2 * kM + 34);
    cudaFuncSetAttribute(passA<Q>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    const int n_tiles = kClips * ((kT + kF - 1) / kF);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int i = 0; i < 3; ++i) passA<Q><<<148, NT, smem>>>(spec, sink, n_tiles);
    cudaEventRecord(e0);
    for (int i = 0; i < 10; ++i) passA<Q><<<148, NT, smem>>>(spec, sink, n_tiles);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    ms /= 10;
    const double bytes = 4.0 * 3 * kRows * kT * kClips;
    int regs = 0;
    cudaFuncAttributes fa;
    cudaFuncGetAttributes(&fa, passA<Q>);
    regs = fa.numRegs;
    std::printf("%s: %d threads, %d registers, %.3f ms, %.0f GB/s  (%s)\n", name, NT, regs, ms, bytes / ms * 1e-6,
                cudaGetErrorString(cudaGetLastError()));
}

static void run_wide(const float* spec, float* sink, int pitch, const char* name) {
    const size_t smem = sizeof(float) * kF * (2 * kM + 34);
    cudaFuncSetAttribute(passA_wide, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    const int n_items = kClips * ((kT + 31) / 32) * 2;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int i = 0; i < 3; ++i) passA_wide<<<148, 512, smem>>>(spec, sink, n_items, pitch);
    cudaEventRecord(e0);
    for (int i = 0; i < 10; ++i) passA_wide<<<148, 512, smem>>>(spec, sink, n_items, pitch);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    ms /= 10;
    const double bytes = 4.0 * 3 * kRows * kT * kClips;
    std::printf("%s: row pitch %d frames, %.3f ms, %.0f GB/s  (%s)\n", name, pitch, ms, bytes / ms * 1e-6,
                cudaGetErrorString(cudaGetLastError()));
}

int main() {
    float *spec, *sink;
    const size_t n = (size_t)3 * kRows * 896 * kClips;   // room for the pitched variants
    cudaMalloc(&spec, n * sizeof(float));
    cudaMalloc(&sink, 148 * 1024 * sizeof(float));
    cudaMemset(spec, 0x3c, n * sizeof(float));           // finite, non-trivial fp32 pattern (0x3c3c3c3c = 0.0115)
    run<32>(spec, sink, "Q=32 (16 warps/SM)");
    run<16>(spec, sink, "Q=16 (32 warps/SM)");
    run_wide(spec, sink, kT, "32-frame row segments, contiguous rows (8-byte aligned)");
    run_wide(spec, sink, 864, "32-frame row segments, rows pitched to 32 bytes");
    run_wide(spec, sink, 896, "32-frame row segments, rows pitched to 128 bytes");
    cudaFree(spec); cudaFree(sink);
    return 0;
}
