// Microbenchmark for the round-2 plan: a mock of ALL three phases of the inverse kernel K2 -- pass A
// (96 / 48 spectrogram loads per thread, expansion, pairing shuffles, radix-Q DFT, inter-pass twiddle, exchange stores),
// pass B (exchange loads, radix-Q DFT, window, frame-buffer stores) and the overlap-add (4 frame-buffer loads + carry per
// output, 128-bit coalesced stores) with their three barriers -- at two occupancies:
//     Q = 32: 512 threads x 128 registers, 16 warps per SM (today's K2)
//     Q = 16: 1024 threads x 64 registers, 32 warps per SM; a thread owns half the points of each radix-32 DFT and the
//             cross-lane radix-2 that recombines the halves is modelled by one shuffle per value in both passes.
// Same memory traffic, same shared-memory footprint, same instruction mix per tile; the numbers are not a transform.
// passA_occupancy.cu showed the memory-facing pass alone gains 4 % from the doubled occupancy; this one asks whether the
// compute phases do.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I ../../audio_intelligence_b200/csrc -o k2_phases_occupancy k2_phases_occupancy.cu
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

constexpr int kM = 1024, kRB = 32, kF = 16, kT = 862, kRows = 1024, kClips = 256, kHop = 512, kN = 2048;
constexpr int kFS = 2 * kM + 34, kNC = kN - kHop, kOutLen = kHop * (kT - 1);

__device__ __forceinline__ float ld(const float* p) {
    float v;
    asm volatile("ld.global.nc.L2::256B.f32 %0, [%1];" : "=f"(v) : "l"(p));
    return v;
}

template <int Q, bool WIDE = false>
__global__ void __launch_bounds__(kF * kRB * (32 / Q), 1)
mock_k2(const float* __restrict__ spec, float* __restrict__ out, const float4* __restrict__ tw4, const float* __restrict__ win,
        const float* __restrict__ ienv, long long* __restrict__ phase_cycles) {
    long long cyc[4] = {0, 0, 0, 0}, stamp = clock64();
#define PHASE_END(i) { const long long now = clock64(); cyc[i] += now - stamp; stamp = now; }
    extern __shared__ __align__(16) float smem[];
    constexpr int PARTS = 32 / Q, NT = kF * kRB * PARTS;
    float* s_win = smem;                                  // [2048]
    float4* s_tw4 = reinterpret_cast<float4*>(smem + kN); // [32][17]
    float* s_x = smem + kN + 4 * 32 * 17;                 // exchange / frame buffers
    float* s_c0 = s_x + kF * kFS + 4;
    float* s_c1 = s_c0 + kNC;
    const int tid = threadIdx.x, part = tid / (kF * kRB), lt = tid % (kF * kRB);
    const int warp = lt >> 5, lane = lt & 31, h = lane >> 4, t = lane & 15;
    const int c = warp;
    const int ja = (c == 0) ? (h ? 16 : 1) : (h ? kRB - c : c);
    for (int i = tid; i < kN; i += NT) s_win[i] = win[i];
    for (int i = tid; i < 32 * 17; i += NT) s_tw4[i] = tw4[i];
    __syncthreads();
    const unsigned long long rowB = 4ull * kT, stepB = (unsigned long long)kRB * rowB * PARTS, planeB = rowB * kRows,
                             plane2B = 2 * planeB;
    constexpr int tiles_per_clip = (kT + 3 + kF - 1) / kF;   // 55: three lead-in frames like K2
    for (int b = blockIdx.x; b < kClips; b += gridDim.x) {
        float *carry_cur = s_c0, *carry_nxt = s_c1;
        for (int i = tid; i < kNC; i += NT) carry_cur[i] = 0.f;
        for (int tile = 0; tile < tiles_per_clip; ++tile) {
            const int t0 = tile * kF - 3;
            // WIDE: lanes along 32 frames (one 128-byte row segment per warp instruction), 16 of the 32 classes per CTA-tile;
            // the other 16 classes and the exchange of frames 16..31 would belong to the cluster partner
            const int tg = WIDE ? (tile >> 1) * 32 - 3 + lane : t0 + t;
            const int jrow = WIDE ? ((warp + 16 * (tile & 1)) == 0 ? 1 : warp + 16 * (tile & 1)) : ja;
            // ---------------- pass A ----------------
            {
                float xr[Q], xi[Q];
                if (tg >= 0 && tg < kT) {
                    unsigned long long a = reinterpret_cast<unsigned long long>(spec) + (unsigned long long)b * 3 * planeB +
                                           (unsigned long long)(jrow - 1 + kRB * part) * rowB + 4ull * tg;
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
                if (PARTS > 1) {
#pragma unroll
                    for (int k = 0; k < Q / 2; ++k) {
                        pre[k].x += __shfl_xor_sync(0xffffffffu, pre[k].y, 1); pim[k].x += __shfl_xor_sync(0xffffffffu, pim[k].y, 1);
                    }
                }
                // inter-pass twiddle (one float4 of two (c, s) pairs per packed value)
                const float4* twa = s_tw4 + (ja & 31) * 17;
#pragma unroll
                for (int k = 0; k < Q / 2; ++k) {
                    const float4 w = twa[(k + part * (Q / 2)) & 15];
                    const float2 wc = make_float2(w.x, w.z), ws = make_float2(w.y, w.w);
                    const float2 nr = p2_fma(pre[k], wc, p2_neg(p2_mul(pim[k], ws)));
                    pim[k] = p2_fma(pre[k], ws, p2_mul(pim[k], wc));
                    pre[k] = nr;
                }
                float* dst = s_x + t * kFS + (h ? 16 * 32 + 16 : 0) + (c % 16) * 32 + part * Q;
#pragma unroll
                for (int k = 0; k < Q / 2; ++k) {
                    *reinterpret_cast<float2*>(dst + 2 * k) = pre[k];
                    *reinterpret_cast<float2*>(dst + kM + 16 + 2 * k) = pim[k];
                }
            }
            __syncthreads();
            PHASE_END(0)
            // ---------------- pass B ----------------
            {
                float xr[Q], xi[Q];
                const float* src = s_x + t * kFS + (h ? 16 : 0) + warp + part * (Q * 32);   // stride-32 words: conflict-free across t? (16-bank skew per h)
#pragma unroll
                for (int q = 0; q < Q; ++q) { xr[q] = src[(q * 32) % kM]; xi[q] = src[kM + 16 + (q * 32) % kM]; }
                float2 pre[Q / 2], pim[Q / 2];
                static_for<0, Q / 2>([&](auto QQ) {
                    constexpr int q = decltype(QQ)::value;
                    dif_first<Q, +1, q>(xr[q], xi[q], xr[q + Q / 2], xi[q + Q / 2], pre[q], pim[q]);
                });
                fft_v<Q / 2, +1, float2>(pre, pim);
                if (PARTS > 1) {
#pragma unroll
                    for (int k = 0; k < Q / 2; ++k) {
                        pre[k].x += __shfl_xor_sync(0xffffffffu, pre[k].y, 1); pim[k].x += __shfl_xor_sync(0xffffffffu, pim[k].y, 1);
                    }
                }
                __syncthreads();   // every thread has read its exchange values: the frame buffers may overwrite them
                float* dst = s_x + t * kFS + 2 * (t & 1) + 2 * (warp + 32 * h) + part * 1024;
#pragma unroll
                for (int k = 0; k < Q / 2; ++k) {
                    const int n = 2 * (warp + 32 * h) + 128 * k;
                    const float2 w0 = *reinterpret_cast<const float2*>(s_win + (n & (kN - 2)));
                    const float2 w1 = *reinterpret_cast<const float2*>(s_win + ((n + 64) & (kN - 2)));
                    *reinterpret_cast<float2*>(dst + ((128 * k) & 1023)) = p2_mul(make_float2(pre[k].x, pim[k].x), w0);
                    *reinterpret_cast<float2*>(dst + ((128 * k + 64) & 1023)) = p2_mul(make_float2(pre[k].y, pim[k].y), w1);
                }
            }
            __syncthreads();
            PHASE_END(1)
            // ---------------- overlap-add ----------------
            {
                float* o = out + (long long)b * kOutLen + (long long)(t0 < 0 ? 0 : t0) * kHop;
                const bool in_range = t0 >= 0 && t0 + kF <= kT - 1;
#pragma unroll
                for (int i = 0; i < 4 / PARTS; ++i) {
                    const int pos = 4 * (tid + NT * i);           // 0 .. 8188: output sample of the tile
                    const int u = pos >> 9, r = pos & 511;        // hop slot, offset inside the hop
                    float4 acc = (pos < kNC) ? *reinterpret_cast<const float4*>(carry_cur + pos) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                    for (int a = 0; a < 4; ++a) {
                        const int f = u - a;
                        if (f >= 0) {
                            const float4 v = *reinterpret_cast<const float4*>(s_x + f * kFS + 2 * (f & 1) + a * kHop + r);
                            acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
                        }
                    }
                    const float4 e = __ldg(reinterpret_cast<const float4*>(ienv + r));
                    acc.x *= e.x; acc.y *= e.y; acc.z *= e.z; acc.w *= e.w;
                    if (in_range) __stcs(reinterpret_cast<float4*>(o + pos), acc);
                }
                // tail carried into the next tile: three more hop slots
                for (int pos = 4 * tid; pos < kNC; pos += 4 * NT) {
                    const int u = 16 + (pos >> 9), r = pos & 511;
                    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                    for (int a = 1; a < 4; ++a) {
                        const int f = u - a;
                        if (f <= 15) {
                            const float4 v = *reinterpret_cast<const float4*>(s_x + f * kFS + 2 * (f & 1) + a * kHop + r);
                            acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
                        }
                    }
                    *reinterpret_cast<float4*>(carry_nxt + pos) = acc;
                }
            }
            __syncthreads();
            PHASE_END(2)
            float* tmp = carry_cur; carry_cur = carry_nxt; carry_nxt = tmp;
        }
    }
    if (tid == 0 && phase_cycles)
        for (int i = 0; i < 3; ++i) phase_cycles[blockIdx.x * 4 + i] = cyc[i];
}

template <int Q, bool WIDE = false>
static void run(const float* spec, float* out, const float4* tw4, const float* win, const float* ienv, const char* name) {
    long long* d_cyc;
    cudaMalloc(&d_cyc, 148 * 4 * sizeof(long long));
    cudaMemset(d_cyc, 0, 148 * 4 * sizeof(long long));
    constexpr int NT = kF * kRB * (32 / Q);
    const size_t smem = sizeof(float) * (kN + 4 * 32 * 17 + kF * kFS + 4 + 2 * kNC);
    cudaFuncSetAttribute(mock_k2<Q, WIDE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int i = 0; i < 3; ++i) mock_k2<Q, WIDE><<<148, NT, smem>>>(spec, out, tw4, win, ienv, d_cyc);
    cudaEventRecord(e0);
    for (int i = 0; i < 10; ++i) mock_k2<Q, WIDE><<<148, NT, smem>>>(spec, out, tw4, win, ienv, d_cyc);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    ms /= 10;
    const double bytes = 4.0 * 3 * kRows * kT * kClips + 4.0 * kOutLen * kClips;
    cudaFuncAttributes fa;
    cudaFuncGetAttributes(&fa, mock_k2<Q, WIDE>);
    std::printf("%s: %d threads, %d registers, %zu B smem, %.3f ms, %.0f GB/s  (%s)\n", name, NT, fa.numRegs, smem, ms,
                bytes / ms * 1e-6, cudaGetErrorString(cudaGetLastError()));
    long long h_cyc[148 * 4];
    cudaMemcpy(h_cyc, d_cyc, sizeof(h_cyc), cudaMemcpyDeviceToHost);
    double sum[3] = {0, 0, 0};
    int n_full = 0;
    for (int b = 0; b < 148; ++b)
        if (b < kClips - 148) { for (int i = 0; i < 3; ++i) sum[i] += (double)h_cyc[b * 4 + i]; ++n_full; }   // CTAs that ran two clips
    const double tot = sum[0] + sum[1] + sum[2];
    std::printf("    phase share (thread 0 of the two-clip CTAs, barrier to barrier): pass A %.1f %%, pass B %.1f %%, overlap-add %.1f %%;"
                " cycles per tile %.0f\n", 100 * sum[0] / tot, 100 * sum[1] / tot, 100 * sum[2] / tot, tot / n_full / (2.0 * 55));
    cudaFree(d_cyc);
}

int main() {
    float *spec, *out, *win, *ienv;
    float4* tw4;
    const size_t n = (size_t)3 * kRows * kT * kClips;
    cudaMalloc(&spec, n * sizeof(float));
    cudaMalloc(&out, (size_t)kOutLen * kClips * sizeof(float));
    cudaMalloc(&win, kN * sizeof(float));
    cudaMalloc(&ienv, kHop * sizeof(float));
    cudaMalloc(&tw4, 32 * 17 * sizeof(float4));
    cudaMemset(spec, 0x3c, n * sizeof(float));
    cudaMemset(win, 0x3c, kN * sizeof(float));
    cudaMemset(ienv, 0x3c, kHop * sizeof(float));
    cudaMemset(tw4, 0x3c, 32 * 17 * sizeof(float4));
    run<32>(spec, out, tw4, win, ienv, "Q=32 (16 warps/SM)");
    run<16>(spec, out, tw4, win, ienv, "Q=16 (32 warps/SM)");
    run<32>(spec, out, tw4, win, ienv, "Q=32 (16 warps/SM)");
    run<16>(spec, out, tw4, win, ienv, "Q=16 (32 warps/SM)");
    run<32, true>(spec, out, tw4, win, ienv, "Q=32, pass A reads 32-frame row segments (cluster-pair plan)");
    run<32, true>(spec, out, tw4, win, ienv, "Q=32, pass A reads 32-frame row segments (cluster-pair plan)");
    return 0;
}
