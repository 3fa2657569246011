// Microbenchmark for the round-2 K2 restructuring: how fast can the inverse kernel's spectrogram reads be STREAMED into
// shared memory by tensor-map TMA (cp.async.bulk.tensor) instead of register loads?  The spectrogram is the reference
// layout [clip][3][1024][T] fp32, T = 862: rows are only 8-byte aligned, which a tensor map cannot express (strides must
// be multiples of 16 bytes).  Work-around measured here: FOUR maps, one per (row mod 4), whose row stride is 4 rows
// (16*T bytes) and whose base is moved back to the previous 16-byte boundary (the frame coordinate is shifted by the
// same 0..3 elements).  One request fetches the box [F frames] x [32 rows of one residue class, 32 rows apart] x
// [3 planes] = what one half-warp of pass A consumes for a tile (F = 16: 6 KB, 96 row segments of 64 bytes).
//
// A dedicated producer thread keeps a ring of SLOTS boxes in flight; 16 consumer warps wait for "their" two boxes of each
// tile, read them like pass A would (96 LDS per lane, lanes along frames) or not at all, and release the slots.
// `spin` adds a busy phase per tile with all consumers idle on the ring (models pass B + overlap-add).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tma_box_stream tma_box_stream.cu
// Run:   ./tma_box_stream [F=16] [SLOTS=10] [mode: 0 wait only, 1 LDS consume] [spin cycles] [l2promo 0..3] [T=862]
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <cuda.h>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e_)); exit(1); } } while (0)

constexpr int kRows = 1024, kClips = 256, kCons = 512;

struct Maps { CUtensorMap m[4]; };

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* b, int n) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(n) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* b, unsigned parity) {
    asm volatile(
        "{\n.reg .pred p;\nW1:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra D1;\nbra W1;\nD1:\n}\n" ::"r"(smem_u32(b)),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned long long* b) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(b)) : "memory");
}
__device__ __forceinline__ void mbar_expect(unsigned long long* b, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_load5(void* dst, const CUtensorMap* map, unsigned long long* bar, int c0, int c1, int c2, int c3,
                                          int c4) {
    asm volatile(
        "cp.async.bulk.tensor.5d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];" ::"r"(
            smem_u32(dst)),
        "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
        : "memory");
}

// Ring of `slots` boxes [3][32][F+4] floats; one "full" mbarrier per box position of a tile (32), used once per tile, so
// a waiter is never more than one phase away (tiles are separated by a CTA barrier).  No producer warp: the warp that
// has consumed box n issues box n + slots into the slot it has just freed (32 % slots == 0, so the barrier of the
// target box position has completed its previous phase: it lies on the same issue chain).
template <int F>
__global__ void __launch_bounds__(F * 32, (F == 8) ? 2 : 1) stream(const __grid_constant__ Maps maps, float* sink, int n_tiles, int T,
                                                                   int slots, int mode, int spin, int toff) {
    extern __shared__ __align__(1024) unsigned char smem[];
    constexpr int FW = F + 4;          // box width: the start coordinate must be a multiple of 4 elements (16 bytes)
    constexpr int BOX = 3 * 32 * FW;  // floats
    constexpr int BPW = (F >= 16) ? 2 : 32 / F;   // boxes per warp (F = 8: four 8-lane groups)
    constexpr int LPB = (F >= 32) ? 32 : F;       // lanes per box
    unsigned long long* full = reinterpret_cast<unsigned long long*>(smem);
    float* ring = reinterpret_cast<float*>(smem + 1024);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid == 0) {
        for (int s = 0; s < 32; ++s) mbar_init(full + s, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const int tiles_per_clip = (T + F - 1) / F;
    auto issue = [&](int tile, int box, int slot) {
        const int b = tile / tiles_per_clip, t0 = (tile % tiles_per_clip) * F + toff;
        const int w = box >> 1, rho = (box & 1) ? 31 - w : w;
        const int r = rho & 3, jr = rho >> 2;
        const int shift = (int)(((long long)r * T * 4) & 15) / 4;
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        mbar_expect(full + box, BOX * 4);
        tma_load5(ring + (size_t)slot * BOX, &maps.m[r], full + box, (t0 + shift) & ~3, jr, 0, 0, b);
    };
    if (tid == 0)
        for (int box = 0; box < slots; ++box) issue(blockIdx.x, box, box);
    const int t = lane % LPB, h = lane / LPB;
    float acc = 0.f;
    unsigned tilecount = 0;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++tilecount) {
        const int t0c = (tile % tiles_per_clip) * F + toff;
        const unsigned par = tilecount & 1u;
        const int ntile = tile + gridDim.x;
        constexpr int ROUNDS = (F >= 32) ? 2 : 1;          // F = 32: the warp's two boxes one after the other
        constexpr int PER = BPW / ROUNDS;                   // boxes consumed at once
#pragma unroll
        for (int hh = 0; hh < ROUNDS; ++hh) {
            const int box0 = BPW * warp + hh * PER;
#pragma unroll
            for (int e = 0; e < PER; ++e) mbar_wait(full + box0 + e, par);
            if (mode == 1) {
                const int box = box0 + ((F >= 32) ? 0 : h);
                const int rho = (box & 1) ? 31 - (box >> 1) : (box >> 1);
                const int off = (t0c + (int)((((long long)(rho & 3) * T * 4) & 15) >> 2)) & 3;
                const float* src = ring + (size_t)(box % slots) * BOX + t + off;
#pragma unroll
                for (int q = 0; q < 32; ++q) {
                    const int qq = (box & 1) ? 31 - q : q;
                    acc += src[(0 * 32 + qq) * FW] * src[(1 * 32 + qq) * FW] + src[(2 * 32 + qq) * FW];
                }
            }
            __syncwarp();
            if (lane == 0) {
                for (int e = 0; e < PER; ++e) {
                    const int nb = box0 + e + slots;     // 32 % slots == 0: same slot
                    if (nb < 32) issue(tile, nb, nb % slots);
                    else if (ntile < n_tiles) issue(ntile, nb - 32, nb % slots);
                }
            }
        }
        if (spin > 0) {
            __syncthreads();
            const long long c0 = clock64();
            while (clock64() - c0 < spin) {}
        }
        __syncthreads();
    }
    if (acc == 123.456f) sink[tid] = acc;
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                             const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main(int argc, char** argv) {
    const int F = argc > 1 ? atoi(argv[1]) : 16;
    const int slots = argc > 2 ? atoi(argv[2]) : 8;   // must divide 32
    const int mode = argc > 3 ? atoi(argv[3]) : 1;
    const int spin = argc > 4 ? atoi(argv[4]) : 0;
    const int promo = argc > 5 ? atoi(argv[5]) : 0;
    const int T = argc > 6 ? atoi(argv[6]) : 862;
    const int toff = argc > 7 ? atoi(argv[7]) : 0;   // frame offset of the tile grid (K2's items start at 29 k - 3)
    const size_t elems = (size_t)kClips * 3 * kRows * T;
    float* spec;
    CK(cudaMalloc(&spec, elems * 4 + 64));
    CK(cudaMemset(spec, 0, elems * 4 + 64));
    float* sink;
    CK(cudaMalloc(&sink, 4096));
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
    EncodeFn encode = (EncodeFn)fn;
    Maps maps;
    for (int r = 0; r < 4; ++r) {
        const long long a = ((long long)r * T * 4) & 15;
        char* base = (char*)spec + (long long)r * T * 4 - a;
        cuuint64_t dims[5] = {(cuuint64_t)(T + a / 4), 8, 32, 3, (cuuint64_t)kClips};
        cuuint64_t strides[4] = {(cuuint64_t)4 * T * 4, (cuuint64_t)32 * T * 4, (cuuint64_t)kRows * T * 4, (cuuint64_t)3 * kRows * T * 4};
        cuuint32_t box[5] = {(cuuint32_t)F + 4, 1, 32, 3, 1};
        cuuint32_t es[5] = {1, 1, 1, 1, 1};
        CUresult rc = encode(&maps.m[r], CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 5, base, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                             CU_TENSOR_MAP_SWIZZLE_NONE, (CUtensorMapL2promotion)promo, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (rc != CUDA_SUCCESS) { printf("encode failed: %d (r=%d)\n", (int)rc, r); return 1; }
    }
    const int tiles_per_clip = (T + F - 1) / F;
    const int n_tiles = kClips * tiles_per_clip;
    const size_t smem = 1024 + (size_t)slots * 3 * 32 * (F + 4) * 4;
    auto run = [&](auto kern) {
        CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        cudaEvent_t e0, e1;
        CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
        float best = 1e9f, sum = 0.f;
        const int reps = 12;
        for (int i = 0; i < reps + 3; ++i) {
            CK(cudaEventRecord(e0));
            kern<<<(F == 8) ? 296 : 148, F >= 32 ? 512 : F * 32, smem>>>(maps, sink, n_tiles, T, slots, mode, spin, toff);
            CK(cudaEventRecord(e1));
            CK(cudaEventSynchronize(e1));
            float ms;
            CK(cudaEventElapsedTime(&ms, e0, e1));
            if (i >= 3) { best = ms < best ? ms : best; sum += ms; }
        }
        CK(cudaGetLastError());
        const double gb = (double)elems * 4 / 1e9;
        printf("F=%d slots=%d (%.0f KB) mode=%d spin=%d promo=%d T=%d toff=%d: best %.3f ms (%.0f GB/s)  mean %.3f ms (%.0f GB/s)\n", F, slots,
               smem / 1024.0, mode, spin, promo, T, toff, best, gb / best * 1e3, sum / reps, gb / (sum / reps) * 1e3);
    };
    if (F == 8) run(stream<8>);
    else if (F == 16) run(stream<16>);
    else run(stream<32>);
    return 0;
}
