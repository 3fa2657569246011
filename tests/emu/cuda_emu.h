// cuda_emu.h -- TEST INFRASTRUCTURE ONLY (never linked into the product library).
//
// A tiny CUDA execution-model emulator used to run the real kernel sources of
// audio_intelligence_b200/csrc on the CPU (this build container has no GPU).
// Every CUDA thread of a block is an OS thread; __syncthreads() is a
// std::barrier over the block; warp shuffles go through a per-warp mailbox.
// Blocks of a grid run one after another.  Device pointers are host pointers.
//
// Only what the kernels need is modelled: threadIdx/blockIdx, dynamic shared
// memory, __syncthreads/__syncwarp, __shfl_sync/__shfl_xor_sync, a few math
// intrinsics, and the cudaMalloc/cudaMemcpy family.
#pragma once
#include <atomic>
#include <barrier>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <condition_variable>
#include <functional>
#include <map>
#include <memory>
#include <mutex>
#include <type_traits>
#include <thread>
#include <vector>

#define __global__
#define __device__
#define __host__
#define __shared__ static   /* blocks run one after another, so block-shared == static */
#define __forceinline__ inline
#define __restrict__
#define __launch_bounds__(...)
#define __align__(n) alignas(n)

struct uint3 { unsigned x, y, z; };
struct dim3 {
    unsigned x, y, z;
    dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};
struct float2 { float x, y; };
struct alignas(16) float4 { float x, y, z, w; };
struct alignas(8) int2 { int x, y; };
static inline int2 make_int2(int x, int y) { return int2{x, y}; }
static inline float2 make_float2(float x, float y) { return float2{x, y}; }
static inline float4 make_float4(float x, float y, float z, float w) { return float4{x, y, z, w}; }

namespace emu {
struct WarpBox {
    std::barrier<> bar;
    uint32_t slot[32];
    explicit WarpBox(int n) : bar(n) {}
};
// mbarrier model: arrival count + transaction bytes + phase bit, kept in a side table keyed by the barrier's address
struct MBar {
    int init = 0, pending = 0;
    long long tx = 0;
    unsigned phase = 0;
};
struct BlockCtx {
    std::mutex mbar_mu;
    std::condition_variable mbar_cv;
    std::map<const void*, MBar> mbars;
    std::mutex named_mu;
    std::unique_ptr<std::barrier<>> named[16];
    std::unique_ptr<std::barrier<>> bar;
    std::vector<std::unique_ptr<WarpBox>> warps;
    unsigned char* smem = nullptr;
    std::atomic<int> or_flag{0};
};
inline thread_local BlockCtx* g_ctx = nullptr;
inline thread_local uint3 g_tid{0, 0, 0};
inline thread_local uint3 g_bid{0, 0, 0};
inline thread_local dim3 g_bdim;
inline thread_local dim3 g_gdim;

// one block of a (possibly multi-dimensional) grid
template <class F>
void launch_at(dim3 bid, dim3 grid, dim3 block, size_t smem_bytes, F&& body) {
    const unsigned nthreads = block.x;
    BlockCtx ctx;
    ctx.bar = std::make_unique<std::barrier<>>(nthreads);
    const unsigned nwarps = (nthreads + 31) / 32;
    for (unsigned w = 0; w < nwarps; ++w) {
        unsigned n = (w + 1 == nwarps) ? nthreads - 32 * w : 32;
        ctx.warps.emplace_back(std::make_unique<WarpBox>((int)n));
    }
    void* p = nullptr;
    if (posix_memalign(&p, 1024, smem_bytes ? smem_bytes : 16) != 0) abort();
    std::memset(p, 0xCD, smem_bytes);  // poison: catches reads of unwritten smem
    ctx.smem = (unsigned char*)p;
    std::vector<std::thread> ts;
    ts.reserve(nthreads);
    for (unsigned t = 0; t < nthreads; ++t) {
        ts.emplace_back([&, t] {
            g_ctx = &ctx;
            g_tid = uint3{t, 0, 0};
            g_bid = uint3{bid.x, bid.y, bid.z};
            g_bdim = block;
            g_gdim = grid;
            body();
        });
    }
    for (auto& th : ts) th.join();
    free(p);
}
template <class F>
void launch(dim3 grid, dim3 block, size_t smem_bytes, F&& body) {
    for (unsigned b = 0; b < grid.x; ++b) launch_at(dim3(b, 0, 0), grid, block, smem_bytes, body);
}
// bar.sync id, count: a barrier over `count` threads, created on first use
inline void named_barrier(int id, int count) {
    BlockCtx& c = *g_ctx;
    {
        std::lock_guard<std::mutex> lk(c.named_mu);
        if (!c.named[id]) c.named[id] = std::make_unique<std::barrier<>>(count);
    }
    c.named[id]->arrive_and_wait();
}
inline void mbar_check(BlockCtx& c, MBar& b) {
    if (b.pending == 0 && b.tx == 0) {
        b.phase ^= 1u;
        b.pending = b.init;
        c.mbar_cv.notify_all();
    }
}
inline void mbar_init(const void* a, int count) {
    BlockCtx& c = *g_ctx;
    std::lock_guard<std::mutex> lk(c.mbar_mu);
    MBar& b = c.mbars[a];
    b.init = b.pending = count; b.tx = 0; b.phase = 0;
}
inline void mbar_arrive(const void* a, long long expect_tx) {
    BlockCtx& c = *g_ctx;
    std::lock_guard<std::mutex> lk(c.mbar_mu);
    MBar& b = c.mbars[a];
    b.tx += expect_tx;
    b.pending -= 1;
    mbar_check(c, b);
}
inline void mbar_complete_tx(const void* a, long long bytes) {
    BlockCtx& c = *g_ctx;
    std::lock_guard<std::mutex> lk(c.mbar_mu);
    MBar& b = c.mbars[a];
    b.tx -= bytes;
    mbar_check(c, b);
}
// mbarrier.try_wait.parity: returns once the phase with the given parity has completed
inline void mbar_wait(const void* a, unsigned parity) {
    BlockCtx& c = *g_ctx;
    std::unique_lock<std::mutex> lk(c.mbar_mu);
    MBar& b = c.mbars[a];
    c.mbar_cv.wait(lk, [&] { return (b.phase & 1u) != (parity & 1u); });
}
inline uint32_t shfl_raw(uint32_t v, int src_lane) {
    WarpBox& wb = *g_ctx->warps[g_tid.x / 32];
    wb.slot[g_tid.x % 32] = v;
    wb.bar.arrive_and_wait();
    uint32_t r = wb.slot[src_lane & 31];
    wb.bar.arrive_and_wait();
    return r;
}
inline unsigned ballot(bool pred) {
    unsigned acc = 0;
    for (int l = 0; l < 32; ++l) acc |= (shfl_raw(pred ? 1u : 0u, l) & 1u) << l;
    return acc;
}
}  // namespace emu

#define threadIdx (emu::g_tid)
#define blockIdx (emu::g_bid)
#define blockDim (emu::g_bdim)
#define gridDim (emu::g_gdim)

#ifdef A2SB_EMU_NO_SYNC   // self-test of the ThreadSanitizer driver: without block barriers it must report races
static inline void __syncthreads() {}
#else
static inline void __syncthreads() { emu::g_ctx->bar->arrive_and_wait(); }
#endif
// barrier + block-wide OR of the predicate (three barrier phases: publish, read, reset)
static inline int __syncthreads_or(int pred) {
    std::atomic<int>& f = emu::g_ctx->or_flag;
    if (pred) f.store(1);
    __syncthreads();
    const int r = f.load();
    __syncthreads();
    if (emu::g_tid.x == 0) f.store(0);
    __syncthreads();
    return r;
}
static inline void __syncwarp(unsigned = 0xffffffffu) {
    emu::g_ctx->warps[emu::g_tid.x / 32]->bar.arrive_and_wait();
}
static inline float __shfl_sync(unsigned, float v, int src) {
    uint32_t u; std::memcpy(&u, &v, 4);
    u = emu::shfl_raw(u, src);
    float r; std::memcpy(&r, &u, 4); return r;
}
static inline float __shfl_xor_sync(unsigned, float v, int mask) {
    uint32_t u; std::memcpy(&u, &v, 4);
    u = emu::shfl_raw(u, (int)(emu::g_tid.x % 32) ^ mask);
    float r; std::memcpy(&r, &u, 4); return r;
}
static inline int __shfl_sync(unsigned, int v, int src) { return (int)emu::shfl_raw((uint32_t)v, src); }

static inline int __any_sync(unsigned, int pred) {
    uint32_t acc = 0;
    for (int l = 0; l < 32; ++l) acc |= emu::shfl_raw(pred ? 1u : 0u, l);
    return acc != 0;
}
static inline unsigned __float_as_uint(float v) { unsigned u; std::memcpy(&u, &v, 4); return u; }
template <class T> static inline T __ldg(const T* p) { return *p; }
static inline int atomicAdd(int* p, int v) {
    return reinterpret_cast<std::atomic<int>*>(p)->fetch_add(v);
}
static inline unsigned atomicAdd(unsigned* p, unsigned v) {
    return reinterpret_cast<std::atomic<unsigned>*>(p)->fetch_add(v);
}

// ---- runtime shim -----------------------------------------------------------
typedef int cudaError_t;
typedef void* cudaStream_t;
enum { cudaSuccess = 0, cudaErrorInvalidValue = 1, cudaErrorMemoryAllocation = 2 };
enum cudaMemcpyKind { cudaMemcpyHostToDevice, cudaMemcpyDeviceToHost, cudaMemcpyDeviceToDevice, cudaMemcpyDefault };
static inline cudaError_t cudaMalloc(void** p, size_t n) { *p = malloc(n ? n : 1); return *p ? 0 : 2; }
static inline cudaError_t cudaFree(void* p) { free(p); return 0; }
static inline cudaError_t cudaMallocHost(void** p, size_t n) { *p = malloc(n ? n : 1); return *p ? 0 : 2; }
static inline cudaError_t cudaFreeHost(void* p) { free(p); return 0; }
static inline cudaError_t cudaMemcpy(void* d, const void* s, size_t n, cudaMemcpyKind) { std::memcpy(d, s, n); return 0; }
static inline cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, cudaMemcpyKind, cudaStream_t) { std::memcpy(d, s, n); return 0; }
static inline cudaError_t cudaMemsetAsync(void* d, int v, size_t n, cudaStream_t) { std::memset(d, v, n); return 0; }
static inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return 0; }
static inline cudaError_t cudaDeviceSynchronize() { return 0; }
static inline cudaError_t cudaGetLastError() { return 0; }
static inline cudaError_t cudaPeekAtLastError() { return 0; }
static inline const char* cudaGetErrorString(cudaError_t) { return "emu"; }
static inline cudaError_t cudaGetDevice(int* d) { *d = 0; return 0; }
static inline cudaError_t cudaSetDevice(int) { return 0; }
