// host_util.h -- host-side helpers shared by the translation units of liba2sb_b200.so:
// error reporting, launch counting and the two launch shapes (persistent, grid-stride).
#pragma once
#include <atomic>
#include <cstdint>
#include <cstdlib>

#include "../../include/a2sb_b200.h"
#include "a2sb_common.cuh"

namespace a2sb {

int fail(int code, const char* fmt, ...);      // sets a2sb_last_error(), returns code
extern std::atomic<long long> g_launches;      // kernels launched by this library
extern std::atomic<long long> g_tma_launches;  // ... of which inverse-kernel launches of the TMA variant

#define A2SB_CUDA(expr)                                                                               \
    do {                                                                                              \
        cudaError_t e_ = (expr);                                                                      \
        if (e_ != cudaSuccess) return ::a2sb::fail(A2SB_ERR_CUDA, "%s: %s", #expr, cudaGetErrorString(e_)); \
    } while (0)

struct LaunchCtx {
    int sm_count;
    int hop;
    int fwd_tile;  // frames per forward tile (8 or 16)
    int inv_tile;  // frames per inverse tile (8 or 16)
};

// (kernel, dynamic shared memory) -> resident CTAs per SM, filled on first launch (a2sb_api.cu); -1 = unknown.
// Keyed per device: the attributes are per device context.
int launch_cache_lookup(const void* kern, size_t smem);
void launch_cache_store(const void* kern, size_t smem, int per_sm);
// Largest dynamic-shared-memory attribute set so far for (kernel, current device).  The attribute is per kernel, not per
// launch, and two plans of one n_fft with different hops share a kernel with different footprints: it is only ever RAISED
// (lowering it for a smaller plan would make the next launch of the larger, already cached plan fail).
size_t launch_smem_attr_get(const void* kern);
void launch_smem_attr_set(const void* kern, size_t smem);

// Thread-block cluster size of the NEXT launch_persistent_n on this thread (0 / 1 = no cluster); consumed by that launch.
// Experiment hook of the forward kernel (A2SB_FWD_CLUSTER: neighbouring CTAs kept in lockstep by a cluster barrier per round).
inline thread_local int tl_next_cluster = 0;

// Launch `kern(args...)` with a persistent grid: min(work, SMs * resident CTAs per SM).
template <class... KArgs, class... Args>
int launch_persistent_n(void (*kern)(KArgs...), long long work, int block, size_t smem, cudaStream_t st, int sm_count,
                        const Args&... args) {
    if (work <= 0) return A2SB_OK;
#ifdef A2SB_EMU
    const long long grid = work < sm_count ? work : sm_count;
    emu::launch(dim3((unsigned)grid), dim3((unsigned)block), smem, [&] { kern(args...); });
    (void)st;
#else
    // Function attributes and occupancy are queried once per (kernel, shared-memory size): on a single 10 s clip the two
    // runtime calls cost as much as the kernel itself.
    int per_sm = launch_cache_lookup(reinterpret_cast<const void*>(kern), smem);
    if (per_sm < 0) {
        if (smem > 48 * 1024 && smem > launch_smem_attr_get(reinterpret_cast<const void*>(kern))) {
            A2SB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            launch_smem_attr_set(reinterpret_cast<const void*>(kern), smem);
        }
        // Experiment hook: shared-memory carve-out in percent (L1 gets the rest).  By default the driver picks the smallest
        // carve-out that holds the kernel's shared memory.
        if (const char* e = std::getenv("A2SB_CARVEOUT")) {
            A2SB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, std::atoi(e)));
        }
        per_sm = 0;
        A2SB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, block, smem));
        if (per_sm < 1) return fail(A2SB_ERR_CUDA, "kernel does not fit on an SM (block %d, smem %zu)", block, smem);
        launch_cache_store(reinterpret_cast<const void*>(kern), smem, per_sm);
    }
    long long grid = (long long)sm_count * per_sm;
    if (grid > work) grid = work;
    const int cluster = tl_next_cluster;
    tl_next_cluster = 0;
    if (cluster > 1 && grid >= cluster) {
        grid -= grid % cluster;
        cudaLaunchConfig_t cfg{};
        cfg.gridDim = dim3((unsigned)grid); cfg.blockDim = dim3((unsigned)block); cfg.dynamicSmemBytes = smem; cfg.stream = st;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension;
        at[0].val.clusterDim.x = (unsigned)cluster; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        cfg.attrs = at; cfg.numAttrs = 1;
        A2SB_CUDA(cudaLaunchKernelEx(&cfg, kern, args...));
    } else {
        kern<<<(unsigned)grid, block, smem, st>>>(args...);
    }
    A2SB_CUDA(cudaGetLastError());
#endif
    g_launches.fetch_add(1);
    return A2SB_OK;
}
template <class P>
int launch_persistent(void (*kern)(const P), long long work, int block, size_t smem, cudaStream_t st, const P& p,
                      int sm_count) {
    return launch_persistent_n(kern, work, block, smem, st, sm_count, p);
}

template <class P>
int launch_grid_stride(void (*kern)(const P), long long total, cudaStream_t st, const P& p, int sm_count) {
    if (total <= 0) return A2SB_OK;
    const int block = 256;
    long long blocks = (total + block - 1) / block;
#ifdef A2SB_EMU
    if (blocks > 2) blocks = 2;
    emu::launch(dim3((unsigned)blocks), dim3(block), 0, [&] { kern(p); });
    (void)st; (void)sm_count;
#else
    const long long cap = (long long)sm_count * 8;  // 8 x 256 threads = full occupancy
    if (blocks > cap) blocks = cap;
    kern<<<(unsigned)blocks, block, 0, st>>>(p);
    A2SB_CUDA(cudaGetLastError());
#endif
    g_launches.fetch_add(1);
    return A2SB_OK;
}

// (RA, RB) of the forward two-pass decomposition M = RA * RB
inline void fwd_radices(int M, int& RA, int& RB) {
    RA = (M >= 1024) ? 32 : 16;
    RB = M / RA;
}

// (RA, RB) of the inverse decomposition (pass A radix RA over the bins, pass B radix RB)
inline void inv_radices(int M, int& RA, int& RB) {
    RA = (M == 256) ? 16 : 32;   // n_fft 4096: 32 x 64, the radix-64 pass B split over lane pairs (istft_inv.cuh)
    RB = M / RA;
}
// frames per inverse tile (the 16-frame exchange of M = 2048 does not fit 227 KB of shared memory)
inline int inv_tile_frames(int M) { return (M >= 2048) ? 8 : 16; }

struct FwdParams;
struct InvParams;
// Per-n_fft kernel families, each compiled in its own translation unit (inst.cu, -DA2SB_INST=k).
int run_fwd_256(const LaunchCtx&, const FwdParams&, cudaStream_t);
int run_fwd_512(const LaunchCtx&, const FwdParams&, cudaStream_t);
int run_fwd_1024(const LaunchCtx&, const FwdParams&, cudaStream_t);
int run_fwd_2048(const LaunchCtx&, const FwdParams&, cudaStream_t);
int run_inv_256(const LaunchCtx&, const InvParams&, cudaStream_t);
int run_inv_512(const LaunchCtx&, const InvParams&, cudaStream_t);
int run_inv_1024(const LaunchCtx&, const InvParams&, cudaStream_t);
int run_inv_2048(const LaunchCtx&, const InvParams&, cudaStream_t);

}  // namespace a2sb
