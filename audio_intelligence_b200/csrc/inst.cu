// inst.cu -- explicit kernel families per n_fft.  Compiled once per family with -DA2SB_INST=k so
// the heavy template instantiations build in parallel; -DA2SB_INST_ALL builds every family in one
// translation unit (used by the CPU emulation build in tests/emu).
#include <cstdlib>

#include "host_util.h"
#include "istft_inv.cuh"
#include "stft_fwd.cuh"

namespace a2sb {

// Forward launch: picks the tile width F (frames per tile; 16/F groups per CTA), the run length
// (consecutive tiles per work item) and the fast / careful kernel variant.
template <int M, int RA, int RB, int F>
static int launch_fwd(const LaunchCtx& cx, FwdParams p, cudaStream_t st) {
    using G = FwdGeom<M, RA, RB, F>;
    const size_t smem = G::smem_bytes(cx.hop);
    const long long T = p.t_end - p.t_begin;
    p.tiles_per_clip = (int)((T + F - 1) / F);
    const long long groups = (long long)cx.sm_count * G::GROUPS;
    // Run length 1: the groups of the grid work on consecutive tiles of a clip at the same time, so the
    // partial sectors at tile seams meet in L2 within microseconds (measured: longer runs are slower).
    int run_best = 1;
    static const int env_run = [] { const char* e = std::getenv("A2SB_FWD_RUN"); return e ? std::atoi(e) : 0; }();
    if (env_run >= 1 && env_run <= 64) run_best = env_run;   // experiments
    p.run = run_best;
    p.items_per_clip = (p.tiles_per_clip + run_best - 1) / run_best;
    p.total_items = (long long)p.items_per_clip * p.batch;
    const long long ctas = (p.total_items + G::GROUPS - 1) / G::GROUPS;
    if (p.epi == kEpiMagPhase && p.pmode == kPowQuarter)
        return launch_persistent(stft_fwd_kernel<M, RA, RB, F, 1>, ctas, G::NT, smem, st, p, cx.sm_count);
    if (p.epi == kEpiMagPhase && p.pmode == kPowNone)
        return launch_persistent(stft_fwd_kernel<M, RA, RB, F, 2>, ctas, G::NT, smem, st, p, cx.sm_count);
    return launch_persistent(stft_fwd_kernel<M, RA, RB, F, 0>, ctas, G::NT, smem, st, p, cx.sm_count);
}

template <int M, int RA, int RB>
static int dispatch_fwd(const LaunchCtx& cx, const FwdParams& p, cudaStream_t st) {
    if constexpr (M >= 2048) return launch_fwd<M, RA, RB, 8>(cx, p, st);   // 16-frame exchange does not fit 227 KB
    else return cx.fwd_tile == 8 ? launch_fwd<M, RA, RB, 8>(cx, p, st) : launch_fwd<M, RA, RB, 16>(cx, p, st);
}

template <int M, int RA, int RB, int F>
static int launch_inv(const LaunchCtx& cx, const InvParams& p, cudaStream_t st) {
    using G = InvGeom<M, RA, RB, F>;
    const size_t smem = G::smem_bytes(cx.hop);
    const bool fast = p.in_kind == kInMagPhase && !p.has_dc && p.svd_fix && p.pmode == kPowFour;
    return fast ? launch_persistent(istft_inv_kernel<M, RA, RB, F, 1>, p.total_items, G::NT, smem, st, p, cx.sm_count)
                : launch_persistent(istft_inv_kernel<M, RA, RB, F, 0>, p.total_items, G::NT, smem, st, p, cx.sm_count);
}

template <int M, int RA, int RB>
static int dispatch_inv(const LaunchCtx& cx, const InvParams& p, cudaStream_t st) {
    // cx.inv_tile == inv_tile_frames(M) unless overridden; RA % 32 == 0 is what the 8-frame layout needs
    if constexpr (M >= 2048) return launch_inv<M, RA, RB, 8>(cx, p, st);   // 16-frame exchange does not fit 227 KB
    else if constexpr (RA % 32 == 0) return cx.inv_tile == 8 ? launch_inv<M, RA, RB, 8>(cx, p, st) : launch_inv<M, RA, RB, 16>(cx, p, st);
    else return launch_inv<M, RA, RB, 16>(cx, p, st);
}

#if defined(A2SB_INST_ALL) || A2SB_INST == 1
int run_fwd_256(const LaunchCtx& c, const FwdParams& p, cudaStream_t s) { return dispatch_fwd<256, 16, 16>(c, p, s); }
#endif
#if defined(A2SB_INST_ALL) || A2SB_INST == 2
int run_fwd_512(const LaunchCtx& c, const FwdParams& p, cudaStream_t s) { return dispatch_fwd<512, 16, 32>(c, p, s); }
#endif
#if defined(A2SB_INST_ALL) || A2SB_INST == 3
int run_fwd_1024(const LaunchCtx& c, const FwdParams& p, cudaStream_t s) { return dispatch_fwd<1024, 32, 32>(c, p, s); }
#endif
#if defined(A2SB_INST_ALL) || A2SB_INST == 7
int run_fwd_2048(const LaunchCtx& c, const FwdParams& p, cudaStream_t s) { return dispatch_fwd<2048, 32, 64>(c, p, s); }
#endif
#if defined(A2SB_INST_ALL) || A2SB_INST == 4
int run_inv_256(const LaunchCtx& c, const InvParams& p, cudaStream_t s) { return dispatch_inv<256, 16, 16>(c, p, s); }
#endif
#if defined(A2SB_INST_ALL) || A2SB_INST == 5
int run_inv_512(const LaunchCtx& c, const InvParams& p, cudaStream_t s) { return dispatch_inv<512, 32, 16>(c, p, s); }
#endif
#if defined(A2SB_INST_ALL) || A2SB_INST == 6
int run_inv_1024(const LaunchCtx& c, const InvParams& p, cudaStream_t s) { return dispatch_inv<1024, 32, 32>(c, p, s); }
#endif

#if defined(A2SB_INST_ALL) || A2SB_INST == 8
int run_inv_2048(const LaunchCtx& c, const InvParams& p, cudaStream_t s) { return dispatch_inv<2048, 64, 32>(c, p, s); }
#endif

}  // namespace a2sb
