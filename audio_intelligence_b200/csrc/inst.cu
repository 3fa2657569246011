// inst.cu -- explicit kernel families per n_fft.  Compiled once per family with -DA2SB_INST=k so
// the heavy template instantiations build in parallel; -DA2SB_INST_ALL builds every family in one
// translation unit (used by the CPU emulation build in tests/emu).
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "host_util.h"
#include "istft_inv.cuh"
#include "stft_fwd.cuh"

namespace a2sb {

// Forward launch: picks the tile width F (frames per tile; 16/F groups per CTA), the run length
// (consecutive tiles per work item) and the fast / careful kernel variant.
template <int M, int RA, int RB, int F, int ROUNDS = 1, int WIDE = 0>
static int launch_fwd(const LaunchCtx& cx, FwdParams p, cudaStream_t st) {
    using G = FwdGeom<M, RA, RB, F, ROUNDS, WIDE>;
    const size_t smem = G::smem_bytes(cx.hop);
    if (smem > 232448)
        return fail(A2SB_ERR_INVALID, "hop_length=%d: a tile's input span does not fit the forward kernel's shared memory (%zu bytes)",
                    cx.hop, smem);
    const long long T = p.t_end - p.t_begin;
    if ((T + F - 1) / F > 0x7fffffffLL) return fail(A2SB_ERR_INVALID, "too many frames per clip (%lld)", T);
    p.tiles_per_clip = (int)((T + F - 1) / F);
    const long long groups = (long long)cx.sm_count * G::GROUPS;
    // Run length 1: the groups of the grid work on consecutive tiles of a clip at the same time, so the
    // partial sectors at tile seams meet in L2 within microseconds (measured: longer runs are slower).
    int run_best = (M == 256 && WIDE) ? 2 : 1;   // n_fft 512, wide 32-frame tiles: runs of 2 tiles 1.08 -> 1.02 ms (every other family loses)
    static const int env_run = [] { const char* e = std::getenv("A2SB_FWD_RUN"); return e ? std::atoi(e) : 0; }();
    if (env_run >= 1 && env_run <= 64) run_best = env_run;   // experiments
    p.run = run_best;
    p.items_per_clip = (p.tiles_per_clip + run_best - 1) / run_best;
    p.total_items = (long long)p.items_per_clip * p.batch;
    const long long ctas = (p.total_items + G::GROUPS - 1) / G::GROUPS;
    // Seam prefetch (stft_fwd.cuh): head seam always; the tail seam of every tile as well only for n_fft = 512 (measured on the
    // round-2 kernels, head only / head + tail: n_fft 512 1.20 / 1.12 ms, 1024 1.03 / 1.10, 2048 1.06 / 1.17, 4096 1.68 / 1.92);
    // the wide pass B (128-byte row segments) wants both: n_fft 512 1.24 / 1.10, 1024 1.18 / 0.99, 2048 two rounds 1.33 / 1.07 --
    // except n_fft 512 in runs of two tiles (head only: 1.02 ms).  prefetch.global.L2::evict_last instead of the plain
    // prefetch changes nothing anywhere; a real load with an evict_last cache hint is 20-100 % slower.
    p.seam = ((M == 256 && !WIDE) || (WIDE && M != 256)) ? 3 : 1;
    static const int env_seam = [] { const char* e = std::getenv("A2SB_SEAM"); return e ? std::atoi(e) : -1; }();
    if (env_seam >= 0) p.seam = env_seam;   // experiments
    // Experiment (A2SB_FWD_CLUSTER=c): launch as clusters of c CTAs (neighbouring tiles) that meet at a cluster barrier after every
    // round, so that both writers of a seam arrive while its prefetched line is still in L2.  Single-group kernels with runs
    // of one tile only (every CTA then knows how many barriers the slowest CTA of its cluster executes).
    static const int env_cluster = [] { const char* e = std::getenv("A2SB_FWD_CLUSTER"); return e ? std::atoi(e) : 0; }();
    p.cluster_barriers = 0;
    if (env_cluster > 1 && env_cluster <= 8 && G::GROUPS == 1 && p.run == 1 && ctas >= cx.sm_count) {
        const long long grid = cx.sm_count - cx.sm_count % env_cluster;        // one CTA per SM for these kernels
        p.cluster_barriers = (int)(((p.total_items + grid - 1) / grid) * ROUNDS);
        tl_next_cluster = env_cluster;
    }
    if (p.out2) {
        // corruption epilogue: shipped chain, default tile geometry of each n_fft only (one extra kernel per family)
        constexpr bool kDefaultGeomC = (M <= 512 && F == 32 && WIDE == 1) || (M == 1024 && F == 16 && ROUNDS == 1) || (M == 2048 && ROUNDS == 2);
        if constexpr (kDefaultGeomC) {
            if (p.pcm || p.wrap_cols > 0 || !(p.epi == kEpiMagPhase && p.pmode == kPowQuarter))
                return fail(A2SB_ERR_INVALID, "the corruption epilogue is built for the shipped forward chain on float32 samples, without wrap padding");
            return launch_persistent(stft_fwd_kernel<M, RA, RB, F, 3, ROUNDS, WIDE>, ctas, G::NT, smem, st, p, cx.sm_count);
        } else {
            return fail(A2SB_ERR_INVALID, "the corruption epilogue is built for the default tile geometry only");
        }
    }
    if (p.pcm) {
        // 16-bit PCM ingest: shipped chain, default tile geometry of each n_fft only (one extra kernel per family)
        constexpr bool kDefaultGeom = (M <= 512 && F == 32 && WIDE == 1) || (M == 1024 && F == 16 && ROUNDS == 1) || (M == 2048 && ROUNDS == 2);
        if constexpr (kDefaultGeom) {
            if (p.epi == kEpiMagPhase && p.pmode == kPowQuarter)
                return launch_persistent(stft_fwd_kernel<M, RA, RB, F, 1, ROUNDS, WIDE, 1>, ctas, G::NT, smem, st, p, cx.sm_count);
            return fail(A2SB_ERR_INVALID, "PCM ingest is built for the shipped forward chain (mag/phase, drop DC, power 0.25)");
        } else {
            return fail(A2SB_ERR_INVALID, "PCM ingest is built for the default tile geometry only (hop too large for it, or A2SB_FWD_TILE set)");
        }
    }
    if (p.epi == kEpiMagPhase && p.pmode == kPowQuarter)
        return launch_persistent(stft_fwd_kernel<M, RA, RB, F, 1, ROUNDS, WIDE>, ctas, G::NT, smem, st, p, cx.sm_count);
    if (p.epi == kEpiMagPhase && p.pmode == kPowNone)
        return launch_persistent(stft_fwd_kernel<M, RA, RB, F, 2, ROUNDS, WIDE>, ctas, G::NT, smem, st, p, cx.sm_count);
    return launch_persistent(stft_fwd_kernel<M, RA, RB, F, 0, ROUNDS, WIDE>, ctas, G::NT, smem, st, p, cx.sm_count);
}

template <int M, int RA, int RB>
static int dispatch_fwd(const LaunchCtx& cx, const FwdParams& p, cudaStream_t st) {
    if constexpr (M >= 2048) {
        // n_fft = 4096: 16-frame tiles in two rounds, 64 x 32 decomposition (512 threads x 128 registers; the one-round kernel
        // needs 8-frame tiles -- 32-byte row segments -- and a radix-64 pass B: 256 threads x 244 registers)
        if (cx.fwd_tile != 8 && p.tw4_alt && FwdGeom<M, 64, 32, 16, 2>::smem_bytes(cx.hop) <= 232448) {
            FwdParams q = p;
            q.tw4 = p.tw4_alt;
            return launch_fwd<M, 64, 32, 16, 2>(cx, q, st);
        }
        return launch_fwd<M, RA, RB, 8>(cx, p, st);   // 16-frame exchange does not fit 227 KB
    } else {
        if constexpr (M <= 512) {
            // n_fft = 512 / 1024: 32-frame tiles (the exchange of 32 frames is what 16 frames cost at n_fft = 2048): two warps per
            // residue class store adjacent 64-byte row segments at the same time.  Falls back to 16 frames when a large hop
            // makes the 32-frame input span too long for shared memory.
            static const int env_wide = [] { const char* e = std::getenv("A2SB_FWD_WIDE"); return e ? std::atoi(e) : 1; }();
            if (cx.fwd_tile == 32 && FwdGeom<M, RA, RB, 32>::smem_bytes(cx.hop) <= 232448)
                return env_wide ? launch_fwd<M, RA, RB, 32, 1, 1>(cx, p, st) : launch_fwd<M, RA, RB, 32>(cx, p, st);
        }
        if constexpr (M == 1024) {
            // n_fft = 2048: 32-frame tiles in two rounds (the 32-frame exchange does not fit; stft_fwd.cuh)
            static const int env_wide = [] { const char* e = std::getenv("A2SB_FWD_WIDE"); return e ? std::atoi(e) : 1; }();
            if (cx.fwd_tile == 32 && FwdGeom<M, RA, RB, 32, 2>::smem_bytes(cx.hop) <= 232448)
                return env_wide ? launch_fwd<M, RA, RB, 32, 2, 1>(cx, p, st) : launch_fwd<M, RA, RB, 32, 2>(cx, p, st);
        }
        return cx.fwd_tile == 8 ? launch_fwd<M, RA, RB, 8>(cx, p, st) : launch_fwd<M, RA, RB, 16>(cx, p, st);
    }
}

// Tensor maps of the local spectrogram buffer for the TMA variant of K2 (layout: istft_inv.cuh::SpecMaps).
static int make_spec_maps(SpecMaps& sm, const InvParams& p, int rows, int RA, int RB, int FW) {
    const long long T = p.spec_T;
    for (int r = 0; r < 4; ++r) {
        const char* row = reinterpret_cast<const char*>(p.spec) + (long long)r * T * 4;
        const long long a = (long long)(reinterpret_cast<uintptr_t>(row) & 15);
        sm.shift[r] = (int)(a / 4);
        const char* base = row - a;
#ifdef A2SB_EMU
        TensorMap5& m = sm.m[r];
        m.base = base;
        const long long dims[5] = {T + a / 4, RB / 4, RA, 3, p.batch};
        const long long strides[4] = {16 * T, 4LL * RB * T, 4LL * rows * T, 12LL * rows * T};
        const int box[5] = {FW, 1, RA, 3, 1};
        for (int i = 0; i < 5; ++i) { m.dims[i] = dims[i]; m.box[i] = box[i]; }
        for (int i = 0; i < 4; ++i) m.strides[i] = strides[i];
#else
        typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                     const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                     CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
        static EncodeFn encode = [] {
            void* fn = nullptr;
            cudaDriverEntryPointQueryResult q;
            if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) != cudaSuccess) fn = nullptr;
            return reinterpret_cast<EncodeFn>(fn);
        }();
        if (!encode) return fail(A2SB_ERR_CUDA, "cuTensorMapEncodeTiled is not available from this driver");
        const cuuint64_t dims[5] = {(cuuint64_t)(T + a / 4), (cuuint64_t)(RB / 4), (cuuint64_t)RA, 3, (cuuint64_t)p.batch};
        const cuuint64_t strides[4] = {(cuuint64_t)(16 * T), (cuuint64_t)(4LL * RB * T), (cuuint64_t)(4LL * rows * T),
                                       (cuuint64_t)(12LL * rows * T)};
        const cuuint32_t box[5] = {(cuuint32_t)FW, 1, (cuuint32_t)RA, 3, 1}, es[5] = {1, 1, 1, 1, 1};
        // L2 promotion 256 B: 0.560 vs 0.570 ms for the bare box stream (tools/microbench/tma_box_stream.cu)
        const CUresult rc = encode(&sm.m[r], CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 5, const_cast<char*>(base), dims, strides, box, es,
                                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (rc != CUDA_SUCCESS) return fail(A2SB_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d) for the spectrogram view", (int)rc);
#endif
    }
    return A2SB_OK;
}

#if defined(A2SB_INV_PROF) && !defined(A2SB_EMU)
// experiments: per-warp cycle counters of the inverse kernel, summed over launches and printed at exit
static long long* g_prof = nullptr;
static void prof_report() {
    if (!g_prof) return;
    std::vector<long long> h(148 * 32 * 8);
    cudaMemcpy(h.data(), g_prof, h.size() * 8, cudaMemcpyDeviceToHost);
    static const char* names[8] = {"job fetch", "E wait(box)", "E work", "X wait(exp)", "X work", "A->barrier", "pass B", "OLA"};
    double tot[8] = {0};
    int n = 0;
    for (int b = 0; b < 148; ++b)
        for (int w = 0; w < 16; ++w) {
            for (int i = 0; i < 8; ++i) tot[i] += (double)h[((size_t)b * 32 + w) * 8 + i];
            ++n;
        }
    double all = 0;
    for (int i = 0; i < 8; ++i) all += tot[i];
    std::fprintf(stderr, "[A2SB_INV_PROF] last launch, mean cycles per warp (%.0f total):\n", all / n);
    for (int i = 0; i < 8; ++i) std::fprintf(stderr, "  %-12s %10.0f  %5.1f %%\n", names[i], tot[i] / n, 100.0 * tot[i] / all);
}
#endif

template <int M, int RA, int RB, int F>
static int launch_inv(const LaunchCtx& cx, const InvParams& p_in, cudaStream_t st) {
    using G = InvGeom<M, RA, RB, F>;
    InvParams p = p_in;
#if defined(A2SB_INV_PROF) && !defined(A2SB_EMU)
    if (!g_prof) { cudaMalloc(&g_prof, 148 * 32 * 8 * sizeof(long long)); cudaMemset(g_prof, 0, 148 * 32 * 8 * sizeof(long long)); std::atexit(prof_report); }
    p.prof = g_prof;
#endif
    const size_t smem = G::smem_bytes(cx.hop);
    const bool fast = p.in_kind == kInMagPhase && !p.has_dc && p.svd_fix && p.pmode == kPowFour;
    SpecMaps maps{};
    if constexpr (G::TMA_OK) {
        // Shipped chain only, experiments (both measured SLOWER than the plain register-load kernel on 256 x 10 s clips, n_fft
        // 2048: 1.04 ms plain, 1.26 ms either way -- DESIGN.md section 8): A2SB_INV_TMA=1: box ring in shared memory filled by
        // tensor-map TMA, pass A as a job queue; A2SB_INV_TMA=2: register loads + one tensor-map L2 prefetch per residue of
        // the NEXT tile, issued when pass A ends.  Default 0: plain register loads.
        static const int env_tma = [] { const char* e = std::getenv("A2SB_INV_TMA"); return e ? std::atoi(e) : 0; }();
        static const int env_slots = [] { const char* e = std::getenv("A2SB_INV_SLOTS"); return e ? std::atoi(e) : 0; }();
        if (fast && env_tma && p.spec_T < (1LL << 31) - 8) {
            if (env_tma == 2) {
                if (int rc = make_spec_maps(maps, p, M, RA, RB, G::FW)) return rc;
                g_tma_launches.fetch_add(1);
                return launch_persistent_n(istft_inv_kernel<M, RA, RB, F, 1, 2>, p.total_items, G::NT, smem, st, cx.sm_count, p, maps, 0);
            }
            const size_t limit = (G::NT <= 256) ? 115712 : 232448;   // (228 KB - 1 KB per CTA) / CTAs per SM
            int slots = 0;
            for (int s = RB; s >= 2; s >>= 1)
                if (RB % s == 0 && G::smem_bytes_tma(cx.hop, s) <= limit) { slots = s; break; }
            if (env_slots >= 2 && RB % env_slots == 0 && G::smem_bytes_tma(cx.hop, env_slots) <= 232448) slots = env_slots;
            if (slots >= 2) {
                if (int rc = make_spec_maps(maps, p, M, RA, RB, G::FW)) return rc;
                g_tma_launches.fetch_add(1);
                return launch_persistent_n(istft_inv_kernel<M, RA, RB, F, 1, 1>, p.total_items, G::NT, G::smem_bytes_tma(cx.hop, slots),
                                           st, cx.sm_count, p, maps, slots);
            }
        }
    }
    if (p.out_pcm) {
        if (!fast) return fail(A2SB_ERR_INVALID, "PCM output is built for the shipped chain only (mag/phase rows 1.., power 4, phase fix)");
        return launch_persistent_n(istft_inv_kernel<M, RA, RB, F, 1, 0, 2>, p.total_items, G::NT, smem, st, cx.sm_count, p, maps, 0);
    }
    if (p.n_mirror > 0) {
        if (!fast) return fail(A2SB_ERR_INVALID, "mirrored output is built for the shipped chain only (mag/phase rows 1.., power 4, phase fix)");
        return launch_persistent_n(istft_inv_kernel<M, RA, RB, F, 1, 0, 1>, p.total_items, G::NT, smem, st, cx.sm_count, p, maps, 0);
    }
    return fast ? launch_persistent_n(istft_inv_kernel<M, RA, RB, F, 1, 0>, p.total_items, G::NT, smem, st, cx.sm_count, p, maps, 0)
                : launch_persistent_n(istft_inv_kernel<M, RA, RB, F, 0, 0>, p.total_items, G::NT, smem, st, cx.sm_count, p, maps, 0);
}

template <int M, int RA, int RB>
static int dispatch_inv(const LaunchCtx& cx, const InvParams& p, cudaStream_t st) {
    // cx.inv_tile == inv_tile_frames(M) unless overridden; RA % 32 == 0 is what the 8-frame layout needs
    if constexpr (M >= 2048) return launch_inv<M, RA, RB, 8>(cx, p, st);   // 16-frame exchange does not fit 227 KB
    if constexpr (M <= 512) {
        // n_fft 512 / 1024: 32-frame tiles (two warps per residue class load adjacent 64-byte row segments at the same time)
        if (cx.inv_tile == 32 && InvGeom<M, RA, RB, 32>::smem_bytes(cx.hop) <= 232448) return launch_inv<M, RA, RB, 32>(cx, p, st);
    }
    if constexpr (M >= 2048) return A2SB_OK;
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
int run_inv_2048(const LaunchCtx& c, const InvParams& p, cudaStream_t s) { return dispatch_inv<2048, 32, 64>(c, p, s); }
#endif

}  // namespace a2sb
