// istft_inv.cuh -- K2: fused power-expand + phase normalise + inverse real FFT + windowed
// overlap-add + envelope normalisation + centre trim.
//
// Replaces, in one pass over HBM, the reference's inverse chain
//   PowerScaleSpectrogram(4) -> SpectrogramAddDCTerm -> SVDFixMagInstPhase -> MagInstPhaseToComplex
//   -> InverseComplexSpectrogram      (A2SB/audio_transforms/transforms.py:121-228; torch.istft)
// The per-bin 2x2 SVD of SVDFixMagInstPhase (transforms.py:144-160) projects [[c,-s],[s,c]] onto
// SO(2); its closed form is (c, s)/sqrt(c^2+s^2) with (0,0) -> (1,0).
//
// Geometry: a work item is (clip, chunk of output hop-blocks).  The CTA sweeps the chunk in tiles
// of F = 16 consecutive frames, carrying the (N - hop)-sample overlap tail in shared memory, so
// every frame is transformed once (plus N/hop-1 warm-up frames per chunk).
//   * pass A (frame-minor threads: lane = 16 frames x {residue j, RB-j}): loads whose lanes run
//     along the spectrogram's fastest (frame) axis -> 64-byte row segments; power expansion and
//     phase normalisation on pairs of bins in packed (FFMA2/FMUL2) registers, MUFU for rcp/rsqrt;
//     the split needs X[M-k] from the partner half-warp (__shfl_xor); radix-RA register iFFT
//     (scalar decimation-in-frequency stage + packed radix-RA/2 of the two half-sequences);
//   * exchange through shared memory (conflict-free both ways);
//   * pass B (frame-major threads): packed twiddle multiply, packed radix-RB/2 + scalar last stage,
//     synthesis window -> the frame's N samples, written over the frame's own exchange region;
//   * overlap-add in ascending frame order (deterministic), * 1/envelope, float4 stores.
#pragma once
#include "a2sb_common.cuh"
#include "fftx2.cuh"
#include "stft_fwd.cuh"  // st_stream
#include "tma.cuh"

namespace a2sb {

#if defined(A2SB_INV_PROF) && !defined(A2SB_EMU)
#define A2SB_PROF_DECL long long pf_[8] = {0, 0, 0, 0, 0, 0, 0, 0}; long long pf_t_ = clock64();
#define A2SB_PROF(i) do { const long long n_ = clock64(); pf_[i] += n_ - pf_t_; pf_t_ = n_; } while (0)
#define A2SB_PROF_FLUSH(p) do { if ((p).prof && (threadIdx.x & 31) == 0) for (int i_ = 0; i_ < 8; ++i_) (p).prof[((long long)blockIdx.x * 32 + (threadIdx.x >> 5)) * 8 + i_] = pf_[i_]; } while (0)
#else
#define A2SB_PROF_DECL
#define A2SB_PROF(i)
#define A2SB_PROF_FLUSH(p)
#endif

#ifndef A2SB_INV_LB
#define A2SB_INV_LB 4   // bin pairs per load batch (6 loads each) issued before the first use; measured 1, 2, 4: 1.285 ms,
                        // 8: 1.307 ms, 16: 1.304 ms -- the LSU queue, not DRAM latency, is what the loads wait on
#endif
#ifndef A2SB_INV_REV
#define A2SB_INV_REV 1  // work items are taken from the END of the batch first: the clips K1 / the network wrote last may
                        // still be in L2 when K2 starts (1.030 -> 1.024 ms on 256 clips; bit-identical; 0 = forward order)
#endif
#ifndef A2SB_INV_PF
#define A2SB_INV_PF 0   // L2 prefetches per 64-byte row segment of the next tile (0..3).  Measured on 256 x 10 s
                        // clips: 0 -> 1.307 ms, 1 -> 1.335 ms, 2 -> 1.368 ms, 3 -> 1.59 ms: the extra LSU requests cost
                        // more than the DRAM latency they hide, so the prefetch is off.
#endif

struct InvParams {
    const float* spec;        // [batch][C][rows][spec_T] local spectrogram buffers
    long long spec_T;         // frames per row of the local buffer
    long long spec_t_first;   // global frame index of local column 0
    long long n_frames;       // global frame count T
    float* out;               // [batch][out_stride]
    long long out_stride;
    long long out_first;      // global (trimmed) sample index of out[b][0]
    long long out_count;      // samples per clip in the local output buffer
    long long hop_begin, hop_end;  // global hop-block range produced by this launch
    int batch;
    int hop;
    int chunk_hops;           // hop-blocks per work item ( = m*F - (N/hop - 1) )
    int chunks_per_clip;
    long long total_items;
    const float* window;      // [N] synthesis window * (1/N)
    const float* wsq;         // [N] window^2 (envelope near clip edges)
    const float* inv_env;     // [hop] 1 / sum_m w^2[r + m*hop] (interior envelope)
    const float4* tw4;        // [RB][RA/2 + 1] inter-pass twiddles: (cos, cos, sin, sin)(+2 pi ja {2k, 2k+1} / M)
    const float2* twN;        // [M/2+1]  (cos, sin)(2 pi k / N)
    int in_kind;              // kInComplex / kInMagPhase
    int has_dc;               // 1: rows are bins 0..M; 0: bins 1..M and DC := 0*row0 (SpectrogramAddDCTerm)
    int svd_fix;              // 1: project (cos, sin) onto the unit circle (SVDFixMagInstPhase)
    int pmode;                // kPowNone / kPowFour / kPowGeneric
    float power, eps;
    long long* prof;          // -DA2SB_INV_PROF: [grid][32 warps][8] cycle counters (experiments only)
    // Fused gather of a sharded result (MIR kernels only): every output vector is also stored at the same offset of
    // n_mirror peer buffers (peer-mapped device memory over NVLink), or -- mirror_mc -- stored ONCE through a multicast
    // address that NVSwitch replicates into every GPU's buffer, this one included (then `out` itself is not written).
    float* mirror[8];
    int n_mirror;
    int mirror_mc;
    int out_pcm;              // 1: `out` is int16 PCM (MIR = 2 kernels)
};

// float -> 16-bit PCM the way libsndfile writes a float buffer to a PCM_16 file with clipping enabled (pcm.c, f2s_clip_array;
// python-soundfile's sf.write -- A2SB/inference/A2SB_inpaint_dataset.py:126 -- opens its files with SFC_SET_CLIPPING):
// scaled = x * 2^31, saturated to int32, rounded to nearest (lrintf), arithmetic shift right by 16.
A2SB_DEV int pcm16_from_float(float x) {
    const float sc = x * 2147483648.0f;
#ifdef A2SB_EMU
    if (!(sc < 2147483647.0f)) return (sc != sc) ? 0 : 0x7FFF;
    if (sc <= -2147483648.0f) return -0x8000;
    return (int)(std::lrintf(sc) >> 16);
#else
    return __float2int_rn(sc) >> 16;   // cvt.rni.s32.f32 saturates (NaN -> 0)
#endif
}

// multimem.st: one store, replicated by the switch into every device buffer bound to the multicast object
A2SB_DEV void st_multicast(float* a, float4 v) {
#ifdef A2SB_EMU
    *reinterpret_cast<float4*>(a) = v;
#else
    asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(a), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
#endif
}
A2SB_DEV void st_multicast(float* a, float v) {
#ifdef A2SB_EMU
    *a = v;
#else
    asm volatile("multimem.st.relaxed.sys.global.f32 [%0], %1;" ::"l"(a), "f"(v) : "memory");
#endif
}

// Tensor maps of the spectrogram for the TMA variant of pass A: one per (row mod 4), because the row pitch (4 * spec_T
// bytes) is in general not a multiple of 16 bytes but four rows are; dimensions (frame, row group jr [4 rows apart],
// q [RB rows apart], plane, clip).  The base of map r is row r's address rounded DOWN to 16 bytes; shift[r] is the
// number of elements that moved it, i.e. frame column j of the local buffer is element j + shift[r] of the map.
struct SpecMaps {
    TensorMap5 m[4];
    int shift[4];
};

template <int M, int RA, int RB, int F>
struct InvGeom {
    static constexpr int N = 2 * M;
    static constexpr int NT = F * RB;            // one pass-A item per thread
    static constexpr bool PAIR_B = (RB == 2 * RA);   // n_fft 4096: one pass-B item (radix RB = 64) per LANE PAIR (below)
    static constexpr int ITEMS_B = PAIR_B ? 0 : RA / RB;      // pass-B items per thread
    static constexpr int CLS = RB / 2;
    static constexpr int FL = (F < 16) ? F : 16;     // frames per lane group (half-warp) of pass A
    static constexpr int FB = F / FL;                // F = 32: a residue class takes FB warps, one per 16-frame block (their
                                                     // 64-byte row segments are adjacent and loaded at the same time)
    static constexpr int CPW = 32 / (2 * FL);    // residue classes per warp
    // Imaginary plane offset inside a frame region and frame region stride.  Kept as tight as the
    // bank skews allow: the kernel's shared memory decides how much of the 256 KB SM array is left as L1, and L1
    // capacity bounds the spectrogram loads in flight (K2 is 30 % slower with 28 KB of L1 than with 60 KB).
    static constexpr int IMOFF = M + ((CPW > 1) ? 32 : 16);
    static constexpr int FS = 2 * IMOFF + 2;   // == 2 mod 32: pass A stores float2 pairs, 16 lanes x 2 banks per wavefront
    static constexpr int TWS = RA / 2 + 1;       // float4 row stride of the inter-pass twiddle table [RB][TWS]
    static_assert(M == RA * RB, "two-pass decomposition");
    static_assert(F == 8 || F == 16 || F == 32, "tile width");
    static_assert((PAIR_B || RA % RB == 0) && NT % 32 == 0 && (NT / 32) * CPW == CLS * FB, "thread mapping");
    static_assert(!PAIR_B || (NT == 2 * F * RA && RA == 32), "pair-split pass B: two threads per (frame, residue) item, radix-32 halves");
    static_assert((M / 2) % 32 == 0, "half-plane offset must keep the 16-bank skew");
    static_assert(CPW == 1 || RA % 32 == 0, "class skew assumes bank-aligned residue blocks");
    static_assert(RB <= 64, "residue classes");
    // Start of the RA-word block of class c (residue c in the lower half h = 0, RB - c -- or RB/2 for
    // c = 0 -- in the upper half h = 1).  The upper half sits 16 banks away.  With two classes per warp
    // (F = 8) the odd classes of a half are stored after its even classes, 8 banks further, so a warp's
    // four lane groups (even/odd class x lower/upper half, 8 frames each) never collide.
    A2SB_HD static constexpr int cblk(int c, int h) {
        return (h ? (RB / 2) * RA + 16 : 0) + ((CPW > 1) ? (c >> 1) * RA + (c & 1) * ((RB / 4) * RA + 8) : c * RA);
    }
    // start of residue ja's block
    A2SB_HD static constexpr int blk(int ja) {
        return (ja == 0) ? cblk(0, 0) : (ja == RB / 2) ? cblk(0, 1) : (ja < RB / 2) ? cblk(ja, 0) : cblk(RB - ja, 1);
    }
    // 16-byte aligned start of frame f's time-domain buffer (aliases its exchange region)
    A2SB_HD static constexpr int fbuf(int f) { return f * FS + 2 * (f & 1); }   // FS == 2 mod 4
    // n_fft = 4096: with the synthesis window (16 KB) in shared memory the footprint is 64 bytes over the 196 KB carve-out
    // and the SM is left with 28 KB of L1; read through L1 instead, the kernel keeps 60 KB.
    // n_fft = 1024: two CTAs per SM; without the 4 KB window table both fit the 164 KB carve-out (92 KB of L1, not 60).
    static constexpr bool WIN_SMEM = (M < 2048 && M != 512);
    static constexpr size_t off_win = 0;
    static constexpr size_t off_tw4 = off_win + (WIN_SMEM ? sizeof(float) * N : 0);
    static constexpr bool TW4_SMEM = (M < 2048);   // n_fft = 4096: inter-pass twiddles (17 KB) through L1 as well -> 164 KB carve-out
    static constexpr size_t off_twN = off_tw4 + (TW4_SMEM ? sizeof(float4) * RB * TWS : 0);
    static constexpr size_t off_x = ((off_twN + sizeof(float2) * (M / 2 + 1) + 15) / 16) * 16;
    static constexpr size_t off_dyn = ((off_x + sizeof(float) * ((size_t)F * FS + 4) + 15) / 16) * 16;
    // dynamic tail: carry[2][N - hop]  (the 1 / sum w^2 table is read from global memory: one float4 per thread and tile)
    A2SB_HD static size_t smem_bytes(int hop) { return off_dyn + sizeof(float) * (2 * (size_t)(N - hop)); }
    // ---- TMA variant: RB "box full" mbarriers + a ring of `slots` boxes behind the carry buffers.
    // A box = what one half-warp of pass A consumes for a tile: [3 planes][RA rows of one residue class][FW frames];
    // FW = F + 4 because a box must start on a 16-byte boundary of the row (tma.cuh) while tiles start at any frame.
    static constexpr int FW = F + 4;
    static constexpr int BOX = 3 * RA * FW;                       // floats per box
    static constexpr unsigned BOX_BYTES = (unsigned)BOX * 4u;     // multiple of 128 for every instantiation
    static constexpr bool TMA_OK = (CPW == 1) && (FB == 1) && (RB % 4 == 0) && (RB <= 32) && (BOX_BYTES % 128 == 0);   // one mbarrier per box position, 256 bytes reserved
    A2SB_HD static size_t ring_off(int hop) { return ((smem_bytes(hop) + 127) / 128) * 128; }
    static constexpr size_t RING_HDR = 1024;   // box-full mbarriers [RB], box-expanded mbarriers [RB] (256 bytes each), job counters
    static size_t smem_bytes_tma(int hop, int slots) { return ring_off(hop) + RING_HDR + (size_t)slots * BOX_BYTES; }
};

enum : int { kInComplex = 0, kInMagPhase = 1 };

// Spectrogram loads.  Lanes run along the frame axis, so a warp instruction reads two 64-byte row
// segments that are only 8-byte aligned; the 32-byte sectors at both ends are shared with the
// neighbouring tiles of the same sweep.  A2SB_INV_LD picks the cache policy.  Measured (256 x 10 s clips):
// 0 ld.global.nc 1.29 ms | 4 +L2::128B 1.29 | 5 +L2::256B 1.22 (default: on a miss L2 fetches the 256-byte
// neighbourhood, which the neighbouring CTAs of the round-robin sweep are about to ask for -- DRAM sees longer
// bursts per page) | 6 +L1::no_allocate 1.55 | 7 +L1::evict_last 1.22.
A2SB_DEV float ld_spec(const float* p) {
#if defined(A2SB_EMU)
    return *p;
#elif A2SB_INV_LD == 1
    return *p;
#elif A2SB_INV_LD == 2
    return __ldcg(p);
#elif A2SB_INV_LD == 4
    float v;
    asm volatile("ld.global.nc.L2::128B.f32 %0, [%1];" : "=f"(v) : "l"(p));
    return v;
#elif A2SB_INV_LD == 5
    float v;
    asm volatile("ld.global.nc.L2::256B.f32 %0, [%1];" : "=f"(v) : "l"(p));
    return v;
#elif A2SB_INV_LD == 6
    float v;
    asm volatile("ld.global.nc.L1::no_allocate.L2::256B.f32 %0, [%1];" : "=f"(v) : "l"(p));
    return v;
#elif A2SB_INV_LD == 7
    float v;
    asm volatile("ld.global.nc.L1::evict_last.L2::256B.f32 %0, [%1];" : "=f"(v) : "l"(p));
    return v;
#elif A2SB_INV_LD == 3
    float v;
    asm volatile("ld.global.L2::evict_last.f32 %0, [%1];" : "=f"(v) : "l"(p));
    return v;
#else
    return __ldg(p);
#endif
}

// (m, cos, sin) -> X = m' * (c', s')   [PowerScale -> SVDFix -> MagInstPhaseToComplex]; careful form:
// any exponent, exact handling of (c, s) -> 0 (SVDFixMagInstPhase maps the zero matrix to the identity).
A2SB_DEV void inv_expand(const InvParams& p, float m, float c, float s, float& xr, float& xi) {
    if (p.pmode == kPowFour) m = m * power_scale_factor<kPowFour>(fabsf(m), p.power, p.eps);
    else if (p.pmode == kPowGeneric) m = m * power_scale_factor<kPowGeneric>(fabsf(m), p.power, p.eps);
    if (p.svd_fix) {
        float n2 = c * c + s * s;
        if (n2 < 1e-30f) {  // rescale before declaring the pair degenerate
            c *= 1.8446744e19f; s *= 1.8446744e19f;
            n2 = c * c + s * s;
        }
        if (n2 > 0.0f) {
            const float rn = rsqrt_approx(n2);
            c *= rn; s *= rn;
        } else {
            c = 1.0f; s = 0.0f;
        }
    }
    xr = m * c;
    xi = m * s;
}

// Shipped chain (power 4 on the magnitude, phase fix) for two bins at once in packed registers:
//   m' = m * |m|^4 / (|m| + eps),  (c', s') = (c, s) / sqrt(c^2 + s^2),  X = m' (c', s').
A2SB_DEV void inv_expand_fast2(float2 m, float2 c, float2 s, float eps, unsigned& minbits, float2& xr, float2& xi) {
    const float2 a2 = p2_mul(m, m);
    const float2 a4 = p2_mul(a2, a2);
    float2 r, rn;
    r.x = rcp_approx(fabsf(m.x) + eps);
    r.y = rcp_approx(fabsf(m.y) + eps);
    const float2 n2 = p2_fma(c, c, p2_mul(s, s));
    minbits = min3u(minbits, __float_as_uint(n2.x), __float_as_uint(n2.y));
    rn.x = rsqrt_approx(n2.x);
    rn.y = rsqrt_approx(n2.y);
    const float2 g = p2_mul(p2_mul(m, p2_mul(a4, r)), rn);
    xr = p2_mul(g, c);
    xi = p2_mul(g, s);
}

// Pair (k, M-k): Zk = E + P, Zm = conj(E - P) with E = Xk + conj(Xm), D = Xk - conj(Xm),
// P = i * conj(W^k) * D, conj(W^k) = (c, s) = (cos, sin)(2 pi k / N).  (0.5 folded into window.)
A2SB_DEV void inv_pair(float xkr, float xki, float xmr, float xmi, float2 w, float& zkr, float& zki, float& zmr,
                       float& zmi) {
    const float er = xkr + xmr, ei = xki - xmi;
    const float dr = xkr - xmr, di = xki + xmi;
    const float pr = -s_fma(w.x, di, w.y * dr);
    const float pi = s_fma(w.x, dr, -(w.y * di));
    zkr = er + pr; zki = ei + pi;
    zmr = er - pr; zmi = pi - ei;
}

// FAST = 1: the shipped chain (mag/phase rows 1..M, power 4, phase fix) through the packed fast
// expansion, falling back to the careful one when a (cos, sin) pair is degenerate.
// FAST = 0: every bin through the careful expansion (complex input, DC row present, any exponent).
// TMA = 1 (FAST only): the spectrogram reaches pass A through tensor-map TMA box loads into a shared-memory ring instead of
// register loads.  There is no producer warp: the warp that has consumed box n of the CTA's box sequence issues box
// n + slots into the slot it has just freed, so the loads of a tile -- and of the first `slots` boxes of the next tile,
// which arrive under pass B and the overlap-add -- are in flight without holding registers or L1 lines.  One mbarrier
// per box position of a tile, used once per tile: tiles are separated by CTA barriers, so no waiter is ever more than
// one phase away; RB % slots == 0 puts the previous use of a box position's barrier on the issuing warp's own chain
// (box n was issued after n - slots was consumed, ... , n + slots - RB), i.e. it has completed and been waited on.
// MIR = 1: the fused-gather variant (InvParams::mirror).  MIR = 2: `out` is a 16-bit PCM buffer (strides and counts in samples).
template <int M, int RA, int RB, int F, int FAST, int TMA, int MIR = 0>
__global__ void __launch_bounds__(F * RB, (F * RB <= 256 && M < 2048) ? 2 : 1)
istft_inv_kernel(const InvParams p, const A2SB_GRID_CONSTANT SpecMaps maps, const int slots) {
    using G = InvGeom<M, RA, RB, F>;
    static_assert(!TMA || (FAST && G::TMA_OK), "TMA variants: shipped chain, one class per warp");
    constexpr bool RING = (TMA == 1);   // TMA == 2: register loads as in the plain variant + tensor-map L2 prefetch of the next tile
    constexpr int kF = F;
    constexpr int N = G::N, NT = G::NT, FS = G::FS, IMOFF = G::IMOFF;
    A2SB_DYN_SMEM(smem);
    float* s_win = reinterpret_cast<float*>(smem + G::off_win);
    float4* s_tw4 = reinterpret_cast<float4*>(smem + G::off_tw4);
    float2* s_twN = reinterpret_cast<float2*>(smem + G::off_twN);
    float* s_x = reinterpret_cast<float*>(smem + G::off_x);
    float* s_dyn = reinterpret_cast<float*>(smem + G::off_dyn);

    const int tid = threadIdx.x;
    const int H = p.hop;
    const int ROV = N / H;            // frames overlapping one output sample
    const int NC = N - H;             // carried overlap tail
    float* s_carry0 = s_dyn;
    float* s_carry1 = s_carry0 + NC;
    const bool cplx = (p.in_kind == kInComplex);
    const int C = cplx ? 2 : 3;
    const int rows = cplx ? M + 1 : (M + p.has_dc);
    const int row_of_k0 = cplx ? 0 : (p.has_dc ? 0 : -1);  // row index of bin k is k + row_of_k0
    const long long plane = (long long)rows * p.spec_T;
    const long long T = p.n_frames;

    if (G::WIN_SMEM) {
        for (int i = tid; i < N; i += NT) s_win[i] = p.window[i];
    }
    if (G::TW4_SMEM) {
        for (int i = tid; i < RB * G::TWS; i += NT) s_tw4[i] = p.tw4[i];
    }
    for (int i = tid; i <= M / 2; i += NT) s_twN[i] = p.twN[i];
    unsigned long long* s_bar = reinterpret_cast<unsigned long long*>(smem + G::ring_off(p.hop));
    unsigned long long* s_exp = s_bar + 32;
    int* s_job = reinterpret_cast<int*>(s_bar + 64);
    float* s_ring = reinterpret_cast<float*>(smem + G::ring_off(p.hop) + G::RING_HDR);
    if (RING && tid == 0) {
        for (int i = 0; i < RB; ++i) { mbar_init(s_bar + i, 1); mbar_init(s_exp + i, 1); }
        s_job[0] = 0; s_job[1] = 0;
        fence_mbar_init();
    }
    __syncthreads();

    const int warp = tid >> 5, lane = tid & 31;
    constexpr int FL = G::FL, FB = G::FB;
    const int h = (lane / FL) & 1, t = (warp % FB) * FL + lane % FL;
    const int c = (warp / FB) * G::CPW + lane / (2 * FL);
    const int ja = (c == 0) ? (h ? RB / 2 : 0) : (h ? RB - c : c);
    // first global frame of tile `tile` of work item `item`, and its clip
    auto tile_origin = [&](long long item, int tile, int& b, long long& t0) {
        if (A2SB_INV_REV) item = p.total_items - 1 - item;
        b = (int)(item / p.chunks_per_clip);
        t0 = p.hop_begin + (long long)(item % p.chunks_per_clip) * p.chunk_hops - (N / p.hop - 1) + (long long)tile * F;
    };
    // (TMA) one thread: box `bx` (0..RB-1: class bx / 2, half bx % 2) of the tile starting at frame t0 of clip b
    auto issue_box = [&](int b, long long t0, int bx) {
        const int cc = bx >> 1, hh = bx & 1;
        const int jj = (cc == 0) ? (hh ? RB / 2 : 0) : (hh ? RB - cc : cc);
        const int rho = (jj + RB - 1) % RB;           // rows rho + RB * q hold bins jj + RB * q (jj = 0: bins RB .. M)
        const int e0 = (int)(t0 - p.spec_t_first) + maps.shift[rho & 3];
        fence_proxy_async();                          // the slot was read through the generic proxy
        tma_load_box5(s_ring + (size_t)(bx % slots) * G::BOX, &maps.m[rho & 3], s_bar + bx, G::BOX_BYTES, e0 & ~3, rho >> 2, 0, 0, b);
    };
    unsigned tile_count = 0;                          // tiles this CTA has started (mbarrier phase parity)
    A2SB_PROF_DECL
    if (RING && tid == 0 && blockIdx.x < p.total_items) {
        int b0; long long t00;
        tile_origin(blockIdx.x, 0, b0, t00);
        for (int bx = 0; bx < slots && bx < RB; ++bx) issue_box(b0, t00, bx);
    }
#ifdef A2SB_CONST_T   // experiment: row / plane strides as compile-time constants (immediate load offsets)
    constexpr unsigned long long rowB = 4ull * A2SB_CONST_T, stepB = (unsigned long long)RB * rowB, planeB = rowB * M,
                                 plane2B = 2ull * planeB;
#else
    const unsigned long long rowB = 4ull * (unsigned long long)p.spec_T;      // bytes between consecutive rows
    const unsigned long long stepB = (unsigned long long)RB * rowB;           // bins ja + RB*q -> ja + RB*(q+1)
    const unsigned long long planeB = 4ull * (unsigned long long)plane, plane2B = 2ull * planeB;
#endif

    // hop = N/4 (every A2SB configuration): the synthesis window is not applied by pass B (RB shared-memory reads and 2 RB
    // multiplies per thread and tile) but by the overlap-add, which touches the same four window columns
    // w[m*H + r .. r+3] for every hop-block it emits: they live in registers for the whole kernel and the add becomes an FMA.
    const bool win_in_ola = (H == N / 4);
    float4 wv[4];
    A2SB_PRAGMA_UNROLL
    for (int m = 0; m < 4; ++m)
        wv[m] = win_in_ola ? __ldg(reinterpret_cast<const float4*>(p.window + m * (N / 4) + (tid * 4) % (N / 4))) : make_float4(0.f, 0.f, 0.f, 0.f);

    // Pass A after the bins X[ja + RB*q] of (frame t, residue ja) are in registers: pairing with the partner residue
    // (real-FFT split), radix-RA inverse DFT over q, inter-pass twiddle, exchange store.  (c, h) = class and half of the
    // lane group; class 0 holds the two self-paired residues 0 and RB/2.  c is warp-uniform (CPW == 1) or the shuffles
    // are executed by every lane (CPW > 1).
    auto pass_a_tail = [&](float (&xr)[RA], float (&xi)[RA], const bool valid, const int c, const int h, const int t, const int ja,
                           const float ny_a, const float ny_b, const float ny_c, const float dc_m) {
        if (c != 0 || G::CPW > 1) {
            // classes exchange X[M-k] / Z[M-k] with the partner lane group; in warp 0 (which also holds
            // class 0) every lane takes part in the shuffles, class-0 lanes just ignore the results
            const bool live = (c != 0);
            A2SB_PRAGMA_UNROLL
            for (int q = 0; q < RA / 2; ++q) {
                const float xmr = __shfl_xor_sync(0xffffffffu, xr[RA - 1 - q], FL);
                const float xmi = __shfl_xor_sync(0xffffffffu, xi[RA - 1 - q], FL);
                float zkr, zki, zmr, zmi;
                inv_pair(xr[q], xi[q], xmr, xmi, s_twN[live ? ja + RB * q : 0], zkr, zki, zmr, zmi);
                // the partner computed Z for my bin ja + RB*(RA-1-q)
                const float br = __shfl_xor_sync(0xffffffffu, zmr, FL);
                const float bi = __shfl_xor_sync(0xffffffffu, zmi, FL);
                if (live) { xr[q] = zkr; xi[q] = zki; xr[RA - 1 - q] = br; xi[RA - 1 - q] = bi; }
            }
        }
        if (c == 0 && h == 0) {
            // ja = 0: k = RB*q pairs with RB*(RA-q); k = 0 pairs with the Nyquist bin M.
            float x0 = xr[0], nyq = RING ? xi[0] : 0.0f;   // ring variant: the expansion job left (DC, Nyquist) in (re, im)[0]
            if (valid && !RING) {
                if (cplx) {
                    nyq = ny_a;
                } else {
                    float xi_unused;
                    inv_expand(p, ny_a, ny_b, ny_c, nyq, xi_unused);
                    if (!p.has_dc) {
                        // SpectrogramAddDCTerm (transforms.py:227): dc = spec[..., :1, :] * 0
                        // (zero, but NaN/Inf in row 0 propagate exactly like the reference).
                        float m = dc_m;
                        if (p.pmode == kPowFour) m = m * power_scale_factor<kPowFour>(fabsf(m), p.power, p.eps);
                        else if (p.pmode == kPowGeneric) m = m * power_scale_factor<kPowGeneric>(fabsf(m), p.power, p.eps);
                        x0 = m * 0.0f;
                    }
                }
            }
            // irfft ignores Im X[0] and Im X[M]
            xr[0] = x0 + nyq;
            xi[0] = x0 - nyq;
            A2SB_PRAGMA_UNROLL
            for (int q = 1; q < RA / 2; ++q)
                inv_pair(xr[q], xi[q], xr[RA - q], xi[RA - q], s_twN[RB * q], xr[q], xi[q], xr[RA - q], xi[RA - q]);
            xr[RA / 2] = 2.0f * xr[RA / 2];  // k = M/2: Z = 2 conj(X)
            xi[RA / 2] = -2.0f * xi[RA / 2];
        } else if (c == 0) {
            // ja = RB/2: k = RB/2 + RB*q pairs with RB/2 + RB*(RA-1-q).
            A2SB_PRAGMA_UNROLL
            for (int q = 0; q < RA / 2; ++q)
                inv_pair(xr[q], xi[q], xr[RA - 1 - q], xi[RA - 1 - q], s_twN[RB / 2 + RB * q], xr[q], xi[q],
                         xr[RA - 1 - q], xi[RA - 1 - q]);
        }
        // radix-RA inverse DFT over q: scalar DIF stage, then the two half-sequences packed
        float2 pre[RA / 2], pim[RA / 2];
        static_for<0, RA / 2>([&](auto Q) {
            constexpr int q = decltype(Q)::value;
            dif_first<RA, +1, q>(xr[q], xi[q], xr[q + RA / 2], xi[q + RA / 2], pre[q], pim[q]);
        });
        fft_v<RA / 2, +1, float2>(pre, pim);  // y[ja*RA + 2k] in .x, y[ja*RA + 2k + 1] in .y
        {
            // The inter-pass twiddle W^(ja*jb) is applied HERE, on outputs jb = 2k, 2k+1 of residue ja, instead of
            // at the start of pass B: pass A waits on HBM and has issue slots to spare, pass B does not
            // (n_fft = 2048: 1.122 -> 1.069 ms).  Same operands and operations as before: bit-identical results.
            const float4* twa = (G::TW4_SMEM ? s_tw4 : p.tw4) + ja * G::TWS;
            A2SB_PRAGMA_UNROLL
            for (int k = 0; k < RA / 2; ++k) {
                const float4 w = G::TW4_SMEM ? twa[k] : __ldg(twa + k);
                const float2 cp = make_float2(w.x, w.y), sp = make_float2(w.z, w.w);
                const float2 tr = p2_fma(pre[k], cp, p2_neg(p2_mul(pim[k], sp)));
                pim[k] = p2_fma(pre[k], sp, p2_mul(pim[k], cp));
                pre[k] = tr;
            }
        }
        float* dst = s_x + t * FS + G::cblk(c, h);
        A2SB_PRAGMA_UNROLL
        for (int k = 0; k < RA / 2; ++k) {     // (dst, IMOFF and 2k are even: 8-byte aligned pairs)
            *reinterpret_cast<float2*>(dst + 2 * k) = pre[k];
            *reinterpret_cast<float2*>(dst + IMOFF + 2 * k) = pim[k];
        }
    };

    for (long long item = blockIdx.x; item < p.total_items; item += gridDim.x) {
        const long long witem = A2SB_INV_REV ? p.total_items - 1 - item : item;
        const int b = (int)(witem / p.chunks_per_clip);
        const long long cb = p.hop_begin + (long long)(witem % p.chunks_per_clip) * p.chunk_hops;
        const long long ce = (cb + p.chunk_hops < p.hop_end) ? cb + p.chunk_hops : p.hop_end;
        const long long tfirst = cb - (ROV - 1);
        const int ntiles = (int)((ce - tfirst + kF - 1) / kF);
        const float* clip = p.spec + (long long)b * C * plane;
        float* clip_out = p.out + (long long)b * p.out_stride;
        float* carry_cur = s_carry0;
        float* carry_nxt = s_carry1;
        for (int i = tid; i < NC; i += NT) carry_cur[i] = 0.0f;
        // (visibility of the zeroed carry is covered by the barriers inside the tile loop)

        for (int tile = 0; tile < ntiles; ++tile, ++tile_count) {
            const long long t0 = tfirst + (long long)tile * kF;
            // ================= pass A: load + expand + split + radix-RA =====================
            if constexpr (RING) {
                // Pass A as a queue of warp-sized jobs, taken in a fixed order by whichever warp is free:
                //   E(bx): box bx has landed -> power expansion + phase normalisation of its RA x F bins, written as
                //          complex values into the exchange at the place the transformed residue will occupy; the slot is
                //          free again after ~RA LDS per lane (not after a whole residue's transform) and is refilled at once;
                //   X(c):  both boxes of class c are expanded -> pairing, radix-RA transform over q, twiddle, in place.
                // Order: E(chunk 0), E(chunk 1), X(chunk 0), E(chunk 2), X(chunk 1), ... (chunks of `slots` boxes): every job
                // waits only on jobs that were handed out before it, so the queue cannot deadlock.
                const unsigned par = tile_count & 1u;
                int nb = b; long long nt0 = t0 + kF;
                bool have_next = tile + 1 < ntiles;
                if (!have_next && item + gridDim.x < p.total_items) { tile_origin(item + gridDim.x, 0, nb, nt0); have_next = true; }
                constexpr int NJ = RB + RB / 2;
                const int blk = 3 * slots / 2, nchunk = RB / slots;
                const int ft = lane % kF, hl = lane / kF;
                const long long tg = t0 + ft;
                const bool valid = tg >= 0 && tg < T;
                for (;;) {
                    int job = 0;
                    if (lane == 0) job = atomicAdd(s_job + par, 1);
                    job = __shfl_sync(0xffffffffu, job, 0);
                    A2SB_PROF(0);
                    if (job >= NJ) break;
                    int ebox = -1, xcls = -1;
                    if (job < slots) ebox = job;
                    else {
                        const int j1 = job - slots, u = j1 / blk, v = j1 - u * blk;
                        if (u < nchunk - 1) { if (v < slots) ebox = (u + 1) * slots + v; else xcls = u * (slots / 2) + (v - slots); }
                        else xcls = (nchunk - 1) * (slots / 2) + (j1 - (nchunk - 1) * blk);
                    }
                    if (ebox >= 0) {
                        const int bx = ebox, cc = bx >> 1, hh = bx & 1;
                        const int jj = (cc == 0) ? (hh ? RB / 2 : 0) : (hh ? RB - cc : cc);
                        const int rho = (jj + RB - 1) % RB;
                        const int e0 = (int)(t0 - p.spec_t_first) + maps.shift[rho & 3];
                        float dcm = 0.0f;    // row 0 of the frame: only its NaN / Inf matter (SpectrogramAddDCTerm, transforms.py:227)
                        if (jj == 0 && hl == 0 && valid) dcm = ld_spec(clip + (tg - p.spec_t_first));
                        mbar_wait(s_bar + bx, par);
                        A2SB_PROF(1);
                        // lane = (frame ft, half hl); the halves take rows 8m + 4 hl + {0..3}: 4 rows = 80 floats = 16 banks apart
                        const float* sb = s_ring + (size_t)(bx % slots) * G::BOX + (e0 & 3) + ft;
                        float* dst = s_x + ft * FS + G::cblk(cc, hh);
                        if (jj != 0) {
                            unsigned minbits = 0x7f800000u;
                            A2SB_PRAGMA_UNROLL
                            for (int i = 0; i < RA / 4; ++i) {
                                const int qq = 8 * (i >> 1) + 4 * hl + 2 * (i & 1);      // rows qq, qq + 1 = bins jj + RB * {qq, qq + 1}
                                float2 m, cs, sn, vr, vi;
                                m.x = sb[(0 * RA + qq) * G::FW]; cs.x = sb[(1 * RA + qq) * G::FW]; sn.x = sb[(2 * RA + qq) * G::FW];
                                m.y = sb[(0 * RA + qq + 1) * G::FW]; cs.y = sb[(1 * RA + qq + 1) * G::FW]; sn.y = sb[(2 * RA + qq + 1) * G::FW];
                                inv_expand_fast2(m, cs, sn, p.eps, minbits, vr, vi);
                                if (!valid) { vr = make_float2(0.0f, 0.0f); vi = vr; }   // outside [0, T): zero fill / the neighbouring row
                                *reinterpret_cast<float2*>(dst + qq) = vr;
                                *reinterpret_cast<float2*>(dst + IMOFF + qq) = vi;
                            }
                            if (valid && minbits < 0x0da24260u /* 1e-30f */) {
                                // a degenerate (cos, sin) pair among this lane's bins: redo them with the careful expansion
                                for (int i = 0; i < RA / 2; ++i) {
                                    const int qq = 8 * (i >> 2) + 4 * hl + (i & 3);
                                    float vr, vi;
                                    inv_expand(p, sb[(0 * RA + qq) * G::FW], sb[(1 * RA + qq) * G::FW], sb[(2 * RA + qq) * G::FW], vr, vi);
                                    dst[qq] = vr; dst[IMOFF + qq] = vi;
                                }
                            }
                        } else {
                            // residue 0: box row qq holds bin RB * (qq + 1) -> position qq + 1.  Position 0 is the DC bin the chain
                            // re-creates; the last row is the Nyquist bin M, whose real part travels in the imaginary slot of
                            // position 0 (irfft ignores Im X[0] and Im X[M]).  One box per tile: careful scalar expansion.
                            for (int i = 0; i < RA / 2; ++i) {
                                const int qq = 8 * (i >> 2) + 4 * hl + (i & 3);
                                float vr = 0.0f, vi = 0.0f;
                                if (valid) inv_expand(p, sb[(0 * RA + qq) * G::FW], sb[(1 * RA + qq) * G::FW], sb[(2 * RA + qq) * G::FW], vr, vi);
                                if (qq + 1 < RA) { dst[qq + 1] = vr; dst[IMOFF + qq + 1] = vi; }
                                else dst[IMOFF] = vr;
                            }
                            if (hl == 0) {
                                float m = dcm;
                                if (p.pmode == kPowFour) m = m * power_scale_factor<kPowFour>(fabsf(m), p.power, p.eps);
                                else if (p.pmode == kPowGeneric) m = m * power_scale_factor<kPowGeneric>(fabsf(m), p.power, p.eps);
                                dst[0] = m * 0.0f;
                            }
                        }
                        __syncwarp();
                        if (lane == 0) {
                            mbar_arrive(s_exp + bx);
                            const int nx = bx + slots;       // RB % slots == 0: same slot
                            if (nx < RB) issue_box(b, t0, nx);
                            else if (have_next) issue_box(nb, nt0, nx - RB);
                        }
                        A2SB_PROF(2);
                    } else {
                        const int cc = xcls;
                        const int jj = (cc == 0) ? (hl ? RB / 2 : 0) : (hl ? RB - cc : cc);
                        mbar_wait(s_exp + 2 * cc, par);
                        mbar_wait(s_exp + 2 * cc + 1, par);
                        A2SB_PROF(3);
                        const float* src = s_x + ft * FS + G::cblk(cc, hl);
                        float xr[RA], xi[RA];
                        A2SB_PRAGMA_UNROLL
                        for (int k = 0; k < RA / 2; ++k) {
                            const float2 r2 = *reinterpret_cast<const float2*>(src + 2 * k);
                            const float2 i2 = *reinterpret_cast<const float2*>(src + IMOFF + 2 * k);
                            xr[2 * k] = r2.x; xr[2 * k + 1] = r2.y; xi[2 * k] = i2.x; xi[2 * k + 1] = i2.y;
                        }
                        pass_a_tail(xr, xi, valid, cc, hl, ft, jj, 0.0f, 1.0f, 0.0f, 0.0f);
                        A2SB_PROF(4);
                    }
                }
            } else {
                const long long tg = t0 + t;
                const bool valid = tg >= 0 && tg < T;
                const float* colp = clip + (tg - p.spec_t_first);
                float xr[RA], xi[RA];  // X[ja + RB*q]
                bool careful = !FAST;
                // The Nyquist row (and row 0, whose NaN/Inf must reach the re-created DC bin) is needed only by the
                // ja = 0 lanes of warp 0.  Issue those loads FIRST, ahead of the bulk loads: left where they are
                // consumed they cost warp 0 -- and, at the barrier, the whole CTA -- one more DRAM round trip per tile.
                float ny_a = 0.0f, ny_b = 1.0f, ny_c = 0.0f, dc_m = 0.0f;
                if (c == 0 && h == 0 && valid) {
                    const float* nrow = colp + (long long)(M + row_of_k0) * p.spec_T;
                    ny_a = ld_spec(nrow);
                    if (!cplx) {
                        ny_b = ld_spec(nrow + plane); ny_c = ld_spec(nrow + 2 * plane);
                        if (!p.has_dc) dc_m = ld_spec(colp);
                    }
                }
                if (FAST) {
                    if (valid) {
                        // rows k - 1 of the three planes; bin 0 (ja == 0, q == 0) has no row: SpectrogramAddDCTerm
                        unsigned minbits = 0x7f800000u;
                        unsigned long long a = reinterpret_cast<unsigned long long>(colp + (long long)(ja - 1) * p.spec_T);
                        // Loads are issued in batches of LB bin pairs (6*LB independent loads in flight per
                        // thread) before the first use: the pass is latency-bound, not bandwidth-bound.
                        constexpr int LB = (RA / 2 >= A2SB_INV_LB) ? A2SB_INV_LB : RA / 2;
                        A2SB_PRAGMA_UNROLL
                        for (int j0 = 0; j0 < RA / 2; j0 += LB) {
                            float2 m[LB], cc[LB], ss[LB];
                            A2SB_PRAGMA_UNROLL
                            for (int jj = 0; jj < LB; ++jj) {
                                const unsigned long long a1 = a + stepB;
                                if (j0 + jj == 0 && ja == 0) { m[jj].x = 0.0f; cc[jj].x = 1.0f; ss[jj].x = 0.0f; }
                                else {
                                    m[jj].x = ld_spec(reinterpret_cast<const float*>(a));
                                    cc[jj].x = ld_spec(reinterpret_cast<const float*>(a + planeB));
                                    ss[jj].x = ld_spec(reinterpret_cast<const float*>(a + plane2B));
                                }
                                m[jj].y = ld_spec(reinterpret_cast<const float*>(a1));
                                cc[jj].y = ld_spec(reinterpret_cast<const float*>(a1 + planeB));
                                ss[jj].y = ld_spec(reinterpret_cast<const float*>(a1 + plane2B));
                                a = a1 + stepB;
                            }
                            A2SB_PRAGMA_UNROLL
                            for (int jj = 0; jj < LB; ++jj) {
                                const int j = j0 + jj;
                                float2 vr, vi;
                                inv_expand_fast2(m[jj], cc[jj], ss[jj], p.eps, minbits, vr, vi);
                                xr[2 * j] = vr.x; xi[2 * j] = vi.x; xr[2 * j + 1] = vr.y; xi[2 * j + 1] = vi.y;
                            }
                        }
                        careful = minbits < 0x0da24260u /* 1e-30f */;
                    } else {
                        A2SB_PRAGMA_UNROLL
                        for (int q = 0; q < RA; ++q) { xr[q] = 0.0f; xi[q] = 0.0f; }
                    }
                    // Next tile of this CTA -- the next tile of this sweep, or the first tile of its next work
                    // item -- : pull its 64-byte row segments into L2 now (one row per lane), so pass A of that
                    // tile waits on L2 instead of DRAM.
                    {
                        long long nt0 = t0 + kF;
                        const float* nclip = clip;
                        bool pf = tile + 1 < ntiles;
#ifdef A2SB_INV_XPF   // experiment: also prefetch the first tile of the next work item (measured slower)
                        if (!pf) {
                            const long long nitem = item + gridDim.x;
                            if (nitem < p.total_items) {
                                nclip = p.spec + (nitem / p.chunks_per_clip) * C * plane;
                                nt0 = p.hop_begin + (nitem % p.chunks_per_clip) * p.chunk_hops - (ROV - 1);
                                pf = true;
                            }
                        }
#endif
                        const long long c0 = nt0 < 0 ? 0 : nt0;
                        const long long c1 = nt0 + kF - 1 < T - 1 ? nt0 + kF - 1 : T - 1;
                        if (pf && c0 <= c1) {
                            const unsigned long long nb =
                                reinterpret_cast<unsigned long long>(nclip + (c0 - p.spec_t_first)) + (unsigned long long)(ja - 1) * rowB;
                            const unsigned last = (unsigned)(c1 - c0) * 4u;
                            A2SB_PRAGMA_UNROLL
                            for (int q = t; q < RA; q += kF) {
                                if (ja == 0 && q == 0) continue;
                                const unsigned long long r0 = nb + (unsigned)q * stepB;
                                A2SB_PRAGMA_UNROLL
                                for (int ch = 0; ch < 3; ++ch) {
#if A2SB_INV_PF >= 1
                                    prefetch_l2(reinterpret_cast<const void*>(r0 + ch * planeB));
#endif
#if A2SB_INV_PF >= 3
                                    prefetch_l2(reinterpret_cast<const void*>(r0 + ch * planeB + (last >> 1)));
#endif
#if A2SB_INV_PF >= 2
                                    prefetch_l2(reinterpret_cast<const void*>(r0 + ch * planeB + last));
#endif
                                }
                            }
                        }
                    }
                }
                if (careful) {
                    // careful expansion of every bin of this frame
                    A2SB_PRAGMA_UNROLL
                    for (int q = 0; q < RA; ++q) {
                        const int k = ja + RB * q;
                        const int row = k + row_of_k0;
                        float vr = 0.0f, vi = 0.0f;
                        if (valid) {
                            if (cplx) {
                                vr = ld_spec(colp + (long long)row * p.spec_T);
                                vi = ld_spec(colp + plane + (long long)row * p.spec_T);
                            } else if (row >= 0) {
                                inv_expand(p, ld_spec(colp + (long long)row * p.spec_T), ld_spec(colp + plane + (long long)row * p.spec_T),
                                           ld_spec(colp + 2 * plane + (long long)row * p.spec_T), vr, vi);
                            }
                        }
                        xr[q] = vr; xi[q] = vi;
                    }
                }
                pass_a_tail(xr, xi, valid, c, h, t, ja, ny_a, ny_b, ny_c, dc_m);
            }
            __syncthreads();  // exchange complete
            A2SB_PROF(5);
            if (RING && tid == 0) s_job[(tile_count & 1u) ^ 1u] = 0;   // the other tile parity's job counter (idle until the next tile)
            if constexpr (TMA == 2) {
                // HBM is idle from here to the end of the tile (pass B and the overlap-add only touch shared memory): pull the
                // NEXT tile's boxes into L2 now -- RB tensor-map prefetches, one per residue, issued by RB threads -- so that its
                // pass A loads hit L2 instead of waiting on DRAM.  (Per-lane prefetch.global.L2 and one bulk prefetch per row
                // segment were measured as losses in round 1: 3072 requests per tile; this is 32.)
                if (tid < RB) {
                    int nb = b; long long nt0 = t0 + kF;
                    bool have_next = tile + 1 < ntiles;
                    if (!have_next && item + gridDim.x < p.total_items) { tile_origin(item + gridDim.x, 0, nb, nt0); have_next = true; }
                    if (have_next) {
                        const int rho = tid;
                        const int e0 = (int)(nt0 - p.spec_t_first) + maps.shift[rho & 3];
                        tma_prefetch_box5(&maps.m[rho & 3], e0 & ~3, rho >> 2, 0, 0, nb);
                    }
                }
            }

            // ================= pass B: twiddle + radix-RB + synthesis window =================
            if constexpr (G::PAIR_B) {
                // RB = 64 (n_fft 4096 as 32 x 64): a radix-64 item would need 128 data registers (256 threads x 244 registers, 8
                // warps per SM -- the round-1 kernel).  Here a LANE PAIR (lanes l, l ^ 16) shares one item (frame f, output
                // residue jb): lane half pp transforms the residues ja = 2a + pp (decimation in time: a radix-32 each, the
                // existing packed routine), the odd half multiplies by W_64^q', and ONE exchange of 16 complex values per lane
                // through __shfl_xor gives each lane the operands of its 32 outputs: pp = 0 emits q'' = i, i + 32 (i < 16),
                // pp = 1 emits q'' = 16 + i, 48 + i.  512 threads x <= 128 registers: 16 warps per SM.
                constexpr int HB = RB / 2;
                const int pp = (lane >> 4) & 1;
                const int pair = (tid >> 5) * 16 + (lane & 15);
                const int f = pair / RA, jb = pair % RA;
                const float* src = s_x + f * FS + jb;
                float er[HB], ei[HB];
                {
                    float2 re[HB / 2], im[HB / 2];
                    A2SB_PRAGMA_UNROLL
                    for (int j = 0; j < HB / 2; ++j) {     // residues 2a + pp for a = 2j (.x) and 2j + 1 (.y)
                        const int o0 = pp ? G::blk(4 * j + 1) : G::blk(4 * j), o1 = pp ? G::blk(4 * j + 3) : G::blk(4 * j + 2);
                        re[j].x = src[o0]; re[j].y = src[o1];
                        im[j].x = src[IMOFF + o0]; im[j].y = src[IMOFF + o1];
                    }
                    __syncthreads();   // the frame buffer below aliases this frame's exchange region, which two warps read
                    fft2x_dit<HB, +1>(re, im, er, ei);      // E_pp[q'], q' = 0 .. 31
                }
                if (pp) {              // T[q'] = W_64^q' E_1[q'], in place
                    static_for<1, HB>([&](auto Q) {
                        constexpr int q = decltype(Q)::value;
                        const float c = kCos64(q), sn = kSin64(q);
                        const float tr = s_fma(er[q], c, -(ei[q] * sn));
                        ei[q] = s_fma(er[q], sn, ei[q] * c);
                        er[q] = tr;
                    });
                }
                float* fb = s_x + G::fbuf(f);
                const int qb = pp * (HB / 2);
                A2SB_PRAGMA_UNROLL
                for (int i = 0; i < HB / 2; ++i) {
                    // pp = 0 sends E_0[16 + i] and receives T[i]; pp = 1 sends T[i] and receives E_0[16 + i]
                    const float sr = pp ? er[i] : er[HB / 2 + i], si = pp ? ei[i] : ei[HB / 2 + i];
                    const float rr = __shfl_xor_sync(0xffffffffu, sr, 16), ri = __shfl_xor_sync(0xffffffffu, si, 16);
                    const float ar = pp ? rr : er[i], ai = pp ? ri : ei[i];                        // E_0[qb + i]
                    const float br = pp ? er[HB / 2 + i] : rr, bi = pp ? ei[HB / 2 + i] : ri;      // T[qb + i]
                    const int n0 = jb + RA * (qb + i), n1 = n0 + RA * HB;
                    float2 z0 = make_float2(ar + br, ai + bi), z1 = make_float2(ar - br, ai - bi);
                    if (!win_in_ola) {
                        const float2 w0 = G::WIN_SMEM ? *reinterpret_cast<const float2*>(s_win + 2 * n0) : __ldg(reinterpret_cast<const float2*>(p.window + 2 * n0));
                        const float2 w1 = G::WIN_SMEM ? *reinterpret_cast<const float2*>(s_win + 2 * n1) : __ldg(reinterpret_cast<const float2*>(p.window + 2 * n1));
                        z0.x *= w0.x; z0.y *= w0.y; z1.x *= w1.x; z1.y *= w1.y;
                    }
                    *reinterpret_cast<float2*>(fb + 2 * n0) = z0;
                    *reinterpret_cast<float2*>(fb + 2 * n1) = z1;
                }
            }
            A2SB_PRAGMA_UNROLL
            for (int u = 0; u < G::ITEMS_B; ++u) {
                const int it = tid + u * NT;
                const int f = it / RA, jb = it % RA;
                const float* src = s_x + f * FS + jb;
                float2 re[RB / 2], im[RB / 2];
                A2SB_PRAGMA_UNROLL
                for (int j = 0; j < RB / 2; ++j) {
                    re[j].x = src[G::blk(2 * j)];
                    re[j].y = src[G::blk(2 * j + 1)];
                    im[j].x = src[IMOFF + G::blk(2 * j)];
                    im[j].y = src[IMOFF + G::blk(2 * j + 1)];
                }
                // the frame buffer below aliases this frame's exchange region, which RA/32 warps read
                if (RA > 32) __syncthreads(); else __syncwarp();
                float zr[RB], zi[RB];
                fft2x_dit<RB, +1>(re, im, zr, zi);  // z[jb + RA*q] = x[2n] + i x[2n+1]
                float* fb = s_x + G::fbuf(f);
                if (win_in_ola) {
                    // hop = N/4: the synthesis window is applied by the overlap-add, from registers (see wv below)
                    A2SB_PRAGMA_UNROLL
                    for (int q = 0; q < RB; ++q) *reinterpret_cast<float2*>(fb + 2 * (jb + RA * q)) = make_float2(zr[q], zi[q]);
                } else {
                    A2SB_PRAGMA_UNROLL
                    for (int q = 0; q < RB; ++q) {
                        const int n = jb + RA * q;
                        const float2 w = G::WIN_SMEM ? *reinterpret_cast<const float2*>(s_win + 2 * n)
                                                     : __ldg(reinterpret_cast<const float2*>(p.window + 2 * n));
                        *reinterpret_cast<float2*>(fb + 2 * n) = make_float2(zr[q] * w.x, zi[q] * w.y);
                    }
                }
            }
            __syncthreads();  // all frames of the tile are in their frame buffers
            A2SB_PROF(6);

            // ================= overlap-add, envelope, trim, store ===========================
            // Frames outside [0, T) were transformed from zeros, so they add nothing.
            // HC = N/4 (every A2SB configuration): hop and overlap count are compile-time, the loops unroll.
            auto ola = [&](auto HC) {
                constexpr int Hc = decltype(HC)::value;
                const int H = Hc ? Hc : p.hop;
                const int ROV = N / H;
                const int NC = N - H;
                // one float4 column r of hop-block hb: carry + the ROV frames that cover it, envelope, store
                auto emit = [&](int hb, int r, float4 ie_int) {
                    const int j4 = hb * H + r;
                    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (j4 < NC) acc = *reinterpret_cast<const float4*>(carry_cur + j4);
                    // frame hb - m contributes its m-th hop-block; ascending frame order, so the fp32 sum does
                    // not depend on where the tile boundary (carry) falls
                    A2SB_PRAGMA_UNROLL
                    for (int m = ROV - 1; m >= 0; --m) {
                        if (m > hb) continue;
                        const float4 v = *reinterpret_cast<const float4*>(s_x + G::fbuf(hb - m) + m * H + r);
                        if (Hc) {   // window applied here: one fused multiply-add per sample, window columns in registers
                            const float4 w = wv[Hc ? m : 0];
                            acc.x = s_fma(v.x, w.x, acc.x); acc.y = s_fma(v.y, w.y, acc.y);
                            acc.z = s_fma(v.z, w.z, acc.z); acc.w = s_fma(v.w, w.w, acc.w);
                        } else {
                            acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
                        }
                    }
                    const long long hg = t0 + hb;  // global hop-block
                    if (hg < cb || hg >= ce) return;
                    // envelope: frames hg-(ROV-1)..hg clipped to [0, T)
                    float4 ie = ie_int;
                    if (hg - (ROV - 1) < 0 || hg >= T) {
                        float e0 = 0.f, e1 = 0.f, e2 = 0.f, e3 = 0.f;
                        for (int m = 0; m < ROV; ++m) {
                            const long long tt = hg - m;
                            if (tt < 0 || tt >= T) continue;
                            const float* w2 = p.wsq + m * H + r;
                            e0 += w2[0]; e1 += w2[1]; e2 += w2[2]; e3 += w2[3];
                        }
                        ie = make_float4(1.0f / e0, 1.0f / e1, 1.0f / e2, 1.0f / e3);
                    }
                    const float4 y = make_float4(acc.x * ie.x, acc.y * ie.y, acc.z * ie.z, acc.w * ie.w);
                    const long long o = hg * H + r - N / 2 - p.out_first;  // local trimmed sample index
                    if constexpr (MIR == 2) {
                        short* d16 = reinterpret_cast<short*>(p.out) + (long long)b * p.out_stride + o;
                        const int q0 = pcm16_from_float(y.x), q1 = pcm16_from_float(y.y), q2 = pcm16_from_float(y.z), q3 = pcm16_from_float(y.w);
                        if (o >= 0 && o + 3 < p.out_count && (reinterpret_cast<uintptr_t>(d16) & 7) == 0) {
                            const int2 pk = make_int2((q0 & 0xffff) | (q1 << 16), (q2 & 0xffff) | (q3 << 16));
#ifdef A2SB_EMU
                            *reinterpret_cast<int2*>(d16) = pk;
#else
                            __stcs(reinterpret_cast<int2*>(d16), pk);
#endif
                        } else {
                            const int v[4] = {q0, q1, q2, q3};
                            for (int e = 0; e < 4; ++e)
                                if (o + e >= 0 && o + e < p.out_count) d16[e] = (short)v[e];
                        }
                        return;
                    }
                    float* dst = clip_out + o;
                    if (o >= 0 && o + 3 < p.out_count && (reinterpret_cast<uintptr_t>(dst) & 15) == 0) {
                        if (!MIR || !p.mirror_mc) {
#ifdef A2SB_EMU
                            *reinterpret_cast<float4*>(dst) = y;
#else
                            __stcs(reinterpret_cast<float4*>(dst), y);
#endif
                        }
                        if constexpr (MIR == 1) {
                            const long long off = dst - p.out;
                            if (p.mirror_mc) st_multicast(p.mirror[0] + off, y);
                            else
                                for (int i = 0; i < p.n_mirror; ++i) *reinterpret_cast<float4*>(p.mirror[i] + off) = y;
                        }
                    } else {
                        const float v[4] = {y.x, y.y, y.z, y.w};
                        for (int e = 0; e < 4; ++e)
                            if (o + e >= 0 && o + e < p.out_count) {
                                if (!MIR || !p.mirror_mc) clip_out[o + e] = v[e];
                                if constexpr (MIR == 1) {
                                    const long long off = dst + e - p.out;
                                    if (p.mirror_mc) st_multicast(p.mirror[0] + off, v[e]);
                                    else
                                        for (int i = 0; i < p.n_mirror; ++i) p.mirror[i][off] = v[e];
                                }
                            }
                    }
                };
                if (Hc) {
                    // compile-time hop: NT*4 is a multiple of H, so a thread keeps its column r and steps over hop-blocks:
                    // kF / hstep iterations exactly (kF % hstep == 0 for every instantiation)
                    const int r = (tid * 4) % H;
                    const int hstep = (NT * 4) / H;
                    const int hb0 = (tid * 4) / H;   // < hstep
                    const float4 ie_int = __ldg(reinterpret_cast<const float4*>(p.inv_env + r));
                    A2SB_PRAGMA_UNROLL
                    for (int it = 0; it < kF / hstep; ++it) emit(hb0 + it * hstep, r, ie_int);
                } else {
                    // run-time hop (any multiple of 4 that divides N, including H > NT*4): walk the tile's kF*H samples
                    for (int j = tid * 4; j < kF * H; j += NT * 4) {
                        const int hb = j / H, r = j - hb * H;
                        emit(hb, r, __ldg(reinterpret_cast<const float4*>(p.inv_env + r)));
                    }
                }
                // new carry: positions kF*H + j, j in [0, NC): hop-blocks kF .. kF + ROV - 2 of the tile's frames
                // (frames hb - d, d descending == ascending frame order)
                for (int j4 = tid * 4; j4 < NC; j4 += NT * 4) {
                    const int hb = kF + j4 / H, r = j4 % H;
                    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
                    A2SB_PRAGMA_UNROLL
                    for (int d = ROV - 1; d >= 1; --d) {
                        const int f = hb - d;
                        if (f < 0 || f >= kF) continue;
                        const float4 v = *reinterpret_cast<const float4*>(s_x + G::fbuf(f) + d * H + r);
                        if (Hc) {
                            const float4 w = wv[Hc ? d : 0];
                            acc.x = s_fma(v.x, w.x, acc.x); acc.y = s_fma(v.y, w.y, acc.y);
                            acc.z = s_fma(v.z, w.z, acc.z); acc.w = s_fma(v.w, w.w, acc.w);
                        } else {
                            acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
                        }
                    }
                    if (kF * H + j4 < NC) {   // only when the tile is shorter than the overlap (tiny n_fft / hop ratios)
                        const float4 v = *reinterpret_cast<const float4*>(carry_cur + kF * H + j4);
                        acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
                    }
                    *reinterpret_cast<float4*>(carry_nxt + j4) = acc;
                }
            };
            if (H == N / 4) ola(std::integral_constant<int, N / 4>{});
            else ola(std::integral_constant<int, 0>{});
            __syncthreads();  // frame buffers and carry_cur consumed; carry_nxt complete
            A2SB_PROF(7);
            float* tmp = carry_cur; carry_cur = carry_nxt; carry_nxt = tmp;
        }
    }
    A2SB_PROF_FLUSH(p);
}

}  // namespace a2sb
