// stft_fwd.cuh -- K1: fused framing + window + real FFT + mag/phase/power-compress epilogue.
//
// Replaces, in one pass over HBM, the reference's forward chain
//   ComplexSpectrogram -> ComplexToMagInstPhase -> SpectrogramDropDCTerm -> PowerScaleSpectrogram
//   (A2SB/audio_transforms/transforms.py:83-118,187-219; torch.stft via torchaudio Spectrogram).
//
// Geometry (n_fft = N = 2M real samples -> M-point complex FFT of z[n] = x[2n] + i x[2n+1],
// M = RA * RB, two register-resident passes):
//   * a CTA holds GROUPS = 16/F independent thread groups; a group sweeps RUNS of consecutive
//     F-frame tiles of one clip in time order (so partially written 32-byte sectors at tile seams
//     are completed in L2 a few microseconds later by the same group);
//   * the tile's input span ((F-1)*hop + N contiguous samples; frames overlap) is brought into
//     shared memory by ONE 1-D TMA bulk copy (cp.async.bulk + mbarrier), issued a tile ahead; clip
//     edges (reflect padding) use an index-mapped loader instead;
//   * pass A (frame-major threads): window folded into the first (scalar, decimation-in-frequency)
//     radix-2 stage, then a packed (FFMA2/FADD2) radix-RA/2 DFT of the two half-sequences;
//   * exchange through shared memory in a layout that is conflict-free on both sides;
//   * pass B (frame-minor threads: lanes run along the F frames of the tile for a residue pair
//     {j, RA-j}): packed twiddle multiply, packed radix-RB/2 + scalar last stage -> Z[jb + RA*q];
//     the real-FFT split needs Z[M-k], which lives in the partner lane group -> one __shfl_xor
//     per value; bins k and M-k are then carried in the two halves of packed registers through
//     magnitude / unit phasor / power compression, and stored with lanes along the frame axis
//     (the output's fastest axis);
//   * bins whose squared magnitude underflows fp32 (digital silence) are detected with one
//     3-input integer min per bin pair and the warp re-emits its rows through a careful scalar
//     path (same path serves the non-shipped output kinds).
// HBM traffic per tile = span read once + 3*M*F floats written once (algorithmic minimum).
//
// Round-2 geometries (FwdGeom<M, RA, RB, F, ROUNDS, WIDE>):
//   F = 32           n_fft 512 / 1024: 32-frame tiles, one 512-thread CTA per SM;
//   WIDE             ... whose pass-B warps hold ONE residue x 32 frames (128-byte row-segment stores; the partner residue's
//                    half-spectrum comes through the class's exchange region instead of a shuffle);
//   ROUNDS = 2       a tile is transformed in two rounds over the same input span, each for the pass-A outputs / residue
//                    classes of one parity: n_fft 4096 gets 16-frame tiles (64 x 32, 512 threads), n_fft 2048 an opt-in
//                    32-frame variant;
//   PCM              the span is 16-bit PCM, decoded by pass A's loads.
#pragma once
#include "a2sb_common.cuh"
#include "fftx2.cuh"
#include "tma.cuh"

namespace a2sb {

struct FwdParams {
    const float* wav;        // [batch][wav_stride] local sample buffers (PCM kernels: int16 samples behind the same pointer,
                             // strides and counts in SAMPLES; `window` then carries the 1/32768 of the decode)
    long long wav_stride;
    long long sample_first;  // global sample index of wav[b][0] (non-zero only when sharded)
    long long n_local;       // samples available in each local buffer
    long long len;           // global clip length L (reflect padding is about 0 and L-1)
    long long t_begin, t_end;  // global frame range computed by this launch
    float* out;              // [batch][C][rows][out_T]
    long long out_T;         // frames per output row
    long long out_t_first;   // global frame index of output column 0
    int batch;
    int hop;
    int tiles_per_clip;      // ceil((t_end - t_begin) / F)
    int run;                 // consecutive tiles per work item
    int items_per_clip;      // ceil(tiles_per_clip / run)
    long long total_items;
    const float* window;     // [N], already multiplied by 0.5 (real-FFT split scale)
    const float4* tw4;       // [RA][RB/2 + 1] pass-B twiddles: (cos q0, cos q1, sin q0, sin q1)(-2 pi jb q / M)
    const float4* tw4_alt;   // n_fft = 4096: the table of the two-round kernel's decomposition (RA = 64, RB = 32), else null
    const void* twS;         // [M/2 + 1] split table: float4 (c, -c, -s, s), or float2 (c, s) when M >= 2048 (shared
                             // memory is short there); (c, s) = (cos, sin)(2 pi k / N)
    long long wrap_at;       // fused wrap padding: output column that follows the last frame ( = T); 0 with wrap_cols = 0
    int wrap_cols;           // number of head frames replicated at columns wrap_at .. wrap_at + wrap_cols - 1
    int cluster_barriers;    // > 0 (experiment, cluster launches only): every CTA executes exactly this many cluster barriers, one
                             // per round, so that the CTAs of a cluster -- neighbouring tiles -- stay in lockstep
    int pcm;                 // 1: wav holds int16 PCM samples (PCM kernels)
    // Corruption epilogue (the masks of A2SB/corruption/corruptions.py applied where the spectrogram is
    // produced): out2 receives x * (1 - mask) + mask * noise * level for the rectangle mask rows [m_row0, m_row1) x
    // columns [m_col0, m_col1) of every [rows][T] slice (tensor coordinates); the clean spectrogram still goes to `out`.
    float* out2;             // same geometry as out; null = no corruption
    const float* noise;      // [batch][C][rows][noise_T]: torch.randn_like(spec) of the caller's generator
    long long noise_T;
    long long m_row0, m_row1, m_col0, m_col1;
    float m_level;
    int seam;                // seam-sector prefetch of the next tile: 1 head, 2 tail of every tile (else: last tile of a clip only), 4 inner
    int epi;                 // kEpiComplex / kEpiMagPhase
    int drop_dc;             // 1: rows are bins 1..M (SpectrogramDropDCTerm), 0: bins 0..M
    int pmode;               // kPowNone / kPowQuarter / kPowGeneric
    float power, eps;
};

enum : int { kEpiComplex = 0, kEpiMagPhase = 1 };

// ROUNDS = 2 (n_fft = 2048: the exchange of 32 frames does not fit shared memory): a 32-frame tile is transformed in two
// rounds over the SAME input span.  Round r runs pass A only for the outputs jb of parity r -- after the first
// decimation-in-frequency stage these are a 16-point DFT of the sums (r = 0) or of the twiddled differences (r = 1), i.e.
// half of pass A's arithmetic, nothing computed twice except the window products -- and pass B for the residue classes
// {jb, RA - jb} of that parity (both members of a class have the same parity).  The exchange holds half the classes of 32
// frames = what all classes of 16 frames cost, and every row is written as two adjacent 64-byte segments at the same
// time: half as many partially written seam atoms per byte.
// WIDE = 1 (F = 32): a pass-B warp holds ONE residue for all 32 frames of the tile (lanes along frames), so every store
// instruction writes one 128-byte row segment and the seam in the middle of the tile disappears (two half-warp stores of
// adjacent 64-byte segments each write partial sectors, and a partially written sector costs a DRAM fill whether or not
// its other half arrives a moment later -- ncu: 2.44 GB read either way).  The partner residue RA - jb lives in the
// neighbouring warp; the real-FFT split gets its Z[M - k] through the class's (already consumed) exchange region
// instead of a shuffle: the same number of shared-memory-pipe instructions, plus two 64-thread named barriers.
template <int M, int RA, int RB, int F, int ROUNDS = 1, int WIDE = 0>
struct FwdGeom {
    static constexpr int N = 2 * M;
    static constexpr int NTG = F * RA / ROUNDS;   // threads per group: one pass-B item per thread (and round)
    static constexpr int FL = WIDE ? 32 : (F < 16) ? F : 16;  // frames per lane group of pass B (half-warp; WIDE: the warp)
    static constexpr int FB = F / FL;             // F > 16: a residue class is spread over FB warps, one per 16-frame block, whose
                                                  // 64-byte row segments are adjacent and stored at the same time (half the seams)
    static constexpr int GROUPS = (M >= 2048 || F > 16) ? 1 : 16 / F;   // independent groups per CTA
    static constexpr bool COMPACT_TWS = (M >= 2048);
    static constexpr size_t TWS_ELEM = COMPACT_TWS ? sizeof(float2) : sizeof(float4);
    static constexpr int NT = NTG * GROUPS;
    static constexpr int ITEMS_A = F * RB / NTG;  // pass-A items per thread (and round)
    static constexpr int CLS = RA / 2;            // residue classes {j, RA-j}
    static constexpr int CLSR = CLS / ROUNDS;     // classes per round
    static constexpr int CPW = WIDE ? 1 : 32 / (2 * FL);   // classes per warp (WIDE: a class takes two warps)
    static constexpr bool SHARED_CLS = WIDE || FB > 1;     // a class's exchange region is read by more than one warp
    static constexpr int QS = 2 * F + 1;          // exchange q-stride (odd: conflict-free pass-A writes)
    static constexpr int CS0 = RB * QS;
    // class stride; when two classes share a warp their lane groups must sit 16 banks apart
    static constexpr int CS = (CPW == 1) ? CS0 : CS0 + ((16 - CS0 % 32) + 32) % 32;
    static constexpr int XPLANE = CLSR * CS;      // floats per exchange plane
    static constexpr int TWS = RB / 2 + 1;        // float4 row stride of the pass-B twiddle table
    static_assert(M == RA * RB && (F * RB) % NTG == 0, "two-pass decomposition");
    static_assert(ROUNDS == 1 || (ROUNDS == 2 && CPW == 1 && ((ITEMS_A == 2 && RA == 32) || (ITEMS_A == 1 && RA == 64))),
                  "two rounds: RA = 32 -> packed radix-16 over frames f, f + F/2; RA = 64 -> one radix-32 item per thread");
    // n_fft = 4096 in two rounds: exchange (135 KB) + input span (78 KB) leave no room for the window and the pass-B twiddles
    // (33 KB): they are read through L1 (the forward kernel has no other use for it)
    static constexpr bool TABLES_SMEM = !(M >= 2048 && ROUNDS > 1);
    static_assert(F == 8 || F == 16 || F == 32 || F == 64, "tile width");
    static_assert(!WIDE || (F == 32 && M < 2048), "wide pass B: 32 lanes along 32 frames");
    static_assert(NTG % 32 == 0 && (NTG / 32) * CPW == CLSR * FB * (WIDE ? 2 : 1), "thread mapping");
    static_assert(SHARED_CLS || CPW * CS >= 32 * RB, "a warp's exchange region must hold its spectra (careful path)");
    static_assert(!SHARED_CLS || (NTG / 32) * 32 * RB <= XPLANE, "careful path: one private parking region per warp");
    static_assert(!WIDE || RB * 32 <= CS, "wide pass B: the class region holds both warps' upper half-spectra");
    // slot of (frame f, half hh) inside the 2F-word record of (class, q): 16-frame blocks, the two halves of a block adjacent,
    // so that the 32 lanes of a pass-B warp (block, both halves) read 32 consecutive words
    A2SB_HD static constexpr int xslot(int f, int hh) { return WIDE ? hh * F + f : (f / FL) * 2 * FL + hh * FL + f % FL; }
    // shared memory carve-up (bytes)
    static constexpr size_t off_win = 0;
    static constexpr size_t off_tw4 = off_win + (TABLES_SMEM ? sizeof(float) * N : 0);
    static constexpr size_t off_twS = off_tw4 + (TABLES_SMEM ? sizeof(float4) * RA * TWS : 0);
    static constexpr size_t off_grp = ((off_twS + TWS_ELEM * (M / 2 + 1) + 15) / 16) * 16;
    // per group: mbarrier (16 B), two exchange planes, input span
    static constexpr size_t g_xre = 16;
    static constexpr size_t g_xim = g_xre + sizeof(float) * XPLANE;
    static constexpr size_t g_in = ((g_xim + sizeof(float) * XPLANE + 127) / 128) * 128;
    A2SB_HD static size_t group_bytes(int hop) { return ((g_in + sizeof(float) * ((size_t)(F - 1) * hop + N) + 127) / 128) * 128; }
    static size_t smem_bytes(int hop) { return ((off_grp + 127) / 128) * 128 + GROUPS * group_bytes(hop); }
};

// Two-round pass A: branch BR of the windowed first decimation-in-frequency stage (BR = 0: u = a + b, BR = 1:
// v = (a - b) W_R^Q) for TWO items (the same residue of two frames, hence the same window values), results packed as
// (item 0, item 1).  Same operations on the same operands as dif_first_windowed: bit-identical values.
template <int R, int DIR, int Q, int BR>
A2SB_DEV void dif_first_windowed_branch(float2 xa0, float2 xa1, float2 xb0, float2 xb1, float2 wa, float2 wb, float2& re, float2& im) {
    const float tr0 = xb0.x * wb.x, ti0 = xb0.y * wb.y, tr1 = xb1.x * wb.x, ti1 = xb1.y * wb.y;
    if (BR == 0) {
        re = make_float2(s_fma(xa0.x, wa.x, tr0), s_fma(xa1.x, wa.x, tr1));
        im = make_float2(s_fma(xa0.y, wa.y, ti0), s_fma(xa1.y, wa.y, ti1));
        return;
    }
    const float2 dr = make_float2(s_fma(xa0.x, wa.x, -tr0), s_fma(xa1.x, wa.x, -tr1));
    const float2 di = make_float2(s_fma(xa0.y, wa.y, -ti0), s_fma(xa1.y, wa.y, -ti1));
    constexpr int tw = (Q * (64 / R)) & 63;
    if (tw == 0) { re = dr; im = di; }
    else if (tw == 16) { if (DIR < 0) { re = di; im = p2_neg(dr); } else { re = p2_neg(di); im = dr; } }
    else {
        const float c = kCos64(tw), sn = (DIR < 0) ? -kSin64(tw) : kSin64(tw);
        re = p2_fma(dr, p2_bc(c), p2_neg(p2_mul(di, p2_bc(sn))));
        im = p2_fma(dr, p2_bc(sn), p2_mul(di, p2_bc(c)));
    }
}

// Same for ONE item, results as scalars (RA = 64: the branch is followed by a radix-32 transform of that item).
template <int R, int DIR, int Q, int BR>
A2SB_DEV void dif_first_windowed_branch1(float2 xa, float2 xb, float2 wa, float2 wb, float& re, float& im) {
    const float tr = xb.x * wb.x, ti = xb.y * wb.y;
    if (BR == 0) { re = s_fma(xa.x, wa.x, tr); im = s_fma(xa.y, wa.y, ti); return; }
    const float dr = s_fma(xa.x, wa.x, -tr), di = s_fma(xa.y, wa.y, -ti);
    constexpr int tw = (Q * (64 / R)) & 63;
    if (tw == 0) { re = dr; im = di; }
    else if (tw == 16) { if (DIR < 0) { re = di; im = -dr; } else { re = -di; im = dr; } }
    else {
        const float c = kCos64(tw), sn = (DIR < 0) ? -kSin64(tw) : kSin64(tw);
        re = s_fma(dr, c, -(di * sn));
        im = s_fma(dr, sn, di * c);
    }
}

A2SB_DEV void st_stream(float* p, float v) {
#ifdef A2SB_EMU
    *p = v;
#else
#ifdef A2SB_PLAIN_STORES
    *p = v;
#else
    __stcs(p, v);  // streaming store: the spectrogram is not re-read by this kernel
#endif
#endif
}

// x * (1 - m) + m * noise * level in the reference's fp32 operation order (corruptions.py:14-15), never contracted.
A2SB_DEV float corrupt_mix(float x, float m, float nz, float level) {
#ifdef A2SB_EMU
    volatile float a = 1.0f - m, b = x * a, c = m * nz, d = c * level, e = b + d;
    return e;
#else
    return __fadd_rn(__fmul_rn(x, __fadd_rn(1.0f, -m)), __fmul_rn(__fmul_rn(m, nz), level));
#endif
}
// Corrupted counterpart of one stored value.  `o` is the element's address in `out`, d2 the byte distance out2 - out, nz the
// address of its noise sample.  Outside the mask the reference's expression returns x + (+-0): the noise sample is only
// needed when x itself is a zero (its sign then decides the sign of the result), so unmasked elements cost no load.
A2SB_DEV void st_corrupt(unsigned long long o, long long d2, const float* nz, bool inside, float nv, float level, float v) {
    // nv: the noise sample, already loaded when `inside`
    float r;
    if (inside) r = corrupt_mix(v, 1.0f, nv, level);
    else if (v == 0.0f) r = corrupt_mix(v, 0.0f, *nz, level);
    else r = v;     // == corrupt_mix(v, 0, finite, level) for v != 0
    st_stream(reinterpret_cast<float*>(o + d2), r);
}

// Careful single-bin emission (any output kind, any power mode; exact handling of |X| -> 0).
A2SB_DEV void fwd_emit(const FwdParams& p, float* __restrict__ clip_out, long long plane, int k, long long col,
                       float xr, float xi, const float* nclip = nullptr, long long nplane = 0) {
    if (p.epi == kEpiComplex) {
        st_stream(clip_out + (long long)k * p.out_T + col, xr);
        st_stream(clip_out + plane + (long long)k * p.out_T + col, xi);
        return;
    }
    const int row = k - p.drop_dc;
    if (row < 0) return;
    // ComplexToMagInstPhase (transforms.py:116-118): mag = sqrt(re^2+im^2); (cos, sin)(atan2(im, re)).
    // Normal bins use exactly the operation sequence of the packed fast path (bit-identical results).
    const float m2 = s_fma(xr, xr, xi * xi);
    float mag, cs, sn;
    if (m2 >= 1e-30f) {
        const float rs = rsqrt_approx(m2);
        mag = m2 * rs; cs = xr * rs; sn = xi * rs;
        // PowerScaleSpectrogram on channel 0 (transforms.py:199-206): m * (|m|^p / (|m| + eps)).
        if (p.pmode == kPowQuarter) {
            const float t = rsqrt_approx(rs), q = rsqrt_approx(t);
            mag = (t * q) * rcp_approx(s_fma(p.eps, rs, 1.0f));
        } else if (p.pmode == kPowGeneric) {
            mag = mag * power_scale_factor<kPowGeneric>(mag, p.power, p.eps);
        }
    } else {
        // |X| below ~1e-15: the square underflows; rescale so the phase stays meaningful,
        // and let the magnitude underflow exactly like the reference's sqrt(re^2+im^2).
        const float xs = xr * 1.8446744e19f, ys = xi * 1.8446744e19f;  // 2^64
        const float m2s = xs * xs + ys * ys;
        mag = sqrtf(m2);
        if (m2s > 0.0f) {
            const float rs = rsqrt_approx(m2s);
            cs = xs * rs; sn = ys * rs;
        } else {  // atan2(0, 0) = 0 -> (cos, sin) = (1, 0)
            cs = 1.0f; sn = 0.0f;
        }
        if (p.pmode == kPowQuarter) mag = mag * power_scale_factor<kPowQuarter>(mag, p.power, p.eps);
        else if (p.pmode == kPowGeneric) mag = mag * power_scale_factor<kPowGeneric>(mag, p.power, p.eps);
    }
    float* o = clip_out + (long long)row * p.out_T + col;
    st_stream(o, mag);
    st_stream(o + plane, cs);
    st_stream(o + 2 * plane, sn);
    if (p.out2 && nclip) {   // corruption epilogue (careful path: rare)
        const long long d2 = (long long)(reinterpret_cast<const char*>(p.out2) - reinterpret_cast<const char*>(p.out));
        const bool in = row >= p.m_row0 && row < p.m_row1 && col >= p.m_col0 && col < p.m_col1;
        const float* nz = nclip + (long long)row * p.noise_T + col;
        st_corrupt(reinterpret_cast<unsigned long long>(o), d2, nz, in, in ? nz[0] : 0.0f, p.m_level, mag);
        st_corrupt(reinterpret_cast<unsigned long long>(o + plane), d2, nz + nplane, in, in ? nz[nplane] : 0.0f, p.m_level, cs);
        st_corrupt(reinterpret_cast<unsigned long long>(o + 2 * plane), d2, nz + 2 * nplane, in, in ? nz[2 * nplane] : 0.0f, p.m_level, sn);
    }
}

// Real-FFT split of the pair (k, M-k) from Z[k] and Z[M-k] (0.5 folded into the window):
//   E = Zk + conj(Zm), O = Zk - conj(Zm), a = c Or + s Oi, b = c Oi - s Or,
//   X[k] = (Er + b, Ei - a),  X[M-k] = (Er - b, -(Ei + a)).
// Returns the two bins in the halves of packed registers: xr = (Re X[k], Re X[M-k]), xi likewise.
A2SB_DEV void fwd_split(float zkr, float zki, float zmr, float zmi, float4 w /* (c, -c, -s, s) */, float2& xr, float2& xi) {
    const float er = zkr + zmr, ei = zki - zmi;
    const float orr = zkr - zmr, oi = zki + zmi;
    const float2 bp = p2_fma(p2_bc(oi), make_float2(w.x, w.y), p2_mul(p2_bc(orr), make_float2(w.z, w.w)));  // (b, -b)
    const float na = s_fma(w.y, orr, w.z * oi);                                                         // -a
    xr = p2_add(p2_bc(er), bp);
    xi = make_float2(ei + na, na - ei);
}

A2SB_DEV void st_stream_at(unsigned long long a, float v) { st_stream(reinterpret_cast<float*>(a), v); }

A2SB_DEV unsigned min3u(unsigned a, unsigned b, unsigned c) {
#ifdef A2SB_EMU
    return a < b ? (a < c ? a : c) : (b < c ? b : c);
#else
    return __vimin3_u32(a, b, c);
#endif
}

// Fast emission of the bin pair (k, M-k) for the shipped chain (mag/cos/sin, power 0.25 on the
// magnitude): packed arithmetic over the two bins, MUFU for rsqrt/rcp.  `a_lo` / `a_hi` point at
// the channel-0 element of the two rows (byte addresses); channels 1 and 2 are planeB / plane2B bytes further.
template <int PMODE>
A2SB_DEV void fwd_pair_values(float2 xr, float2 xi, float eps, unsigned& minbits, float2& mag, float2& cs, float2& sn) {
    const float2 m2 = p2_fma(xr, xr, p2_mul(xi, xi));
    minbits = min3u(minbits, __float_as_uint(m2.x), __float_as_uint(m2.y));
    float2 rs;
    rs.x = rsqrt_approx(m2.x);
    rs.y = rsqrt_approx(m2.y);
    mag = p2_mul(m2, rs);
    cs = p2_mul(xr, rs);
    sn = p2_mul(xi, rs);
    if (PMODE == kPowQuarter) {
        // m * m^(1/4) / (m + eps) = m^(1/4) / (1 + eps / m), with 1/m = rs:
        float2 t, q, r;
        t.x = rsqrt_approx(rs.x); t.y = rsqrt_approx(rs.y);    // m^(1/2)
        q.x = rsqrt_approx(t.x); q.y = rsqrt_approx(t.y);      // m^(-1/4)
        const float2 d = p2_fma(p2_bc(eps), rs, p2_bc(1.0f));
        r.x = rcp_approx(d.x); r.y = rcp_approx(d.y);
        mag = p2_mul(p2_mul(t, q), r);
    }
}
template <int PMODE>
A2SB_DEV void fwd_emit_pair_fast(float2 xr, float2 xi, float eps, unsigned& minbits, bool valid, unsigned long long a_lo,
                                 unsigned long long a_hi, unsigned long long planeB, unsigned long long plane2B) {
    float2 mag, cs, sn;
    fwd_pair_values<PMODE>(xr, xi, eps, minbits, mag, cs, sn);
#ifdef A2SB_EXP_NOSTORE
    if (valid && mag.x == 123.456f) {
#else
    if (valid) {
#endif
        st_stream_at(a_lo, mag.x); st_stream_at(a_lo + planeB, cs.x); st_stream_at(a_lo + plane2B, sn.x);
        st_stream_at(a_hi, mag.y); st_stream_at(a_hi + planeB, cs.y); st_stream_at(a_hi + plane2B, sn.y);
    }
}

template <int I, int NN, class Fn>
A2SB_DEV void static_for(Fn&& f) {
    if constexpr (I < NN) {
        f(std::integral_constant<int, I>{});
        static_for<I + 1, NN>(f);
    }
}

A2SB_DEV void group_sync(int groups, int g, int nthreads) {
#ifdef A2SB_EMU
    if (groups == 1) __syncthreads(); else emu::named_barrier(g + 1, nthreads);
#else
    if (groups == 1) __syncthreads();
    else asm volatile("bar.sync %0, %1;" ::"r"(g + 1), "r"(nthreads) : "memory");
#endif
}

// timing-only barrier over the CTAs of a thread-block cluster (no data is exchanged)
A2SB_DEV void cluster_lockstep() {
#ifndef A2SB_EMU
    asm volatile("barrier.cluster.arrive.relaxed.aligned;\n\tbarrier.cluster.wait.aligned;" ::: "memory");
#endif
}

// barrier over the two warps of a residue class (WIDE pass B)
A2SB_DEV void pair_sync(int id) {
#ifdef A2SB_EMU
    emu::named_barrier(id, 64);
#else
    asm volatile("bar.sync %0, 64;" ::"r"(id) : "memory");
#endif
}

// FAST = 1 / 2: mag/phase output with power 0.25 (the shipped chain) / without power scaling through
// the packed fast path, the careful path being a rare fallback.  FAST = 0: every bin through the
// careful path (complex output, generic exponents).
// PCM = 1: the waveform is 16-bit PCM (what librosa.load / soundfile decode to float32 by an exact division by 32768,
// A2SB/datasets/datasets.py:231): the span travels and sits in shared memory as int16 (half the HBM / PCIe bytes of the
// input side), pass A converts on the fly, and the 2^-15 rides on the window table -- a power of two, so the spectrogram is
// bit-identical to the one computed from the decoded float32 samples.
template <int M, int RA, int RB, int F, int FAST, int ROUNDS = 1, int WIDE = 0, int PCM = 0>
__global__ void __launch_bounds__(FwdGeom<M, RA, RB, F, ROUNDS, WIDE>::NT, (FwdGeom<M, RA, RB, F, ROUNDS, WIDE>::NT <= 256 && M < 2048) ? 2 : 1)
stft_fwd_kernel(const FwdParams p) {
    using G = FwdGeom<M, RA, RB, F, ROUNDS, WIDE>;
    using in_t = std::conditional_t<PCM != 0, short, float>;
    constexpr int N = G::N, NT = G::NT, NTG = G::NTG, QS = G::QS, CS = G::CS, GROUPS = G::GROUPS;
    A2SB_DYN_SMEM(smem);
    const float* s_win = G::TABLES_SMEM ? reinterpret_cast<const float*>(smem + G::off_win) : p.window;
    const float4* s_tw4 = G::TABLES_SMEM ? reinterpret_cast<const float4*>(smem + G::off_tw4) : p.tw4;
    auto ld2 = [&](const float* q) -> float2 {   // window pair: shared memory, or global memory through L1
        if constexpr (G::TABLES_SMEM) return *reinterpret_cast<const float2*>(q);
        else return __ldg(reinterpret_cast<const float2*>(q));
    };
    unsigned char* s_twS_raw = smem + G::off_twS;
    // split-table entry k as (c, -c, -s, s)
    auto twS_at = [&](int k) -> float4 {
        if (G::COMPACT_TWS) {
            const float2 w = reinterpret_cast<const float2*>(s_twS_raw)[k];
            return make_float4(w.x, -w.x, -w.y, w.y);
        }
        return reinterpret_cast<const float4*>(s_twS_raw)[k];
    };

    const int tid = threadIdx.x;
    const int g = tid / NTG, gt = tid - g * NTG;
    const int H = p.hop;
    const int span = (F - 1) * H + N;
    unsigned char* gbase = smem + ((G::off_grp + 127) / 128) * 128 + (size_t)g * G::group_bytes(H);
    unsigned long long* s_bar = reinterpret_cast<unsigned long long*>(gbase);
    float* s_xre = reinterpret_cast<float*>(gbase + G::g_xre);
    float* s_xim = reinterpret_cast<float*>(gbase + G::g_xim);
    in_t* s_in = reinterpret_cast<in_t*>(gbase + G::g_in);
    // z[n] = x[2n] + i x[2n+1] at an even sample offset of the span
    auto ldz = [&](const in_t* q) -> float2 {
        if constexpr (PCM) {
            const unsigned v = *reinterpret_cast<const unsigned*>(q);
            return make_float2((float)(short)(v & 0xffffu), (float)(short)(v >> 16));
        } else {
            return *reinterpret_cast<const float2*>(q);
        }
    };

    const int C = (p.epi == kEpiComplex) ? 2 : 3;
    const int rows = (p.epi == kEpiComplex) ? M + 1 : (M + 1 - p.drop_dc);
    const long long plane = (long long)rows * p.out_T;

    // ---- tables -> shared memory; barrier init
    if constexpr (G::TABLES_SMEM) {
        float* w = reinterpret_cast<float*>(smem + G::off_win);
        float4* t4 = reinterpret_cast<float4*>(smem + G::off_tw4);
        for (int i = tid; i < N; i += NT) w[i] = p.window[i];
        for (int i = tid; i < RA * G::TWS; i += NT) t4[i] = p.tw4[i];
    }
    if (G::COMPACT_TWS) {
        for (int i = tid; i <= M / 2; i += NT) reinterpret_cast<float2*>(s_twS_raw)[i] = reinterpret_cast<const float2*>(p.twS)[i];
    } else {
        for (int i = tid; i <= M / 2; i += NT) reinterpret_cast<float4*>(s_twS_raw)[i] = reinterpret_cast<const float4*>(p.twS)[i];
    }
    if (gt == 0) { mbar_init(s_bar, 1); fence_mbar_init(); }
    __syncthreads();

    // Work items are (clip, run of p.run consecutive tiles); this group owns items gid, gid + gstride, ...
    const long long gid = (long long)blockIdx.x * GROUPS + g, gstride = (long long)gridDim.x * GROUPS;
    // slot s of this group -> (clip b, first global frame t0); false when the slot is empty / past the end
    auto locate = [&](long long s, int& b, long long& t0, bool& done) -> bool {
        const long long item = gid + (s / p.run) * gstride;
        done = item >= p.total_items;
        if (done) return false;
        b = (int)(item / p.items_per_clip);
        const long long tile = (item % p.items_per_clip) * p.run + (s % p.run);
        if (tile >= p.tiles_per_clip) return false;
        t0 = p.t_begin + tile * F;
        return true;
    };
    // Loads the input span of the tile into s_in.  Returns true if an asynchronous TMA copy was
    // issued (completion on s_bar), false if the span was filled synchronously by the group.
    auto load_span = [&](int b, long long t0) -> bool {
        const long long g0 = t0 * H - N / 2;  // global sample index of s_in[0]
        const in_t* clip = reinterpret_cast<const in_t*>(p.wav) + (long long)b * p.wav_stride;
        const long long l0 = g0 - p.sample_first;
        const in_t* src = clip + l0;
        const bool inside = g0 >= 0 && g0 + span <= p.len && l0 >= 0 && l0 + span <= p.n_local &&
                            ((reinterpret_cast<uintptr_t>(src) & 15) == 0) && ((span * sizeof(in_t)) % 16 == 0);
        if (inside) {
            if (gt == 0) {
                fence_proxy_async();
                bulk_g2s(s_in, src, (unsigned)(span * sizeof(in_t)), s_bar);
            }
            return true;
        }
        for (int i = gt; i < span; i += NTG) {
            long long gg = reflect_index(g0 + i, p.len);
            gg = gg < 0 ? 0 : (gg >= p.len ? p.len - 1 : gg);  // only reachable for masked frames
            long long l = gg - p.sample_first;
            l = l < 0 ? 0 : (l >= p.n_local ? p.n_local - 1 : l);
            s_in[i] = clip[l];
        }
        return false;
    };
    // first non-empty slot at or after s
    auto next_slot = [&](long long& s, int& b, long long& t0) -> bool {
        for (;; ++s) {
            bool done;
            if (locate(s, b, t0, done)) return true;
            if (done) return false;
        }
    };

    long long s = 0;
    int b = 0;
    long long t0 = 0;
    bool have = next_slot(s, b, t0);
    bool cur_async = false;
    unsigned phase = 0;
    if (have) cur_async = load_span(b, t0);
    group_sync(GROUPS, g, NTG);  // (synchronous) span visible

    // pass-B identity of this thread
    const int warp = gt >> 5, lane = gt & 31;
    constexpr int FL = G::FL, FB = G::FB;
    const int wslot = WIDE ? warp / 2 : warp / FB;  // class slot of the warp within a round
    const int h = WIDE ? (warp & 1) : (lane / FL) & 1, t = WIDE ? lane : (warp % FB) * FL + lane % FL;
    const int xb = (WIDE ? wslot : wslot * G::CPW + lane / (2 * FL)) * CS + G::xslot(t, h);

    int n_cluster_barriers = 0;
    while (have) {
        if (cur_async) { mbar_wait(s_bar, phase); phase ^= 1u; }
        const int cur_b = b;
        const long long cur_t0 = t0;
        bool next_async = false;
        if constexpr (ROUNDS > 1) {   // the next tile is located up front: every round prefetches its rows' seam sectors
            ++s;
            have = next_slot(s, b, t0);
        }
        // ---- rounds of the tile (ROUNDS == 1: one trip).  Rolled on purpose: pass B exists once in the instruction stream and
        // takes the round's class / residue as run-time values.  (The body keeps the indentation of the tile loop.)
#pragma unroll 1
      for (int rnd = 0; rnd < ROUNDS; ++rnd) {
        // class of this thread in this round (two rounds: classes of parity rnd) and its residue
        const int c = (ROUNDS > 1) ? 2 * wslot + rnd : WIDE ? wslot : wslot * G::CPW + lane / (2 * FL);
        const int wc = (ROUNDS > 1 || WIDE) ? c : wslot;   // warp-uniform; 0: the warp holds class 0 (F = 8: together with class 1)
        const int jb = (c == 0) ? (h ? RA / 2 : 0) : (h ? RA - c : c);

        // ================= pass A: window + radix-RA over q of z[ja + RB*q] =================
        if constexpr (ROUNDS > 1 && RA == 32) {
            // items (f0, ja) and (f0 + F/2, ja) in the halves of packed registers; outputs jb = 2k + rnd
            const int f0 = gt / RB, ja = gt % RB;
            const in_t* fin0 = s_in + f0 * H + 2 * ja;
            const in_t* fin1 = fin0 + (F / 2) * H;
            const float* win = s_win + 2 * ja;
            float2 re[RA / 2], im[RA / 2];
            auto first = [&](auto BR) {
                static_for<0, RA / 2>([&](auto Q) {
                    constexpr int q = decltype(Q)::value;
                    const float2 wa = ld2(win + 2 * RB * q);
                    const float2 wb = ld2(win + 2 * RB * (q + RA / 2));
                    dif_first_windowed_branch<RA, -1, q, decltype(BR)::value>(
                        ldz(fin0 + 2 * RB * q), ldz(fin1 + 2 * RB * q), ldz(fin0 + 2 * RB * (q + RA / 2)),
                        ldz(fin1 + 2 * RB * (q + RA / 2)), wa, wb, re[q], im[q]);
                });
            };
            if (rnd == 0) first(std::integral_constant<int, 0>{}); else first(std::integral_constant<int, 1>{});
            fft_v<RA / 2, -1, float2>(re, im);
            const int a0 = ja * QS + G::xslot(f0, 0), a1 = ja * QS + G::xslot(f0 + F / 2, 0);
            A2SB_PRAGMA_UNROLL
            for (int k = 0; k < RA / 2; ++k) {
                // residue j = 2k + rnd: class cc = min(j, RA - j) (0 for j = 0, RA/2), upper half when j >= RA/2; the
                // class slot of a round is cc >> 1.  Even round: k = 0 -> (0, 0), k = RA/4 -> (0, 1).
                const int je = 2 * k, jo = 2 * k + 1;
                const int cce = (je == 0 || je == RA / 2) ? 0 : (je < RA / 2 ? je : RA - je), hhe = (je == 0) ? 0 : (je >= RA / 2 ? 1 : 0);
                const int cco = (jo < RA / 2 ? jo : RA - jo), hho = (jo > RA / 2) ? 1 : 0;
                const int off = rnd ? (cco >> 1) * CS + hho * FL : (cce >> 1) * CS + hhe * FL;
                s_xre[a0 + off] = re[k].x; s_xre[a1 + off] = re[k].y;
                s_xim[a0 + off] = im[k].x; s_xim[a1 + off] = im[k].y;
            }
        } else if constexpr (ROUNDS > 1) {
            // RA = 64: one item (f, ja) per thread; the branch of the first stage leaves a 32-point transform, done as a
            // scalar decimation-in-frequency stage + packed radix-16 x 2.  Output j' of it is residue jb = 2 j' + rnd.
            const int f = gt / RB, ja = gt % RB;
            const in_t* fin = s_in + f * H + 2 * ja;
            const float* win = s_win + 2 * ja;
            float br_[RA / 2], bi_[RA / 2];
            auto first = [&](auto BR) {
                static_for<0, RA / 2>([&](auto Q) {
                    constexpr int q = decltype(Q)::value;
                    dif_first_windowed_branch1<RA, -1, q, decltype(BR)::value>(
                        ldz(fin + 2 * RB * q), ldz(fin + 2 * RB * (q + RA / 2)),
                        ld2(win + 2 * RB * q), ld2(win + 2 * RB * (q + RA / 2)), br_[q], bi_[q]);
                });
            };
            if (rnd == 0) first(std::integral_constant<int, 0>{}); else first(std::integral_constant<int, 1>{});
            float2 re[RA / 4], im[RA / 4];
            static_for<0, RA / 4>([&](auto Q) {
                constexpr int q = decltype(Q)::value;
                dif_first<RA / 2, -1, q>(br_[q], bi_[q], br_[q + RA / 4], bi_[q + RA / 4], re[q], im[q]);
            });
            fft_v<RA / 4, -1, float2>(re, im);   // pair k: outputs j' = 2k (.x), 2k + 1 (.y)
            const int a0 = ja * QS + G::xslot(f, 0);
            A2SB_PRAGMA_UNROLL
            for (int k = 0; k < RA / 4; ++k) {
                A2SB_PRAGMA_UNROLL
                for (int o = 0; o < 2; ++o) {
                    const int j0 = 2 * (2 * k + o), j1 = j0 + 1;   // residue in round 0 / round 1
                    const int cc0 = (j0 == 0 || j0 == RA / 2) ? 0 : (j0 < RA / 2 ? j0 : RA - j0), hh0 = (j0 == 0) ? 0 : (j0 >= RA / 2 ? 1 : 0);
                    const int cc1 = (j1 < RA / 2 ? j1 : RA - j1), hh1 = (j1 > RA / 2) ? 1 : 0;
                    const int off = rnd ? (cc1 >> 1) * CS + hh1 * FL : (cc0 >> 1) * CS + hh0 * FL;
                    s_xre[a0 + off] = o ? re[k].y : re[k].x;
                    s_xim[a0 + off] = o ? im[k].y : im[k].x;
                }
            }
        } else {
        A2SB_PRAGMA_UNROLL
        for (int u = 0; u < G::ITEMS_A; ++u) {
            const int item = gt + u * NTG;
            const int f = item / RB, ja = item % RB;
            const in_t* fin = s_in + f * H + 2 * ja;
            const float* win = s_win + 2 * ja;
            float2 re[RA / 2], im[RA / 2];
            static_for<0, RA / 2>([&](auto Q) {
                constexpr int q = decltype(Q)::value;
                const float2 xa = ldz(fin + 2 * RB * q);
                const float2 wa = ld2(win + 2 * RB * q);
                const float2 xc = ldz(fin + 2 * RB * (q + RA / 2));
                const float2 wc = ld2(win + 2 * RB * (q + RA / 2));
                dif_first_windowed<RA, -1, q>(xa, wa, xc, wc, re[q], im[q]);
            });
            fft_v<RA / 2, -1, float2>(re, im);
            // pair k holds outputs jb = 2k (.x) and 2k + 1 (.y); element (residue jb, q = ja) of frame f
            A2SB_PRAGMA_UNROLL
            for (int k = 0; k < RA / 2; ++k) {
                A2SB_PRAGMA_UNROLL
                for (int o = 0; o < 2; ++o) {
                    const int j = 2 * k + o;
                    const int cc = (j == 0 || j == RA / 2) ? 0 : (j < RA / 2 ? j : RA - j);
                    const int hh = (j == 0) ? 0 : (j == RA / 2 ? 1 : (j < RA / 2 ? 0 : 1));
                    const int addr = cc * CS + ja * QS + G::xslot(f, hh);
                    s_xre[addr] = o ? re[k].y : re[k].x;
                    s_xim[addr] = o ? im[k].y : im[k].x;
                }
            }
        }
        }
        group_sync(GROUPS, g, NTG);  // exchange complete; after the last round s_in is free

        // prefetch the next tile's span while pass B runs
        if (rnd == ROUNDS - 1) {
            if constexpr (ROUNDS == 1) {
                ++s;
                have = next_slot(s, b, t0);
            }
            if (have) next_async = load_span(b, t0);
        }

        // ================= pass B: twiddle + radix-RB, split, epilogue =======================
        {
            float zr[RB], zi[RB];
            {
                float2 re[RB / 2], im[RB / 2];
                A2SB_PRAGMA_UNROLL
                for (int j = 0; j < RB / 2; ++j) {
                    re[j].x = s_xre[xb + (2 * j) * QS];
                    re[j].y = s_xre[xb + (2 * j + 1) * QS];
                    im[j].x = s_xim[xb + (2 * j) * QS];
                    im[j].y = s_xim[xb + (2 * j + 1) * QS];
                }
                const float4* tw = s_tw4 + jb * G::TWS;
                A2SB_PRAGMA_UNROLL
                for (int j = 0; j < RB / 2; ++j) {
                    const float4 w = G::TABLES_SMEM ? tw[j] : __ldg(tw + j);
                    const float2 cp = make_float2(w.x, w.y), sp = make_float2(w.z, w.w);
                    const float2 tr = p2_fma(re[j], cp, p2_neg(p2_mul(im[j], sp)));
                    im[j] = p2_fma(re[j], sp, p2_mul(im[j], cp));
                    re[j] = tr;
                }
                fft2x_dit<RB, -1>(re, im, zr, zi);  // zr/zi[q] = Z[jb + RA*q]
            }
            const long long tg = cur_t0 + t;
            const bool valid = tg < p.t_end;
            const long long col = tg - p.out_t_first;
            float* clip_out = p.out + (long long)cur_b * C * plane;
            // (corruption epilogue, careful path) this clip's noise tensor
            const long long nplane_c = (long long)rows * p.noise_T;
            const float* nclip = (FAST == 3 && p.noise) ? p.noise + (long long)cur_b * C * nplane_c : nullptr;
            bool careful = !FAST;
            if (FAST) {
                constexpr int PM = (FAST == 2) ? kPowNone : kPowQuarter;
                constexpr bool CORR = (FAST == 3);   // shipped chain + corruption epilogue (FwdParams::out2)
#ifdef A2SB_CONST_T   // experiment: row / plane strides as compile-time constants (immediate store offsets)
                constexpr unsigned long long rowB = 4ull * A2SB_CONST_T, planeB = rowB * M, plane2B = 2ull * planeB,
                                             stepB = (unsigned long long)RA * rowB;
#else
                const unsigned T32 = (unsigned)p.out_T;
                const unsigned long long planeB = 4ull * (unsigned long long)plane, plane2B = 2ull * planeB;
                const unsigned long long rowB = 4ull * T32, stepB = (unsigned long long)RA * rowB;
#endif
                const int drop = p.drop_dc;
                unsigned minbits = 0x7f800000u;
                const float eps = p.eps;
                const unsigned long long base = reinterpret_cast<unsigned long long>(clip_out + col);
                // corruption epilogue: byte distance to the second output, this clip's noise column, rectangle tests
                const long long d2 = CORR ? (long long)(reinterpret_cast<const char*>(p.out2) - reinterpret_cast<const char*>(p.out)) : 0;
                const long long nplane = (long long)rows * p.noise_T;
                const float* ncol = CORR ? p.noise + (long long)cur_b * C * nplane + col : nullptr;
                const bool col_in = CORR && col >= p.m_col0 && col < p.m_col1;
                // one pair of bins (rows r_lo, r_hi of the tensor) at addresses a_lo / a_hi: values, clean stores and -- CORR --
                // the corrupted counterparts
                auto emit_pair = [&](float2 xr, float2 xi, unsigned& mbits, bool ok, unsigned long long a_lo, unsigned long long a_hi,
                                     int r_lo, int r_hi) {
                    if constexpr (!CORR) {
                        fwd_emit_pair_fast<PM>(xr, xi, eps, mbits, ok, a_lo, a_hi, planeB, plane2B);
                    } else {
                        // the six noise samples are requested BEFORE the values are computed (independent loads in flight instead
                        // of one dependent load per store); outside the mask they are only needed for exact zeros (st_corrupt)
                        const bool in_lo = ok && col_in && r_lo >= p.m_row0 && r_lo < p.m_row1;
                        const bool in_hi = ok && col_in && r_hi >= p.m_row0 && r_hi < p.m_row1;
                        const float* n_lo = ncol + (long long)r_lo * p.noise_T;
                        const float* n_hi = ncol + (long long)r_hi * p.noise_T;
                        float z0 = 0.f, z1 = 0.f, z2 = 0.f, z3 = 0.f, z4 = 0.f, z5 = 0.f;
                        if (in_lo) { z0 = __ldg(n_lo); z1 = __ldg(n_lo + nplane); z2 = __ldg(n_lo + 2 * nplane); }
                        if (in_hi) { z3 = __ldg(n_hi); z4 = __ldg(n_hi + nplane); z5 = __ldg(n_hi + 2 * nplane); }
                        float2 mag, cs, sn;
                        fwd_pair_values<PM>(xr, xi, eps, mbits, mag, cs, sn);
                        if (ok) {
                            st_stream_at(a_lo, mag.x); st_stream_at(a_lo + planeB, cs.x); st_stream_at(a_lo + plane2B, sn.x);
                            st_stream_at(a_hi, mag.y); st_stream_at(a_hi + planeB, cs.y); st_stream_at(a_hi + plane2B, sn.y);
                            st_corrupt(a_lo, d2, n_lo, in_lo, z0, p.m_level, mag.x);
                            st_corrupt(a_lo + planeB, d2, n_lo + nplane, in_lo, z1, p.m_level, cs.x);
                            st_corrupt(a_lo + plane2B, d2, n_lo + 2 * nplane, in_lo, z2, p.m_level, sn.x);
                            st_corrupt(a_hi, d2, n_hi, in_hi, z3, p.m_level, mag.y);
                            st_corrupt(a_hi + planeB, d2, n_hi + nplane, in_hi, z4, p.m_level, cs.y);
                            st_corrupt(a_hi + plane2B, d2, n_hi + 2 * nplane, in_hi, z5, p.m_level, sn.y);
                        }
                    }
                };
#ifndef A2SB_NO_SEAM_PREFETCH
                // Rows are only 8-byte aligned (T*4 is not a multiple of 32), so the first and last 32-byte
                // sector of every 4F-byte row segment is written partially, and L2 has to fill such a sector
                // from DRAM before it can merge the write; left on the store path those fills throttle the
                // LSU (measured: 2.0 ms -> 0.84 ms for the same kernel when T*4 % 32 == 0).  So pull the
                // head seam sectors of this group's NEXT tile into L2 now, a tile ahead of its stores: one lane
                // per row, only where the seam really is unaligned.
                // Two rounds: the lead is ONE ROUND, not one tile -- round 0 fetches the seams of the rows round 1 of this tile
                // writes, round 1 those of round 0 of the next tile.  (A line prefetched a whole 32-frame tile = 20 us ahead is
                // gone again when its stores arrive: 60 MB of streaming stores pass through L2 in that time; measured 1.72 ms
                // with any policy at that distance.)
                const bool pf_same = ((ROUNDS > 1) && rnd == 0) || (ROUNDS == 1 && (p.seam & 8));   // (8: experiment, this tile's own seams)
                const int pb = pf_same ? cur_b : b;
                const long long pt0 = pf_same ? cur_t0 : t0;
                const int pc = (ROUNDS > 1) ? 2 * wslot + (1 - rnd) : c;
                const int pjb = (ROUNDS > 1) ? ((pc == 0) ? (h ? RA / 2 : 0) : (h ? RA - pc : pc)) : jb;
                if (pf_same || have) {
                    long long nlast = pt0 + F;
                    if (nlast > p.t_end) nlast = p.t_end;
                    const unsigned tail_off = (unsigned)(nlast - pt0 - 1) * 4u;   // byte offset of the last frame
                    // The seam sector at the tile's tail is the head seam of the NEXT tile of the row, whose own group
                    // prefetches it at the same time: fetching it here as well doubled the fill reads (measured
                    // 1.17 -> 1.05 ms without).  Only the last tile of a clip has no successor to do it.
                    // (Measured per kernel family: without the tail prefetch n_fft = 2048 goes 1.17 -> 1.06 ms, but
                    // n_fft = 512 / 1024 / 4096, whose smaller CTAs run two per SM and drift apart more, go
                    // 1.58 -> 1.95, 2.38 -> 2.54 and 2.65 -> 3.21 ms: they keep prefetching both seams.)
                    const bool clip_tail = nlast == p.t_end || (p.seam & 2);
                    const unsigned long long nb =
                        reinterpret_cast<unsigned long long>(p.out + (long long)pb * C * plane + (pt0 - p.out_t_first));
                    // F > 16: the lanes of 16-frame block tb look after the seam at the head of THEIR block; the inner seams
                    // (tb > 0) are written from both sides by two warps of this CTA at the same time and are prefetched only
                    // on request (p.seam & 4)
                    const int tb = t / FL;
                    if (tb == 0 || ((p.seam & 4) && pt0 + tb * FL < nlast)) {
                        A2SB_PRAGMA_UNROLL
                        for (int qq = t % FL; qq < RB / 2; qq += FL) {
                            const int k = pjb + RA * qq;
                            A2SB_PRAGMA_UNROLL
                            for (int hl = 0; hl < 2; ++hl) {
                                const int row = (hl ? M - k : k) - drop;
                                if (row < 0) continue;
                                A2SB_PRAGMA_UNROLL
                                for (int ch = 0; ch < 3; ++ch) {
                                    const unsigned long long a = nb + (unsigned)row * rowB + ch * planeB + (unsigned)(tb * FL) * 4u;
#ifndef A2SB_SEAM_MASK
#define A2SB_SEAM_MASK 31u   // experiment: 63u = treat the 64-byte L2 / DRAM atom as the unit (tools/microbench/store_pattern.cu)
#endif
                                    if ((p.seam & 1) && (a & A2SB_SEAM_MASK)) {
                                        prefetch_seam(reinterpret_cast<const void*>(a));
                                        if constexpr (CORR) prefetch_seam(reinterpret_cast<const void*>(a + d2));
                                    }
                                    if (tb == 0 && clip_tail && ((a + tail_off + 4u) & A2SB_SEAM_MASK)) {
                                        prefetch_seam(reinterpret_cast<const void*>(a + tail_off));
                                        if constexpr (CORR) prefetch_seam(reinterpret_cast<const void*>(a + tail_off + d2));
                                    }
                                }
                            }
                        }
                    }
                }
#endif
                if (wc != 0) {
                    unsigned long long a_lo = base + (unsigned)(jb - drop) * rowB;
                    unsigned long long a_hi = base + (unsigned)(M - jb - drop) * rowB;
                    float* sre = s_xre + wslot * CS;
                    float* sim = s_xim + wslot * CS;
                    if constexpr (WIDE) {
                        // partner residue RA - jb = the other warp of the class: swap the upper half-spectra through the class
                        // region of the exchange, once both warps have read their records out of it
                        pair_sync(1 + wslot);
                        A2SB_PRAGMA_UNROLL
                        for (int q = 0; q < RB / 2; ++q) {
                            sre[(h * (RB / 2) + q) * 32 + lane] = zr[RB / 2 + q];
                            sim[(h * (RB / 2) + q) * 32 + lane] = zi[RB / 2 + q];
                        }
                        pair_sync(1 + wslot);
                    }
                    A2SB_PRAGMA_UNROLL
                    for (int q = 0; q < RB / 2; ++q) {
                        const float zmr = WIDE ? sre[((h ^ 1) * (RB / 2) + (RB / 2 - 1 - q)) * 32 + lane] : __shfl_xor_sync(0xffffffffu, zr[RB - 1 - q], FL);
                        const float zmi = WIDE ? sim[((h ^ 1) * (RB / 2) + (RB / 2 - 1 - q)) * 32 + lane] : __shfl_xor_sync(0xffffffffu, zi[RB - 1 - q], FL);
                        float2 xr, xi;
                        fwd_split(zr[q], zi[q], zmr, zmi, twS_at(jb + RA * q), xr, xi);
                        emit_pair(xr, xi, minbits, valid, a_lo, a_hi, jb + RA * q - drop, M - jb - RA * q - drop);
                        a_lo += stepB; a_hi -= stepB;
                    }
                } else {
                    // Warp 0 holds class 0 = the two self-paired residues jb = 0 (h = 0) and jb = RA/2
                    // (h = 1), whose partner bin M-k is one of the thread's own outputs; with F = 8 its
                    // upper lanes hold class 1, which exchanges through the shuffle like every other class.
                    A2SB_PRAGMA_UNROLL
                    for (int q = 0; q < RB / 2; ++q) {
                        float zmr = 0.0f, zmi = 0.0f;
                        if (G::CPW > 1) {
                            zmr = __shfl_xor_sync(0xffffffffu, zr[RB - 1 - q], FL);
                            zmi = __shfl_xor_sync(0xffffffffu, zi[RB - 1 - q], FL);
                        }
                        if (c == 0) {
                            zmr = h ? zr[RB - 1 - q] : zr[(RB - q) % RB];
                            zmi = h ? zi[RB - 1 - q] : zi[(RB - q) % RB];
                        }
                        const int k = jb + RA * q;
                        float2 xr, xi;
                        fwd_split(zr[q], zi[q], zmr, zmi, twS_at(k), xr, xi);
                        const bool self0 = (q == 0 && jb == 0);         // k = 0 / M are emitted below
                        const bool pv = valid && !self0;
                        unsigned mb = 0x7f800000u;
                        emit_pair(xr, xi, mb, pv, base + (unsigned)(k - drop) * rowB, base + (unsigned)(M - k - drop) * rowB, k - drop, M - k - drop);
                        if (!self0) minbits = mb < minbits ? mb : minbits;
                    }
                    if (valid && jb == 0) {
                        // k = 0: X[0] = Zr + Zi (DC), X[M] = Zr - Zi (Nyquist); k = M/2: X = conj(Z); window carries 0.5.
                        fwd_emit(p, clip_out, plane, 0, col, 2.0f * (zr[0] + zi[0]), 0.0f, nclip, nplane_c);
                        fwd_emit(p, clip_out, plane, M, col, 2.0f * (zr[0] - zi[0]), 0.0f, nclip, nplane_c);
                        fwd_emit(p, clip_out, plane, M / 2, col, 2.0f * zr[RB / 2], -2.0f * zi[RB / 2], nclip, nplane_c);
                    }
                }
                // squared magnitudes below 1e-30 (or NaN-free zeros): redo this warp's rows carefully
                careful = __any_sync(0xffffffffu, valid && minbits < 0x0da24260u /* 1e-30f */);
            }
            // Careful emission of this warp's rows at output column `ocol` for the lanes with `ok` set.  The warp parks its
            // spectra in its own exchange region (only this warp reads it in pass B, and every lane has finished those
            // reads) and walks the bins in a rolled loop.
            float* wre = s_xre + (G::SHARED_CLS ? warp * 32 * RB : warp * G::CPW * CS);
            float* wim = s_xim + (G::SHARED_CLS ? warp * 32 * RB : warp * G::CPW * CS);
            // where the partner residue's spectrum is parked: the same warp's partner lanes, or (WIDE) the partner warp
            const float* qre = (WIDE && c != 0) ? s_xre + (warp ^ 1) * 32 * RB : wre;
            const float* qim = (WIDE && c != 0) ? s_xim + (warp ^ 1) * 32 * RB : wim;
            auto park = [&] {
                __syncwarp();
                A2SB_PRAGMA_UNROLL
                for (int q = 0; q < RB; ++q) {
                    wre[q * 32 + lane] = zr[q];
                    wim[q * 32 + lane] = zi[q];
                }
                __syncwarp();
            };
            auto careful_emit = [&](const long long ocol, const bool ok) {
                if (ok) {
                    const int pl = (c != 0 && !WIDE) ? (lane ^ FL) : lane;
                    for (int q = 0; q < RB / 2; ++q) {
                        const int qm = (c != 0 || h) ? RB - 1 - q : (RB - q) % RB;
                        const float zkr = wre[q * 32 + lane], zki = wim[q * 32 + lane];
                        const float zmr = qre[qm * 32 + pl], zmi = qim[qm * 32 + pl];
                        const int k = jb + RA * q;
                        if (k == 0) {
                            fwd_emit(p, clip_out, plane, 0, ocol, 2.0f * (zkr + zki), 0.0f, nclip, nplane_c);
                            fwd_emit(p, clip_out, plane, M, ocol, 2.0f * (zkr - zki), 0.0f, nclip, nplane_c);
                            continue;
                        }
                        float2 xr, xi;
                        fwd_split(zkr, zki, zmr, zmi, twS_at(k), xr, xi);
                        fwd_emit(p, clip_out, plane, k, ocol, xr.x, xi.x, nclip, nplane_c);
                        fwd_emit(p, clip_out, plane, M - k, ocol, xr.y, xi.y, nclip, nplane_c);
                    }
                    if (c == 0 && h == 0)
                        fwd_emit(p, clip_out, plane, M / 2, ocol, 2.0f * wre[(RB / 2) * 32 + lane], -2.0f * wim[(RB / 2) * 32 + lane], nclip, nplane_c);
                }
            };
            // multidiffusion_pad_inputs fused (A2SB/diffusion.py:67-83): the first wrap_cols frames are also the padding that
            // follows column wrap_at.  Only the first tiles of a clip get here (a tile-uniform branch); their head lanes
            // re-emit through the careful path, whose normal-bin arithmetic is the fast path's (bit-identical values).
            // ONE inlined copy of the rolled emission serves both uses (two copies cost K1 2 %: instruction-cache footprint).
            const bool do_wrap = p.wrap_cols > 0 && cur_t0 < p.wrap_cols;
            const bool need = careful || do_wrap;
            bool parked = need;
            if constexpr (G::SHARED_CLS) {
                // A class region is read by several warps, so a warp may park its spectra only after EVERY warp has finished
                // its pass-B reads: the tile-closing barrier doubles as the vote, and the (rare) careful tiles pay more.
                parked = __syncthreads_or(parked) != 0;
            }
            if (parked) {
                if (need || WIDE) park();            // WIDE: the partner warp reads this warp's parked spectrum
                if constexpr (WIDE) __syncthreads();
                if (need) {
#pragma unroll 1
                    for (int pass = careful ? 0 : 1; pass < (do_wrap ? 2 : 1); ++pass)
                        careful_emit(pass ? col + p.wrap_at : col, pass ? (valid && tg < p.wrap_cols) : valid);
                }
                if constexpr (G::SHARED_CLS) __syncthreads();
            }
        }
        if constexpr (!G::SHARED_CLS) group_sync(GROUPS, g, NTG);  // exchange free; synchronous span (if any) visible
        if (p.cluster_barriers > 0) { cluster_lockstep(); ++n_cluster_barriers; }
      }  // rounds
        cur_async = next_async;
    }
    for (; n_cluster_barriers < p.cluster_barriers; ++n_cluster_barriers) cluster_lockstep();   // CTAs with fewer tiles keep the count
}

}  // namespace a2sb
