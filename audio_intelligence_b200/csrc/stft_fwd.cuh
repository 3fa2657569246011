// stft_fwd.cuh -- K1: fused framing + window + real FFT + mag/phase/power-compress epilogue.
//
// Replaces, in one pass over HBM, the reference's forward chain
//   ComplexSpectrogram -> ComplexToMagInstPhase -> SpectrogramDropDCTerm -> PowerScaleSpectrogram
//   (A2SB/audio_transforms/transforms.py:83-118,187-219; torch.stft via torchaudio Spectrogram).
//
// Geometry (n_fft = N = 2M real samples -> M-point complex FFT of z[n] = x[2n] + i x[2n+1]):
//   * one CTA sweeps tiles of F = 16 consecutive frames of one clip (persistent, grid = k*148);
//   * the tile's input span ((F-1)*hop + N contiguous samples, frames overlap) is brought into
//     shared memory by ONE 1-D TMA bulk copy, prefetched a tile ahead; clip edges (reflect
//     padding) use an index-mapped loader instead;
//   * pass A (frame-major threads): window, radix-RA register FFT over q of z[ja + RB*q];
//   * exchange through shared memory in a layout that makes both sides conflict-free;
//   * pass B (frame-minor threads: lane = 16 frames x {residue j, RA-j}): twiddle, radix-RB
//     register FFT -> Z[jb + RA*q]; the real-FFT split needs Z[M-k], which lives in the partner
//     half-warp -> one __shfl_xor per value; then magnitude / unit phasor / power compression
//     in registers and stores whose lanes run along the frame axis (the output's fastest axis),
//     i.e. 64-byte row segments per half-warp.
// HBM traffic per tile = span read once + 3*M*F floats written once (algorithmic minimum).
#pragma once
#include "a2sb_common.cuh"
#include "radix.cuh"
#include "tma.cuh"

namespace a2sb {

constexpr int kF = 16;  // frames per tile == lanes along the frame axis

struct FwdParams {
    const float* wav;        // [batch][wav_stride] local sample buffers
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
    int tiles_per_clip;
    long long total_tiles;
    const float* window;     // [N], already multiplied by 0.5 (real-FFT split scale)
    const float2* twM;       // [M]      exp(-2 pi i m / M)
    const float2* twN;       // [M/2+1]  (cos, sin)(2 pi k / N)
    int drop_dc;             // 1: rows are bins 1..M (SpectrogramDropDCTerm), 0: bins 0..M
    float power, eps;
};

enum : int { kEpiComplex = 0, kEpiMagPhase = 1 };

template <int M, int RA, int RB>
struct FwdGeom {
    static constexpr int N = 2 * M;
    static constexpr int NT = kF * RA;            // one pass-B item per thread
    static constexpr int ITEMS_A = RB / RA;       // pass-A items per thread
    static constexpr int QS = (RB == 32) ? 33 : 34;  // exchange q-stride (conflict-free writes)
    static constexpr int CLS = RA / 2;            // residue classes {j, RA-j}
    static constexpr int XPLANE = CLS * RB * QS;  // floats per exchange plane
    static_assert(M == RA * RB, "two-pass decomposition");
    static_assert(RB % RA == 0 && NT % 32 == 0 && NT / 32 == CLS, "thread mapping");
    // shared memory carve-up (bytes)
    static constexpr size_t off_bar = 0;
    static constexpr size_t off_win = 16;
    static constexpr size_t off_twM = off_win + sizeof(float) * N;
    static constexpr size_t off_twN = off_twM + sizeof(float2) * M;
    static constexpr size_t off_xre = off_twN + ((sizeof(float2) * (M / 2 + 1) + 15) / 16) * 16;
    static constexpr size_t off_xim = off_xre + sizeof(float) * XPLANE;
    static constexpr size_t off_in = ((off_xim + sizeof(float) * XPLANE + 127) / 128) * 128;
    static size_t smem_bytes(int hop) { return off_in + sizeof(float) * ((size_t)(kF - 1) * hop + N); }
};

A2SB_DEV void st_stream(float* p, float v) {
#ifdef A2SB_EMU
    *p = v;
#else
    __stcs(p, v);  // streaming store: the spectrogram is not re-read by this kernel
#endif
}

// One output bin -> global memory.  xr/xi already carry the final scale.
template <int EPI, int PMODE>
A2SB_DEV void fwd_emit(const FwdParams& p, float* __restrict__ clip_out, long long plane, int rows, int k,
                       long long col, float xr, float xi) {
    if (EPI == kEpiComplex) {
        st_stream(clip_out + (long long)k * p.out_T + col, xr);
        st_stream(clip_out + plane + (long long)k * p.out_T + col, xi);
        return;
    }
    const int row = k - p.drop_dc;
    if (row < 0) return;
    // ComplexToMagInstPhase (transforms.py:116-118): mag = sqrt(re^2+im^2); (cos, sin)(atan2(im, re)).
    const float m2 = xr * xr + xi * xi;
    float mag, cs, sn;
    if (m2 >= 1e-30f) {
        const float rs = rsqrt_approx(m2);
        mag = m2 * rs; cs = xr * rs; sn = xi * rs;
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
    }
    // PowerScaleSpectrogram on channel 0 (transforms.py:199-206): m * (|m|^p / (|m| + eps)).
    if (PMODE != kPowNone) mag = mag * power_scale_factor<PMODE>(mag, p.power, p.eps);
    float* o = clip_out + (long long)row * p.out_T + col;
    st_stream(o, mag);
    st_stream(o + plane, cs);
    st_stream(o + 2 * plane, sn);
    (void)rows;
}

// Pair (k, M-k), 1 <= k < M/2 (or k == M/2 via the same formula), from Z[k] and Z[M-k].
// With the 0.5 folded into the window:  E = Zk + conj(Zm),  O = Zk - conj(Zm),
//   X[k]   = (Er + b, Ei - a),  X[M-k] = (Er - b, -(Ei + a)),
//   a = c*Or + s*Oi, b = c*Oi - s*Or, (c, s) = (cos, sin)(2 pi k / N).
template <int EPI, int PMODE, int M>
A2SB_DEV void fwd_emit_pair(const FwdParams& p, float* clip_out, long long plane, int rows, int k, long long col,
                            float zkr, float zki, float zmr, float zmi, const float2* s_twN) {
    const float er = zkr + zmr, ei = zki - zmi;
    const float orr = zkr - zmr, oi = zki + zmi;
    const float2 w = s_twN[k];
    const float a = w.x * orr + w.y * oi;
    const float b = w.x * oi - w.y * orr;
    fwd_emit<EPI, PMODE>(p, clip_out, plane, rows, k, col, er + b, ei - a);
    fwd_emit<EPI, PMODE>(p, clip_out, plane, rows, M - k, col, er - b, -(ei + a));
}

template <int M, int RA, int RB, int EPI, int PMODE>
__global__ void __launch_bounds__(kF * RA, (kF * RA <= 256) ? 2 : 1) stft_fwd_kernel(const FwdParams p) {
    using G = FwdGeom<M, RA, RB>;
    constexpr int N = G::N, NT = G::NT, QS = G::QS;
    A2SB_DYN_SMEM(smem);
    unsigned long long* s_bar = reinterpret_cast<unsigned long long*>(smem + G::off_bar);
    float* s_win = reinterpret_cast<float*>(smem + G::off_win);
    float2* s_twM = reinterpret_cast<float2*>(smem + G::off_twM);
    float2* s_twN = reinterpret_cast<float2*>(smem + G::off_twN);
    float* s_xre = reinterpret_cast<float*>(smem + G::off_xre);
    float* s_xim = reinterpret_cast<float*>(smem + G::off_xim);
    float* s_in = reinterpret_cast<float*>(smem + G::off_in);

    const int tid = threadIdx.x;
    const int H = p.hop;
    const int span = (kF - 1) * H + N;
    const int C = (EPI == kEpiComplex) ? 2 : 3;
    const int rows = (EPI == kEpiComplex) ? M + 1 : (M + 1 - p.drop_dc);
    const long long plane = (long long)rows * p.out_T;

    // ---- tables -> shared memory; barrier init
    for (int i = tid; i < N; i += NT) s_win[i] = p.window[i];
    for (int i = tid; i < M; i += NT) s_twM[i] = p.twM[i];
    for (int i = tid; i <= M / 2; i += NT) s_twN[i] = p.twN[i];
    if (tid == 0) { mbar_init(s_bar, 1); fence_mbar_init(); }

    // Loads the input span of `tile` into s_in.  Returns true if an asynchronous TMA copy was
    // issued (completion on s_bar), false if the span was filled synchronously by all threads.
    auto load_span = [&](long long tile) -> bool {
        const int b = (int)(tile / p.tiles_per_clip);
        const long long t0 = p.t_begin + (long long)(tile % p.tiles_per_clip) * kF;
        const long long g0 = t0 * H - N / 2;  // global sample index of s_in[0]
        const float* clip = p.wav + (long long)b * p.wav_stride;
        const long long l0 = g0 - p.sample_first;
        const float* src = clip + l0;
        const bool inside = g0 >= 0 && g0 + span <= p.len && l0 >= 0 && l0 + span <= p.n_local &&
                            ((reinterpret_cast<uintptr_t>(src) & 15) == 0) && (H % 4 == 0);
        if (inside) {
            if (tid == 0) {
                fence_proxy_async();
                bulk_g2s(s_in, src, (unsigned)(span * sizeof(float)), s_bar);
            }
            return true;
        }
        for (int i = tid; i < span; i += NT) {
            long long g = reflect_index(g0 + i, p.len);
            g = g < 0 ? 0 : (g >= p.len ? p.len - 1 : g);  // only reachable for masked frames
            long long l = g - p.sample_first;
            l = l < 0 ? 0 : (l >= p.n_local ? p.n_local - 1 : l);
            s_in[i] = clip[l];
        }
        return false;
    };

    long long tile = blockIdx.x;
    bool cur_async = false;
    unsigned phase = 0;
    __syncthreads();  // barrier init visible before any TMA is issued
    if (tile < p.total_tiles) cur_async = load_span(tile);
    __syncthreads();  // tables + (synchronous) span visible

    for (; tile < p.total_tiles; tile += gridDim.x) {
        const int b = (int)(tile / p.tiles_per_clip);
        const long long t0 = p.t_begin + (long long)(tile % p.tiles_per_clip) * kF;
        if (cur_async) { mbar_wait(s_bar, phase); phase ^= 1u; }

        // ================= pass A: window + radix-RA over q of z[ja + RB*q] =================
        A2SB_PRAGMA_UNROLL
        for (int u = 0; u < G::ITEMS_A; ++u) {
            const int item = tid + u * NT;
            const int f = item / RB, ja = item % RB;
            const float* fin = s_in + f * H;
            float re[RA], im[RA];
            A2SB_PRAGMA_UNROLL
            for (int q = 0; q < RA; ++q) {
                const int n = ja + RB * q;
                const float2 v = *reinterpret_cast<const float2*>(fin + 2 * n);
                const float2 w = *reinterpret_cast<const float2*>(s_win + 2 * n);
                re[q] = v.x * w.x;
                im[q] = v.y * w.y;
            }
            fft_reg<RA, -1>(re, im);
            // y[ja*RA + jb] is pass B's element (residue jb, q = ja) of frame f.
            A2SB_PRAGMA_UNROLL
            for (int jb = 0; jb < RA; ++jb) {
                const int c = (jb == 0 || jb == RA / 2) ? 0 : (jb < RA / 2 ? jb : RA - jb);
                const int h = (jb == 0) ? 0 : (jb == RA / 2 ? 1 : (jb < RA / 2 ? 0 : 1));
                const int addr = c * (RB * QS) + ja * QS + h * kF + f;
                s_xre[addr] = re[jb];
                s_xim[addr] = im[jb];
            }
        }
        __syncthreads();  // exchange complete; s_in is free

        // prefetch the next tile's span while pass B runs
        const long long next = tile + gridDim.x;
        bool next_async = false;
        if (next < p.total_tiles) next_async = load_span(next);

        // ================= pass B: twiddle + radix-RB, split, epilogue =======================
        {
            const int warp = tid >> 5, lane = tid & 31;
            const int h = lane >> 4, t = lane & (kF - 1);
            const int c = warp;
            const int jb = (c == 0) ? (h ? RA / 2 : 0) : (h ? RA - c : c);
            float re[RB], im[RB];
            const int base = c * (RB * QS) + lane;
            A2SB_PRAGMA_UNROLL
            for (int q = 0; q < RB; ++q) {
                re[q] = s_xre[base + q * QS];
                im[q] = s_xim[base + q * QS];
            }
            A2SB_PRAGMA_UNROLL
            for (int q = 1; q < RB; ++q) {
                const float2 w = s_twM[jb * q];
                const float r = re[q] * w.x - im[q] * w.y;
                im[q] = re[q] * w.y + im[q] * w.x;
                re[q] = r;
            }
            fft_reg<RB, -1>(re, im);  // re/im[q] = Z[jb + RA*q]

            const long long tg = t0 + t;
            const bool valid = tg < p.t_end;
            const long long col = tg - p.out_t_first;
            float* clip_out = p.out + (long long)b * C * plane;
            if (c != 0) {
                A2SB_PRAGMA_UNROLL
                for (int q = 0; q < RB / 2; ++q) {
                    const float zmr = __shfl_xor_sync(0xffffffffu, re[RB - 1 - q], kF);
                    const float zmi = __shfl_xor_sync(0xffffffffu, im[RB - 1 - q], kF);
                    if (valid)
                        fwd_emit_pair<EPI, PMODE, M>(p, clip_out, plane, rows, jb + RA * q, col, re[q], im[q], zmr,
                                                     zmi, s_twN);
                }
            } else {
                // class 0 holds the two self-paired residues: jb = 0 (h = 0) and jb = RA/2 (h = 1).
                A2SB_PRAGMA_UNROLL
                for (int q = 0; q < RB / 2; ++q) {
                    const float zmr = h ? re[RB - 1 - q] : re[(RB - q) % RB];
                    const float zmi = h ? im[RB - 1 - q] : im[(RB - q) % RB];
                    if (!valid) continue;
                    if (q == 0 && h == 0) {
                        // k = 0: X[0] = Zr + Zi (DC), X[M] = Zr - Zi (Nyquist); window carries 0.5.
                        fwd_emit<EPI, PMODE>(p, clip_out, plane, rows, 0, col, 2.0f * (re[0] + im[0]), 0.0f);
                        fwd_emit<EPI, PMODE>(p, clip_out, plane, rows, M, col, 2.0f * (re[0] - im[0]), 0.0f);
                    } else {
                        fwd_emit_pair<EPI, PMODE, M>(p, clip_out, plane, rows, jb + RA * q, col, re[q], im[q], zmr,
                                                     zmi, s_twN);
                    }
                }
                if (valid && h == 0)  // k = M/2 pairs with itself: X = 2*conj(Z)
                    fwd_emit<EPI, PMODE>(p, clip_out, plane, rows, M / 2, col, 2.0f * re[RB / 2], -2.0f * im[RB / 2]);
            }
        }
        __syncthreads();  // exchange free; synchronous span (if any) visible
        cur_async = next_async;
    }
}

}  // namespace a2sb
