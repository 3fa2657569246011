// dft_generic.cuh -- STFT / iSTFT for transform lengths the radix kernels do not cover (ETTA's STFT helper
// defaults to num_fft = 1023, an odd length chosen to get exactly 512 bins; ETTA/stable_audio_tools/models/adp.py:1510-1590).
// The transform is evaluated as a plain DFT from a cos / sin table of the N-th roots of unity (index k*n mod N kept
// incrementally, no multiplication), O(N^2) per frame: 1 GFLOP per second of audio at N = 1023 -- a convenience path for a
// module the shipped configs do not instantiate, not a hot path.  torch.stft(center=True, pad_mode="reflect", onesided) and
// torch.istft(center=True, length=...) semantics for ANY n_fft >= 2, any hop >= 1, any window (padded to n_fft by the caller).
#pragma once
#include "a2sb_common.cuh"

namespace a2sb {

struct GenParams {
    const float* wav;       // forward: [batch][wav_stride]
    long long wav_stride, len;
    const float* window;    // [N] analysis / synthesis window, centre-padded to N, any normalisation folded in by the caller
    float* spec;            // [batch][2][K][T]  (re, im planes; frames fastest), K = N/2 + 1
    long long T;
    int N, K, hop;
    int batch;
    float* frames;          // inverse scratch: [batch][T][N] windowed time-domain frames
    float* out;             // inverse: [batch][out_len]
    long long out_len;      // samples produced per clip, starting at position N/2 of the overlap-added signal
};

constexpr int kGenTF = 8;      // frames per CTA
constexpr int kGenNT = 128;    // threads per CTA = bins (forward) / samples (inverse) per CTA

A2SB_DEV long long gen_reflect(long long i, long long len) {   // torch reflect padding (pad < len)
    if (i < 0) i = -i;
    if (i >= len) i = 2 * (len - 1) - i;
    return i;
}

// table of the N-th roots of unity in shared memory: (cos, sin)(2 pi m / N), rounded once from double
A2SB_DEV void gen_roots(float2* s_cs, int N) {
    for (int m = threadIdx.x; m < N; m += blockDim.x) {
#ifdef A2SB_EMU
        const double a = 2.0 * 3.14159265358979323846 * (double)m / (double)N;
        s_cs[m] = make_float2((float)std::cos(a), (float)std::sin(a));
#else
        double sn, cs;
        sincospi(2.0 * (double)m / (double)N, &sn, &cs);
        s_cs[m] = make_float2((float)cs, (float)sn);
#endif
    }
}

// grid: (ceil(T / TF), ceil(K / NT), batch).  Shared: roots [N] float2 + windowed frames [TF][N].
__global__ void __launch_bounds__(kGenNT) dft_fwd_generic_kernel(const GenParams p) {
    A2SB_DYN_SMEM(smem);
    float2* s_cs = reinterpret_cast<float2*>(smem);
    float* s_x = reinterpret_cast<float*>(smem + sizeof(float2) * p.N);
    const int N = p.N, b = blockIdx.z;
    const long long t0 = (long long)blockIdx.x * kGenTF;
    gen_roots(s_cs, N);
    const float* clip = p.wav + (long long)b * p.wav_stride;
    for (int i = threadIdx.x; i < kGenTF * N; i += blockDim.x) {
        const int f = i / N, n = i - f * N;
        const long long t = t0 + f;
        float v = 0.0f;
        if (t < p.T) v = clip[gen_reflect(t * p.hop + n - N / 2, p.len)] * p.window[n];
        s_x[i] = v;
    }
    __syncthreads();
    const int k = blockIdx.y * kGenNT + threadIdx.x;
    if (k >= p.K) return;
    float ar[kGenTF], ai[kGenTF];
    A2SB_PRAGMA_UNROLL
    for (int f = 0; f < kGenTF; ++f) { ar[f] = 0.0f; ai[f] = 0.0f; }
    // two partial sums (first / second half of n) keep the rounding error of the N-term dot products at sqrt(2) less
    float br[kGenTF], bi[kGenTF];
    A2SB_PRAGMA_UNROLL
    for (int f = 0; f < kGenTF; ++f) { br[f] = 0.0f; bi[f] = 0.0f; }
    int idx = 0;
    const int half = N / 2;
    for (int n = 0; n < N; ++n) {
        const float2 w = s_cs[idx];          // exp(-2 pi i k n / N) = (cos, -sin)
        if (n < half) {
            A2SB_PRAGMA_UNROLL
            for (int f = 0; f < kGenTF; ++f) { const float x = s_x[f * N + n]; ar[f] = fmaf(x, w.x, ar[f]); ai[f] = fmaf(-x, w.y, ai[f]); }
        } else {
            A2SB_PRAGMA_UNROLL
            for (int f = 0; f < kGenTF; ++f) { const float x = s_x[f * N + n]; br[f] = fmaf(x, w.x, br[f]); bi[f] = fmaf(-x, w.y, bi[f]); }
        }
        idx += k;
        if (idx >= N) idx -= N;
    }
    float* o = p.spec + ((long long)b * 2 * p.K + k) * p.T;
    A2SB_PRAGMA_UNROLL
    for (int f = 0; f < kGenTF; ++f) {
        const long long t = t0 + f;
        if (t < p.T) { o[t] = ar[f] + br[f]; o[(long long)p.K * p.T + t] = ai[f] + bi[f]; }
    }
}

// Inverse, step 1: frames[b][t][n] = w[n] / N * sum_k c_k (Re X[k] cos(2 pi k n / N) - Im X[k] sin(2 pi k n / N)), c_0 = 1,
// c_k = 2, c_{N/2} = 1 for even N (irfft: the imaginary parts of the DC and Nyquist bins do not contribute).
// grid: (ceil(T / TF), ceil(N / NT), batch).  Shared: roots [N] float2 + spectrum tile [K][TF] float2.
__global__ void __launch_bounds__(kGenNT) dft_inv_generic_frames_kernel(const GenParams p) {
    A2SB_DYN_SMEM(smem);
    float2* s_cs = reinterpret_cast<float2*>(smem);
    float2* s_X = reinterpret_cast<float2*>(smem + sizeof(float2) * p.N);
    const int N = p.N, K = p.K, b = blockIdx.z;
    const long long t0 = (long long)blockIdx.x * kGenTF;
    gen_roots(s_cs, N);
    const float* re = p.spec + (long long)b * 2 * K * p.T;
    const float* im = re + (long long)K * p.T;
    for (int i = threadIdx.x; i < K * kGenTF; i += blockDim.x) {
        const int k = i / kGenTF, f = i - k * kGenTF;
        const long long t = t0 + f;
        float2 v = make_float2(0.0f, 0.0f);
        if (t < p.T) {
            const float ck = (k == 0 || (N % 2 == 0 && k == N / 2)) ? 1.0f : 2.0f;
            v = make_float2(ck * re[(long long)k * p.T + t], ck * im[(long long)k * p.T + t]);
        }
        s_X[i] = v;
    }
    __syncthreads();
    const int n = blockIdx.y * kGenNT + threadIdx.x;
    if (n >= N) return;
    float acc[kGenTF], acd[kGenTF];
    A2SB_PRAGMA_UNROLL
    for (int f = 0; f < kGenTF; ++f) { acc[f] = 0.0f; acd[f] = 0.0f; }
    int idx = 0;
    const int half = K / 2;
    for (int k = 0; k < K; ++k) {
        const float2 w = s_cs[idx];          // exp(+2 pi i k n / N) = (cos, sin)
        if (k < half) {
            A2SB_PRAGMA_UNROLL
            for (int f = 0; f < kGenTF; ++f) { const float2 x = s_X[k * kGenTF + f]; acc[f] = fmaf(x.x, w.x, fmaf(-x.y, w.y, acc[f])); }
        } else {
            A2SB_PRAGMA_UNROLL
            for (int f = 0; f < kGenTF; ++f) { const float2 x = s_X[k * kGenTF + f]; acd[f] = fmaf(x.x, w.x, fmaf(-x.y, w.y, acd[f])); }
        }
        idx += n;
        if (idx >= N) idx -= N;
    }
    const float scale = p.window[n] / (float)N;
    A2SB_PRAGMA_UNROLL
    for (int f = 0; f < kGenTF; ++f) {
        const long long t = t0 + f;
        if (t < p.T) p.frames[((long long)b * p.T + t) * N + n] = (acc[f] + acd[f]) * scale;
    }
}

// Inverse, step 2: overlap-add in ascending frame order, divided by the overlap-added squared window, trimmed by N/2 at the
// head (torch.istft center=True); samples past the last frame are zero (torch pads them).  One thread per output sample.
__global__ void __launch_bounds__(256) dft_inv_generic_ola_kernel(const GenParams p) {
    const long long total = (long long)p.batch * p.out_len;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long b = i / p.out_len, j = i - b * p.out_len;
        const long long pos = j + p.N / 2;
        long long t_hi = pos / p.hop;
        if (t_hi > p.T - 1) t_hi = p.T - 1;
        long long t_lo = (pos - p.N + p.hop) / p.hop;     // smallest t with t*hop + N > pos
        if (pos - p.N + 1 <= 0) t_lo = 0;
        float y = 0.0f, e = 0.0f;
        for (long long t = t_lo; t <= t_hi; ++t) {
            const long long n = pos - t * p.hop;
            if (n < 0 || n >= p.N) continue;
            y += p.frames[(b * p.T + t) * p.N + n];
            const float w = p.window[n];
            e += w * w;
        }
        p.out[i] = (e > 0.0f) ? y / e : 0.0f;
    }
}

}  // namespace a2sb
