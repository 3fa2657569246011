// a2sb_api.cu -- C-ABI entry points of liba2sb_b200.so (see include/a2sb_b200.h).
// Host-side plan/table construction, argument validation (same error conditions as the
// reference path raises through torch), launch geometry, and kernel dispatch.
#include "../../include/a2sb_b200.h"

#include <atomic>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <map>
#include <mutex>
#include <string>
#include <tuple>
#include <vector>

#include "host_util.h"
#include "istft_inv.cuh"
#include "dft_generic.cuh"
#include "masks.cuh"
#include "pointwise.cuh"
#include "segments.cuh"
#include "stft_fwd.cuh"

namespace a2sb {
thread_local std::string g_err;
std::atomic<long long> g_launches{0};
std::atomic<long long> g_tma_launches{0};

namespace {
struct LaunchKey {
    const void* kern;
    size_t smem;
    int dev;
    bool operator<(const LaunchKey& o) const { return std::tie(kern, smem, dev) < std::tie(o.kern, o.smem, o.dev); }
};
std::mutex g_launch_mu;
std::map<LaunchKey, int> g_launch_cache;
std::map<std::pair<const void*, int>, size_t> g_smem_attr;   // (kernel, device) -> largest attribute set
int current_device() {
    int d = 0;
    cudaGetDevice(&d);
    return d;
}
}  // namespace

int launch_cache_lookup(const void* kern, size_t smem) {
    std::lock_guard<std::mutex> lk(g_launch_mu);
    auto it = g_launch_cache.find(LaunchKey{kern, smem, current_device()});
    return it == g_launch_cache.end() ? -1 : it->second;
}
void launch_cache_store(const void* kern, size_t smem, int per_sm) {
    std::lock_guard<std::mutex> lk(g_launch_mu);
    g_launch_cache[LaunchKey{kern, smem, current_device()}] = per_sm;
}

size_t launch_smem_attr_get(const void* kern) {
    std::lock_guard<std::mutex> lk(g_launch_mu);
    auto it = g_smem_attr.find({kern, current_device()});
    return it == g_smem_attr.end() ? 0 : it->second;
}
void launch_smem_attr_set(const void* kern, size_t smem) {
    std::lock_guard<std::mutex> lk(g_launch_mu);
    size_t& v = g_smem_attr[{kern, current_device()}];
    if (smem > v) v = smem;
}

int fail(int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_err = buf;
    return code;
}
}  // namespace a2sb

namespace {
using a2sb::fail;
using a2sb::g_launches;
using a2sb::launch_grid_stride;

int device_sm_count() {
#ifdef A2SB_EMU
    return 3;  // the emulation runs a 3-CTA persistent grid
#else
    static std::atomic<int> cached[64];              // per device ordinal; 0 = not queried yet
    int dev = 0, n = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return a2sb::kSMs;
    if (dev >= 0 && dev < 64 && (n = cached[dev].load(std::memory_order_relaxed)) > 0) return n;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) return a2sb::kSMs;
    if (dev >= 0 && dev < 64) cached[dev].store(n, std::memory_order_relaxed);
    return n;
#endif
}

bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

}  // namespace

struct a2sb_plan {
    int n_fft = 0, win_length = 0, hop = 0, M = 0;
    bool inverse_ok = false;   // hop is a multiple of 4 that divides n_fft (forward-only plan otherwise)
    int sm_count = 0;
    std::vector<float> h_w;  // analysis/synthesis window padded to n_fft (torch.stft centre-pads it)
    float* d_win_fwd = nullptr;   // 0.5 * w
    float* d_win_fwd_pcm = nullptr;  // 0.5 * w / 32768 (16-bit PCM ingest: the decode's scale rides on the window)
    float* d_win_inv = nullptr;   // w / n_fft
    float* d_wsq = nullptr;       // w^2
    float* d_inv_env = nullptr;   // 1 / sum_m w^2[r + m*hop]
    float2* d_twM = nullptr;      // exp(-2 pi i m / M)
    float2* d_twN = nullptr;      // (cos, sin)(2 pi k / n_fft), k <= M/2
    float4* d_tw4f = nullptr;     // forward pass-B twiddle pairs [RA][RB/2 + 1]
    float4* d_tw4f2 = nullptr;    // n_fft = 4096: the same for the two-round kernel's decomposition (64 x 32)
    void* d_twS = nullptr;        // split table (c, -c, -s, s)(2 pi k / n_fft), k <= M/2 (float2 (c, s) when M >= 2048)
    float4* d_tw4i = nullptr;     // inverse inter-pass twiddle pairs [RB][RA/2 + 1] (applied at the end of pass A)
    int fwd_tile = 16;            // frames per forward tile (32 for n_fft <= 1024; A2SB_FWD_TILE=8|16|32)
    int inv_tile = 16;            // frames per inverse tile (8 for n_fft = 4096; A2SB_INV_TILE=8|16)
    // lazily allocated staging for a2sb_roundtrip_host
    struct Lane {
        cudaStream_t stream = nullptr;
        float *d_wav = nullptr, *d_spec = nullptr, *d_out = nullptr;
        long long clips = 0, len = 0;
    } lanes[8];
    int n_lanes = 3;              // lanes used by a2sb_roundtrip_host (A2SB_E2E_LANES=1..8)
};

// Process-wide caps on the persistent grids of K1 / K2 (0 = one CTA per SM slot as usual): lets a caller run K1 and K2 of
// different pieces CONCURRENTLY on disjoint halves of the GPU (sharding.PeerLongClipRoundTrip, overlap mode).
static std::atomic<int> g_grid_limit_fwd{0}, g_grid_limit_inv{0};
static int limited_sm_count(int sm_count, int limit) { return (limit > 0 && limit < sm_count) ? limit : sm_count; }

extern "C" {

int a2sb_set_grid_limit(int max_ctas_forward, int max_ctas_inverse) {
    if (max_ctas_forward < 0 || max_ctas_inverse < 0) return fail(A2SB_ERR_INVALID, "negative grid limit");
    g_grid_limit_fwd.store(max_ctas_forward); g_grid_limit_inv.store(max_ctas_inverse);
    return A2SB_OK;
}

const char* a2sb_last_error(void) { return a2sb::g_err.c_str(); }
int a2sb_version(void) { return 100; }
int a2sb_is_device_build(void) {
#ifdef A2SB_EMU
    return 0;
#else
    return 1;
#endif
}
int64_t a2sb_launch_count(void) { return g_launches.load(); }
int64_t a2sb_tma_launch_count(void) { return a2sb::g_tma_launches.load(); }
int64_t a2sb_num_frames(int64_t len, int hop_length) { return a2sb::num_frames(len, hop_length); }
int64_t a2sb_istft_length(int64_t n_frames, int hop_length) { return (int64_t)hop_length * (n_frames - 1); }

int a2sb_plan_create(a2sb_plan** out, int n_fft, int win_length, int hop, const float* h_window) {
    if (!out) return fail(A2SB_ERR_INVALID, "plan pointer is null");
    *out = nullptr;
    if (n_fft != 512 && n_fft != 1024 && n_fft != 2048 && n_fft != 4096)
        return fail(A2SB_ERR_INVALID, "n_fft=%d unsupported (supported: 512, 1024, 2048, 4096)", n_fft);
    if (win_length < 1 || win_length > n_fft)
        return fail(A2SB_ERR_INVALID, "win_length=%d must be in [1, n_fft=%d]", win_length, n_fft);
    // The inverse kernel needs a hop that is a multiple of 4 and divides n_fft (an integer number of overlapping frames per
    // sample); the forward kernel only needs an even hop (pairs of samples are loaded together) -- the multi-resolution
    // STFT loss of ETTA's auraloss uses hops such as 50 / 120 / 240 (auraloss.py:363-372).  Such plans are forward-only.
    if (hop < 2 || hop % 2 != 0 || hop > n_fft)
        return fail(A2SB_ERR_INVALID, "hop_length=%d must be an even number in [2, n_fft=%d] (and a multiple of 4 that divides n_fft for "
                    "the inverse transform)", hop, n_fft);
    a2sb_plan* pl = new a2sb_plan();
    pl->n_fft = n_fft; pl->win_length = win_length; pl->hop = hop; pl->M = n_fft / 2;
    pl->inverse_ok = (hop % 4 == 0 && n_fft % hop == 0);
    pl->sm_count = device_sm_count();
    const int N = n_fft, M = n_fft / 2;
    // window, centre-padded to n_fft like torch.stft (functional.py:508: left = (n_fft - win_length) // 2)
    pl->h_w.assign(N, 0.0f);
    const int left = (N - win_length) / 2;
    for (int n = 0; n < win_length; ++n)
        pl->h_w[left + n] = h_window ? h_window[n]
                                     : (float)(0.5 - 0.5 * std::cos(2.0 * M_PI * (double)n / (double)win_length));
    std::vector<float> wf(N), wfp(N), wi(N), wsq(N), ienv(hop);
    for (int n = 0; n < N; ++n) {
        wf[n] = 0.5f * pl->h_w[n];
        wfp[n] = wf[n] * (1.0f / 32768.0f);
        wi[n] = pl->h_w[n] / (float)N;
        wsq[n] = pl->h_w[n] * pl->h_w[n];
    }
    for (int r = 0; r < hop && pl->inverse_ok; ++r) {
        float e = 0.0f;
        for (int m = 0; m < N / hop; ++m) e += wsq[r + m * hop];
        ienv[r] = 1.0f / e;  // interior envelope; NOLA violations are rejected per call
    }
    std::vector<float2> twM(M), twN(M / 2 + 1);
    for (int m = 0; m < M; ++m) {
        const double a = -2.0 * M_PI * (double)m / (double)M;
        twM[m] = make_float2((float)std::cos(a), (float)std::sin(a));
    }
    for (int k = 0; k <= M / 2; ++k) {
        const double a = 2.0 * M_PI * (double)k / (double)N;
        twN[k] = make_float2((float)std::cos(a), (float)std::sin(a));
    }
    int fRA = 0, fRB = 0;
    a2sb::fwd_radices(M, fRA, fRB);
    const int tws = fRB / 2 + 1;
    std::vector<float4> tw4f((size_t)fRA * tws, make_float4(0.f, 0.f, 0.f, 0.f)), twS(M / 2 + 1);
    for (int jb = 0; jb < fRA; ++jb)
        for (int j = 0; j < fRB / 2; ++j) {
            const double a0 = -2.0 * M_PI * (double)jb * (double)(2 * j) / (double)M;
            const double a1 = -2.0 * M_PI * (double)jb * (double)(2 * j + 1) / (double)M;
            tw4f[(size_t)jb * tws + j] = make_float4((float)std::cos(a0), (float)std::cos(a1), (float)std::sin(a0), (float)std::sin(a1));
        }
    for (int k = 0; k <= M / 2; ++k) {
        const double a = 2.0 * M_PI * (double)k / (double)N;
        const float c = (float)std::cos(a), sn = (float)std::sin(a);
        twS[k] = make_float4(c, -c, -sn, sn);
    }
    std::vector<float4> tw4f2;
    if (M == 2048) {
        const int RA2 = 64, RB2 = 32, tws2 = RB2 / 2 + 1;
        tw4f2.assign((size_t)RA2 * tws2, make_float4(0.f, 0.f, 0.f, 0.f));
        for (int jb = 0; jb < RA2; ++jb)
            for (int j = 0; j < RB2 / 2; ++j) {
                const double a0 = -2.0 * M_PI * (double)jb * (double)(2 * j) / (double)M;
                const double a1 = -2.0 * M_PI * (double)jb * (double)(2 * j + 1) / (double)M;
                tw4f2[(size_t)jb * tws2 + j] = make_float4((float)std::cos(a0), (float)std::cos(a1), (float)std::sin(a0), (float)std::sin(a1));
            }
    }
    int iRA = 0, iRB = 0;
    a2sb::inv_radices(M, iRA, iRB);
    // inter-pass twiddles of the inverse transform, in the orientation pass A uses them: row = residue ja (iRB rows),
    // entry k = outputs jb = 2k, 2k+1 of that residue's radix-iRA transform
    const int twsi = iRA / 2 + 1;
    std::vector<float4> tw4i((size_t)iRB * twsi, make_float4(0.f, 0.f, 0.f, 0.f));
    for (int ja = 0; ja < iRB; ++ja)
        for (int k = 0; k < iRA / 2; ++k) {
            const double a0 = 2.0 * M_PI * (double)ja * (double)(2 * k) / (double)M;
            const double a1 = 2.0 * M_PI * (double)ja * (double)(2 * k + 1) / (double)M;
            tw4i[(size_t)ja * twsi + k] = make_float4((float)std::cos(a0), (float)std::cos(a1), (float)std::sin(a0), (float)std::sin(a1));
        }
    pl->fwd_tile = (M <= 512) ? 32 : 16;   // n_fft 2048: the two-round 32-frame kernel is opt-in (measured slower: 1.10 vs 1.06 ms)
    if (const char* e = std::getenv("A2SB_FWD_TILE")) { const int v = std::atoi(e); pl->fwd_tile = (v == 8 || v == 32) ? v : 16; }
    pl->inv_tile = a2sb::inv_tile_frames(M);
    if (const char* e = std::getenv("A2SB_INV_TILE")) {
        int iRA0 = 0, iRB0 = 0;
        a2sb::inv_radices(M, iRA0, iRB0);
        if (std::atoi(e) == 8 && iRA0 % 32 == 0) pl->inv_tile = 8;
        if (std::atoi(e) == 32 && M <= 512) pl->inv_tile = 32;     // experiment: 32-frame inverse tiles for n_fft 512 / 1024
        if (std::atoi(e) == 16 && M < 2048) pl->inv_tile = 16;
    }
    auto up = [&](void** d, const void* h, size_t bytes) -> int {
        A2SB_CUDA(cudaMalloc(d, bytes));
        A2SB_CUDA(cudaMemcpy(*d, h, bytes, cudaMemcpyHostToDevice));
        return A2SB_OK;
    };
    int rc = A2SB_OK;
    if ((rc = up((void**)&pl->d_win_fwd, wf.data(), sizeof(float) * N)) ||
        (rc = up((void**)&pl->d_win_fwd_pcm, wfp.data(), sizeof(float) * N)) ||
        (rc = up((void**)&pl->d_win_inv, wi.data(), sizeof(float) * N)) ||
        (rc = up((void**)&pl->d_wsq, wsq.data(), sizeof(float) * N)) ||
        (rc = up((void**)&pl->d_inv_env, ienv.data(), sizeof(float) * hop)) ||
        (rc = up((void**)&pl->d_twM, twM.data(), sizeof(float2) * M)) ||
        (rc = up((void**)&pl->d_twN, twN.data(), sizeof(float2) * (M / 2 + 1))) ||
        (rc = up((void**)&pl->d_tw4f, tw4f.data(), sizeof(float4) * tw4f.size())) ||
        (rc = tw4f2.empty() ? A2SB_OK : up((void**)&pl->d_tw4f2, tw4f2.data(), sizeof(float4) * tw4f2.size())) ||
        (rc = (M >= 2048) ? up((void**)&pl->d_twS, twN.data(), sizeof(float2) * twN.size())
                          : up((void**)&pl->d_twS, twS.data(), sizeof(float4) * twS.size())) ||
        (rc = up((void**)&pl->d_tw4i, tw4i.data(), sizeof(float4) * tw4i.size()))) {
        a2sb_plan_destroy(pl);
        return rc;
    }
    *out = pl;
    return A2SB_OK;
}

int a2sb_plan_destroy(a2sb_plan* pl) {
    if (!pl) return A2SB_OK;
    cudaFree(pl->d_win_fwd); cudaFree(pl->d_win_fwd_pcm); cudaFree(pl->d_win_inv); cudaFree(pl->d_wsq);
    cudaFree(pl->d_inv_env); cudaFree(pl->d_twM); cudaFree(pl->d_twN); cudaFree(pl->d_tw4f); cudaFree(pl->d_tw4f2); cudaFree(pl->d_twS); cudaFree(pl->d_tw4i);
    for (auto& ln : pl->lanes) {
        cudaFree(ln.d_wav); cudaFree(ln.d_spec); cudaFree(ln.d_out);
#ifndef A2SB_EMU
        if (ln.stream) cudaStreamDestroy(ln.stream);
#endif
    }
    delete pl;
    return A2SB_OK;
}

}  // extern "C"

namespace {

using namespace a2sb;

// torch.istft: `window overlap add min` check over the trimmed envelope (ATen SpectralOps istft).
bool nola_ok(const a2sb_plan* pl, long long T) {
    const int N = pl->n_fft, H = pl->hop;
    const long long begin = N / 2, end = N / 2 + (long long)H * (T - 1);
    auto env = [&](long long J) {
        float e = 0.0f;
        long long t_hi = J / H;
        if (t_hi > T - 1) t_hi = T - 1;
        for (long long t = t_hi; t >= 0 && t * H + N > J; --t) e += pl->h_w[J - t * H] * pl->h_w[J - t * H];
        return e;
    };
    auto bad = [&](long long a, long long b) {
        for (long long J = a; J < b; ++J)
            if (std::fabs(env(J)) < 1e-11f) return true;
        return false;
    };
    if (end - begin <= 6LL * N) return !bad(begin, end);
    return !(bad(begin, begin + 2LL * N) || bad(begin + 2LL * N, begin + 2LL * N + H) || bad(end - 2LL * N, end));
}

}  // namespace

extern "C" {

static int forward_impl(a2sb_plan* pl, const a2sb_fwd_args* a, int pcm, const a2sb_corrupt_args* c);

int a2sb_stft_forward(a2sb_plan* pl, const a2sb_fwd_args* a) { return forward_impl(pl, a, 0, nullptr); }
int a2sb_stft_forward_pcm16(a2sb_plan* pl, const a2sb_fwd_args* a) { return forward_impl(pl, a, 1, nullptr); }
int a2sb_stft_forward_corrupt(a2sb_plan* pl, const a2sb_fwd_args* a, const a2sb_corrupt_args* c) {
    if (!c) return fail(A2SB_ERR_INVALID, "null corruption args");
    if (!c->d_out_corrupt || !c->d_noise) return fail(A2SB_ERR_INVALID, "null device pointer (corrupted output / noise)");
    return forward_impl(pl, a, 0, c);
}

static int forward_impl(a2sb_plan* pl, const a2sb_fwd_args* a, int pcm, const a2sb_corrupt_args* c) {
    if (!pl || !a) return fail(A2SB_ERR_INVALID, "null plan/args");
    if (a->batch < 0 || a->len < 0) return fail(A2SB_ERR_INVALID, "negative size");
    if (a->batch > 0x7fffffffLL) return fail(A2SB_ERR_INVALID, "batch %lld exceeds 2^31-1", (long long)a->batch);
    const int N = pl->n_fft, H = pl->hop;
    // torch.stft(center=True, pad_mode='reflect') requires pad < len (functional.py:508 -> F.pad):
    if (a->len <= N / 2)
        return fail(A2SB_ERR_INVALID,
                    "Argument #4: Padding size should be less than the corresponding input dimension, but got: "
                    "padding (%d, %d) at dimension 2 of input of length %lld",
                    N / 2, N / 2, (long long)a->len);
    const long long T = a2sb::num_frames(a->len, H);
    if (a->t_begin < 0 || a->t_end > T || a->t_begin > a->t_end)
        return fail(A2SB_ERR_INVALID, "frame range [%lld, %lld) outside [0, %lld)", (long long)a->t_begin,
                    (long long)a->t_end, T);
    if (a->out_kind != A2SB_KIND_COMPLEX && a->out_kind != A2SB_KIND_MAGPHASE)
        return fail(A2SB_ERR_INVALID, "bad out_kind %d", a->out_kind);
    if (a->batch == 0 || a->t_begin == a->t_end) return A2SB_OK;
    if (!a->d_wav || !a->d_out) return fail(A2SB_ERR_INVALID, "null device pointer");
    if (a->n_local < 1 || a->sample_first < 0 || a->sample_first + a->n_local > a->len)
        return fail(A2SB_ERR_INVALID, "local sample window [%lld, +%lld) outside the clip", (long long)a->sample_first,
                    (long long)a->n_local);
    // every sample the requested frames touch (after reflection) must be in the local buffer
    {
        long long lo = a->t_begin * H - N / 2, hi = (a->t_end - 1) * H + N / 2 - 1;
        long long need_lo = lo < 0 ? 0 : lo, need_hi = hi >= a->len ? a->len - 1 : hi;
        if (lo < 0 && -lo > need_hi) need_hi = -lo;
        if (hi >= a->len && 2 * (a->len - 1) - hi < need_lo) need_lo = 2 * (a->len - 1) - hi;
        if (need_lo < a->sample_first || need_hi >= a->sample_first + a->n_local)
            return fail(A2SB_ERR_INVALID, "frames [%lld, %lld) need samples [%lld, %lld] but the buffer holds [%lld, %lld)",
                        (long long)a->t_begin, (long long)a->t_end, need_lo, need_hi, (long long)a->sample_first,
                        (long long)(a->sample_first + a->n_local));
    }
    FwdParams p{};
    p.wav = a->d_wav; p.wav_stride = a->wav_stride; p.sample_first = a->sample_first; p.n_local = a->n_local;
    p.len = a->len; p.t_begin = a->t_begin; p.t_end = a->t_end;
    if (a->out_pitch != 0 && a->out_pitch < a->t_end - a->t_begin)
        return fail(A2SB_ERR_INVALID, "out_pitch %lld smaller than the %lld frames written per row", (long long)a->out_pitch,
                    (long long)(a->t_end - a->t_begin));
    p.out = a->d_out; p.out_T = a->out_pitch ? a->out_pitch : a->t_end - a->t_begin; p.out_t_first = a->t_begin;
    if (a->wrap_cols < 0 || a->wrap_cols > T) return fail(A2SB_ERR_INVALID, "wrap_cols %lld outside [0, %lld]", (long long)a->wrap_cols, T);
    if (a->wrap_cols > 0 && (a->t_begin != 0 || a->t_end != T || p.out_T < T + a->wrap_cols))
        return fail(A2SB_ERR_INVALID, "wrap_cols needs the whole clip in one launch and out_pitch >= T + wrap_cols (%lld < %lld)",
                    (long long)p.out_T, (long long)(T + a->wrap_cols));
    p.wrap_cols = (int)a->wrap_cols; p.wrap_at = a->wrap_cols > 0 ? T : 0;
    p.batch = (int)a->batch; p.hop = H;
    p.window = pcm ? pl->d_win_fwd_pcm : pl->d_win_fwd; p.tw4 = pl->d_tw4f; p.tw4_alt = pl->d_tw4f2; p.twS = pl->d_twS;
    p.pcm = pcm;
    if (c) {
        // rectangle in tensor coordinates with python slice semantics (negative bounds count from the end), like a2sb_rect_mask
        const long long n_rows = (a->out_kind == A2SB_KIND_MAGPHASE) ? pl->M + 1 - (a->drop_dc ? 1 : 0) : pl->M + 1, n_cols = a->t_end - a->t_begin;
        auto clampi = [](long long v, long long hi) { if (v < 0) v += hi; return v < 0 ? 0 : (v > hi ? hi : v); };
        p.out2 = c->d_out_corrupt; p.noise = c->d_noise; p.noise_T = c->noise_pitch ? c->noise_pitch : n_cols;
        if (p.noise_T < n_cols) return fail(A2SB_ERR_INVALID, "noise_pitch %lld smaller than the %lld frames per row", (long long)p.noise_T, n_cols);
        p.m_row0 = clampi(c->row0, n_rows); p.m_row1 = clampi(c->row1, n_rows);
        p.m_col0 = clampi(c->col0, n_cols); p.m_col1 = clampi(c->col1, n_cols);
        p.m_level = c->level;
    }
    p.epi = (a->out_kind == A2SB_KIND_MAGPHASE) ? kEpiMagPhase : kEpiComplex;
    p.drop_dc = (a->out_kind == A2SB_KIND_MAGPHASE) ? (a->drop_dc ? 1 : 0) : 0;
    p.pmode = (a->out_kind == A2SB_KIND_MAGPHASE && a->power_on) ? (a->power == 0.25f ? kPowQuarter : kPowGeneric) : kPowNone;
    p.power = a->power; p.eps = a->eps;
    cudaStream_t st = (cudaStream_t)a->stream;
    const a2sb::LaunchCtx cx{limited_sm_count(pl->sm_count, g_grid_limit_fwd.load()), pl->hop, pl->fwd_tile, pl->inv_tile};
    switch (pl->M) {
        case 256: return a2sb::run_fwd_256(cx, p, st);
        case 512: return a2sb::run_fwd_512(cx, p, st);
        case 1024: return a2sb::run_fwd_1024(cx, p, st);
        case 2048: return a2sb::run_fwd_2048(cx, p, st);
    }
    return fail(A2SB_ERR_INVALID, "unsupported n_fft");
}

static int inverse_impl(a2sb_plan* pl, const a2sb_inv_args* a, int mirror_mode, int n_mirrors, float* const* d_mirrors);

int a2sb_istft_inverse(a2sb_plan* pl, const a2sb_inv_args* a) { return inverse_impl(pl, a, 0, 0, nullptr); }
int a2sb_istft_inverse_pcm16(a2sb_plan* pl, const a2sb_inv_args* a) { return inverse_impl(pl, a, -1, 0, nullptr); }

int a2sb_istft_inverse_mirrored(a2sb_plan* pl, const a2sb_inv_args* a, int mode, int n_mirrors, float* const* d_mirrors) {
    if (mode != A2SB_MIRROR_PEERS && mode != A2SB_MIRROR_MULTICAST) return fail(A2SB_ERR_INVALID, "bad mirror mode %d", mode);
    if (n_mirrors < 1 || n_mirrors > 8 || !d_mirrors) return fail(A2SB_ERR_INVALID, "1..8 mirror buffers expected, got %d", n_mirrors);
    if (mode == A2SB_MIRROR_MULTICAST && n_mirrors != 1) return fail(A2SB_ERR_INVALID, "multicast mode takes ONE (multicast) address");
    for (int i = 0; i < n_mirrors; ++i)
        if (!d_mirrors[i] || (reinterpret_cast<uintptr_t>(d_mirrors[i]) & 15) != (reinterpret_cast<uintptr_t>(a ? a->d_wav : nullptr) & 15))
            return fail(A2SB_ERR_INVALID, "mirror %d is null or not aligned like d_wav (mod 16 bytes)", i);
    return inverse_impl(pl, a, mode, n_mirrors, d_mirrors);
}

static int inverse_impl(a2sb_plan* pl, const a2sb_inv_args* a, int mirror_mode, int n_mirrors, float* const* d_mirrors) {
    if (!pl || !a) return fail(A2SB_ERR_INVALID, "null plan/args");
    if (!pl->inverse_ok)
        return fail(A2SB_ERR_INVALID, "hop_length=%d: the inverse transform needs a multiple of 4 that divides n_fft=%d", pl->hop, pl->n_fft);
    const int N = pl->n_fft, H = pl->hop, ROV = N / H;
    const long long T = a->n_frames;
    if (a->batch < 0 || T < 1) return fail(A2SB_ERR_INVALID, "bad sizes (batch %lld, frames %lld)", (long long)a->batch, T);
    if (a->batch > 0x7fffffffLL) return fail(A2SB_ERR_INVALID, "batch %lld exceeds 2^31-1", (long long)a->batch);
    if (a->in_kind != A2SB_KIND_COMPLEX && a->in_kind != A2SB_KIND_MAGPHASE)
        return fail(A2SB_ERR_INVALID, "bad in_kind %d", a->in_kind);
    if (!nola_ok(pl, T)) return fail(A2SB_ERR_NOLA, "window overlap add min: 1");
    const long long total_out = (long long)H * (T - 1);
    if (a->out_first < 0 || a->out_count < 0 || a->out_first + a->out_count > total_out)
        return fail(A2SB_ERR_INVALID, "output range [%lld, +%lld) outside [0, %lld)", (long long)a->out_first,
                    (long long)a->out_count, total_out);
    if (a->out_first % H != 0 || (a->out_count % H != 0 && a->out_first + a->out_count != total_out))
        return fail(A2SB_ERR_INVALID, "sharded output ranges must be multiples of hop_length");
    if (a->batch == 0 || a->out_count == 0) return A2SB_OK;
    if (!a->d_spec || !a->d_wav) return fail(A2SB_ERR_INVALID, "null device pointer");
    // untrimmed sample J = out + N/2 lives in hop-block J / H
    const long long hop_begin = (a->out_first + N / 2) / H;
    const long long hop_end = (a->out_first + a->out_count + N / 2 + H - 1) / H;
    {   // frames needed: [hop_begin - (ROV-1), hop_end - 1] clipped to [0, T)
        long long f_lo = hop_begin - (ROV - 1), f_hi = hop_end - 1;
        if (f_lo < 0) f_lo = 0;
        if (f_hi > T - 1) f_hi = T - 1;
        if (f_lo < a->spec_t_first || f_hi >= a->spec_t_first + a->spec_T)
            return fail(A2SB_ERR_INVALID, "output needs frames [%lld, %lld] but the buffer holds [%lld, %lld)", f_lo, f_hi,
                        (long long)a->spec_t_first, (long long)(a->spec_t_first + a->spec_T));
    }
    InvParams p{};
    p.spec = a->d_spec; p.spec_T = a->spec_T; p.spec_t_first = a->spec_t_first; p.n_frames = T;
    p.out = a->d_wav; p.out_stride = a->wav_stride; p.out_first = a->out_first; p.out_count = a->out_count;
    p.hop_begin = hop_begin; p.hop_end = hop_end;
    p.batch = (int)a->batch; p.hop = H;
    // Chunking: a work item is m tiles (m*16 frames, of which N/hop - 1 are warm-up frames recomputed to
    // seed the overlap-add).  Items are dealt round-robin, so the CTAs of the grid read NEIGHBOURING frame
    // ranges of the same rows at the same time; measured on 256 x 10 s clips: m = 16 -> 1.79 ms,
    // m = 4 -> 1.63 ms, m = 2 -> 1.43 ms (10% recompute), m = 1 -> 1.59 ms (23% recompute): DRAM page and
    // L2 sector locality of the 64-byte row segments outweighs the recompute.  Re-measured on the final kernel:
    // m = 2 -> 1.076 ms, 3 -> 1.26, 4 -> 1.25, 6 and 8 -> 1.43.
    const long long HT = hop_end - hop_begin;
    const int kF = pl->inv_tile;
    // 32 frames per item for the 16-frame kernels; the 8-frame kernel (n_fft 4096) takes 16: its CTAs come back to the next 32
    // bytes of a row only a tile later, when the line L2 fetched for the first 32 is gone again -- ncu 9.6 GB of DRAM reads for
    // 2.7 GB of spectrogram at 4 tiles per item; measured 1 / 2 / 3 / 4 / 8 tiles: 2.10 / 1.58 / 1.59 / 1.95 / 2.51 ms
    // (the L2 promotion size -- none / 128 B / 256 B -- changes nothing)
    int m_best = (kF == 8) ? 2 : 32 / kF;
    if ((long long)m_best * kF - (ROV - 1) < 1) m_best = (ROV - 1) / kF + 1;
    static const int env_m = [] { const char* e = std::getenv("A2SB_INV_M"); return e ? std::atoi(e) : 0; }();
    if (env_m >= 1 && env_m <= 64) m_best = env_m;   // experiments
    long long ch = (long long)m_best * kF - (ROV - 1);
    if (ch < 1) return fail(A2SB_ERR_INVALID, "n_fft / hop_length = %d too large for the fused inverse kernel", ROV);
    p.chunk_hops = (int)ch;
    if ((HT + ch - 1) / ch > 0x7fffffffLL) return fail(A2SB_ERR_INVALID, "too many hop-blocks per clip (%lld)", HT);
    p.chunks_per_clip = (int)((HT + ch - 1) / ch);
    p.total_items = (long long)p.chunks_per_clip * a->batch;
    p.window = pl->d_win_inv; p.wsq = pl->d_wsq; p.inv_env = pl->d_inv_env; p.tw4 = pl->d_tw4i; p.twN = pl->d_twN;
    p.in_kind = (a->in_kind == A2SB_KIND_MAGPHASE) ? kInMagPhase : kInComplex;
    p.has_dc = (p.in_kind == kInMagPhase) ? (a->has_dc ? 1 : 0) : 1;
    p.svd_fix = (p.in_kind == kInMagPhase && a->phase_fix) ? 1 : 0;
    p.pmode = (p.in_kind == kInMagPhase && a->power_on) ? (a->power == 4.0f ? kPowFour : kPowGeneric) : kPowNone;
    p.power = a->power; p.eps = a->eps;
    p.n_mirror = n_mirrors; p.mirror_mc = (mirror_mode == A2SB_MIRROR_MULTICAST) ? 1 : 0; p.out_pcm = (mirror_mode == -1) ? 1 : 0;
    for (int i = 0; i < n_mirrors; ++i) p.mirror[i] = d_mirrors[i];
    cudaStream_t st = (cudaStream_t)a->stream;
    const a2sb::LaunchCtx cx{limited_sm_count(pl->sm_count, g_grid_limit_inv.load()), pl->hop, pl->fwd_tile, pl->inv_tile};
    switch (pl->M) {
        case 256: return a2sb::run_inv_256(cx, p, st);
        case 512: return a2sb::run_inv_512(cx, p, st);
        case 1024: return a2sb::run_inv_1024(cx, p, st);
        case 2048: return a2sb::run_inv_2048(cx, p, st);
    }
    return fail(A2SB_ERR_INVALID, "unsupported n_fft");
}

// ---- any-length STFT / iSTFT (dft_generic.cuh) ----
static int gen_check(int n_fft, int hop, int64_t batch) {
    if (n_fft < 2 || n_fft > 8192) return fail(A2SB_ERR_INVALID, "n_fft=%d outside [2, 8192]", n_fft);
    if (hop < 1) return fail(A2SB_ERR_INVALID, "hop_length=%d must be positive", hop);
    if (batch < 0 || batch > 65535) return fail(A2SB_ERR_INVALID, "batch %lld outside [0, 65535]", (long long)batch);
    return A2SB_OK;
}

int a2sb_dft_generic_forward(const float* d_wav, int64_t batch, int64_t len, int64_t wav_stride, int n_fft, int hop,
                             const float* d_window, float* d_spec, void* stream) {
    if (int rc = gen_check(n_fft, hop, batch)) return rc;
    if (len <= n_fft / 2)
        return fail(A2SB_ERR_INVALID, "Argument #4: Padding size should be less than the corresponding input dimension, but got: "
                    "padding (%d, %d) at dimension 2 of input of length %lld", n_fft / 2, n_fft / 2, (long long)len);
    if (batch == 0) return A2SB_OK;
    if (!d_wav || !d_window || !d_spec) return fail(A2SB_ERR_INVALID, "null device pointer");
    a2sb::GenParams p{};
    p.wav = d_wav; p.wav_stride = wav_stride; p.len = len; p.window = d_window; p.spec = d_spec;
    p.N = n_fft; p.K = n_fft / 2 + 1; p.hop = hop; p.batch = (int)batch;
    p.T = 1 + (len + 2 * (n_fft / 2) - n_fft) / hop;        // torch.stft(center=True)
    const size_t smem = sizeof(float2) * n_fft + sizeof(float) * a2sb::kGenTF * n_fft;
    const dim3 grid((unsigned)((p.T + a2sb::kGenTF - 1) / a2sb::kGenTF), (unsigned)((p.K + a2sb::kGenNT - 1) / a2sb::kGenNT), (unsigned)batch);
#ifdef A2SB_EMU
    for (unsigned z = 0; z < grid.z; ++z) for (unsigned y = 0; y < grid.y; ++y) for (unsigned x = 0; x < grid.x; ++x)
        emu::launch_at(dim3(x, y, z), grid, dim3(a2sb::kGenNT), smem, [&] { a2sb::dft_fwd_generic_kernel(p); });
#else
    if (smem > 48 * 1024) A2SB_CUDA(cudaFuncSetAttribute(a2sb::dft_fwd_generic_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    a2sb::dft_fwd_generic_kernel<<<grid, a2sb::kGenNT, smem, (cudaStream_t)stream>>>(p);
    A2SB_CUDA(cudaGetLastError());
#endif
    g_launches.fetch_add(1);
    return A2SB_OK;
}

int a2sb_dft_generic_inverse(const float* d_spec, int64_t batch, int64_t n_frames, int n_fft, int hop, const float* d_window,
                             float* d_frames, float* d_out, int64_t out_len, void* stream) {
    if (int rc = gen_check(n_fft, hop, batch)) return rc;
    if (n_frames < 1 || out_len < 0) return fail(A2SB_ERR_INVALID, "bad sizes (frames %lld, out_len %lld)", (long long)n_frames, (long long)out_len);
    if (batch == 0 || out_len == 0) return A2SB_OK;
    if (!d_spec || !d_window || !d_frames || !d_out) return fail(A2SB_ERR_INVALID, "null device pointer");
    a2sb::GenParams p{};
    p.spec = const_cast<float*>(d_spec); p.window = d_window; p.frames = d_frames; p.out = d_out; p.out_len = out_len;
    p.N = n_fft; p.K = n_fft / 2 + 1; p.hop = hop; p.batch = (int)batch; p.T = n_frames;
    const size_t smem = sizeof(float2) * n_fft + sizeof(float2) * a2sb::kGenTF * p.K;
    const dim3 grid((unsigned)((p.T + a2sb::kGenTF - 1) / a2sb::kGenTF), (unsigned)((n_fft + a2sb::kGenNT - 1) / a2sb::kGenNT), (unsigned)batch);
#ifdef A2SB_EMU
    for (unsigned z = 0; z < grid.z; ++z) for (unsigned y = 0; y < grid.y; ++y) for (unsigned x = 0; x < grid.x; ++x)
        emu::launch_at(dim3(x, y, z), grid, dim3(a2sb::kGenNT), smem, [&] { a2sb::dft_inv_generic_frames_kernel(p); });
    emu::launch(dim3(2), dim3(256), 0, [&] { a2sb::dft_inv_generic_ola_kernel(p); });
#else
    if (smem > 48 * 1024) A2SB_CUDA(cudaFuncSetAttribute(a2sb::dft_inv_generic_frames_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    a2sb::dft_inv_generic_frames_kernel<<<grid, a2sb::kGenNT, smem, (cudaStream_t)stream>>>(p);
    A2SB_CUDA(cudaGetLastError());
    const long long total = (long long)batch * out_len;
    long long blocks = (total + 255) / 256;
    const long long cap = (long long)device_sm_count() * 8;
    if (blocks > cap) blocks = cap;
    a2sb::dft_inv_generic_ola_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(p);
    A2SB_CUDA(cudaGetLastError());
#endif
    g_launches.fetch_add(2);
    return A2SB_OK;
}

int a2sb_pointwise(int op, const float* d_in, float* d_out, int64_t n, int channels, uint32_t chan_mask, float power,
                   float eps, void* stream) {
    if (op < 0 || op > 5) return fail(A2SB_ERR_INVALID, "bad pointwise op %d", op);
    if (n < 0 || channels < 1 || channels > 32) return fail(A2SB_ERR_INVALID, "bad sizes");
    if (n == 0) return A2SB_OK;
    if (!d_in || !d_out) return fail(A2SB_ERR_INVALID, "null device pointer");
    PwParams p{d_in, d_out, (long long)n, channels, chan_mask, power, eps, op};
    return launch_grid_stride(pointwise_kernel, (long long)n, (cudaStream_t)stream, p, device_sm_count());
}

int a2sb_griffinlim_update(const float* d_rebuilt, const float* d_tprev, const float* d_mag, float* d_product, int64_t batch,
                           int64_t n, float momentum, void* stream) {
    if (n < 0 || batch < 0) return fail(A2SB_ERR_INVALID, "negative size");
    if (n == 0 || batch == 0) return A2SB_OK;
    if (!d_rebuilt || !d_mag || !d_product) return fail(A2SB_ERR_INVALID, "null device pointer");
    GlParams p{d_rebuilt, d_tprev, d_mag, d_product, (long long)n, (long long)batch * n, momentum};
    return launch_grid_stride(griffinlim_update_kernel, p.total, (cudaStream_t)stream, p, device_sm_count());
}

int a2sb_wrap_pad(const float* d_in, float* d_out, int64_t nrows, int64_t width, int64_t out_width, int use_const,
                  float pad_const, void* stream) {
    if (nrows < 0 || width < 1 || out_width < width || out_width - width > width)
        return fail(A2SB_ERR_INVALID, "bad pad geometry (width %lld -> %lld)", (long long)width, (long long)out_width);
    PadParams p{d_in, d_out, (long long)nrows, (long long)width, (long long)out_width, use_const, pad_const,
                (long long)nrows * out_width};
    if (p.total == 0) return A2SB_OK;
    if (!d_in || !d_out) return fail(A2SB_ERR_INVALID, "null device pointer");
    if (p.out_width % 4 == 0 && (reinterpret_cast<uintptr_t>(d_out) & 15) == 0) {
        p.total /= 4;
        p.d_ow = a2sb::make_divmod(p.out_width / 4);
        return p.total < (1LL << 31) ? launch_grid_stride(wrap_pad_kernel<4, true>, p.total, (cudaStream_t)stream, p, device_sm_count())
                                     : launch_grid_stride(wrap_pad_kernel<4, false>, p.total, (cudaStream_t)stream, p, device_sm_count());
    }
    p.d_ow = a2sb::make_divmod(p.out_width);
    return p.total < (1LL << 31) ? launch_grid_stride(wrap_pad_kernel<1, true>, p.total, (cudaStream_t)stream, p, device_sm_count())
                                 : launch_grid_stride(wrap_pad_kernel<1, false>, p.total, (cudaStream_t)stream, p, device_sm_count());
}

static int seg_common(SegParams& p, const float* in, float* out, int64_t batch, int64_t rows, int64_t width, int win,
                      int hop) {
    if (batch < 0 || rows < 0 || width < 1 || win < 1 || hop < 1 || hop > win)
        return fail(A2SB_ERR_INVALID, "bad segment geometry (width %lld, win %d, hop %d)", (long long)width, win, hop);
    if (batch > 0x7fffffffLL) return fail(A2SB_ERR_INVALID, "batch %lld exceeds 2^31-1", (long long)batch);
    p.in = in; p.out = out; p.rows = rows; p.width = width; p.batch = (int)batch; p.win = win; p.hop = hop;
    p.num_hops = (width - (win - hop)) / hop;  // diffusion.py:33
    if (width < win) p.num_hops = 0;
    p.col_off = 0; p.col_cnt = width; p.out_pitch = width;
    return A2SB_OK;
}

// launch-invariant divisors of the index decomposition; returns true when every dividend fits 31 bits
static bool seg_divisors(SegParams& p, int vec, long long total) {
    p.d_wv = a2sb::make_divmod(p.win / vec);
    p.d_cv = a2sb::make_divmod(p.width / vec);
    p.d_rows = a2sb::make_divmod(p.rows);
    p.d_hops = a2sb::make_divmod(p.num_hops);
    p.d_hop = a2sb::make_divmod(p.hop);
    return total < (1LL << 31) && p.width < (1LL << 30) && p.rows < (1LL << 31);
}

#define A2SB_SEG_DISPATCH(KERN, PARAMS, SEGP, V4, F32, ST)                                                                   \
    ((V4) ? ((F32) ? launch_grid_stride(KERN<4, true>, (SEGP).total, (ST), (PARAMS), device_sm_count())                      \
                   : launch_grid_stride(KERN<4, false>, (SEGP).total, (ST), (PARAMS), device_sm_count()))                    \
          : ((F32) ? launch_grid_stride(KERN<1, true>, (SEGP).total, (ST), (PARAMS), device_sm_count())                      \
                   : launch_grid_stride(KERN<1, false>, (SEGP).total, (ST), (PARAMS), device_sm_count())))

int a2sb_segment_gather(const float* d_x, float* d_seg, int64_t batch, int64_t rows, int64_t width, int win, int hop,
                        void* stream) {
    SegParams p{};
    if (int rc = seg_common(p, d_x, d_seg, batch, rows, width, win, hop)) return rc;
    const long long elems = (long long)batch * p.num_hops * rows * win;
    if (elems == 0) return A2SB_OK;
    if (!d_x || !d_seg) return fail(A2SB_ERR_INVALID, "null device pointer");
    const bool v4 = win % 4 == 0 && hop % 4 == 0 && width % 4 == 0 && aligned16(d_x) && aligned16(d_seg);
    p.total = v4 ? elems / 4 : elems;
    const bool f32 = seg_divisors(p, v4 ? 4 : 1, p.total);
    return A2SB_SEG_DISPATCH(segment_gather_kernel, p, p, v4, f32, (cudaStream_t)stream);
}

int a2sb_segment_blend(const float* d_seg, float* d_out, int64_t batch, int64_t rows, int64_t width, int win, int hop,
                       void* stream) {
    return a2sb_segment_blend_window(d_seg, d_out, batch, rows, width, win, hop, 0, width, width, stream);
}

int a2sb_segment_blend_window(const float* d_seg, float* d_out, int64_t batch, int64_t rows, int64_t width, int win, int hop,
                              int64_t col_off, int64_t col_cnt, int64_t out_pitch, void* stream) {
    SegParams p{};
    if (int rc = seg_common(p, d_seg, d_out, batch, rows, width, win, hop)) return rc;
    if (col_off < 0 || col_cnt < 0 || col_off + col_cnt > width || out_pitch < col_cnt)
        return fail(A2SB_ERR_INVALID, "bad blend window (columns [%lld, +%lld) of %lld, pitch %lld)", (long long)col_off,
                    (long long)col_cnt, (long long)width, (long long)out_pitch);
    p.col_off = col_off; p.col_cnt = col_cnt; p.out_pitch = out_pitch;
    const long long elems = (long long)batch * rows * col_cnt;
    if (elems == 0) return A2SB_OK;
    if (!d_seg || !d_out) return fail(A2SB_ERR_INVALID, "null device pointer");
    const bool v4 = win % 4 == 0 && hop % 4 == 0 && width % 4 == 0 && col_off % 4 == 0 && col_cnt % 4 == 0 && out_pitch % 4 == 0 &&
                    aligned16(d_seg) && aligned16(d_out);
    p.total = v4 ? elems / 4 : elems;
    const bool f32 = seg_divisors(p, v4 ? 4 : 1, p.total);
    p.d_cv = a2sb::make_divmod(col_cnt / (v4 ? 4 : 1));
    return A2SB_SEG_DISPATCH(segment_blend_kernel, p, p, v4, f32, (cudaStream_t)stream);
}

int a2sb_segment_blend_step(const float* d_seg, const a2sb_step_args* a, int64_t batch, int64_t rows, int64_t width, int win,
                            int hop, void* stream) {
    if (!a) return fail(A2SB_ERR_INVALID, "null args");
    StepParams sp{};
    if (int rc = seg_common(sp.seg, d_seg, nullptr, batch, rows, width, win, hop)) return rc;
    const long long elems = (long long)batch * rows * width;
    if (elems == 0) return A2SB_OK;
    if (!d_seg || !a->d_x_t || !a->d_x_1 || !a->d_pred_x0 || !a->d_x_next) return fail(A2SB_ERR_INVALID, "null device pointer");
    sp.x_t = a->d_x_t; sp.x_1 = a->d_x_1; sp.mask = a->d_mask; sp.noise_post = a->d_noise_post; sp.noise_mask = a->d_noise_mask;
    sp.pred_x0 = a->d_pred_x0; sp.x_next = a->d_x_next;
    sp.std_fwd_t = a->std_fwd_t; sp.mu_x0 = a->mu_x0; sp.mu_xt = a->mu_xt; sp.sd_post = a->sd_post; sp.std_sb = a->std_sb;
    sp.mask_pred_x0 = a->mask_pred_x0;
    const bool v4 = win % 4 == 0 && hop % 4 == 0 && width % 4 == 0 && aligned16(d_seg) && aligned16(a->d_pred_x0) &&
                    aligned16(a->d_x_next) && aligned16(a->d_x_t) && aligned16(a->d_x_1) && aligned16(a->d_mask) &&
                    aligned16(a->d_noise_post) && aligned16(a->d_noise_mask);
    sp.seg.total = v4 ? elems / 4 : elems;
    const bool f32 = seg_divisors(sp.seg, v4 ? 4 : 1, sp.seg.total);
    return A2SB_SEG_DISPATCH(segment_blend_step_kernel, sp, sp.seg, v4, f32, (cudaStream_t)stream);
}

static int mask_common(MaskParams& p, int64_t slices, int64_t rows, int64_t width, int64_t row0, int64_t row1, int64_t col0,
                       int64_t col1) {
    if (slices < 0 || rows < 0 || width < 0) return fail(A2SB_ERR_INVALID, "negative size");
    // python slice semantics of the reference (mask[:, a:b, c:d] = 1): a negative bound counts from the end
    // (corruptions.py:155-158 slices with whatever int() produced), then clamp into range; empty if reversed
    auto clampi = [](int64_t v, int64_t hi) { if (v < 0) v += hi; return v < 0 ? 0 : (v > hi ? hi : v); };
    p.rows = rows; p.width = width;
    p.row0 = clampi(row0, rows); p.row1 = clampi(row1, rows);
    p.col0 = clampi(col0, width); p.col1 = clampi(col1, width);
    p.total = (long long)slices * rows * width;
    p.d_width = a2sb::make_divmod(width);
    p.d_rows = a2sb::make_divmod(rows);
    return A2SB_OK;
}

int a2sb_rect_mask(float* d_mask, int64_t slices, int64_t rows, int64_t width, int64_t row0, int64_t row1, int64_t col0,
                   int64_t col1, void* stream) {
    MaskParams p{};
    if (int rc = mask_common(p, slices, rows, width, row0, row1, col0, col1)) return rc;
    if (p.total == 0) return A2SB_OK;
    if (!d_mask) return fail(A2SB_ERR_INVALID, "null device pointer");
    p.mask_out = d_mask;
    return p.total < (1LL << 31) ? launch_grid_stride(rect_mask_kernel<true>, p.total, (cudaStream_t)stream, p, device_sm_count())
                                 : launch_grid_stride(rect_mask_kernel<false>, p.total, (cudaStream_t)stream, p, device_sm_count());
}

int a2sb_mask_with_noise(const float* d_x, const float* d_mask, const float* d_noise, float* d_out, int64_t n, float level,
                         void* stream) {
    if (n < 0) return fail(A2SB_ERR_INVALID, "negative size");
    if (n == 0) return A2SB_OK;
    if (!d_x || !d_mask || !d_noise || !d_out) return fail(A2SB_ERR_INVALID, "null device pointer");
    MaskParams p{};
    p.x = d_x; p.mask_in = d_mask; p.noise = d_noise; p.out = d_out; p.total = n; p.level = level;
    p.rows = 1; p.width = n;
    return launch_grid_stride(mask_noise_kernel, p.total, (cudaStream_t)stream, p, device_sm_count());
}

int a2sb_mask_fill(const float* d_x, const float* d_noise, float* d_out, float* d_mask, int64_t slices, int64_t rows,
                   int64_t width, int64_t row0, int64_t row1, int64_t col0, int64_t col1, float level, void* stream) {
    MaskParams p{};
    if (int rc = mask_common(p, slices, rows, width, row0, row1, col0, col1)) return rc;
    if (p.total == 0) return A2SB_OK;
    if (!d_x || !d_noise || !d_out) return fail(A2SB_ERR_INVALID, "null device pointer");
    p.x = d_x; p.noise = d_noise; p.out = d_out; p.mask_out = d_mask; p.level = level;
    const bool v4 = width % 4 == 0 && aligned16(d_x) && aligned16(d_noise) && aligned16(d_out) && aligned16(d_mask);
    if (v4) {
        p.total /= 4;
        p.d_width = a2sb::make_divmod(width / 4);
        return p.total < (1LL << 31) ? launch_grid_stride(mask_fill_kernel<4, true>, p.total, (cudaStream_t)stream, p, device_sm_count())
                                     : launch_grid_stride(mask_fill_kernel<4, false>, p.total, (cudaStream_t)stream, p, device_sm_count());
    }
    return p.total < (1LL << 31) ? launch_grid_stride(mask_fill_kernel<1, true>, p.total, (cudaStream_t)stream, p, device_sm_count())
                                 : launch_grid_stride(mask_fill_kernel<1, false>, p.total, (cudaStream_t)stream, p, device_sm_count());
}

int a2sb_mask_fill_padded(const float* d_x, int64_t in_pitch, const float* d_noise, float* d_out, float* d_mask, int64_t slices,
                          int64_t rows, int64_t width, int64_t out_width, int64_t row0, int64_t row1, int64_t col0, int64_t col1,
                          float level, void* stream) {
    MaskPadParams q{};
    if (int rc = mask_common(q.m, slices, rows, width, row0, row1, col0, col1)) return rc;
    if (in_pitch < width || out_width < width || out_width - width > width)
        return fail(A2SB_ERR_INVALID, "bad padded mask geometry (width %lld, in pitch %lld, out width %lld)", (long long)width,
                    (long long)in_pitch, (long long)out_width);
    if (q.m.total == 0) return A2SB_OK;
    if (!d_x || !d_noise || !d_out) return fail(A2SB_ERR_INVALID, "null device pointer");
    q.m.x = d_x; q.m.noise = d_noise; q.m.out = d_out; q.m.mask_out = d_mask; q.m.level = level;
    q.in_pitch = in_pitch; q.out_width = out_width;
    auto al8 = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 7) == 0; };
    const bool v2 = width % 2 == 0 && in_pitch % 2 == 0 && out_width % 2 == 0 && al8(d_x) && al8(d_noise) && al8(d_out) && al8(d_mask);
    if (v2) {
        q.m.total /= 2;
        q.m.d_width = a2sb::make_divmod(width / 2);
        return q.m.total < (1LL << 31) ? launch_grid_stride(mask_fill_padded_kernel<2, true>, q.m.total, (cudaStream_t)stream, q, device_sm_count())
                                       : launch_grid_stride(mask_fill_padded_kernel<2, false>, q.m.total, (cudaStream_t)stream, q, device_sm_count());
    }
    return q.m.total < (1LL << 31) ? launch_grid_stride(mask_fill_padded_kernel<1, true>, q.m.total, (cudaStream_t)stream, q, device_sm_count())
                                   : launch_grid_stride(mask_fill_padded_kernel<1, false>, q.m.total, (cudaStream_t)stream, q, device_sm_count());
}

int a2sb_zero_segment_windows(const float* d_row, int64_t n, int win_length, int32_t* d_centres, int32_t* d_lr,
                              int32_t* d_count, int max_out, void* stream) {
    if (n < 1) return fail(A2SB_ERR_INVALID, "Input must be a non-empty 1D tensor.");
    if (n > 0x7fffffffLL) return fail(A2SB_ERR_INVALID, "row too long (%lld)", (long long)n);
    if (win_length < 1 || max_out < 1) return fail(A2SB_ERR_INVALID, "bad win_length / max_out");
    if (!d_row || !d_centres || !d_lr || !d_count) return fail(A2SB_ERR_INVALID, "null device pointer");
    ZeroSegParams p{d_row, (long long)n, win_length, d_centres, d_lr, d_count, max_out};
#ifdef A2SB_EMU
    emu::launch(dim3(1), dim3(kZeroSegThreads), 0, [&] { zero_segment_kernel(p); });
    (void)stream;
#else
    zero_segment_kernel<<<1, kZeroSegThreads, 0, (cudaStream_t)stream>>>(p);
    A2SB_CUDA(cudaGetLastError());
#endif
    g_launches.fetch_add(1);
    return A2SB_OK;
}

// Not thread-safe per plan: the staging lanes belong to the plan (one caller at a time; the Python wrapper holds a lock).
static int roundtrip_host_impl(a2sb_plan* pl, const void* h_wav, int64_t batch, int64_t len, void* h_wav_out, float* h_spec,
                               float power_fwd, float power_inv, float eps, int phase_fix, int pcm);
static int roundtrip_host_any(a2sb_plan* pl, const void* h_wav, int64_t batch, int64_t len, void* h_wav_out, float* h_spec,
                              float power_fwd, float power_inv, float eps, int phase_fix, int pcm);

int a2sb_roundtrip_host(a2sb_plan* pl, const float* h_wav, int64_t batch, int64_t len, float* h_wav_out, float* h_spec,
                        float power_fwd, float power_inv, float eps, int phase_fix) {
    return roundtrip_host_any(pl, h_wav, batch, len, h_wav_out, h_spec, power_fwd, power_inv, eps, phase_fix, 0);
}

int a2sb_roundtrip_host_pcm16(a2sb_plan* pl, const int16_t* h_pcm, int64_t batch, int64_t len, int16_t* h_pcm_out, float* h_spec,
                              float power_fwd, float power_inv, float eps, int phase_fix) {
    return roundtrip_host_any(pl, h_pcm, batch, len, h_pcm_out, h_spec, power_fwd, power_inv, eps, phase_fix, 1);
}

static int roundtrip_host_any(a2sb_plan* pl, const void* h_wav, int64_t batch, int64_t len, void* h_wav_out, float* h_spec,
                              float power_fwd, float power_inv, float eps, int phase_fix, int pcm) {
    if (!pl) return fail(A2SB_ERR_INVALID, "null plan");
    const int rc = roundtrip_host_impl(pl, h_wav, batch, len, h_wav_out, h_spec, power_fwd, power_inv, eps, phase_fix, pcm);
    if (rc != A2SB_OK) {
        // copies into the caller's host buffers may still be in flight: drain the lanes before reporting the error
        const std::string keep = a2sb::g_err;
        for (auto& ln : pl->lanes)
            if (ln.stream) cudaStreamSynchronize(ln.stream);
        a2sb::g_err = keep;
    }
    return rc;
}

static int roundtrip_host_impl(a2sb_plan* pl, const void* h_wav_v, int64_t batch, int64_t len, void* h_wav_out_v, float* h_spec,
                               float power_fwd, float power_inv, float eps, int phase_fix, int pcm) {
    // pcm: both host buffers hold 16-bit PCM; the lanes' sample buffers are then half as large (allocated for float32)
    const char* h_wav = static_cast<const char*>(h_wav_v);
    char* h_wav_out = static_cast<char*>(h_wav_out_v);
    const size_t sb = pcm ? sizeof(short) : sizeof(float);
    if (batch <= 0) return A2SB_OK;
    if (!h_wav || !h_wav_out) return fail(A2SB_ERR_INVALID, "null host pointer");
    const int H = pl->hop, M = pl->M;
    if (!pl->inverse_ok) return fail(A2SB_ERR_INVALID, "forward-only plan (hop_length=%d)", pl->hop);
    if (len <= pl->n_fft / 2) return fail(A2SB_ERR_INVALID, "clip shorter than n_fft/2");
    const long long T = a2sb::num_frames(len, H), out_len = (long long)H * (T - 1);
    const long long spec_clip = 3LL * M * T;
    // clip groups sized to ~128 MB of spectrogram so copies and kernels of different groups overlap
    long long group_mb = 128;   // measured 64: 11.7 ms, 128: 10.9, 256: 11.2, 512: 11.5 per 256-clip step (PCIe floor 9.1)
    if (const char* e = std::getenv("A2SB_E2E_GROUP_MB")) { const long long v = std::atoll(e); if (v >= 8 && v <= 8192) group_mb = v; }
    long long group = (group_mb << 20) / (spec_clip * (long long)sizeof(float));
    if (group < 1) group = 1;
    if (group > batch) group = batch;
    if (const char* e = std::getenv("A2SB_E2E_LANES")) { const int v = std::atoi(e); if (v >= 1 && v <= 8) pl->n_lanes = v; }
    const int n_lanes = pl->n_lanes;
    for (int l = 0; l < n_lanes; ++l) {
        auto& ln = pl->lanes[l];
#ifndef A2SB_EMU
        if (!ln.stream) A2SB_CUDA(cudaStreamCreateWithFlags(&ln.stream, cudaStreamNonBlocking));
#endif
        if (ln.clips < group || ln.len != len) {
            cudaFree(ln.d_wav); cudaFree(ln.d_spec); cudaFree(ln.d_out);
            ln.d_wav = ln.d_spec = ln.d_out = nullptr;
            ln.clips = 0; ln.len = 0;   // a failed cudaMalloc below must not leave a lane that passes the reuse check
            A2SB_CUDA(cudaMalloc((void**)&ln.d_wav, sizeof(float) * group * len));
            A2SB_CUDA(cudaMalloc((void**)&ln.d_spec, sizeof(float) * group * spec_clip));
            A2SB_CUDA(cudaMalloc((void**)&ln.d_out, sizeof(float) * group * out_len));
            ln.clips = group; ln.len = len;
        }
    }
    // Group schedule: ceil(batch / group) groups of (almost) equal size, round-robin over the lanes.  Measured and rejected
    // (256 x 10 s clips, 10.67 ms per call = 42 GB/s each way at the same time): small first / last groups (1, 2, 4, ...
    // clips) to shorten the pipeline's fill and drain (+0.26 ms: every extra group costs ~40 us), and one stream per
    // pipeline stage (upload / kernels / download) ordered by events over a ring of 3-6 buffer sets (10.66-10.73 ms,
    // no change), groups growing geometrically from 2 clips to 256 MB - 1 GB and shrinking again (11.2 - 12.2 vs 11.05 ms:
    // larger groups lose even without their fill / drain cost) -- the call is bound by the two PCIe directions, not by how
    // the copies are queued.
    std::vector<long long> sizes;
    {
        const long long n_groups = (batch + group - 1) / group;
        for (long long i = 0; i < n_groups; ++i) sizes.push_back(batch / n_groups + (i < batch % n_groups ? 1 : 0));
    }
    int li = 0;
    long long b0 = 0;
    for (size_t gi = 0; gi < sizes.size(); b0 += sizes[gi], ++gi, li = (li + 1) % n_lanes) {
        auto& ln = pl->lanes[li];
        const long long nb = sizes[gi];
        A2SB_CUDA(cudaMemcpyAsync(ln.d_wav, h_wav + sb * b0 * len, sb * nb * len, cudaMemcpyHostToDevice, ln.stream));
        a2sb_fwd_args fa{};
        fa.d_wav = ln.d_wav; fa.batch = nb; fa.len = len; fa.wav_stride = len; fa.sample_first = 0; fa.n_local = len;
        fa.t_begin = 0; fa.t_end = T; fa.d_out = ln.d_spec; fa.out_pitch = 0; fa.out_kind = A2SB_KIND_MAGPHASE; fa.drop_dc = 1;
        fa.power_on = 1; fa.power = power_fwd; fa.eps = eps; fa.stream = ln.stream; fa.wrap_cols = 0;
        if (int rc = pcm ? a2sb_stft_forward_pcm16(pl, &fa) : a2sb_stft_forward(pl, &fa)) return rc;
        if (h_spec)
            A2SB_CUDA(cudaMemcpyAsync(h_spec + b0 * spec_clip, ln.d_spec, sizeof(float) * nb * spec_clip,
                                      cudaMemcpyDeviceToHost, ln.stream));
        a2sb_inv_args ia{};
        ia.d_spec = ln.d_spec; ia.batch = nb; ia.n_frames = T; ia.spec_T = T; ia.spec_t_first = 0;
        ia.in_kind = A2SB_KIND_MAGPHASE; ia.has_dc = 0; ia.phase_fix = phase_fix; ia.power_on = 1;
        ia.power = power_inv; ia.eps = eps; ia.d_wav = ln.d_out; ia.wav_stride = out_len; ia.out_first = 0;
        ia.out_count = out_len; ia.stream = ln.stream;
        if (int rc = pcm ? a2sb_istft_inverse_pcm16(pl, &ia) : a2sb_istft_inverse(pl, &ia)) return rc;
        A2SB_CUDA(cudaMemcpyAsync(h_wav_out + sb * b0 * out_len, ln.d_out, sb * nb * out_len, cudaMemcpyDeviceToHost, ln.stream));
    }
    for (int l = 0; l < n_lanes; ++l) A2SB_CUDA(cudaStreamSynchronize(pl->lanes[l].stream));
    return A2SB_OK;
}

}  // extern "C"
