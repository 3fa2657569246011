// a2sb_common.cuh -- shared definitions for the A2SB spectral-transform kernels (sm_100a).
//
// The same kernel sources compile two ways:
//   * nvcc -gencode arch=compute_100a,code=sm_100a  -> the product library (liba2sb_b200.so)
//   * g++ -DA2SB_EMU (tests/emu/cuda_emu.h)          -> a CPU emulation used ONLY by tests/ to
//     check index math and synchronisation structure in a container without a GPU.
#pragma once

#ifdef A2SB_EMU
#include "cuda_emu.h"
#define A2SB_HD
#define A2SB_DEV static inline
#define A2SB_DYN_SMEM(name) unsigned char* name = emu::g_ctx->smem
#define A2SB_PRAGMA_UNROLL
#define A2SB_GRID_CONSTANT
#else
#include <cuda_runtime.h>
#include <cstdint>
#define A2SB_HD __host__ __device__
#define A2SB_DEV static __device__ __forceinline__
#define A2SB_DYN_SMEM(name) extern __shared__ __align__(1024) unsigned char name[]
#define A2SB_PRAGMA_UNROLL _Pragma("unroll")
#define A2SB_GRID_CONSTANT __grid_constant__
#endif

#include "twiddle64.h"

#ifndef A2SB_INV_LD
#define A2SB_INV_LD 5   // spectrogram load policy of K2 (istft_inv.cuh::ld_spec); 5 = ld.global.nc.L2::256B
#endif

namespace a2sb {

constexpr int kSMs = 148;  // B200: 2 dies x 74 SMs; grids are sized in multiples of this

// ---- bit-exact index arithmetic (shared with host code) ---------------------------------------

// Number of STFT frames for center=True: T = 1 + L / hop   (torch/functional.py:508 `stft`).
A2SB_HD inline long long num_frames(long long len, int hop) { return 1 + len / hop; }

// reflect-pad source index (pad_mode="reflect", no edge repeat): i in (-L, 2L-1)
A2SB_HD inline long long reflect_index(long long i, long long len) {
    if (i < 0) i = -i;
    if (i >= len) i = 2 * (len - 1) - i;
    return i;
}

// ---- fast math with explicit, documented accuracy ---------------------------------------------
#ifdef A2SB_EMU
A2SB_DEV float rsqrt_approx(float x) { return 1.0f / std::sqrt(x); }
A2SB_DEV float rcp_approx(float x) { return 1.0f / x; }
A2SB_DEV float lg2_approx(float x) { return std::log2(x); }
A2SB_DEV float ex2_approx(float x) { return std::exp2(x); }
#else
// MUFU.RSQ / MUFU.RCP / MUFU.LG2 / MUFU.EX2: <= ~2^-22 relative error; denormals flushed.
A2SB_DEV float rsqrt_approx(float x) { float r; asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
A2SB_DEV float rcp_approx(float x) { float r; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
A2SB_DEV float lg2_approx(float x) { float r; asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
A2SB_DEV float ex2_approx(float x) { float r; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
#endif

// |m|^p / (|m| + eps): the scale factor of PowerScaleSpectrogram
// (A2SB/audio_transforms/transforms.py:199-200).  PMODE selects a closed form for the two
// shipped exponents (configs/ensemble_2split_sampling.yaml:64-69,114-119).
enum : int { kPowGeneric = 0, kPowQuarter = 1, kPowFour = 2, kPowNone = 3 };

template <int PMODE>
A2SB_DEV float power_scale_factor(float a /* = |m| >= 0 */, float power, float eps) {
    float num;
    if (PMODE == kPowQuarter) {
        // a^(1/4) = rsqrt(rsqrt(a)); a == 0 -> rsqrt(+inf) = 0, matching pow(0, .25) = 0.
        num = rsqrt_approx(rsqrt_approx(a));
    } else if (PMODE == kPowFour) {
        const float a2 = a * a;
        num = a2 * a2;
    } else {
        // exp2(p * log2(a)); a == 0 -> exp2(-inf) = 0 for p > 0.
        num = ex2_approx(power * lg2_approx(a));
    }
    return num * rcp_approx(a + eps);
}

// ---- division by a launch-invariant integer ------------------------------------------------------------
// The streaming kernels turn a linear work index into (line, column, ...) coordinates; a runtime 64-bit
// divide costs ~100 instructions per element.  For dividends below 2^31 (checked on the host) the quotient is
// (n * m) >> k with m = ceil(2^k / d), k = 31 + ceil(log2 d)  (Granlund & Montgomery): one wide multiply and a shift.
struct DivMod {
    unsigned long long d;   // divisor (>= 1)
    unsigned m;             // magic multiplier ceil(2^k / d) (< 2^32)
    unsigned k;             // shift
};
inline DivMod make_divmod(long long d) {
    DivMod r{};
    r.d = (unsigned long long)(d < 1 ? 1 : d);
    unsigned l = 0;
    while ((1ull << l) < r.d) ++l;            // ceil(log2 d)
    r.k = 31 + l;
    r.m = (unsigned)(((1ull << r.k) + r.d - 1) / r.d);   // < 2^32 for every d < 2^31 (k <= 62)
    return r;
}
// q = n / d, r = n % d.  FAST32: 0 <= n < 2^31 and d < 2^31.
template <bool FAST32>
A2SB_HD inline void divmod(long long n, const DivMod& dm, long long& q, long long& r) {
    if (FAST32) {
        const unsigned nn = (unsigned)n;
        const unsigned qq = (unsigned)(((unsigned long long)nn * dm.m) >> dm.k);   // one 32x32->64 multiply
        q = (long long)qq;
        r = (long long)(nn - qq * (unsigned)dm.d);
    } else {
        q = n / (long long)dm.d;
        r = n - q * (long long)dm.d;
    }
}

// Pack/unpack helpers for launch parameter blocks -----------------------------------------------
struct Span {
    const float* ptr;        // local buffer
    long long first;         // global sample/frame index of ptr[0]
    long long count;         // number of valid elements in the local buffer
};

}  // namespace a2sb
