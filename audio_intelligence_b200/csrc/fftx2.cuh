// fftx2.cuh -- register-resident radix-16 / radix-32 complex DFTs built on Blackwell's packed
// fp32 pipe (FFMA2 / FADD2 / FMUL2, PTX fma.rn.f32x2, sm_100+).
//
// The kernels in this library are issue-bound, not lane-bound, so the butterflies are arranged so
// that ONE packed instruction does the same butterfly of TWO independent 16-point sub-transforms:
//   * a 32-point DFT is two 16-point DFTs (A over even inputs, B over odd inputs, or -- in the
//     decimation-in-frequency form -- over the sum / twiddled difference of the two input halves)
//     plus one scalar radix-2 stage;
//   * the two sub-transforms live in the .x and .y halves of the same 64-bit register pairs, and
//     real and imaginary parts are kept in separate arrays, so every stage of the 16-point part is
//     a packed op with a scalar (broadcast) twiddle, and multiplications by +-i are free.
// All loops have compile-time trip counts; after unrolling every index and twiddle is a
// compile-time constant and the arrays are scalarised into registers.
#pragma once
#include "a2sb_common.cuh"

namespace a2sb {

#ifdef A2SB_EMU
A2SB_DEV float2 p2_fma(float2 a, float2 b, float2 c) { return make_float2(std::fma(a.x, b.x, c.x), std::fma(a.y, b.y, c.y)); }
A2SB_DEV float2 p2_add(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
A2SB_DEV float2 p2_mul(float2 a, float2 b) { return make_float2(a.x * b.x, a.y * b.y); }
A2SB_DEV float s_fma(float a, float b, float c) { return std::fma(a, b, c); }
#else
A2SB_DEV float2 p2_fma(float2 a, float2 b, float2 c) { return __ffma2_rn(a, b, c); }
A2SB_DEV float2 p2_add(float2 a, float2 b) { return __fadd2_rn(a, b); }
A2SB_DEV float2 p2_mul(float2 a, float2 b) { return __fmul2_rn(a, b); }
A2SB_DEV float s_fma(float a, float b, float c) { return __fmaf_rn(a, b, c); }
#endif
A2SB_DEV float2 p2_bc(float s) { return make_float2(s, s); }
A2SB_DEV float2 p2_neg(float2 a) { return make_float2(-a.x, -a.y); }
A2SB_DEV float2 p2_sub(float2 a, float2 b) { return p2_add(a, p2_neg(b)); }

// Lane-type traits: V = float (one transform) or float2 (two transforms, packed).
A2SB_DEV float v_add(float a, float b) { return a + b; }
A2SB_DEV float v_sub(float a, float b) { return a - b; }
A2SB_DEV float v_fma(float a, float s, float c) { return s_fma(a, s, c); }     // a*s + c
A2SB_DEV float v_fms(float a, float s, float c) { return s_fma(a, s, -c); }    // a*s - c
A2SB_DEV float2 v_add(float2 a, float2 b) { return p2_add(a, b); }
A2SB_DEV float2 v_sub(float2 a, float2 b) { return p2_sub(a, b); }
A2SB_DEV float2 v_fma(float2 a, float s, float2 c) { return p2_fma(a, p2_bc(s), c); }
A2SB_DEV float2 v_fms(float2 a, float s, float2 c) { return p2_fma(a, p2_bc(s), p2_neg(c)); }

A2SB_HD constexpr int ilog2c(int v) { return v <= 1 ? 0 : 1 + ilog2c(v >> 1); }
A2SB_HD constexpr int bitrevc(int v, int bits) {
    int r = 0;
    for (int i = 0; i < bits; ++i) r |= ((v >> i) & 1) << (bits - 1 - i);
    return r;
}

// One radix-2 butterfly (a, b) -> (a + W b, a - W b), W = exp(DIR * 2 pi i * tw / 64), on lane type V.
// 4 ops for W in {1, -+i}, 6 ops otherwise (the second output is formed as 2a - (a + W b)).
template <int DIR, int TW, class V>
A2SB_DEV void bfly(V& ar, V& ai, V& br, V& bi) {
    constexpr int tw = TW & 63;
    if (tw == 0) {
        const V r0 = v_add(ar, br), i0 = v_add(ai, bi);
        br = v_sub(ar, br); bi = v_sub(ai, bi);
        ar = r0; ai = i0;
    } else if (tw == 16) {  // W = -i (forward) / +i (inverse):  W b = (bi, -br) / (-bi, br)
        V r0, i0, r1, i1;
        if (DIR < 0) { r0 = v_add(ar, bi); i0 = v_sub(ai, br); r1 = v_sub(ar, bi); i1 = v_add(ai, br); }
        else         { r0 = v_sub(ar, bi); i0 = v_add(ai, br); r1 = v_add(ar, bi); i1 = v_sub(ai, br); }
        ar = r0; ai = i0; br = r1; bi = i1;
    } else if (tw == 8 || tw == 24) {
        // W = h(+-1 -+ i) patterns: W b = h * (p, q) with p, q sums/differences of (br, bi).
        constexpr float h = 0.70710678118654752440f;
        V p, q;
        if (tw == 8) {  // forward W = h(1 - i): (br + bi, bi - br); inverse W = h(1 + i): (br - bi, bi + br)
            if (DIR < 0) { p = v_add(br, bi); q = v_sub(bi, br); }
            else         { p = v_sub(br, bi); q = v_add(bi, br); }
        } else {        // forward W = h(-1 - i): (bi - br, -(br + bi)); inverse W = h(-1 + i): (-(br + bi), br - bi)
            if (DIR < 0) { p = v_sub(bi, br); q = v_add(br, bi); }
            else         { p = v_add(br, bi); q = v_sub(br, bi); }
        }
        const bool nq = (tw == 24 && DIR < 0);   // q enters with a minus sign
        const bool np = (tw == 24 && DIR > 0);   // p enters with a minus sign
        const V r0 = v_fma(p, np ? -h : h, ar), r1 = v_fma(p, np ? h : -h, ar);
        const V i0 = v_fma(q, nq ? -h : h, ai), i1 = v_fma(q, nq ? h : -h, ai);
        ar = r0; ai = i0; br = r1; bi = i1;
    } else {
        const float c = kCos64(tw);
        const float s = (DIR < 0) ? -kSin64(tw) : kSin64(tw);  // W = c + i s
        V r0 = v_fma(br, c, ar);
        r0 = v_fma(bi, -s, r0);
        V i0 = v_fma(br, s, ai);
        i0 = v_fma(bi, c, i0);
        br = v_fms(ar, 2.0f, r0);
        bi = v_fms(ai, 2.0f, i0);
        ar = r0; ai = i0;
    }
}

template <int R, int DIR, class V, int LEN, int HALF, int IDX>
A2SB_DEV void fft_stage_groups(V (&re)[R], V (&im)[R]) {
    // IDX enumerates the R/2 butterflies of the stage: group g = IDX / HALF, position k = IDX % HALF.
    if constexpr (IDX < R / 2) {
        constexpr int g = (IDX / HALF) * LEN, k = IDX % HALF;
        bfly<DIR, k*(64 / LEN), V>(re[g + k], im[g + k], re[g + k + HALF], im[g + k + HALF]);
        fft_stage_groups<R, DIR, V, LEN, HALF, IDX + 1>(re, im);
    }
}

template <int R, int DIR, class V, int S>
A2SB_DEV void fft_stages(V (&re)[R], V (&im)[R]) {
    if constexpr (S <= ilog2c(R)) {
        constexpr int len = 1 << S, half = len >> 1;
        fft_stage_groups<R, DIR, V, len, half, 0>(re, im);
        fft_stages<R, DIR, V, S + 1>(re, im);
    }
}

// Natural-order in, natural-order out R-point DFT (R = 2..32) on lane type V, decimation in time.
template <int R, int DIR, class V>
A2SB_DEV void fft_v(V (&re)[R], V (&im)[R]) {
    constexpr int LOG = ilog2c(R);
    static_assert((1 << LOG) == R && R >= 2 && R <= 32, "radix must be a power of two <= 32");
    A2SB_PRAGMA_UNROLL
    for (int i = 0; i < R; ++i) {
        const int j = bitrevc(i, LOG);
        if (j > i) {
            V t = re[i]; re[i] = re[j]; re[j] = t;
            t = im[i]; im[i] = im[j]; im[j] = t;
        }
    }
    fft_stages<R, DIR, V, 1>(re, im);
}

// ---- R-point DFT of one sequence = packed R/2-point DFT of two sub-sequences + one scalar stage --

template <int R, int DIR, int K>
A2SB_DEV void fft2x_dit_last(float2 (&re)[R / 2], float2 (&im)[R / 2], float (&ore)[R], float (&oim)[R]) {
    if constexpr (K < R / 2) {
        float ar = re[K].x, ai = im[K].x, br = re[K].y, bi = im[K].y;
        bfly<DIR, K*(64 / R), float>(ar, ai, br, bi);   // W_R^K
        ore[K] = ar; oim[K] = ai; ore[K + R / 2] = br; oim[K + R / 2] = bi;
        fft2x_dit_last<R, DIR, K + 1>(re, im, ore, oim);
    }
}
// Decimation in time.  in: pair j holds (x[2j], x[2j+1]) in the (.x, .y) halves of (re[j], im[j]).
// out: X[k] in ore/oim[k], natural order.
template <int R, int DIR>
A2SB_DEV void fft2x_dit(float2 (&re)[R / 2], float2 (&im)[R / 2], float (&ore)[R], float (&oim)[R]) {
    fft_v<R / 2, DIR, float2>(re, im);   // .x: DFT(even inputs), .y: DFT(odd inputs)
    fft2x_dit_last<R, DIR, 0>(re, im, ore, oim);
}

// Decimation in frequency, first (scalar) stage for index Q < R/2:
//   u = a + b, v = (a - b) W_R^Q  with a = x[Q], b = x[Q + R/2];  X[2k] = DFT(u)[k], X[2k+1] = DFT(v)[k].
// Stores (u, v) into the (.x, .y) halves of (re, im).
template <int R, int DIR, int Q>
A2SB_DEV void dif_first(float ar, float ai, float br, float bi, float2& re, float2& im) {
    const float ur = ar + br, ui = ai + bi;
    const float dr = ar - br, di = ai - bi;
    float vr, vi;
    constexpr int tw = (Q * (64 / R)) & 63;
    if (tw == 0) { vr = dr; vi = di; }
    else if (tw == 16) { if (DIR < 0) { vr = di; vi = -dr; } else { vr = -di; vi = dr; } }
    else {
        const float c = kCos64(tw), s = (DIR < 0) ? -kSin64(tw) : kSin64(tw);
        vr = s_fma(dr, c, -(di * s));
        vi = s_fma(dr, s, di * c);
    }
    re = make_float2(ur, vr);
    im = make_float2(ui, vi);
}
// Same with the analysis window folded in: a = (xa.x wa.x, xa.y wa.y), b likewise (3 ops per
// component instead of 4).
template <int R, int DIR, int Q>
A2SB_DEV void dif_first_windowed(float2 xa, float2 wa, float2 xb, float2 wb, float2& re, float2& im) {
    const float tr = xb.x * wb.x, ti = xb.y * wb.y;
    const float ur = s_fma(xa.x, wa.x, tr), ui = s_fma(xa.y, wa.y, ti);
    const float dr = s_fma(xa.x, wa.x, -tr), di = s_fma(xa.y, wa.y, -ti);
    float vr, vi;
    constexpr int tw = (Q * (64 / R)) & 63;
    if (tw == 0) { vr = dr; vi = di; }
    else if (tw == 16) { if (DIR < 0) { vr = di; vi = -dr; } else { vr = -di; vi = dr; } }
    else {
        const float c = kCos64(tw), s = (DIR < 0) ? -kSin64(tw) : kSin64(tw);
        vr = s_fma(dr, c, -(di * s));
        vi = s_fma(dr, s, di * c);
    }
    re = make_float2(ur, vr);
    im = make_float2(ui, vi);
}

}  // namespace a2sb
