// radix.cuh -- register-resident radix-R complex DFT butterflies (R = 2..64, power of two).
//
// Each CUDA thread holds R complex points in registers (re[], im[]).  All loops have
// compile-time trip counts and every index is a compile-time constant after unrolling, so the
// arrays are scalarised into registers and twiddles become immediates.  Input and output are in
// natural order; the bit reversal is a static register renaming.
#pragma once
#include "a2sb_common.cuh"

namespace a2sb {

A2SB_HD constexpr int ilog2(int v) { return v <= 1 ? 0 : 1 + ilog2(v >> 1); }
A2SB_HD constexpr int bitrev(int v, int bits) {
    int r = 0;
    for (int i = 0; i < bits; ++i) r |= ((v >> i) & 1) << (bits - 1 - i);
    return r;
}

// DIR = -1: forward (e^{-2 pi i nk/R});  DIR = +1: inverse (unnormalised).
template <int R, int DIR>
A2SB_DEV void fft_reg(float (&re)[R], float (&im)[R]) {
    constexpr int LOG = ilog2(R);
    static_assert((1 << LOG) == R && R >= 2 && R <= 64, "radix must be a power of two <= 64");
    // static bit-reversal permutation (register renaming)
    A2SB_PRAGMA_UNROLL
    for (int i = 0; i < R; ++i) {
        const int j = bitrev(i, LOG);
        if (j > i) {
            float t = re[i]; re[i] = re[j]; re[j] = t;
            t = im[i]; im[i] = im[j]; im[j] = t;
        }
    }
    // decimation-in-time stages
    A2SB_PRAGMA_UNROLL
    for (int s = 1; s <= LOG; ++s) {
        const int len = 1 << s, half = len >> 1;
        A2SB_PRAGMA_UNROLL
        for (int g = 0; g < R; g += len) {
            A2SB_PRAGMA_UNROLL
            for (int k = 0; k < half; ++k) {
                const int i0 = g + k, i1 = g + k + half;
                const int tw = k * (64 / len);  // index into the 64-point table
                float br, bi;
                if (tw == 0) {
                    br = re[i1]; bi = im[i1];
                } else if (tw == 16) {  // W = -i (forward) / +i (inverse)
                    if (DIR < 0) { br = im[i1]; bi = -re[i1]; }
                    else         { br = -im[i1]; bi = re[i1]; }
                } else if (tw == 8) {   // W = (1 -/+ i)/sqrt2
                    const float h = 0.70710678118654752440f;
                    if (DIR < 0) { br = h * (re[i1] + im[i1]); bi = h * (im[i1] - re[i1]); }
                    else         { br = h * (re[i1] - im[i1]); bi = h * (im[i1] + re[i1]); }
                } else if (tw == 24) {  // W = (-1 -/+ i)/sqrt2
                    const float h = 0.70710678118654752440f;
                    if (DIR < 0) { br = h * (im[i1] - re[i1]); bi = -h * (re[i1] + im[i1]); }
                    else         { br = -h * (re[i1] + im[i1]); bi = h * (re[i1] - im[i1]); }
                } else {
                    const float c = kCos64(tw);
                    const float sn = (DIR < 0) ? -kSin64(tw) : kSin64(tw);  // W = c + i*sn
                    br = re[i1] * c - im[i1] * sn;
                    bi = re[i1] * sn + im[i1] * c;
                }
                const float ar = re[i0], ai = im[i0];
                re[i0] = ar + br; im[i0] = ai + bi;
                re[i1] = ar - br; im[i1] = ai - bi;
            }
        }
    }
}

}  // namespace a2sb
