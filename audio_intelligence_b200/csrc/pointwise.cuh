// pointwise.cuh -- the reference's per-bin transform ops as standalone kernels, for callers that
// apply them one at a time (e.g. `pl_module.inv_transforms[0](...)`,
// A2SB/A2SB_lightning_module.py:501).  The canonical chains never launch these: they are fused
// into K1's epilogue / K2's prologue.  All tensors are contiguous [C, n] with n = rows*frames.
#pragma once
#include "a2sb_common.cuh"

namespace a2sb {

enum : int {
    kOpComplexToMagPhase = 0,  // ComplexToMagInstPhase   transforms.py:108-118   [2,n] -> [3,n]
    kOpMagPhaseToComplex = 1,  // MagInstPhaseToComplex   transforms.py:121-132   [3,n] -> [2,n]
    kOpPhaseFix = 2,           // SVDFixMagInstPhase      transforms.py:135-160   [3,n] -> [3,n]
    kOpPowerScale = 3,         // PowerScaleSpectrogram   transforms.py:187-207   [C,n] -> [C,n]
    // other consumers of the forward / inverse kernels:
    kOpComplexToMagAngle = 4,  // ETTA STFT.encode (adp.py:1548-1551: torch.abs, torch.angle) and auraloss STFTLoss.stft
                               // (auraloss.py:373-381: sqrt(clamp(re^2 + im^2, min=eps)), torch.angle)   [2,n] -> [2,n]
    kOpPolarToComplex = 5,     // ETTA STFT.decode (adp.py:1566-1567: m cos(phi), m sin(phi))              [2,n] -> [2,n]
};

struct PwParams {
    const float* in;
    float* out;
    long long n;        // elements per channel
    int channels;       // C (power scale)
    unsigned chan_mask; // bit c set: channel c is scaled (power scale)
    float power, eps;
    int op;
};

__global__ void __launch_bounds__(256) pointwise_kernel(const PwParams p) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < p.n; i += stride) {
        if (p.op == kOpComplexToMagPhase) {
            const float xr = p.in[i], xi = p.in[p.n + i];
            const float m2 = xr * xr + xi * xi;
            float mag = sqrtf(m2), cs, sn;
            float a = xr, b = xi, q = m2;
            if (q < 1e-30f) { a *= 1.8446744e19f; b *= 1.8446744e19f; q = a * a + b * b; }
            if (q > 0.0f) { const float rs = rsqrt_approx(q); cs = a * rs; sn = b * rs; }
            else { cs = 1.0f; sn = 0.0f; }
            p.out[i] = mag; p.out[p.n + i] = cs; p.out[2 * p.n + i] = sn;
        } else if (p.op == kOpMagPhaseToComplex) {
            const float m = p.in[i];
            p.out[i] = m * p.in[p.n + i];
            p.out[p.n + i] = m * p.in[2 * p.n + i];
        } else if (p.op == kOpPhaseFix) {
            float c = p.in[p.n + i], s = p.in[2 * p.n + i];
            float n2 = c * c + s * s;
            if (n2 < 1e-30f) { c *= 1.8446744e19f; s *= 1.8446744e19f; n2 = c * c + s * s; }
            if (n2 > 0.0f) { const float rn = rsqrt_approx(n2); c *= rn; s *= rn; }
            else { c = 1.0f; s = 0.0f; }
            p.out[i] = p.in[i]; p.out[p.n + i] = c; p.out[2 * p.n + i] = s;
        } else if (p.op == kOpComplexToMagAngle) {
            const float xr = p.in[i], xi = p.in[p.n + i];
            float m2 = xr * xr + xi * xi;
            if (p.eps > 0.0f && !(m2 >= p.eps)) m2 = p.eps;      // torch.clamp(min=eps) (NaN propagates like torch's)
            p.out[i] = (p.eps > 0.0f) ? sqrtf(m2) : hypotf(xr, xi);   // torch.abs(complex) is hypot
            p.out[p.n + i] = atan2f(xi, xr);
        } else if (p.op == kOpPolarToComplex) {
            const float m = p.in[i], ph = p.in[p.n + i];
            float sn, cs;
            sincosf(ph, &sn, &cs);
            p.out[i] = m * cs; p.out[p.n + i] = m * sn;
        } else {
            for (int ch = 0; ch < p.channels; ++ch) {
                float v = p.in[ch * p.n + i];
                if ((p.chan_mask >> ch) & 1u) v = v * power_scale_factor<kPowGeneric>(fabsf(v), p.power, p.eps);
                p.out[ch * p.n + i] = v;
            }
        }
    }
}

// One phase update of (fast) Griffin-Lim -- the loop body of `griffinlim`
// (A2SB/audio_transforms/transforms.py:351-362, after torchaudio.functional.griffinlim):
//     angles  = rebuilt - momentum * tprev                (complex; momentum already m / (1 + m))
//     angles  = angles / (|angles| + 1e-16)
//     product = specgram * angles                         (input of the next istft)
// rebuilt / tprev / product are [batch][2][n] (re, im planes); mag is [batch][n].  `tprev` may be null (first
// iteration).
struct GlParams {
    const float* rebuilt;
    const float* tprev;
    const float* mag;
    float* product;
    long long n;        // bins x frames per batch item
    long long total;    // batch * n
    float momentum;
};

__global__ void __launch_bounds__(256) griffinlim_update_kernel(const GlParams p) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < p.total; i += stride) {
        const long long b = i / p.n, j = i - b * p.n;
        const long long re = 2 * b * p.n + j, im = re + p.n;
        float ar = p.rebuilt[re], ai = p.rebuilt[im];
        if (p.tprev) {
            ar = ar - p.tprev[re] * p.momentum;
            ai = ai - p.tprev[im] * p.momentum;
        }
#ifdef A2SB_EMU
        const float mod = std::hypot(ar, ai) + 1e-16f;
#else
        const float mod = hypotf(ar, ai) + 1e-16f;
#endif
        const float m = p.mag[i];
        p.product[re] = m * (ar / mod);
        p.product[im] = m * (ai / mod);
    }
}

}  // namespace a2sb
