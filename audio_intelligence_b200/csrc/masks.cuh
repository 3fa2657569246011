// masks.cuh -- the integer / mask logic adjacent to the transform path.
//
//   * rect_mask_kernel / mask_noise_kernel / mask_fill_kernel: the corruption masks of
//     A2SB/corruption/corruptions.py (UpsampleMask :26-51, ExtensionMask :60-79, InpaintMask :90-117,
//     TimestampedSegmentInpaintMaskTransform :147-160 -- all axis-aligned rectangles) and
//     mask_with_noise (:14-15), x*(1-mask) + mask*noise*level, in the reference's fp32 operation order
//     (no fused multiply-add), so that the result is bit-identical for the same noise tensor.
//   * zero_segment_kernel: find_middle_of_zero_segments (A2SB/utils.py:54-81) and the window clamp
//     of the fast-inpaint sampler (A2SB/A2SB_lightning_module.py:161-174), integer-exact.
#pragma once
#include "a2sb_common.cuh"

namespace a2sb {

struct MaskParams {
    const float* x;       // [slices][rows][width]
    const float* noise;   // same shape (torch.randn_like on the caller's generator)
    const float* mask_in; // arbitrary mask (mask_noise_kernel)
    float* out;           // may alias nothing; same shape
    float* mask_out;      // rectangle mask (may be null)
    long long rows, width;
    long long row0, row1, col0, col1;   // mask == 1 on rows [row0, row1) x cols [col0, col1)
    long long total;      // slices * rows * width
    float level;
    DivMod d_width, d_rows;
};

#ifdef A2SB_EMU
A2SB_DEV float mul_rn(float a, float b) { volatile float r = a * b; return r; }
A2SB_DEV float add_rn(float a, float b) { volatile float r = a + b; return r; }
#else
A2SB_DEV float mul_rn(float a, float b) { return __fmul_rn(a, b); }   // never contracted into an FMA
A2SB_DEV float add_rn(float a, float b) { return __fadd_rn(a, b); }
#endif

// x * (1 - m) + m * noise * level, evaluated left to right like the reference expression
A2SB_DEV float mask_mix(float x, float m, float nz, float level) {
    return add_rn(mul_rn(x, add_rn(1.0f, -m)), mul_rn(mul_rn(m, nz), level));
}

template <bool F32>
A2SB_DEV float rect_value(const MaskParams& p, long long i) {
    long long line, w, slice, r;
    divmod<F32>(i, p.d_width, line, w);
    divmod<F32>(line, p.d_rows, slice, r);
    return (r >= p.row0 && r < p.row1 && w >= p.col0 && w < p.col1) ? 1.0f : 0.0f;
}

// mask only (get_upsample_mask / get_extension_mask / get_inpainting_mask)
template <bool F32>
__global__ void __launch_bounds__(256) rect_mask_kernel(const MaskParams p) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < p.total; i += stride) p.mask_out[i] = rect_value<F32>(p, i);
}

// mask_with_noise with a caller-supplied mask tensor
__global__ void __launch_bounds__(256) mask_noise_kernel(const MaskParams p) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < p.total; i += stride)
        p.out[i] = mask_mix(p.x[i], p.mask_in[i], p.noise[i], p.level);
}

// rectangle mask + mask_with_noise in one pass (the mask is written only if requested).  VEC = 4: 128-bit accesses
// (width % 4 == 0, 16-byte aligned pointers; p.total and p.d_width then count float4 columns).
template <int VEC, bool F32>
__global__ void __launch_bounds__(256) mask_fill_kernel(const MaskParams p) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < p.total; i += stride) {
        if (VEC == 1) {
            const float m = rect_value<F32>(p, i);
            p.out[i] = mask_mix(p.x[i], m, p.noise[i], p.level);
            if (p.mask_out) p.mask_out[i] = m;
        } else {
            long long line, wv, slice, r;
            divmod<F32>(i, p.d_width, line, wv);
            divmod<F32>(line, p.d_rows, slice, r);
            const bool row_in = r >= p.row0 && r < p.row1;
            const float4 x = reinterpret_cast<const float4*>(p.x)[i], nz = reinterpret_cast<const float4*>(p.noise)[i];
            const long long w = wv * 4;
            float4 m;
            m.x = (row_in && w + 0 >= p.col0 && w + 0 < p.col1) ? 1.0f : 0.0f;
            m.y = (row_in && w + 1 >= p.col0 && w + 1 < p.col1) ? 1.0f : 0.0f;
            m.z = (row_in && w + 2 >= p.col0 && w + 2 < p.col1) ? 1.0f : 0.0f;
            m.w = (row_in && w + 3 >= p.col0 && w + 3 < p.col1) ? 1.0f : 0.0f;
            float4 o;
            o.x = mask_mix(x.x, m.x, nz.x, p.level);
            o.y = mask_mix(x.y, m.y, nz.y, p.level);
            o.z = mask_mix(x.z, m.z, nz.z, p.level);
            o.w = mask_mix(x.w, m.w, nz.w, p.level);
            reinterpret_cast<float4*>(p.out)[i] = o;
            if (p.mask_out) reinterpret_cast<float4*>(p.mask_out)[i] = m;
        }
    }
}

// mask_fill for the segment-windowed sampler: x may be row-pitched (K1's aligned / wrap-padded output), noise is the
// contiguous [.., width] tensor drawn by torch.randn_like, and both outputs are written with rows of `out_width`
// columns whose tail replicates the head -- what multidiffusion_pad_inputs (A2SB/diffusion.py:67-83) would append to the
// filled spectrogram and to the mask (padding_constant = None) afterwards, saving its two passes.
struct MaskPadParams {
    MaskParams m;           // m.width = valid columns; m.total = slices * rows * width; m.d_width / m.d_rows as usual
    long long in_pitch;     // elements between rows of x
    long long out_width;    // elements between rows of out / mask_out ( >= width; columns width .. out_width-1 = head)
};

// VEC = 2: 64-bit accesses (width, pitches and the wrap width even, 8-byte aligned pointers; m.total and m.d_width then
// count column pairs).  The shipped geometry (862 -> 896 frames) is even but not a multiple of 4.
template <int VEC, bool F32>
__global__ void __launch_bounds__(256) mask_fill_padded_kernel(const MaskPadParams q) {
    const MaskParams& p = q.m;
    const long long stride = (long long)gridDim.x * blockDim.x;
    const long long wrap = q.out_width - p.width;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < p.total; i += stride) {
        long long line, w, slice, r;
        divmod<F32>(i, p.d_width, line, w);
        divmod<F32>(line, p.d_rows, slice, r);
        w *= VEC;
        const bool row_in = r >= p.row0 && r < p.row1;
        float xv[VEC], nz[VEC], m[VEC], v[VEC];
        if (VEC == 2) {
            const float2 x2 = *reinterpret_cast<const float2*>(p.x + line * q.in_pitch + w);
            const float2 n2 = *reinterpret_cast<const float2*>(p.noise + line * p.width + w);
            xv[0] = x2.x; xv[VEC - 1] = x2.y; nz[0] = n2.x; nz[VEC - 1] = n2.y;
        } else {
            xv[0] = p.x[line * q.in_pitch + w]; nz[0] = p.noise[line * p.width + w];
        }
        A2SB_PRAGMA_UNROLL
        for (int e = 0; e < VEC; ++e) {
            m[e] = (row_in && w + e >= p.col0 && w + e < p.col1) ? 1.0f : 0.0f;
            v[e] = mask_mix(xv[e], m[e], nz[e], p.level);
        }
        float* o = p.out + line * q.out_width + w;
        float* mo = p.mask_out ? p.mask_out + line * q.out_width + w : nullptr;
        if (VEC == 2) {
            *reinterpret_cast<float2*>(o) = make_float2(v[0], v[VEC - 1]);
            if (w < wrap) *reinterpret_cast<float2*>(o + p.width) = make_float2(v[0], v[VEC - 1]);
            if (mo) {
                *reinterpret_cast<float2*>(mo) = make_float2(m[0], m[VEC - 1]);
                if (w < wrap) *reinterpret_cast<float2*>(mo + p.width) = make_float2(m[0], m[VEC - 1]);
            }
        } else {
            *o = v[0];
            if (w < wrap) o[p.width] = v[0];
            if (mo) { *mo = m[0]; if (w < wrap) mo[p.width] = m[0]; }
        }
    }
}

struct ZeroSegParams {
    const float* row;   // [n] values (the reference passes 1 - mask[0, 0, 0])
    long long n;
    int win_length;
    int* centres;       // [max_out]
    int* lr;            // [max_out][2]: window [l, r) of length win_length clamped into [0, n]
    int* count;         // [1]: number of zero segments found (may exceed max_out; only max_out are written)
    int max_out;
};

constexpr int kZeroSegThreads = 1024;

// One CTA.  diff[i] = row[i] - row[i-1] with row[-1] = 1 (utils.py:68); a segment starts where
// diff == -1 and ends one before where diff == +1; an array ending in 0 closes its last segment
// at n - 1 (:75-76).  k-th start pairs with k-th end; centre = int((start + end) / 2) evaluated in
// fp32 like torch's true division followed by .int() (:79).
__global__ void __launch_bounds__(kZeroSegThreads) zero_segment_kernel(const ZeroSegParams p) {
    __shared__ int s_ns[kZeroSegThreads], s_ne[kZeroSegThreads];
    const int tid = threadIdx.x;
    const long long chunk = (p.n + kZeroSegThreads - 1) / kZeroSegThreads;
    const long long lo = (long long)tid * chunk;
    const long long hi = lo + chunk < p.n ? lo + chunk : p.n;
    int ns = 0, ne = 0;
    {
        float prev = lo == 0 ? 1.0f : (lo <= p.n ? p.row[lo - 1] : 0.0f);
        for (long long i = lo; i < hi; ++i) {
            const float v = p.row[i];
            const float d = v - prev;
            ns += (d == -1.0f);
            ne += (d == 1.0f);
            prev = v;
        }
        if (hi == p.n && lo < hi && p.row[p.n - 1] == 0.0f) ne += 1;   // trailing zero segment
    }
    s_ns[tid] = ns;
    s_ne[tid] = ne;
    __syncthreads();
    // inclusive Hillis-Steele scans over the per-thread counts
    for (int off = 1; off < kZeroSegThreads; off <<= 1) {
        const int a = tid >= off ? s_ns[tid - off] : 0;
        const int b = tid >= off ? s_ne[tid - off] : 0;
        __syncthreads();
        s_ns[tid] += a;
        s_ne[tid] += b;
        __syncthreads();
    }
    int ks = s_ns[tid] - ns, ke = s_ne[tid] - ne;   // exclusive prefixes
    const int total_s = s_ns[kZeroSegThreads - 1], total_e = s_ne[kZeroSegThreads - 1];
    const int total = total_s < total_e ? total_s : total_e;
    {
        float prev = lo == 0 ? 1.0f : (lo <= p.n ? p.row[lo - 1] : 0.0f);
        for (long long i = lo; i < hi; ++i) {
            const float v = p.row[i];
            const float d = v - prev;
            if (d == -1.0f) { if (ks < p.max_out) p.lr[2 * ks] = (int)i; ++ks; }
            if (d == 1.0f) { if (ke < p.max_out) p.lr[2 * ke + 1] = (int)(i - 1); ++ke; }
            prev = v;
        }
        if (hi == p.n && lo < hi && p.row[p.n - 1] == 0.0f) { if (ke < p.max_out) p.lr[2 * ke + 1] = (int)(p.n - 1); }
    }
    __syncthreads();
    if (tid == 0) *p.count = total;
    for (int k = tid; k < total && k < p.max_out; k += kZeroSegThreads) {
        const long long sum = (long long)p.lr[2 * k] + (long long)p.lr[2 * k + 1];
        const int c = (int)((float)sum / 2.0f);
        // A2SB_lightning_module.py:162-171
        int l = (int)((float)c - (float)p.win_length / 2.0f);
        int r = (int)((float)c + (float)p.win_length / 2.0f);
        if (l < 0) { r -= l; l = 0; }
        if ((long long)r > p.n) { l -= (int)(r - p.n); r = (int)p.n; }
        p.centres[k] = c;
        p.lr[2 * k] = l;
        p.lr[2 * k + 1] = r;
    }
}

}  // namespace a2sb
