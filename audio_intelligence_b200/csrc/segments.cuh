// segments.cuh -- K3/K4: MultiDiffusion segment windowing and overlap blend along the frame axis,
// plus the wrap-pad that precedes them.
//
// Reference (A2SB/diffusion.py):
//   multidiffusion_pad_inputs :67-83   pad the frame axis by copying the HEAD of the signal
//   get_multidiffusion_vf     :27-64   nn.Unfold -> "(b l) c h w" segments, network, then
//                                      sequential `vf_t[l:r] += seg_l ; counts[l:r] += 1`, vf_t / counts
// K3 gathers the overlapping segments (same (b l) order as the reference's rearrange).
// K4 sums the <= ceil(win/hop) contributing segments of every column in ascending segment order
// (the reference's order, so the fp32 result is bit-identical) and divides by the closed-form
// overlap count.  Both are pure HBM streaming kernels: 128-bit accesses, grid-stride.
#pragma once
#include "a2sb_common.cuh"

namespace a2sb {

struct SegParams {
    const float* in;
    float* out;
    long long rows;      // c*h rows per batch item
    long long width;     // W (padded frame axis)
    int batch;
    int win, hop;
    long long num_hops;  // L = (W - (win - hop)) / hop
    long long total;     // work items (vectors)
};

template <int VEC>
struct VecT;
template <> struct VecT<1> { using type = float; };
template <> struct VecT<4> { using type = float4; };

// segments[(b*L + l), row, w] = x[b, row, l*hop + w]
template <int VEC>
__global__ void __launch_bounds__(256) segment_gather_kernel(const SegParams p) {
    using V = typename VecT<VEC>::type;
    const long long wv = p.win / VEC;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < p.total;
         i += (long long)gridDim.x * blockDim.x) {
        const long long w = (i % wv) * VEC;
        long long r = i / wv;
        const long long row = r % p.rows;
        r /= p.rows;
        const long long l = r % p.num_hops;
        const long long b = r / p.num_hops;
        const float* src = p.in + (b * p.rows + row) * p.width + l * p.hop + w;
        float* dst = p.out + ((b * p.num_hops + l) * p.rows + row) * p.win + w;
        *reinterpret_cast<V*>(dst) = *reinterpret_cast<const V*>(src);
    }
}

A2SB_DEV void vadd(float& a, const float& v) { a += v; }
A2SB_DEV void vadd(float4& a, const float4& v) { a.x += v.x; a.y += v.y; a.z += v.z; a.w += v.w; }
A2SB_DEV float vdiv(const float& a, float d) { return a / d; }
A2SB_DEV float4 vdiv(const float4& a, float d) { return make_float4(a.x / d, a.y / d, a.z / d, a.w / d); }
A2SB_DEV void vzero(float& a) { a = 0.0f; }
A2SB_DEV void vzero(float4& a) { a = make_float4(0.f, 0.f, 0.f, 0.f); }

// out[b, row, col] = (sum over l ascending of seg[(b*L+l), row, col - l*hop]) / count(col)
template <int VEC>
__global__ void __launch_bounds__(256) segment_blend_kernel(const SegParams p) {
    using V = typename VecT<VEC>::type;
    const long long cv = p.width / VEC;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < p.total;
         i += (long long)gridDim.x * blockDim.x) {
        const long long col = (i % cv) * VEC;
        long long r = i / cv;
        const long long row = r % p.rows;
        const long long b = r / p.rows;
        // segments l with l*hop <= col < l*hop + win, 0 <= l < L
        long long l_hi = col / p.hop;
        if (l_hi > p.num_hops - 1) l_hi = p.num_hops - 1;
        long long l_lo = (col - p.win + p.hop) / p.hop;  // ceil((col - win + 1) / hop) for col >= win - hop
        if (col < p.win) l_lo = 0;
        V acc;
        vzero(acc);
        for (long long l = l_lo; l <= l_hi; ++l) {
            const float* src = p.in + ((b * p.num_hops + l) * p.rows + row) * p.win + (col - l * p.hop);
            vadd(acc, *reinterpret_cast<const V*>(src));
        }
        const float cnt = (float)(l_hi >= l_lo ? (l_hi - l_lo + 1) : 0);
        *reinterpret_cast<V*>(p.out + (b * p.rows + row) * p.width + col) = vdiv(acc, cnt);
    }
}

struct PadParams {
    const float* in;
    float* out;
    long long nrows;      // b*c*h
    long long width;      // W
    long long out_width;  // W + pad
    int use_const;
    float pad_const;
    long long total;
};

// out[row, w] = w < W ? in[row, w] : (const ? pad_const (with NaN propagation of in*0) : in[row, w - W])
__global__ void __launch_bounds__(256) wrap_pad_kernel(const PadParams p) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < p.total;
         i += (long long)gridDim.x * blockDim.x) {
        const long long w = i % p.out_width, row = i / p.out_width;
        float v;
        if (w < p.width) {
            v = p.in[row * p.width + w];
        } else {
            v = p.in[row * p.width + (w - p.width)];
            if (p.use_const) v = v * 0.0f + p.pad_const;  // diffusion.py:77 `padding*0+padding_constant`
        }
        p.out[i] = v;
    }
}

}  // namespace a2sb
