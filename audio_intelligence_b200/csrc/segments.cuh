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
    DivMod d_wv, d_cv, d_rows, d_hops, d_hop;   // win/VEC, width/VEC, rows, num_hops, hop (host: seg_divisors)
    // blend output window (sharded long audio: a rank keeps its owned columns only and writes them straight into the
    // pre-padded state buffer of the next step): columns [col_off, col_off + col_cnt) -> out[(b*rows+row)*out_pitch + col - col_off].
    // d_cv divides by col_cnt/VEC.  Whole output: col_off = 0, col_cnt = out_pitch = width.
    long long col_off, col_cnt, out_pitch;
};

template <int VEC>
struct VecT;
template <> struct VecT<1> { using type = float; };
template <> struct VecT<4> { using type = float4; };

// segments[(b*L + l), row, w] = x[b, row, l*hop + w]
template <int VEC, bool F32>
__global__ void __launch_bounds__(256) segment_gather_kernel(const SegParams p) {
    using V = typename VecT<VEC>::type;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < p.total;
         i += (long long)gridDim.x * blockDim.x) {
        long long r, w, row, l, b;
        divmod<F32>(i, p.d_wv, r, w);
        w *= VEC;
        divmod<F32>(r, p.d_rows, r, row);
        divmod<F32>(r, p.d_hops, b, l);
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
// (b, row, col) of work item i and the contributing segment range [l_lo, l_hi]
template <int VEC, bool F32>
A2SB_DEV void blend_coords(const SegParams& p, long long i, long long& b, long long& row, long long& col, long long& l_lo,
                           long long& l_hi) {
    long long r, rem;
    divmod<F32>(i, p.d_cv, r, col);
    col = col * VEC + p.col_off;
    divmod<F32>(r, p.d_rows, b, row);
    // segments l with l*hop <= col < l*hop + win, 0 <= l < L
    divmod<F32>(col, p.d_hop, l_hi, rem);
    if (l_hi > p.num_hops - 1) l_hi = p.num_hops - 1;
    l_lo = 0;
    if (col >= p.win) divmod<F32>(col - p.win + p.hop, p.d_hop, l_lo, rem);  // ceil((col - win + 1) / hop)
}

// The (at most two, for win <= 2 hop) segment loads of a work item are issued before the first add: with the loads inside
// a variable-trip loop the second waited for the first (2.26 -> 2.10 ms at config-3 size, 77 % -> 83 % of the copy peak).
// A2SB_BLEND_U work items per thread and iteration: measured 1 / 2 / 4 -> 5.37 / 5.40 / 5.29 TB/s, so 1.
template <int VEC, bool F32>
__global__ void __launch_bounds__(256) segment_blend_kernel(const SegParams p) {
    using V = typename VecT<VEC>::type;
#ifndef A2SB_BLEND_U
#define A2SB_BLEND_U 1
#endif
    constexpr int U = A2SB_BLEND_U;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i0 = (long long)blockIdx.x * blockDim.x + threadIdx.x; i0 < p.total; i0 += U * stride) {
        long long b[U], row[U], col[U], l_lo[U], l_hi[U];
        const float* src[U];
        V first[U], second[U];
        A2SB_PRAGMA_UNROLL
        for (int u = 0; u < U; ++u) {
            const long long i = i0 + u * stride;
            vzero(first[u]);
            vzero(second[u]);
            l_lo[u] = 0; l_hi[u] = -1;
            if (i < p.total) {
                blend_coords<VEC, F32>(p, i, b[u], row[u], col[u], l_lo[u], l_hi[u]);
                src[u] = p.in + ((b[u] * p.num_hops + l_lo[u]) * p.rows + row[u]) * p.win + (col[u] - l_lo[u] * p.hop);
                if (l_hi[u] >= l_lo[u]) first[u] = *reinterpret_cast<const V*>(src[u]);
                if (l_hi[u] > l_lo[u]) second[u] = *reinterpret_cast<const V*>(src[u] + p.rows * p.win - p.hop);
            }
        }
        A2SB_PRAGMA_UNROLL
        for (int u = 0; u < U; ++u) {
            if (i0 + u * stride >= p.total) continue;
            V acc;
            vzero(acc);
            if (l_hi[u] >= l_lo[u]) vadd(acc, first[u]);      // ascending segment order, starting from +0 like the reference
            if (l_hi[u] > l_lo[u]) vadd(acc, second[u]);
            for (long long l = l_lo[u] + 2; l <= l_hi[u]; ++l)    // win > 2 hop
                vadd(acc, *reinterpret_cast<const V*>(src[u] + (l - l_lo[u]) * (p.rows * p.win - p.hop)));
            const float cnt = (float)(l_hi[u] >= l_lo[u] ? (l_hi[u] - l_lo[u] + 1) : 0);
            *reinterpret_cast<V*>(p.out + (b[u] * p.rows + row[u]) * p.out_pitch + (col[u] - p.col_off)) = vdiv(acc, cnt);
        }
    }
}


// ---- K4s: overlap blend fused with one reverse step of the bridge sampler -------------------------------------
// A2SB/A2SB_lightning_module.py:127-144 (ddpm_sample) around get_multidiffusion_vf, with the schedule scalars of
// A2SB/diffusion.py:153-168:
//     vf      = blend(segments)                                  (K4)
//     pred_x0 = x_t - std_fwd_t * vf                             (get_pred_x0, diffusion.py:165-168)
//     pred_x0 = pred_x0 * mask + (1 - mask) * x_1                (if mask and mask_pred_x0)
//     x_prev  = mu_x0 * pred_x0 + mu_xt * x_t [+ sd_post * n1]   (p_posterior, diffusion.py:153-163)
//     x_next  = (1 - mask) * (x_1 [+ std_sb * n2]) + mask * x_prev   (if mask)
// Each line is evaluated with separately rounded fp32 multiplies and adds in the reference's operand order
// (torch evaluates them as separate elementwise kernels), so the step is bit-identical to the reference for the
// same network outputs and noise tensors.  One pass over HBM instead of ~10 full-tensor passes and a D2H copy.
struct StepParams {
    SegParams seg;            // seg.in = network outputs per segment; seg.out unused
    const float* x_t;         // [batch][rows][width]
    const float* x_1;
    const float* mask;        // may be null
    const float* noise_post;  // n1, may be null (ot_ode or t_prev == 0)
    const float* noise_mask;  // n2, may be null (ot_ode or no mask)
    float* pred_x0;
    float* x_next;
    float std_fwd_t, mu_x0, mu_xt, sd_post, std_sb;
    int mask_pred_x0;
};

A2SB_DEV float f_mul(float a, float b) {
#ifdef A2SB_EMU
    volatile float r = a * b; return r;
#else
    return __fmul_rn(a, b);
#endif
}
A2SB_DEV float f_add(float a, float b) {
#ifdef A2SB_EMU
    volatile float r = a + b; return r;
#else
    return __fadd_rn(a, b);
#endif
}

A2SB_DEV void sampler_step(const StepParams& p, float vf, float xt, float x1, float m, float n1, float n2, float& pred,
                           float& xn) {
    pred = f_add(xt, -f_mul(p.std_fwd_t, vf));
    float om = 1.0f;
    if (p.mask) {
        om = f_add(1.0f, -m);
        if (p.mask_pred_x0) pred = f_add(f_mul(pred, m), f_mul(om, x1));
    }
    float xp = f_add(f_mul(p.mu_x0, pred), f_mul(p.mu_xt, xt));
    if (p.noise_post) xp = f_add(xp, f_mul(p.sd_post, n1));
    xn = xp;
    if (p.mask) {
        float xtrue = x1;
        if (p.noise_mask) xtrue = f_add(xtrue, f_mul(p.std_sb, n2));
        xn = f_add(f_mul(om, xtrue), f_mul(m, xp));
    }
}

template <int VEC, bool F32>
__global__ void __launch_bounds__(256) segment_blend_step_kernel(const StepParams sp) {
    using V = typename VecT<VEC>::type;
    const SegParams& p = sp.seg;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < p.total;
         i += (long long)gridDim.x * blockDim.x) {
        long long b, row, col, l_lo, l_hi;
        blend_coords<VEC, F32>(p, i, b, row, col, l_lo, l_hi);
        V acc;
        vzero(acc);
        for (long long l = l_lo; l <= l_hi; ++l) {
            const float* src = p.in + ((b * p.num_hops + l) * p.rows + row) * p.win + (col - l * p.hop);
            vadd(acc, *reinterpret_cast<const V*>(src));
        }
        const float cnt = (float)(l_hi >= l_lo ? (l_hi - l_lo + 1) : 0);
        const V vf = vdiv(acc, cnt);
        const long long base = (b * p.rows + row) * p.width + col;
        V zero;
        vzero(zero);
        const V xtv = *reinterpret_cast<const V*>(sp.x_t + base), x1v = *reinterpret_cast<const V*>(sp.x_1 + base);
        const V mv = sp.mask ? *reinterpret_cast<const V*>(sp.mask + base) : zero;
        const V n1v = sp.noise_post ? *reinterpret_cast<const V*>(sp.noise_post + base) : zero;
        const V n2v = sp.noise_mask ? *reinterpret_cast<const V*>(sp.noise_mask + base) : zero;
        const float* vfe = reinterpret_cast<const float*>(&vf);
        const float* xte = reinterpret_cast<const float*>(&xtv);
        const float* x1e = reinterpret_cast<const float*>(&x1v);
        const float* me = reinterpret_cast<const float*>(&mv);
        const float* n1e = reinterpret_cast<const float*>(&n1v);
        const float* n2e = reinterpret_cast<const float*>(&n2v);
        float pr[VEC], xn[VEC];
        A2SB_PRAGMA_UNROLL
        for (int e = 0; e < VEC; ++e) sampler_step(sp, vfe[e], xte[e], x1e[e], me[e], n1e[e], n2e[e], pr[e], xn[e]);
        V prv, xnv;
        A2SB_PRAGMA_UNROLL
        for (int e = 0; e < VEC; ++e) { reinterpret_cast<float*>(&prv)[e] = pr[e]; reinterpret_cast<float*>(&xnv)[e] = xn[e]; }
        *reinterpret_cast<V*>(sp.pred_x0 + base) = prv;
        *reinterpret_cast<V*>(sp.x_next + base) = xnv;
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
    DivMod d_ow;          // out_width
};

// out[row, w] = w < W ? in[row, w] : (const ? pad_const (with NaN propagation of in*0) : in[row, w - W])
A2SB_DEV float wrap_pad_value(const PadParams& p, long long row, long long w) {
    if (w < p.width) return p.in[row * p.width + w];
    float v = p.in[row * p.width + (w - p.width)];
    if (p.use_const) v = v * 0.0f + p.pad_const;  // diffusion.py:77 `padding*0+padding_constant`
    return v;
}

// VEC = 4: out_width % 4 == 0 and a 16-byte aligned output (p.total and p.d_ow then count float4 columns): one 128-bit
// store per thread.  The source row is only 4-byte aligned when the width is odd (config 3: 310,079 frames), so the four
// values are loaded one by one -- a warp still reads one contiguous 512-byte run.
template <int VEC, bool F32>
__global__ void __launch_bounds__(256) wrap_pad_kernel(const PadParams p) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < p.total;
         i += (long long)gridDim.x * blockDim.x) {
        long long w, row;
        divmod<F32>(i, p.d_ow, row, w);
        if (VEC == 1) {
            p.out[i] = wrap_pad_value(p, row, w);
        } else {
            float4 v;
            v.x = wrap_pad_value(p, row, 4 * w + 0);
            v.y = wrap_pad_value(p, row, 4 * w + 1);
            v.z = wrap_pad_value(p, row, 4 * w + 2);
            v.w = wrap_pad_value(p, row, 4 * w + 3);
            reinterpret_cast<float4*>(p.out)[i] = v;
        }
    }
}

}  // namespace a2sb
