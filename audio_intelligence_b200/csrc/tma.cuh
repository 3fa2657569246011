// tma.cuh -- mbarrier, 1-D bulk asynchronous copy (TMA, SASS UBLKCP), tensor-map box loads (SASS UTMALDG) and L2
// prefetch wrappers.
#pragma once
#include "a2sb_common.cuh"
#ifndef A2SB_EMU
#include <cuda.h>   // CUtensorMap (type only; cuTensorMapEncodeTiled is fetched through cudaGetDriverEntryPoint)
#endif

namespace a2sb {

#ifdef A2SB_EMU
// Emulation: copies are performed synchronously by the issuing thread; mbarriers are modelled (arrival count, transaction
// bytes, phase) so that producer / consumer protocols between warps are really exercised.
A2SB_DEV void mbar_init(unsigned long long* b, int n) { emu::mbar_init(b, n); }
A2SB_DEV void fence_mbar_init() {}
A2SB_DEV void fence_proxy_async() {}
A2SB_DEV void mbar_wait(unsigned long long* b, unsigned parity) { emu::mbar_wait(b, parity); }
A2SB_DEV void mbar_arrive(unsigned long long* b) { emu::mbar_arrive(b, 0); }
A2SB_DEV void bulk_g2s(void* dst, const void* src, unsigned bytes, unsigned long long* b) {
    emu::mbar_arrive(b, bytes);
    std::memcpy(dst, src, bytes);
    emu::mbar_complete_tx(b, bytes);
}
A2SB_DEV void prefetch_l2(const void*) {}
A2SB_DEV void prefetch_seam(const void*) {}

// 5-D tensor map of fp32 elements (dimension 0 contiguous); strides in bytes for dimensions 1..4
struct TensorMap5 {
    const char* base;
    long long dims[5];
    long long strides[4];
    int box[5];
};
// Box load with the hardware's out-of-bounds rule: elements whose coordinate falls outside [0, dim) read as zero.
A2SB_DEV void tma_load_box5(void* dst_smem, const TensorMap5* m, unsigned long long* bar, unsigned nbytes, int c0, int c1, int c2,
                            int c3, int c4) {
    const long long bytes = 4LL * m->box[0] * m->box[1] * m->box[2] * m->box[3] * m->box[4];
    if (bytes != (long long)nbytes || (c0 & 3) != 0 || (reinterpret_cast<uintptr_t>(m->base) & 15) != 0) {
        std::fprintf(stderr, "emu: bad TMA box (bytes %lld vs %u, c0 %d)\n", bytes, nbytes, c0);
        std::abort();
    }
    emu::mbar_arrive(bar, bytes);
    float* d = static_cast<float*>(dst_smem);
    const int c[5] = {c0, c1, c2, c3, c4};
    for (int i4 = 0; i4 < m->box[4]; ++i4)
        for (int i3 = 0; i3 < m->box[3]; ++i3)
            for (int i2 = 0; i2 < m->box[2]; ++i2)
                for (int i1 = 0; i1 < m->box[1]; ++i1)
                    for (int i0 = 0; i0 < m->box[0]; ++i0) {
                        const long long x[5] = {c[0] + i0, c[1] + i1, c[2] + i2, c[3] + i3, c[4] + i4};
                        bool in = true;
                        for (int k = 0; k < 5; ++k) in = in && x[k] >= 0 && x[k] < m->dims[k];
                        float v = 0.0f;
                        if (in) {
                            const char* a = m->base + 4 * x[0];
                            for (int k = 1; k < 5; ++k) a += x[k] * m->strides[k - 1];
                            std::memcpy(&v, a, 4);
                        }
                        *d++ = v;
                    }
    emu::mbar_complete_tx(bar, bytes);
}
A2SB_DEV void tma_prefetch_box5(const TensorMap5* m, int c0, int, int, int, int) {
    if ((c0 & 3) != 0 || (reinterpret_cast<uintptr_t>(m->base) & 15) != 0) { std::fprintf(stderr, "emu: bad TMA prefetch box\n"); std::abort(); }
}
#else
A2SB_DEV unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }

A2SB_DEV void mbar_init(unsigned long long* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
A2SB_DEV void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
// Order prior generic-proxy shared-memory accesses before subsequent async-proxy (TMA) accesses.
A2SB_DEV void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

A2SB_DEV void mbar_wait(unsigned long long* bar, unsigned parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra WAIT_DONE;\n"
        "bra WAIT_LOOP;\n"
        "WAIT_DONE:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}

// plain arrival (release semantics at CTA scope)
A2SB_DEV void mbar_arrive(unsigned long long* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// One elected thread: arm the barrier with the byte count and issue the bulk copy
// global -> shared (16-byte aligned addresses, size a multiple of 16).
A2SB_DEV void bulk_g2s(void* dst_smem, const void* src_gmem, unsigned bytes, unsigned long long* bar) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst_smem)),
        "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}
A2SB_DEV void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
// Seam-sector prefetch of the forward kernel.  A2SB_SEAM_PF (experiments): 1 = prefetch with the evict_last priority,
// 2 = a real load with the evict_last priority (result discarded).
A2SB_DEV void prefetch_seam(const void* p) {
#if A2SB_SEAM_PF == 1
    asm volatile("prefetch.global.L2::evict_last [%0];" ::"l"(p));
#elif A2SB_SEAM_PF == 2
    unsigned v;
    unsigned long long pol;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
    asm volatile("ld.global.L2::cache_hint.u32 %0, [%1], %2;" : "=r"(v) : "l"(p), "l"(pol));
#else
    asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
#endif
}

// Tensor-map TMA.  Hardware rules that shape the callers (measured, tools/microbench/tma_debug.cu): global strides are
// multiples of 16 bytes, and the byte address of the box start (base + 4 * c0) must be 16-byte aligned as well -- a box
// whose first coordinate is not a multiple of 4 fp32 elements raises "illegal instruction".
using TensorMap5 = CUtensorMap;
// One elected thread: arm the barrier with the box size and issue the box load (coordinates in elements, dimension 0 first).
A2SB_DEV void tma_load_box5(void* dst_smem, const TensorMap5* map, unsigned long long* bar, unsigned bytes, int c0, int c1, int c2,
                            int c3, int c4) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
    asm volatile(
        "cp.async.bulk.tensor.5d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];" ::"r"(
            smem_u32(dst_smem)),
        "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
        : "memory");
}
// One thread: pull the box into L2 (no shared-memory destination, no completion tracking).
A2SB_DEV void tma_prefetch_box5(const TensorMap5* map, int c0, int c1, int c2, int c3, int c4) {
    asm volatile("cp.async.bulk.prefetch.tensor.5d.L2.global.tile [%0, {%1, %2, %3, %4, %5}];" ::"l"(map), "r"(c0), "r"(c1), "r"(c2),
                 "r"(c3), "r"(c4)
                 : "memory");
}
#endif

}  // namespace a2sb
