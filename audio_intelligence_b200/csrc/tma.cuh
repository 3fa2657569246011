// tma.cuh -- mbarrier + 1-D bulk asynchronous copy (TMA, SASS UBLKCP) and L2 prefetch wrappers.
#pragma once
#include "a2sb_common.cuh"

namespace a2sb {

#ifdef A2SB_EMU
// Emulation: the copy is performed synchronously by the issuing thread; callers always have a
// __syncthreads() between the issue and the first consumer, so waits are no-ops.
A2SB_DEV void mbar_init(unsigned long long*, int) {}
A2SB_DEV void fence_mbar_init() {}
A2SB_DEV void fence_proxy_async() {}
A2SB_DEV void mbar_wait(unsigned long long*, unsigned) {}
A2SB_DEV void bulk_g2s(void* dst, const void* src, unsigned bytes, unsigned long long*) {
    std::memcpy(dst, src, bytes);
}
A2SB_DEV void prefetch_l2(const void*) {}
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
#endif

}  // namespace a2sb
