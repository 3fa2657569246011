// inst.cu -- explicit kernel families per n_fft.  Compiled once per family with -DA2SB_INST=k so
// the heavy template instantiations build in parallel; -DA2SB_INST_ALL builds every family in one
// translation unit (used by the CPU emulation build in tests/emu).
#include "host_util.h"
#include "istft_inv.cuh"
#include "stft_fwd.cuh"

namespace a2sb {

template <int M, int RA, int RB>
static int dispatch_fwd(const LaunchCtx& cx, const FwdParams& p, int kind, int power_on, float power, cudaStream_t st) {
    using G = FwdGeom<M, RA, RB>;
    const size_t smem = G::smem_bytes(cx.hop);
    if (kind == A2SB_KIND_COMPLEX)
        return launch_persistent(stft_fwd_kernel<M, RA, RB, kEpiComplex, kPowNone>, p.total_tiles, G::NT, smem, st, p,
                                 cx.sm_count);
    if (!power_on)
        return launch_persistent(stft_fwd_kernel<M, RA, RB, kEpiMagPhase, kPowNone>, p.total_tiles, G::NT, smem, st, p,
                                 cx.sm_count);
    if (power == 0.25f)
        return launch_persistent(stft_fwd_kernel<M, RA, RB, kEpiMagPhase, kPowQuarter>, p.total_tiles, G::NT, smem, st,
                                 p, cx.sm_count);
    return launch_persistent(stft_fwd_kernel<M, RA, RB, kEpiMagPhase, kPowGeneric>, p.total_tiles, G::NT, smem, st, p,
                             cx.sm_count);
}

template <int M, int RA, int RB>
static int dispatch_inv(const LaunchCtx& cx, const InvParams& p, int kind, int power_on, float power, cudaStream_t st) {
    using G = InvGeom<M, RA, RB>;
    const size_t smem = G::smem_bytes(cx.hop);
    if (kind == A2SB_KIND_COMPLEX)
        return launch_persistent(istft_inv_kernel<M, RA, RB, kInComplex, kPowNone>, p.total_items, G::NT, smem, st, p,
                                 cx.sm_count);
    if (!power_on)
        return launch_persistent(istft_inv_kernel<M, RA, RB, kInMagPhase, kPowNone>, p.total_items, G::NT, smem, st, p,
                                 cx.sm_count);
    if (power == 4.0f)
        return launch_persistent(istft_inv_kernel<M, RA, RB, kInMagPhase, kPowFour>, p.total_items, G::NT, smem, st, p,
                                 cx.sm_count);
    return launch_persistent(istft_inv_kernel<M, RA, RB, kInMagPhase, kPowGeneric>, p.total_items, G::NT, smem, st, p,
                             cx.sm_count);
}

#if defined(A2SB_INST_ALL) || A2SB_INST == 1
int run_fwd_256(const LaunchCtx& c, const FwdParams& p, int k, int on, float pw, cudaStream_t s) { return dispatch_fwd<256, 16, 16>(c, p, k, on, pw, s); }
#endif
#if defined(A2SB_INST_ALL) || A2SB_INST == 2
int run_fwd_512(const LaunchCtx& c, const FwdParams& p, int k, int on, float pw, cudaStream_t s) { return dispatch_fwd<512, 16, 32>(c, p, k, on, pw, s); }
#endif
#if defined(A2SB_INST_ALL) || A2SB_INST == 3
int run_fwd_1024(const LaunchCtx& c, const FwdParams& p, int k, int on, float pw, cudaStream_t s) { return dispatch_fwd<1024, 32, 32>(c, p, k, on, pw, s); }
#endif
#if defined(A2SB_INST_ALL) || A2SB_INST == 4
int run_inv_256(const LaunchCtx& c, const InvParams& p, int k, int on, float pw, cudaStream_t s) { return dispatch_inv<256, 16, 16>(c, p, k, on, pw, s); }
#endif
#if defined(A2SB_INST_ALL) || A2SB_INST == 5
int run_inv_512(const LaunchCtx& c, const InvParams& p, int k, int on, float pw, cudaStream_t s) { return dispatch_inv<512, 32, 16>(c, p, k, on, pw, s); }
#endif
#if defined(A2SB_INST_ALL) || A2SB_INST == 6
int run_inv_1024(const LaunchCtx& c, const InvParams& p, int k, int on, float pw, cudaStream_t s) { return dispatch_inv<1024, 32, 32>(c, p, k, on, pw, s); }
#endif

}  // namespace a2sb
