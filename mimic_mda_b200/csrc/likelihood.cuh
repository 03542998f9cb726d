// likelihood.cuh -- device-side building blocks shared by the batched kernel
// (lnpost_kernel.cu) and the device-resident ensemble stepper (ensemble_kernel.cu):
// mbarrier / bulk-copy PTX wrappers and the register-tiled chi-square inner loop.
#pragma once
#include "mda_kernels.cuh"

#include <math_constants.h>

namespace mda {

// ---- PTX helpers (sm_90+: mbarrier transactions, bulk async copies) ---------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
            smem_u32(dst)),
        "l"(src), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "MDA_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra MDA_DONE;\n"
        "bra MDA_WAIT;\n"
        "MDA_DONE:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void cp_async_8(void *dst, const void *src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(smem_u32(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit_wait_all() {
    asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory");
}

__device__ __forceinline__ double finish(double chi, bool outside, int mode) {
    if (mode == kChi) return chi;
    const double lnl = chi * -0.5;                       // == chi / (-2.0) exactly
    return (mode == kLnPost && outside) ? -CUDART_INF : lnl;
}


// chi^2 contributions of `valid` angles staged angle-pair-major at p ([pairs][1+L] double2:
// data pair, then the pair of every distribution) for WPT parameter vectors held in
// registers.  Operation order per walker is the reference's: r_j = d_j - a_0 D_0j, then
// r_j -= a_l D_lj for ascending l (C/ChiSquare.c:188-202), chi += r_j^2 for ascending j
// (C/ChiSquare.c:204-210).  chi[] is carried across calls (angle chunks).
template <int L, int WPT, int UNROLL = 2>
__device__ __forceinline__ void accumulate_chi(const double2 *p, int valid, const double (&a)[WPT][L],
                                               double (&chi)[WPT]) {
    constexpr int ROW = L + 1;
    const int pairs = valid >> 1;
#pragma unroll UNROLL
    for (int j2 = 0; j2 < pairs; ++j2, p += ROW) {
        const double2 dv = p[0];
        double r0[WPT], r1[WPT];
#pragma unroll
        for (int k = 0; k < WPT; ++k) { r0[k] = dv.x; r1[k] = dv.y; }
#pragma unroll
        for (int l = 0; l < L; ++l) {
            const double2 D = p[1 + l];
#pragma unroll
            for (int k = 0; k < WPT; ++k) {
                r0[k] = fma(-a[k][l], D.x, r0[k]);     // C/ChiSquare.c:190,200
                r1[k] = fma(-a[k][l], D.y, r1[k]);
            }
        }
#pragma unroll
        for (int k = 0; k < WPT; ++k) {
            chi[k] = fma(r0[k], r0[k], chi[k]);        // C/ChiSquare.c:209, ascending j
            chi[k] = fma(r1[k], r1[k], chi[k]);
        }
    }
    if (valid & 1) {                                   // last, unpaired angle of the bin
        const double dv = p[0].x;
        double r0[WPT];
#pragma unroll
        for (int k = 0; k < WPT; ++k) r0[k] = dv;
#pragma unroll
        for (int l = 0; l < L; ++l) {
            const double Dx = p[1 + l].x;
#pragma unroll
            for (int k = 0; k < WPT; ++k) r0[k] = fma(-a[k][l], Dx, r0[k]);
        }
#pragma unroll
        for (int k = 0; k < WPT; ++k) chi[k] = fma(r0[k], r0[k], chi[k]);
    }
}

}  // namespace mda
