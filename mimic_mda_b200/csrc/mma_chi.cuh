// mma_chi.cuh -- the chi-square contraction on the fp64 tensor path (DMMA.8x8x4), shared by the device-resident
// samplers (ensemble_mma_kernel.cu, ensemble_ws_kernel.cu) and the batched log-posterior (lnpost_mma_kernel.cu).
//
// Constants per bin and angle tile t (8 angles, zero padded), 8 (1 + L) doubles (see mda_kernels.cuh):
//   rows 0..R  (R = L % 4): [8] data, then -D_l for l < R;   k-step ks < L / 4: [32] B fragment, lane (g, tig) holds
//   -D_{R + 4 ks + tig}[8 t + g].  The distributions are stored negated, so the residual d - sum a_l D_l
//   (C/ChiSquare.c:188-202) is a pure multiply-accumulate onto d.
#pragma once
#include "likelihood.cuh"

namespace mda {

__device__ __forceinline__ void dmma884(double &c0, double &c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}
// chi^2 of WT tiles of 8 walkers (A fragments a[][], leading elements lead[][]) against NTC angle tiles
// (NTC = 0: nt of them, a run-time count).  Constants are fetched one tile ahead of their use.
template <int L, int WT, int NTC>
__device__ __forceinline__ void chi_tiles(const double *sC, int nt, int lane, const double (&a)[WT][L / 4],
                                          const double (&lead)[WT][(L % 4) > 0 ? (L % 4) : 1], double (&chi0)[WT], double (&chi1)[WT]) {
    constexpr int Q = L / 4, R = L % 4, RR = R > 0 ? R : 1, TS = 8 * (L + 1);
    const int tig = lane & 3;
    const int n = NTC > 0 ? NTC : nt;
    const double *tp = sC;
    double2 dv = *reinterpret_cast<const double2 *>(tp + 2 * tig);
    double2 Di[RR];
#pragma unroll
    for (int l = 0; l < R; ++l) Di[l] = *reinterpret_cast<const double2 *>(tp + 8 * (1 + l) + 2 * tig);
    double Bf[Q];
#pragma unroll
    for (int ks = 0; ks < Q; ++ks) Bf[ks] = tp[8 * (1 + R) + 32 * ks + lane];
#pragma unroll(NTC > 0 ? NTC : 2)
    for (int at = 0; at < n; ++at) {
        tp = sC + min(at + 1, n - 1) * TS;
        const double2 dv_n = *reinterpret_cast<const double2 *>(tp + 2 * tig);
        double2 Di_n[RR];
#pragma unroll
        for (int l = 0; l < R; ++l) Di_n[l] = *reinterpret_cast<const double2 *>(tp + 8 * (1 + l) + 2 * tig);
        double Bf_n[Q];
#pragma unroll
        for (int ks = 0; ks < Q; ++ks) Bf_n[ks] = tp[8 * (1 + R) + 32 * ks + lane];
        double c0[WT], c1[WT];
#pragma unroll
        for (int k = 0; k < WT; ++k) {
            c0[k] = dv.x;
            c1[k] = dv.y;
#pragma unroll
            for (int l = 0; l < R; ++l) {
                c0[k] = fma(lead[k][l], Di[l].x, c0[k]);
                c1[k] = fma(lead[k][l], Di[l].y, c1[k]);
            }
        }
#pragma unroll
        for (int ks = 0; ks < Q; ++ks)
#pragma unroll
            for (int k = 0; k < WT; ++k) dmma884(c0[k], c1[k], a[k][ks], Bf[ks]);
#pragma unroll
        for (int k = 0; k < WT; ++k) {
            chi0[k] = fma(c0[k], c0[k], chi0[k]);
            chi1[k] = fma(c1[k], c1[k], chi1[k]);
        }
        dv = dv_n;
#pragma unroll
        for (int l = 0; l < R; ++l) Di[l] = Di_n[l];
#pragma unroll
        for (int ks = 0; ks < Q; ++ks) Bf[ks] = Bf_n[ks];
    }
}

}  // namespace mda
