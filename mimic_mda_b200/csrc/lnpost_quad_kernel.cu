// lnpost_quad_kernel.cu -- the batched log-posterior from the bins' triangular factors (MDA_BATCH_IMPL=quad).
//
// Same call as lnpost_tile_kernel (all bins x walkers in one launch: box prior of mda.py:1574-1576, then
// calculateLnLiklihood / calculateChi, C/ChiSquare.c:114-213), with chi^2(a) = |T [a; -1]|^2 (quad_chi.cuh):
// (L+1)(L+2)/2 multiply-adds per evaluation whatever the number of angles, so the launch is bound by its
// 8 (L+1) bytes per evaluation of HBM traffic instead of by the fp64 pipe.  One thread per walker; a warp's parameter
// rows are one contiguous run of memory.  Not the default of the batched entry points: those keep the reference's
// angle-by-angle arithmetic (bit-for-bit independent of tiling); this variant agrees with it to ~1e-14 relative.
#include "quad_chi.cuh"

#include <math_constants.h>

namespace mda {

namespace {

constexpr int kThreads = 256;

template <int L>
__global__ void __launch_bounds__(kThreads) lnpost_quad_kernel(const BatchArgs a) {
    __shared__ __align__(16) double sTri[quad_padded_doubles(L)];
    const int bin = blockIdx.y;
    load_factor<L>(sTri, a.quad + (size_t)bin * quad_stride(L), threadIdx.x, kThreads);
    __syncthreads();
    const int w = blockIdx.x * kThreads + threadIdx.x;
    if (w >= a.walkers) return;
    const double *p = a.params + ((size_t)bin * a.walkers + w) * L;
    double q[L];
    bool outside = false;
#pragma unroll
    for (int l = 0; l < L; ++l) {
        q[l] = __ldcs(p + l);
        outside = outside | (q[l] < a.lo[l]) | (q[l] > a.hi[l]);              // a NaN compares false, as in mda.py:1575
    }
    const double chi = chi_from_factor<L>(sTri, q);
    double r = a.mode == kChi ? chi : chi * -0.5;                             // C/ChiSquare.c:171 / :212
    if (a.mode == kLnPost && outside) r = -CUDART_INF;                        // mda.py:1576
    a.out[(size_t)bin * a.walkers + w] = r;
}

using QuadKernel = void (*)(const BatchArgs);

QuadKernel pick(int n_l) {
    switch (n_l) {
        case 1: return lnpost_quad_kernel<1>;
        case 2: return lnpost_quad_kernel<2>;
        case 3: return lnpost_quad_kernel<3>;
        case 4: return lnpost_quad_kernel<4>;
        case 5: return lnpost_quad_kernel<5>;
        case 6: return lnpost_quad_kernel<6>;
        case 7: return lnpost_quad_kernel<7>;
        case 8: return lnpost_quad_kernel<8>;
        case 9: return lnpost_quad_kernel<9>;
        case 10: return lnpost_quad_kernel<10>;
        case 11: return lnpost_quad_kernel<11>;
        case 12: return lnpost_quad_kernel<12>;
        case 13: return lnpost_quad_kernel<13>;
        case 14: return lnpost_quad_kernel<14>;
        case 15: return lnpost_quad_kernel<15>;
        case 16: return lnpost_quad_kernel<16>;
        default: return nullptr;
    }
}

}  // namespace

bool lnpost_quad_supported(int n_l) { return n_l >= 1 && n_l <= kMaxRegL; }

cudaError_t launch_batch_quad(const BatchArgs &args, cudaStream_t stream) {
    if (args.bins <= 0 || args.walkers <= 0) return cudaSuccess;
    QuadKernel k = pick(args.n_l);
    if (!k || !args.quad || args.bins > 65535) return cudaErrorInvalidValue;
    const dim3 grid((unsigned)((args.walkers + kThreads - 1) / kThreads), (unsigned)args.bins);
    k<<<grid, kThreads, 0, stream>>>(args);
    return cudaGetLastError();
}

}  // namespace mda
