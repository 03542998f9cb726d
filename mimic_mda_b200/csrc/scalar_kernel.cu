// scalar_kernel.cu -- the reference's scalar entry points on the GPU.
//
// One Ex bin, one parameter vector per call: calculateChi (C/ChiSquare.c:114-172),
// calculateLnLiklihood (:176-213), calculateLnLiklihoodResids (:216-252).  A call is a
// single small launch whose parameter vector and result live in mapped pinned host
// memory, so no separate copies are enqueued; the cost is launch + sync latency, which
// is why bulk work belongs on the batched entry points instead.
//
// Threads own angles: thread j builds r_j with the reference's ascending-l update order
// and stores it; thread 0 then accumulates r_j^2 in ascending j from 0, as the
// reference's last loop does.
#include "mda_kernels.cuh"

namespace mda {

namespace {

__global__ void __launch_bounds__(256) scalar_kernel(const double *__restrict__ consts, int n, int n_l,
                                                     const double *params, double *result,
                                                     double *__restrict__ resid, int mode) {
    extern __shared__ double s_a[];                       // [n_l]
    const int tid = threadIdx.x;
    for (int l = tid; l < n_l; l += blockDim.x) s_a[l] = params[l];
    __syncthreads();
    for (int j = tid; j < n; j += blockDim.x) {
        double r = consts[j];
        for (int l = 0; l < n_l; ++l) r = fma(-s_a[l], consts[(size_t)(1 + l) * n + j], r);
        resid[j] = r;
    }
    __syncthreads();
    if (tid == 0) {
        double chi = 0.0;
        for (int j = 0; j < n; ++j) chi = fma(resid[j], resid[j], chi);
        result[0] = (mode == kChi) ? chi : chi * -0.5;
    }
}

}  // namespace

cudaError_t launch_scalar(const double *consts, int n, int n_l, const double *params, double *result,
                          double *resid, int mode, cudaStream_t stream) {
    int threads = 32;
    while (threads < n && threads < 256) threads <<= 1;
    const size_t smem = sizeof(double) * (size_t)(n_l > 0 ? n_l : 1);
    scalar_kernel<<<1, threads, smem, stream>>>(consts, n, n_l, params, result, resid, mode);
    return cudaGetLastError();
}

}  // namespace mda
