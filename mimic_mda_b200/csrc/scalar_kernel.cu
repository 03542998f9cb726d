// scalar_kernel.cu -- the reference's scalar entry points on the GPU.
//
// One Ex bin, one parameter vector per call: calculateChi (C/ChiSquare.c:114-172),
// calculateLnLiklihood (:176-213), calculateLnLiklihoodResids (:216-252).  A call is a
// single small launch whose parameter vector and result live in mapped pinned host
// memory, so no separate copies are enqueued.  The host does not synchronise the stream either: thread 0 writes the
// result, fences at system scope and then writes the call's ticket into a mapped word the host spins on (chisq_abi.cu:
// scalar_eval) -- a stream synchronisation's wake-up was a third of the call.  What is left is launch latency, which
// is why bulk work belongs on the batched entry points instead.
//
// Threads own angles: thread j builds r_j with the reference's ascending-l update order
// and stores it; thread 0 then accumulates r_j^2 in ascending j from 0, as the
// reference's last loop does.
#include "mda_kernels.cuh"

namespace mda {

namespace {

constexpr int kScalarValueL = 32;          // parameter vectors up to this length travel in the kernel's argument block
constexpr int kScalarSmemPts = 4096;       // residuals of up to this many angles are kept in shared memory for the ordered sum

struct ScalarParams { double a[kScalarValueL]; };

// BY_VALUE: the parameters are kernel arguments (no read of mapped host memory over PCIe before anything can start).
template <bool BY_VALUE>
__global__ void __launch_bounds__(256) scalar_kernel(const double *__restrict__ consts, int n, int n_l, const ScalarParams pv,
                                                     const double *params, double *result, double *__restrict__ resid,
                                                     int mode, int resid_in_smem, unsigned long long *done,
                                                     unsigned long long ticket) {
    extern __shared__ double s_mem[];                     // [n_l] parameters, then [n] residuals if they fit
    double *s_a = s_mem, *s_r = s_mem + n_l;
    const int tid = threadIdx.x;
    for (int l = tid; l < n_l; l += blockDim.x) s_a[l] = BY_VALUE ? pv.a[l] : params[l];
    __syncthreads();
    for (int j = tid; j < n; j += blockDim.x) {
        double r = consts[j];
        for (int l = 0; l < n_l; ++l) r = fma(-s_a[l], consts[(size_t)(1 + l) * n + j], r);     // C/ChiSquare.c:188-202
        resid[j] = r;
        if (resid_in_smem) s_r[j] = r;
    }
    __syncthreads();
    if (tid == 0) {
        const double *r = resid_in_smem ? s_r : resid;
        double chi = 0.0;
#pragma unroll 8
        for (int j = 0; j < n; ++j) chi = fma(r[j], r[j], chi);                                   // ascending j, C/ChiSquare.c:204-210
        result[0] = (mode == kChi) ? chi : chi * -0.5;
        __threadfence_system();                           // the result before the ticket, as the host sees them
        *reinterpret_cast<volatile unsigned long long *>(done) = ticket;
    }
}

}  // namespace

// params_host: the caller's vector (read on the host when it fits the argument block); params_dev: its mapped device alias.
cudaError_t launch_scalar(const double *consts, int n, int n_l, const double *params_host, const double *params_dev, double *result,
                          double *resid, int mode, unsigned long long *done, unsigned long long ticket, cudaStream_t stream) {
    int threads = 32;
    while (threads < n && threads < 256) threads <<= 1;
    const int in_smem = n <= kScalarSmemPts ? 1 : 0;
    const size_t smem = sizeof(double) * ((size_t)(n_l > 0 ? n_l : 1) + (in_smem ? (size_t)n : 0));
    ScalarParams pv;
    if (n_l <= kScalarValueL) {
        for (int l = 0; l < kScalarValueL; ++l) pv.a[l] = l < n_l ? params_host[l] : 0.0;
        scalar_kernel<true><<<1, threads, smem, stream>>>(consts, n, n_l, pv, nullptr, result, resid, mode, in_smem, done, ticket);
    } else {
        scalar_kernel<false><<<1, threads, smem, stream>>>(consts, n, n_l, pv, params_dev, result, resid, mode, in_smem, done, ticket);
    }
    return cudaGetLastError();
}

}  // namespace mda
