// peaks.cu -- roofline denominators measured on the device in use.
//
// MEASURED_PEAKS.json carries an HBM copy rate and a bf16 GEMM rate but no fp64
// figure, and the likelihood kernel is bound by the fp64 FMA pipe.
// measure_fp64_peak times dependency-free DFMA streams on every SM (16 independent
// accumulators per thread, 8 warps per scheduler), 2 flop per FMA, and reports the better
// of two operand patterns: (a) acc = fma(acc, x, y) with x, y in registers, and (b)
// acc = fma(a, U, acc) with `a` re-used by consecutive DFMAs and U in a uniform register.
// (b) needs the fewest register-file reads per DFMA and is the true pipe peak
// (36.3 vs 33.8 TFLOP/s on a B200): register-operand bandwidth, not the pipe, is what a
// DFMA with three distinct register operands runs into (profiles/experiments/).
#include "mda_kernels.cuh"

namespace mda {

namespace {

constexpr int kChains = 16;
constexpr int kInner = 64;

__global__ void __launch_bounds__(256) dfma_stream_kernel(double *sink, int outer, double x, double y) {
    double acc[kChains];
#pragma unroll
    for (int c = 0; c < kChains; ++c) acc[c] = (double)(threadIdx.x + c) * 1e-3;
    for (int o = 0; o < outer; ++o) {
#pragma unroll
        for (int i = 0; i < kInner; ++i) {
#pragma unroll
            for (int c = 0; c < kChains; ++c) acc[c] = fma(acc[c], x, y);
        }
    }
    double s = 0.0;
#pragma unroll
    for (int c = 0; c < kChains; ++c) s += acc[c];
    if (s == 0.123456789) sink[blockIdx.x * blockDim.x + threadIdx.x] = s;   // never true; keeps the math
}

// pattern (b): the multiplicands come from uniform global loads that are loop invariant,
// which ptxas keeps in uniform registers (DFMA R, R.reuse, UR, R)
__global__ void __launch_bounds__(256) dfma_uniform_kernel(double *sink, int outer, const double *__restrict__ in) {
    double acc[kChains], a[kChains], b[kChains];
#pragma unroll
    for (int c = 0; c < kChains; ++c) {
        acc[c] = in[c] + (double)threadIdx.x;
        a[c] = in[16 + c] * (1.0 + 1e-12 * threadIdx.x);
        b[c] = in[32 + c];
    }
    for (int o = 0; o < outer; ++o) {
#pragma unroll
        for (int i = 0; i < kInner / 4; ++i) {
#pragma unroll
            for (int q = 0; q < 4; ++q) {
#pragma unroll
                for (int c = 0; c < kChains; ++c) acc[c] = fma(a[(c / 4 + q) % kChains], b[(i + c) % kChains], acc[c]);
            }
        }
    }
    double s = 0.0;
#pragma unroll
    for (int c = 0; c < kChains; ++c) s += acc[c];
    if (s == 0.123456789) sink[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

}  // namespace

cudaError_t measure_fp64_peak(int repeats, int sm_count, double *tflops) {
    *tflops = 0.0;
    const int threads = 256, blocks = sm_count * 8, outer = 256;
    double *sink = nullptr;
    cudaError_t e = cudaMalloc(&sink, sizeof(double) * (size_t)threads * blocks);
    if (e != cudaSuccess) return e;
    cudaEvent_t t0, t1;
    cudaEventCreate(&t0);
    cudaEventCreate(&t1);
    double *uni = nullptr;
    e = cudaMalloc(&uni, sizeof(double) * 64);
    if (e != cudaSuccess) { cudaFree(sink); return e; }
    double host_uni[64];
    for (int i = 0; i < 64; ++i) host_uni[i] = 1e-3 * (1.0 + 1e-3 * i);
    cudaMemcpy(uni, host_uni, sizeof(host_uni), cudaMemcpyHostToDevice);
    dfma_stream_kernel<<<blocks, threads>>>(sink, 8, 0.999999, 1e-9);       // warm-up
    dfma_uniform_kernel<<<blocks, threads>>>(sink, 8, uni);
    double best = 0.0;
    for (int r = 0; r < 2 * (repeats > 0 ? repeats : 1); ++r) {
        cudaEventRecord(t0);
        if (r & 1) dfma_uniform_kernel<<<blocks, threads>>>(sink, outer, uni);
        else dfma_stream_kernel<<<blocks, threads>>>(sink, outer, 0.999999, 1e-9);
        cudaEventRecord(t1);
        e = cudaEventSynchronize(t1);
        if (e != cudaSuccess) break;
        float ms = 0.f;
        cudaEventElapsedTime(&ms, t0, t1);
        const double flop = 2.0 * (double)threads * blocks * (double)outer * kInner * kChains;
        const double tf = flop / ((double)ms * 1e-3) / 1e12;
        if (tf > best) best = tf;
    }
    cudaEventDestroy(t0);
    cudaEventDestroy(t1);
    cudaFree(sink);
    cudaFree(uni);
    if (e == cudaSuccess) e = cudaGetLastError();
    *tflops = best;
    return e;
}

cudaError_t measure_hbm_copy(size_t bytes, int repeats, double *gbs) {
    *gbs = 0.0;
    void *a = nullptr, *b = nullptr;
    cudaError_t e = cudaMalloc(&a, bytes);
    if (e != cudaSuccess) return e;
    e = cudaMalloc(&b, bytes);
    if (e != cudaSuccess) { cudaFree(a); return e; }
    cudaMemset(a, 1, bytes);
    cudaMemset(b, 2, bytes);
    cudaEvent_t t0, t1;
    cudaEventCreate(&t0);
    cudaEventCreate(&t1);
    double best = 0.0;
    for (int r = 0; r < (repeats > 0 ? repeats : 1) + 1; ++r) {
        cudaEventRecord(t0);
        cudaMemcpyAsync(b, a, bytes, cudaMemcpyDeviceToDevice);
        cudaEventRecord(t1);
        e = cudaEventSynchronize(t1);
        if (e != cudaSuccess) break;
        float ms = 0.f;
        cudaEventElapsedTime(&ms, t0, t1);
        const double rate = 2.0 * (double)bytes / ((double)ms * 1e-3) / 1e9;
        if (r > 0 && rate > best) best = rate;
    }
    cudaEventDestroy(t0);
    cudaEventDestroy(t1);
    cudaFree(a);
    cudaFree(b);
    if (e == cudaSuccess) e = cudaGetLastError();
    *gbs = best;
    return e;
}

}  // namespace mda
