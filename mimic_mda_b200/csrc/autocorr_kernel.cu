// autocorr_kernel.cu -- integrated autocorrelation time on the device-resident chain.
//
// The reference asks its sampler for `get_autocorr_time(c=CONFIG["ACorr WindSize"], fast=CONFIG["ACorr Use FFT"])`
// (mda.py:1304-1308) and writes the times into the diagnostics file (mda.py:253-272).  emcee 2.x computes it
// from the ENSEMBLE-MEAN series -- np.mean(sampler.chain, axis=0), one value per step and parameter -- as the
// normalised autocorrelation function summed over a self-consistent window (Sokal).  The chain of a run is
// gigabytes and lives in HBM; the result is one number per bin and parameter, so the two reductions run here:
//   ensemble_mean_kernel   series[col][k] = mean over the walkers of kept sample k (col = bin * L + l), the
//                          walkers added one after another as numpy's reduction over the leading axis does
//   series_centre_kernel   y = x - mean(x) over the first n samples of a column (sequential sum, as numpy)
//   acf_kernel             acf[col][lag] = sum_{t < n - lag} y[t] y[t + lag] for the lags the window search can
//                          reach -- the linear autocorrelation emcee obtains from a zero-padded FFT, evaluated
//                          directly (fewer roundings; only n / (2c) of the n lags are ever read)
// The window search itself is a few hundred scalar operations per bin and runs on the host (chisq_abi.cu).
#include "mda_kernels.cuh"

namespace mda {

namespace {

constexpr int kMeanThreads = 256;
constexpr int kLagThreads = 256;       // lags per CTA of acf_kernel, one per thread
constexpr int kTile = 1024;            // samples of y staged per pass

// Thread (k, bin, l), l fastest: consecutive threads read consecutive doubles of a walker's row.
__global__ void __launch_bounds__(kMeanThreads) ensemble_mean_kernel(const double *__restrict__ chain, long long kept, int bins,
                                                                     int walkers, int n_l, double *__restrict__ series,
                                                                     long long stride) {
    const long long idx = (long long)blockIdx.x * kMeanThreads + threadIdx.x;
    if (idx >= kept * bins * n_l) return;
    const int l = (int)(idx % n_l);
    const long long slab = idx / n_l;                      // k * bins + bin
    const int bin = (int)(slab % bins);
    const long long k = slab / bins;
    const double *p = chain + (size_t)slab * walkers * n_l + l;
    double sum = 0.0;
    int w = 0;
    for (; w + 8 <= walkers; w += 8) {                     // eight loads in flight, added in walker order
        double x[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) x[i] = __ldcs(p + (size_t)(w + i) * n_l);
#pragma unroll
        for (int i = 0; i < 8; ++i) sum += x[i];
    }
    for (; w < walkers; ++w) sum += __ldcs(p + (size_t)w * n_l);
    series[((size_t)bin * n_l + l) * stride + k] = sum / (double)walkers;
}

// One warp per column: lane 0 sums the first n samples in order (numpy's mean over the leading axis of
// [steps, dim]), then the warp writes y = x - mean.
__global__ void __launch_bounds__(128) series_centre_kernel(const double *__restrict__ series, long long stride, int n, int cols,
                                                            double *__restrict__ centred) {
    const int col = blockIdx.x * 4 + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (col >= cols) return;
    const double *x = series + (size_t)col * stride;
    double sum = 0.0;
    for (int t0 = 0; t0 < n; t0 += 32) {                   // a coalesced load, then 32 ordered additions on lane 0
        const double mine = t0 + lane < n ? x[t0 + lane] : 0.0;
        const int m = min(32, n - t0);
        for (int i = 0; i < m; ++i) {
            const double xi = __shfl_sync(0xffffffffu, mine, i);
            sum += xi;
        }
    }
    const double mean = sum / (double)n;
    double *y = centred + (size_t)col * stride;
    for (int t = lane; t < n; t += 32) y[t] = x[t] - mean;
}

// grid = (cols, ceil(n_lags / kLagThreads)); thread = one lag.  Per pass a tile of y and the tile shifted by
// the CTA's first lag (plus kLagThreads of halo, zero beyond n) sit in shared memory: y[t] is a broadcast read,
// y[t + lag] a conflict-free one.
__global__ void __launch_bounds__(kLagThreads) acf_kernel(const double *__restrict__ centred, long long stride, int n, int n_lags,
                                                          double *__restrict__ acf) {
    __shared__ double sa[kTile];
    __shared__ double sb[kTile + kLagThreads];
    const int col = blockIdx.x;
    const int lag0 = blockIdx.y * kLagThreads;
    const int lag = lag0 + threadIdx.x;
    const double *y = centred + (size_t)col * stride;
    double acc = 0.0;
    for (int t0 = 0; t0 + lag0 < n; t0 += kTile) {         // later tiles only pair with zeros for every lag of this CTA
        for (int i = threadIdx.x; i < kTile; i += kLagThreads) sa[i] = t0 + i < n ? y[t0 + i] : 0.0;
        for (int i = threadIdx.x; i < kTile + kLagThreads; i += kLagThreads) sb[i] = t0 + lag0 + i < n ? y[t0 + lag0 + i] : 0.0;
        __syncthreads();
        const double *b = sb + threadIdx.x;
#pragma unroll 8
        for (int i = 0; i < kTile; ++i) acc = fma(sa[i], b[i], acc);
        __syncthreads();
    }
    if (lag < n_lags) acf[(size_t)col * n_lags + lag] = acc;
}

}  // namespace

cudaError_t launch_ensemble_mean(const double *chain, long long kept, int bins, int walkers, int n_l, double *series,
                                 long long stride, cudaStream_t stream) {
    const long long total = kept * bins * n_l;
    if (total <= 0) return cudaSuccess;
    const long long grid = (total + kMeanThreads - 1) / kMeanThreads;
    if (grid > 0x7fffffffLL) return cudaErrorInvalidValue;
    ensemble_mean_kernel<<<(unsigned int)grid, kMeanThreads, 0, stream>>>(chain, kept, bins, walkers, n_l, series, stride);
    return cudaGetLastError();
}

cudaError_t launch_series_acf(const double *series, long long stride, int n, int cols, int n_lags, double *centred, double *acf,
                              cudaStream_t stream) {
    if (cols <= 0 || n <= 0 || n_lags <= 0) return cudaSuccess;
    series_centre_kernel<<<(cols + 3) / 4, 128, 0, stream>>>(series, stride, n, cols, centred);
    cudaError_t err = cudaGetLastError();
    if (err != cudaSuccess) return err;
    const dim3 grid((unsigned int)cols, (unsigned int)((n_lags + kLagThreads - 1) / kLagThreads));
    acf_kernel<<<grid, kLagThreads, 0, stream>>>(centred, stride, n, n_lags, acf);
    return cudaGetLastError();
}

}  // namespace mda
