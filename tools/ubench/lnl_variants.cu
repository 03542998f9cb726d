// lnl_variants.cu -- inner-loop experiment (not product code): the same likelihood
// evaluated with the bin constants (A) staged in shared memory and read with LDS.128,
// (B) passed by value as a kernel parameter and read with LDCU into uniform registers.
// One bin, many walkers, so that only the inner loop matters.
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>

constexpr int L = 9, ROW = L + 1, PAIRS = 30;          // C3/C4: 9 dists, 60 angles
struct Consts { double2 v[PAIRS * ROW]; };

template <int WPT>
__global__ void __launch_bounds__(128) k_smem(const double2 *gc, const double *params, double *out, int walkers) {
    __shared__ double2 s[PAIRS * ROW];
    for (int i = threadIdx.x; i < PAIRS * ROW; i += blockDim.x) s[i] = gc[i];
    __syncthreads();
    double a[WPT][L], chi[WPT];
    const int w0 = (blockIdx.x * blockDim.x + threadIdx.x) * WPT;
#pragma unroll
    for (int k = 0; k < WPT; ++k) { chi[k] = 0;
#pragma unroll
        for (int l = 0; l < L; ++l) a[k][l] = params[(size_t)(w0 + k) * L + l]; }
    const double2 *p = s;
#pragma unroll 2
    for (int j = 0; j < PAIRS; ++j, p += ROW) {
        double r0[WPT], r1[WPT];
        const double2 dv = p[0];
#pragma unroll
        for (int k = 0; k < WPT; ++k) { r0[k] = dv.x; r1[k] = dv.y; }
#pragma unroll
        for (int l = 0; l < L; ++l) {
            const double2 D = p[1 + l];
#pragma unroll
            for (int k = 0; k < WPT; ++k) { r0[k] = fma(-a[k][l], D.x, r0[k]); r1[k] = fma(-a[k][l], D.y, r1[k]); }
        }
#pragma unroll
        for (int k = 0; k < WPT; ++k) { chi[k] = fma(r0[k], r0[k], chi[k]); chi[k] = fma(r1[k], r1[k], chi[k]); }
    }
#pragma unroll
    for (int k = 0; k < WPT; ++k) out[w0 + k] = chi[k] * -0.5;
}

// JB angle pairs in flight per walker; DFMAs ordered so that consecutive ones share a[k][l]
template <int WPT, int JB>
__global__ void __launch_bounds__(128) k_cparam(const __grid_constant__ Consts c, const double *params, double *out, int walkers) {
    double a[WPT][L], chi[WPT];
    const int w0 = (blockIdx.x * blockDim.x + threadIdx.x) * WPT;
#pragma unroll
    for (int k = 0; k < WPT; ++k) { chi[k] = 0;
#pragma unroll
        for (int l = 0; l < L; ++l) a[k][l] = params[(size_t)(w0 + k) * L + l]; }
    for (int j = 0; j < PAIRS; j += JB) {
        double r[WPT][2 * JB];
#pragma unroll
        for (int q = 0; q < JB; ++q) {
            const double2 dv = c.v[(j + q) * ROW];
#pragma unroll
            for (int k = 0; k < WPT; ++k) { r[k][2 * q] = dv.x; r[k][2 * q + 1] = dv.y; }
        }
#pragma unroll
        for (int l = 0; l < L; ++l) {
            double2 D[JB];
#pragma unroll
            for (int q = 0; q < JB; ++q) D[q] = c.v[(j + q) * ROW + 1 + l];
#pragma unroll
            for (int k = 0; k < WPT; ++k) {
#pragma unroll
                for (int q = 0; q < JB; ++q) {
                    r[k][2 * q] = fma(-a[k][l], D[q].x, r[k][2 * q]);
                    r[k][2 * q + 1] = fma(-a[k][l], D[q].y, r[k][2 * q + 1]);
                }
            }
        }
#pragma unroll
        for (int k = 0; k < WPT; ++k)
#pragma unroll
            for (int q = 0; q < 2 * JB; ++q) chi[k] = fma(r[k][q], r[k][q], chi[k]);
    }
#pragma unroll
    for (int k = 0; k < WPT; ++k) out[w0 + k] = chi[k] * -0.5;
}

template <typename F>
float time_it(F launch) {
    cudaEvent_t t0, t1; cudaEventCreate(&t0); cudaEventCreate(&t1);
    launch(); cudaDeviceSynchronize();
    float best = 1e30f;
    for (int r = 0; r < 5; ++r) {
        cudaEventRecord(t0); launch(); cudaEventRecord(t1); cudaEventSynchronize(t1);
        float ms; cudaEventElapsedTime(&ms, t0, t1); if (ms < best) best = ms;
    }
    return best;
}

int main() {
    const int walkers = 148 * 128 * 4 * 16;            // 1.2M walkers
    std::vector<double> hp((size_t)walkers * L);
    srand(1);
    for (auto &x : hp) x = 0.3 * rand() / RAND_MAX;
    Consts hc;
    for (int i = 0; i < PAIRS * ROW; ++i) { hc.v[i].x = 1.0 + 0.01 * (rand() % 100); hc.v[i].y = 1.0 + 0.01 * (rand() % 100); }
    double *dp, *o1, *o2; double2 *dc;
    cudaMalloc(&dp, hp.size() * 8); cudaMalloc(&o1, walkers * 8); cudaMalloc(&o2, walkers * 8); cudaMalloc(&dc, sizeof(hc));
    cudaMemcpy(dp, hp.data(), hp.size() * 8, cudaMemcpyHostToDevice);
    cudaMemcpy(dc, &hc, sizeof(hc), cudaMemcpyHostToDevice);
    const double flop = 2.0 * 60 * (L + 1) * walkers;
    auto report = [&](const char *name, float ms) { printf("%-34s %8.3f ms %7.2f TFLOP/s\n", name, ms, flop / (ms * 1e-3) / 1e12); };
    report("smem LDS.128  WPT=2", time_it([&] { k_smem<2><<<walkers / 256, 128>>>(dc, dp, o1, walkers); }));
    report("smem LDS.128  WPT=4", time_it([&] { k_smem<4><<<walkers / 512, 128>>>(dc, dp, o1, walkers); }));
    report("cparam LDCU   WPT=1 JB=2", time_it([&] { k_cparam<1, 2><<<walkers / 128, 128>>>(hc, dp, o2, walkers); }));
    report("cparam LDCU   WPT=2 JB=1", time_it([&] { k_cparam<2, 1><<<walkers / 256, 128>>>(hc, dp, o2, walkers); }));
    report("cparam LDCU   WPT=2 JB=2", time_it([&] { k_cparam<2, 2><<<walkers / 256, 128>>>(hc, dp, o2, walkers); }));
    report("cparam LDCU   WPT=2 JB=3", time_it([&] { k_cparam<2, 3><<<walkers / 256, 128>>>(hc, dp, o2, walkers); }));
    report("cparam LDCU   WPT=4 JB=1", time_it([&] { k_cparam<4, 1><<<walkers / 512, 128>>>(hc, dp, o2, walkers); }));
    report("cparam LDCU   WPT=4 JB=2", time_it([&] { k_cparam<4, 2><<<walkers / 512, 128>>>(hc, dp, o2, walkers); }));
    report("cparam LDCU   WPT=1 JB=5", time_it([&] { k_cparam<1, 5><<<walkers / 128, 128>>>(hc, dp, o2, walkers); }));
    report("cparam LDCU   WPT=1 JB=3", time_it([&] { k_cparam<1, 3><<<walkers / 128, 128>>>(hc, dp, o2, walkers); }));
    k_smem<2><<<walkers / 256, 128>>>(dc, dp, o1, walkers);
    k_cparam<2, 2><<<walkers / 256, 128>>>(hc, dp, o2, walkers);
    std::vector<double> h1(walkers), h2(walkers);
    cudaMemcpy(h1.data(), o1, walkers * 8, cudaMemcpyDeviceToHost);
    cudaMemcpy(h2.data(), o2, walkers * 8, cudaMemcpyDeviceToHost);
    int bad = 0; for (int i = 0; i < walkers; ++i) bad += (h1[i] != h2[i]);
    printf("bitwise mismatches smem vs cparam: %d of %d (%s)\n", bad, walkers, cudaGetErrorString(cudaGetLastError()));
    return 0;
}
