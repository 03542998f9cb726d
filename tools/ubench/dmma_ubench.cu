// dmma_ubench.cu -- is the fp64 tensor path (mma.sync ... f64) a way past the DFMA
// register-operand limit on sm_100a?  (not product code)
//   t0: DMMA m8n8k4 issue rate, 8 independent accumulators per warp
//   t1: m16n8k4 / m16n8k8 / m16n8k16 (sm_90+ shapes)
//   t2: DMMA + DFMA interleaved: do the two share a pipe?
//   t3: dependent DMMA chain (latency)
//   t4: the C4 likelihood (L=9, N=60) with the constants as B fragments in registers
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include <cuda_runtime.h>

__device__ __forceinline__ void dmma884(double &d0, double &d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}
__device__ __forceinline__ void dmma1684(double (&c)[4], double a0, double a1, double b) {
    asm volatile("mma.sync.aligned.m16n8k4.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};"
                 : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3]) : "d"(a0), "d"(a1), "d"(b));
}
__device__ __forceinline__ void dmma1688(double (&c)[4], const double (&a)[4], double b0, double b1) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3])
                 : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(b0), "d"(b1));
}
__device__ __forceinline__ void dmma16816(double (&c)[4], const double (&a)[8], const double (&b)[4]) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, {%0,%1,%2,%3};"
                 : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3])
                 : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(a[4]), "d"(a[5]), "d"(a[6]), "d"(a[7]),
                   "d"(b[0]), "d"(b[1]), "d"(b[2]), "d"(b[3]));
}

// V: 0 m8n8k4, 1 m16n8k4, 2 m16n8k8, 3 m16n8k16, 4 m8n8k4 dependent chain,
//    10+x: per 8 m8n8k4 also 4*x independent DFMAs
template <int V>
__global__ void __launch_bounds__(256) k_rate(double *sink, int outer, const double *in) {
    double c[8][4], a[8], b[4], f[16];
#pragma unroll
    for (int i = 0; i < 8; ++i) { a[i] = in[i] + 1e-9 * threadIdx.x;
#pragma unroll
        for (int j = 0; j < 4; ++j) c[i][j] = in[8 + i] * j; }
#pragma unroll
    for (int i = 0; i < 4; ++i) b[i] = in[16 + i];
#pragma unroll
    for (int i = 0; i < 16; ++i) f[i] = in[20 + i];
    const double x = in[40], y = in[41];
    for (int o = 0; o < outer; ++o) {
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            if (V == 0 || V >= 10) {
#pragma unroll
                for (int i = 0; i < 8; ++i) dmma884(c[i][0], c[i][1], a[i], b[r]);
                if (V >= 10) {
#pragma unroll
                    for (int i = 0; i < 4 * (V - 10); ++i) f[i % 16] = fma(f[i % 16], x, y);
                }
            } else if (V == 1) {
#pragma unroll
                for (int i = 0; i < 8; ++i) dmma1684(c[i], a[i], a[(i + 1) & 7], b[r]);
            } else if (V == 2) {
#pragma unroll
                for (int i = 0; i < 8; ++i) { const double aa[4] = {a[i], a[(i + 1) & 7], a[(i + 2) & 7], a[(i + 3) & 7]}; dmma1688(c[i], aa, b[r], b[(r + 1) & 3]); }
            } else if (V == 3) {
#pragma unroll
                for (int i = 0; i < 8; ++i) dmma16816(c[i], a, b);
            } else if (V == 4) {
                dmma884(c[0][0], c[0][1], a[0], b[r]);
            }
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) s += c[i][j];
#pragma unroll
    for (int i = 0; i < 16; ++i) s += f[i];
    if (s == 0.12345) sink[blockIdx.x * blockDim.x + threadIdx.x] = s;
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

// 8 independent DMMAs + NI dependent-free integer multiply-adds per iteration: does a DMMA hold the issue port?
template <int NI>
__global__ void __launch_bounds__(256) k_mix_int(double *sink, int outer, const double *in) {
    double c[8][2], a[8], b;
    unsigned u[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { a[i] = in[i] + 1e-9 * threadIdx.x; c[i][0] = in[8 + i]; c[i][1] = in[9 + i]; u[i] = threadIdx.x * 7 + i; }
    b = in[16];
    for (int o = 0; o < outer; ++o) {
#pragma unroll
        for (int r = 0; r < 4; ++r) {
#pragma unroll
            for (int i = 0; i < 8; ++i) dmma884(c[i][0], c[i][1], a[i], b);
#pragma unroll
            for (int i = 0; i < NI; ++i) u[i % 8] = __umulhi(u[i % 8], 0xD2511F53u) ^ u[(i + 3) % 8];
        }
    }
    double acc = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) acc += c[i][0] + c[i][1] + u[i];
    sink[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

template <int V>
void rate(const char *name, double *sink, const double *in, int sms, double fma_per_warp_iter, double dfma_per_thread_iter = 0) {
    const int threads = 256, blocks = sms * 8, outer = 256;
    const float ms = time_it([&] { k_rate<V><<<blocks, threads>>>(sink, outer, in); });
    const double warps = (double)blocks * threads / 32;
    const double mma_fma = fma_per_warp_iter * 4.0 * outer * warps;
    const double vec_fma = dfma_per_thread_iter * 4.0 * outer * warps * 32;
    printf("%-34s %8.3f ms  mma %7.2f TFLOP/s  dfma %7.2f TFLOP/s  sum %7.2f\n", name, ms, 2 * mma_fma / (ms * 1e-3) / 1e12,
           2 * vec_fma / (ms * 1e-3) / 1e12, 2 * (mma_fma + vec_fma) / (ms * 1e-3) / 1e12);
}

// ---- t4: likelihood, L = 9, N = 60 (8 angle tiles of 8, 4 zero-padded angles) ------------------
constexpr int L = 9, NT = 8, NPAD = 64;
// consts: d[NPAD], D[L][NPAD] dense rows
// MODE 0: K = 8 by DMMA (l = 1..8), C init = d - a0 D0 by DFMA
// MODE 1: K = 12 by DMMA (1, -a0..-a8, 0, 0), C init = 0
template <int MODE, int WT>      // WT walker tiles (of 8) in flight per warp
__global__ void __launch_bounds__(256) k_lnl(const double *gc, const double *params, double *out, int walkers, int reps) {
    const int lane = threadIdx.x & 31, g = lane >> 2, tig = lane & 3;
    constexpr int KS = MODE == 0 ? 2 : 3;
    double B[NT][KS];
    __shared__ double2 sInit[2][NPAD / 2];            // d pairs, D0 pairs
#pragma unroll
    for (int t = 0; t < NT; ++t) {
#pragma unroll
        for (int ks = 0; ks < KS; ++ks) {
            const int col = t * 8 + g;
            const int k = ks * 4 + tig;                // row of the K dimension
            double v;
            if (MODE == 0) v = gc[(size_t)(1 + 1 + k) * NPAD + col];                   // D_{1+k}
            else v = k < 10 ? gc[(size_t)k * NPAD + col] : 0.0;                        // d, D_0..D_8, 0, 0
            B[t][ks] = v;
        }
    }
    for (int i = threadIdx.x; i < NPAD / 2; i += blockDim.x) {
        sInit[0][i] = make_double2(gc[2 * i], gc[2 * i + 1]);
        sInit[1][i] = make_double2(gc[NPAD + 2 * i], gc[NPAD + 2 * i + 1]);
    }
    __syncthreads();
    const int warp_global = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int n_warps = (gridDim.x * blockDim.x) >> 5;
    for (int rep = 0; rep < reps; ++rep)
    for (int w0 = warp_global * 8 * WT; w0 < walkers; w0 += n_warps * 8 * WT) {
        double A[WT][KS], a0[WT], chi[WT];
#pragma unroll
        for (int q = 0; q < WT; ++q) {
            const double *p = params + (size_t)(w0 + q * 8 + g) * L;
            chi[q] = 0.0;
            if (MODE == 0) {
                a0[q] = p[0];
                A[q][0] = -p[1 + tig];
                A[q][1] = -p[5 + tig];
            } else {
                a0[q] = 0;
                A[q][0] = tig == 0 ? 1.0 : -p[tig - 1];
                A[q][1] = -p[3 + tig];
                A[q][2] = tig < 2 ? -p[7 + tig] : 0.0;
            }
        }
#pragma unroll
        for (int t = 0; t < NT; ++t) {
            double c0[WT], c1[WT];
            if (MODE == 0) {
                const double2 dv = sInit[0][t * 4 + tig], D0 = sInit[1][t * 4 + tig];
#pragma unroll
                for (int q = 0; q < WT; ++q) { c0[q] = fma(-a0[q], D0.x, dv.x); c1[q] = fma(-a0[q], D0.y, dv.y); }
            } else {
#pragma unroll
                for (int q = 0; q < WT; ++q) { c0[q] = 0.0; c1[q] = 0.0; }
            }
#pragma unroll
            for (int ks = 0; ks < KS; ++ks)
#pragma unroll
                for (int q = 0; q < WT; ++q) dmma884(c0[q], c1[q], A[q][ks], B[t][ks]);
#pragma unroll
            for (int q = 0; q < WT; ++q) { chi[q] = fma(c0[q], c0[q], chi[q]); chi[q] = fma(c1[q], c1[q], chi[q]); }
        }
#pragma unroll
        for (int q = 0; q < WT; ++q) {
            double v = chi[q];
            v += __shfl_xor_sync(0xffffffffu, v, 1);
            v += __shfl_xor_sync(0xffffffffu, v, 2);
            if (tig == 0) out[w0 + q * 8 + g] = v * -0.5;
        }
    }
}

int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    const int sms = p.multiProcessorCount;
    double h[64]; for (int i = 0; i < 64; ++i) h[i] = 1.0 + 1e-9 * i;
    double *in, *sink; cudaMalloc(&in, sizeof(h)); cudaMalloc(&sink, 8 * 256 * sms * 8);
    cudaMemcpy(in, h, sizeof(h), cudaMemcpyHostToDevice);
    printf("%s, %d SMs\n", p.name, sms);
    rate<0>("m8n8k4 x8 independent", sink, in, sms, 8 * 256.0);
    rate<1>("m16n8k4 x8", sink, in, sms, 8 * 512.0);
    rate<2>("m16n8k8 x8", sink, in, sms, 8 * 1024.0);
    rate<3>("m16n8k16 x8", sink, in, sms, 8 * 2048.0);
    rate<11>("m8n8k4 x8 + 4 DFMA", sink, in, sms, 8 * 256.0, 4);
    rate<12>("m8n8k4 x8 + 8 DFMA", sink, in, sms, 8 * 256.0, 8);
    rate<14>("m8n8k4 x8 + 16 DFMA", sink, in, sms, 8 * 256.0, 16);
    rate<18>("m8n8k4 x8 + 32 DFMA", sink, in, sms, 8 * 256.0, 32);
    rate<26>("m8n8k4 x8 + 64 DFMA", sink, in, sms, 8 * 256.0, 64);
    {
        k_rate<4><<<sms, 128>>>(sink, 64, in);
        const int outer = 4096;
        const float ms = time_it([&] { k_rate<4><<<sms, 128>>>(sink, outer, in); });
        printf("dependent m8n8k4 chain: %.2f ns per DMMA (x SM clock = latency in cycles)\n", ms * 1e6 / (4.0 * outer));
    }

    {
        const int threads = 256, blocks = sms * 8, outer = 256;
        auto mix = [&](const char *name, float ms, int ni) {
            const double warps = (double)blocks * threads / 32;
            printf("%-34s %8.3f ms  mma %7.2f TFLOP/s  imad %6.2f Tinstr/s (thread level)\n", name, ms, 2 * 8 * 256.0 * 4.0 * outer * warps / (ms * 1e-3) / 1e12,
                   ni * 4.0 * outer * warps * 32 / (ms * 1e-3) / 1e12);
        };
        mix("m8n8k4 x8 + 0 IMAD", time_it([&] { k_mix_int<0><<<blocks, threads>>>(sink, outer, in); }), 0);
        mix("m8n8k4 x8 + 32 IMAD", time_it([&] { k_mix_int<32><<<blocks, threads>>>(sink, outer, in); }), 32);
        mix("m8n8k4 x8 + 64 IMAD", time_it([&] { k_mix_int<64><<<blocks, threads>>>(sink, outer, in); }), 64);
        mix("m8n8k4 x8 + 128 IMAD", time_it([&] { k_mix_int<128><<<blocks, threads>>>(sink, outer, in); }), 128);
        mix("m8n8k4 x8 + 256 IMAD", time_it([&] { k_mix_int<256><<<blocks, threads>>>(sink, outer, in); }), 256);
    }
    // ---- likelihood
    const int walkers = 148 * 256 * 32;               // 1.2M walkers
    std::vector<double> hp((size_t)walkers * L), hc((size_t)(L + 1) * NPAD, 0.0);
    srand(1);
    for (auto &x : hp) x = 0.3 * rand() / RAND_MAX;
    for (int r = 0; r < L + 1; ++r) for (int j = 0; j < 60; ++j) hc[(size_t)r * NPAD + j] = (r == 0 ? 20.0 : 1.0) + 0.01 * (rand() % 100);
    double *dp, *o, *dc;
    cudaMalloc(&dp, hp.size() * 8); cudaMalloc(&o, (size_t)walkers * 8); cudaMalloc(&dc, hc.size() * 8);
    cudaMemcpy(dp, hp.data(), hp.size() * 8, cudaMemcpyHostToDevice);
    cudaMemcpy(dc, hc.data(), hc.size() * 8, cudaMemcpyHostToDevice);
    const int reps = 16;
    const double flop = 2.0 * 60 * (L + 1) * walkers * reps;
    std::vector<double> ho(walkers);
    auto check = [&](const char *name, float ms) {
        cudaMemcpy(ho.data(), o, (size_t)walkers * 8, cudaMemcpyDeviceToHost);
        double worst = 0;
        for (int w = 0; w < walkers; w += 997) {
            double chi = 0;
            for (int j = 0; j < 60; ++j) {
                double r = hc[j];
                for (int l = 0; l < L; ++l) r = fma(-hp[(size_t)w * L + l], hc[(size_t)(1 + l) * NPAD + j], r);
                chi = fma(r, r, chi);
            }
            const double want = chi * -0.5;
            worst = fmax(worst, fabs(ho[w] - want) / fabs(want));
        }
        printf("%-34s %8.3f ms %7.2f TFLOP/s (algorithmic)  max rel err %.2e  %s\n", name, ms, flop / (ms * 1e-3) / 1e12, worst,
               cudaGetErrorString(cudaGetLastError()));
    };
    const int blocks = sms * 2;
    check("lnl K=8 DMMA + init DFMA, WT=1", time_it([&] { k_lnl<0, 1><<<blocks, 256>>>(dc, dp, o, walkers, reps); }));
    check("lnl K=8 DMMA + init DFMA, WT=2", time_it([&] { k_lnl<0, 2><<<blocks, 256>>>(dc, dp, o, walkers, reps); }));
    check("lnl K=8 DMMA + init DFMA, WT=4", time_it([&] { k_lnl<0, 4><<<blocks, 256>>>(dc, dp, o, walkers, reps); }));
    check("lnl K=12 DMMA, WT=1", time_it([&] { k_lnl<1, 1><<<blocks, 256>>>(dc, dp, o, walkers, reps); }));
    check("lnl K=12 DMMA, WT=2", time_it([&] { k_lnl<1, 2><<<blocks, 256>>>(dc, dp, o, walkers, reps); }));
    check("lnl K=12 DMMA, WT=4", time_it([&] { k_lnl<1, 4><<<blocks, 256>>>(dc, dp, o, walkers, reps); }));
    return 0;
}
