// chi_loop_ubench.cu -- the math-warp loop of the DMMA stepper in isolation (not product code):
// WT tiles of 8 walkers x NT angle tiles, constants fragment-major in shared memory, L = 9
// (R = 1 leading row by DFMA, Q = 2 k-steps by DMMA).  Reports cycles per item per warp and the
// fraction of the fp64 issue floor (16 cycles per DMMA + 2 per DFMA, per scheduler).
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void dmma884(double &c0, double &c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
constexpr int L = 9, Q = 2, R = 1, TS = 8 * (L + 1), NT = 8;

// V: 0 = as in the stepper (prefetch one tile ahead, unroll 2); 1 = no prefetch, unroll 2; 2 = full unroll, no prefetch;
//    3 = full unroll + ks-outer order; 4 = prefetch, unroll 4
template <int WT, int V>
__global__ void __launch_bounds__(512) k_chi(const double *gc, const double *gq, double *out, long long *cyc, int items) {
    extern __shared__ double sm[];
    double *sC = sm, *sQ = sm + NT * TS;
    for (int i = threadIdx.x; i < NT * TS; i += blockDim.x) sC[i] = gc[i];
    for (int i = threadIdx.x; i < 256 * L; i += blockDim.x) sQ[i] = gq[i];
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g = lane >> 2, tig = lane & 3;
    double acc = 0;
    const long long t0 = clock64();
    for (int it = 0; it < items; ++it) {
        const int item = (warp + it) & (256 / (8 * WT) - 1);
        double A[WT][Q], lead[WT][R];
#pragma unroll
        for (int k = 0; k < WT; ++k) {
            const double *row = sQ + (item * 8 * WT + 8 * k + g) * L;
            lead[k][0] = row[0];
#pragma unroll
            for (int ks = 0; ks < Q; ++ks) A[k][ks] = row[R + 4 * ks + tig];
        }
        double chi0[WT], chi1[WT];
#pragma unroll
        for (int k = 0; k < WT; ++k) { chi0[k] = 0; chi1[k] = 0; }
        auto body = [&](const double2 dv, const double2 Di, const double (&Bf)[Q]) {
            double c0[WT], c1[WT];
#pragma unroll
            for (int k = 0; k < WT; ++k) { c0[k] = fma(lead[k][0], Di.x, dv.x); c1[k] = fma(lead[k][0], Di.y, dv.y); }
            if (V == 3) {
#pragma unroll
                for (int k = 0; k < WT; ++k)
#pragma unroll
                    for (int ks = 0; ks < Q; ++ks) dmma884(c0[k], c1[k], A[k][ks], Bf[ks]);
            } else {
#pragma unroll
                for (int ks = 0; ks < Q; ++ks)
#pragma unroll
                    for (int k = 0; k < WT; ++k) dmma884(c0[k], c1[k], A[k][ks], Bf[ks]);
            }
#pragma unroll
            for (int k = 0; k < WT; ++k) { chi0[k] = fma(c0[k], c0[k], chi0[k]); chi1[k] = fma(c1[k], c1[k], chi1[k]); }
        };
        auto load = [&](int at, double2 &dv, double2 &Di, double (&Bf)[Q]) {
            const double *tp = sC + at * TS;
            dv = *reinterpret_cast<const double2 *>(tp + 2 * tig);
            Di = *reinterpret_cast<const double2 *>(tp + 8 + 2 * tig);
#pragma unroll
            for (int ks = 0; ks < Q; ++ks) Bf[ks] = tp[8 * (1 + R) + 32 * ks + lane];
        };
        if (V == 0 || V == 4) {
            double2 dv, Di; double Bf[Q];
            load(0, dv, Di, Bf);
            if (V == 0) {
#pragma unroll 2
                for (int at = 0; at < NT; ++at) {
                    double2 dvn, Din; double Bfn[Q];
                    load(min(at + 1, NT - 1), dvn, Din, Bfn);
                    body(dv, Di, Bf);
                    dv = dvn; Di = Din;
#pragma unroll
                    for (int ks = 0; ks < Q; ++ks) Bf[ks] = Bfn[ks];
                }
            } else {
#pragma unroll 4
                for (int at = 0; at < NT; ++at) {
                    double2 dvn, Din; double Bfn[Q];
                    load(min(at + 1, NT - 1), dvn, Din, Bfn);
                    body(dv, Di, Bf);
                    dv = dvn; Di = Din;
#pragma unroll
                    for (int ks = 0; ks < Q; ++ks) Bf[ks] = Bfn[ks];
                }
            }
        } else if (V == 1) {
#pragma unroll 2
            for (int at = 0; at < NT; ++at) { double2 dv, Di; double Bf[Q]; load(at, dv, Di, Bf); body(dv, Di, Bf); }
        } else {
#pragma unroll
            for (int at = 0; at < NT; ++at) { double2 dv, Di; double Bf[Q]; load(at, dv, Di, Bf); body(dv, Di, Bf); }
        }
#pragma unroll
        for (int k = 0; k < WT; ++k) acc += chi0[k] + chi1[k];
    }
    const long long t1 = clock64();
    if (lane == 0) cyc[blockIdx.x * (blockDim.x >> 5) + warp] = t1 - t0;
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

template <int WT, int V>
void run(const char *name, int warps, const double *gc, const double *gq, double *out, long long *cyc) {
    const int items = 2048 / WT, smem = (NT * TS + 256 * L) * 8;
    cudaFuncSetAttribute(k_chi<WT, V>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    k_chi<WT, V><<<148, 32 * warps, smem>>>(gc, gq, out, cyc, items);
    k_chi<WT, V><<<148, 32 * warps, smem>>>(gc, gq, out, cyc, items);
    long long h[16];
    cudaMemcpy(h, cyc, sizeof(long long) * warps, cudaMemcpyDeviceToHost);
    const double per_item = (double)h[0] / items;
    const double floor_cycles = WT * NT * (Q * 16.0 + 4 * 2.0) * (warps / 4.0);      // per item, the scheduler shared by warps/4 warps
    printf("%-44s WT=%d warps/SM=%2d  %8.1f cycles per item  issue floor %7.1f  -> %5.1f %%  %s\n", name, WT, warps, per_item, floor_cycles,
           100.0 * floor_cycles / per_item, cudaGetErrorString(cudaDeviceSynchronize()));
}

int main() {
    double *gc, *gq, *out; long long *cyc;
    cudaMalloc(&gc, NT * TS * 8); cudaMalloc(&gq, 256 * L * 8); cudaMalloc(&out, 148 * 512 * 8); cudaMalloc(&cyc, 148 * 16 * 8);
    double hc[NT * TS], hq[256 * L];
    for (int i = 0; i < NT * TS; ++i) hc[i] = 1.0 + 1e-3 * (i % 97);
    for (int i = 0; i < 256 * L; ++i) hq[i] = 0.1 + 1e-4 * (i % 89);
    cudaMemcpy(gc, hc, sizeof(hc), cudaMemcpyHostToDevice); cudaMemcpy(gq, hq, sizeof(hq), cudaMemcpyHostToDevice);
    run<4, 0>("prefetch, unroll 2 (stepper)", 8, gc, gq, out, cyc);
    run<4, 0>("prefetch, unroll 2 (stepper)", 4, gc, gq, out, cyc);
    run<4, 0>("prefetch, unroll 2 (stepper)", 16, gc, gq, out, cyc);
    run<4, 1>("no prefetch, unroll 2", 8, gc, gq, out, cyc);
    run<4, 2>("full unroll", 8, gc, gq, out, cyc);
    run<4, 3>("full unroll, k-step inner", 8, gc, gq, out, cyc);
    run<4, 4>("prefetch, unroll 4", 8, gc, gq, out, cyc);
    run<2, 0>("prefetch, unroll 2", 8, gc, gq, out, cyc);
    run<2, 0>("prefetch, unroll 2", 16, gc, gq, out, cyc);
    run<2, 2>("full unroll", 8, gc, gq, out, cyc);
    run<2, 2>("full unroll", 16, gc, gq, out, cyc);
    run<1, 2>("full unroll", 16, gc, gq, out, cyc);
    return 0;
}
