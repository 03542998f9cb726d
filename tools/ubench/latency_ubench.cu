// latency_ubench.cu -- dependent-issue latencies on sm_100a that the stepper's serial phases are made of
// (not product code).  One warp per SM unless noted; cycles per operation from clock64 around a chain.
#include <cstdio>
#include <cuda_runtime.h>

template <int OP>
__global__ void k_chain(double *out, long long *cyc, const double *in, int n) {
    __shared__ double sm[1024];
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) sm[i] = in[i & 63] + i;
    __syncthreads();
    double x = in[0], y = in[1], z = in[2];
    int idx = threadIdx.x & 31;
    unsigned u = threadIdx.x;
    bool p = false;
    const long long t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < n; ++i) {
        if (OP == 0) x = fma(x, y, z);                       // DFMA
        if (OP == 1) x = __dadd_rn(x, y);                    // DADD
        if (OP == 2) x = __dmul_rn(x, y);                    // DMUL
        if (OP == 3) { p = p || (x < y + i); }               // DSETP chained through the predicate (+ a DADD off the chain)
        if (OP == 4) { idx = (int)sm[idx & 1023] & 31; }     // LDS.64 -> F2I -> address
        if (OP == 5) { x = __shfl_xor_sync(0xffffffffu, x, 1); }   // 2 x SHFL
        if (OP == 6) { u = u * 0xD2511F53u + 12345u; }       // IMAD
        if (OP == 7) { u = __umulhi(u, 0xD2511F53u) ^ 0x9E3779B9u; }   // IMAD.HI + LOP3
        if (OP == 8) { x = (double)(float)x; }               // F2F down + up
        if (OP == 9) { x = (double)(int)(u + i) ; u += (unsigned)x; }  // I2F.F64 + F2I
    }
    const long long t1 = clock64();
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
    out[blockIdx.x * blockDim.x + threadIdx.x] = x + idx + u + (p ? 1 : 0);
}

__global__ void k_bar(long long *cyc, int n) {
    const long long t0 = clock64();
    for (int i = 0; i < n; ++i) __syncthreads();
    const long long t1 = clock64();
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

// ping-pong between two warps through two mbarriers: one hop = arrive -> the other warp's try_wait succeeds
__global__ void k_hop(long long *cyc, int n) {
    __shared__ __align__(8) unsigned long long bar[2];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int b = 0; b < 2; ++b) {
            unsigned a = (unsigned)__cvta_generic_to_shared(&bar[b]);
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(a));
        }
    }
    __syncthreads();
    const unsigned mine = (unsigned)__cvta_generic_to_shared(&bar[warp]), other = (unsigned)__cvta_generic_to_shared(&bar[warp ^ 1]);
    const long long t0 = clock64();
    for (int i = 0; i < n; ++i) {
        if (warp == 0) {
            if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(other) : "memory");
        }
        unsigned ok = 0;
        while (!ok) asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }" : "=r"(ok) : "r"(mine), "r"(i & 1) : "memory");
        if (warp == 1) {
            __syncwarp();
            if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(other) : "memory");
        }
    }
    const long long t1 = clock64();
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

int main() {
    double h[64];
    for (int i = 0; i < 64; ++i) h[i] = 1.0 + 1e-9 * i;
    double *in, *out;
    long long *cyc, hc;
    cudaMalloc(&in, sizeof(h)); cudaMalloc(&out, 8 * 1024 * 8); cudaMalloc(&cyc, 64);
    cudaMemcpy(in, h, sizeof(h), cudaMemcpyHostToDevice);
    const int n = 4096;
    const char *names[] = {"DFMA", "DADD", "DMUL", "DSETP.OR chain", "LDS.64 + F2I + addr", "SHFL x2 (64-bit)", "IMAD", "IMAD.HI + LOP3", "F2F.F32.F64 + F2F.F64.F32", "I2F.F64 + F2I + IADD"};
#define RUN(OP) { k_chain<OP><<<1, 32>>>(out, cyc, in, n); k_chain<OP><<<1, 32>>>(out, cyc, in, n); cudaMemcpy(&hc, cyc, 8, cudaMemcpyDeviceToHost); \
                  printf("%-28s %6.1f cycles per dependent op (1 warp)\n", names[OP], (double)hc / n); }
    RUN(0) RUN(1) RUN(2) RUN(3) RUN(4) RUN(5) RUN(6) RUN(7) RUN(8) RUN(9)
    for (int threads = 64; threads <= 1024; threads *= 2) {
        k_bar<<<1, threads>>>(cyc, n); cudaMemcpy(&hc, cyc, 8, cudaMemcpyDeviceToHost);
        printf("__syncthreads, %4d threads  %6.1f cycles\n", threads, (double)hc / n);
    }
    k_hop<<<1, 64>>>(cyc, n); cudaMemcpy(&hc, cyc, 8, cudaMemcpyDeviceToHost);
    printf("mbarrier round trip (2 hops) %6.1f cycles  %s\n", (double)hc / n, cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
