// dfma_ubench.cu -- what limits DFMA issue on sm_100a?  (not product code)
//   v0: acc = fma(acc, x, y)            x,y loop-invariant         (the "peak" pattern)
//   v1: acc[c] = fma(a[c], b[c], acc[c]) three distinct 64-bit register operands per DFMA
//   v2: acc[c] = fma(a[c], B, acc[c])    B shared by consecutive DFMAs (operand-reuse friendly)
//   v3: snake: consecutive DFMAs alternate sharing slot A or slot B
//   v4: dependent chain (latency)
#include <cstdio>
#include <cuda_runtime.h>

constexpr int C = 16;

template <int V>
__global__ void __launch_bounds__(256) k(double *sink, int outer, const double *in) {
    double acc[C], a[C], b[C];
#pragma unroll
    for (int c = 0; c < C; ++c) { acc[c] = in[c] + threadIdx.x; a[c] = in[16 + c]; b[c] = in[32 + c]; }
    double x = in[48], y = in[49];
    for (int o = 0; o < outer; ++o) {
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            if (V == 0) {
#pragma unroll
                for (int c = 0; c < C; ++c) acc[c] = fma(acc[c], x, y);
            } else if (V == 1) {
#pragma unroll
                for (int c = 0; c < C; ++c) acc[c] = fma(a[c], b[(c + i) % C], acc[c]);
            } else if (V == 2) {
#pragma unroll
                for (int c = 0; c < C; ++c) acc[c] = fma(a[c], b[i % C], acc[c]);
            } else if (V == 3) {
                // 4 walkers (a0..a3) x 4 D values: snake through the 4x4 grid
#pragma unroll
                for (int d = 0; d < 4; ++d) {
#pragma unroll
                    for (int w = 0; w < 4; ++w) {
                        const int ww = (d & 1) ? 3 - w : w;
                        acc[d * 4 + ww] = fma(a[ww + 4 * (i & 3)], b[d + 4 * (i & 3)], acc[d * 4 + ww]);
                    }
                }
            } else if (V == 4) {
                acc[0] = fma(acc[0], x, y);
            }
        }
    }
    double s = 0;
#pragma unroll
    for (int c = 0; c < C; ++c) s += acc[c];
    if (s == 0.12345) sink[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int V>
void run(const char *name, double *sink, const double *in, int sms) {
    const int threads = 256, blocks = sms * 8, outer = 128;
    cudaEvent_t t0, t1;
    cudaEventCreate(&t0); cudaEventCreate(&t1);
    k<V><<<blocks, threads>>>(sink, 4, in);
    float best = 1e30f;
    for (int r = 0; r < 5; ++r) {
        cudaEventRecord(t0);
        k<V><<<blocks, threads>>>(sink, outer, in);
        cudaEventRecord(t1); cudaEventSynchronize(t1);
        float ms; cudaEventElapsedTime(&ms, t0, t1);
        if (ms < best) best = ms;
    }
    const double n = (V == 4 ? 1.0 : (double)C) * 16.0 * outer * (double)threads * blocks;
    printf("%-28s %8.3f ms  %7.2f TFLOP/s\n", name, best, 2.0 * n / (best * 1e-3) / 1e12);
    if (V == 4) {
        // one warp per SMSP so the chain latency is exposed
        k<4><<<sms, 128>>>(sink, outer, in);
        cudaEventRecord(t0);
        k<4><<<sms, 128>>>(sink, outer * 16, in);
        cudaEventRecord(t1); cudaEventSynchronize(t1);
        float ms; cudaEventElapsedTime(&ms, t0, t1);
        printf("   dependent chain: %.2f ns per DFMA (x SM clock = latency in cycles)\n", ms * 1e6 / (16.0 * outer * 16));
    }
}

int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    double h[64]; for (int i = 0; i < 64; ++i) h[i] = 1.0 + 1e-9 * i;
    double *in, *sink; cudaMalloc(&in, sizeof(h)); cudaMalloc(&sink, 8 * 256 * p.multiProcessorCount * 8);
    cudaMemcpy(in, h, sizeof(h), cudaMemcpyHostToDevice);
    printf("%s, %d SMs\n", p.name, p.multiProcessorCount);
    run<0>("v0 acc=fma(acc,x,y)", sink, in, p.multiProcessorCount);
    run<1>("v1 three distinct operands", sink, in, p.multiProcessorCount);
    run<2>("v2 shared B operand", sink, in, p.multiProcessorCount);
    run<3>("v3 snake A/B sharing", sink, in, p.multiProcessorCount);
    run<4>("v4 latency", sink, in, p.multiProcessorCount);
    return 0;
}
