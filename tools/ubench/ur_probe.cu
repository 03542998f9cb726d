// does ptxas move loop-variant warp-uniform shared-memory values into uniform registers?
#include <cuda_runtime.h>
template <int WPT, int MODE>
__global__ void __launch_bounds__(128) probe(const double *g, double *out, int n) {
    extern __shared__ double2 s[];
    for (int i = threadIdx.x; i < n * 10; i += blockDim.x) s[i] = reinterpret_cast<const double2 *>(g)[i];
    __syncthreads();
    double a[WPT][9], chi[WPT];
#pragma unroll
    for (int k = 0; k < WPT; ++k) { chi[k] = 0; 
#pragma unroll
        for (int l = 0; l < 9; ++l) a[k][l] = g[(threadIdx.x * WPT + k) * 9 + l]; }
    const double2 *p = s;
    for (int j = 0; j < n; ++j, p += 10) {
        double r0[WPT], r1[WPT];
        double2 dv = p[0];
        if (MODE == 1) { dv.x = __shfl_sync(0xffffffffu, dv.x, 0); dv.y = __shfl_sync(0xffffffffu, dv.y, 0); }
#pragma unroll
        for (int k = 0; k < WPT; ++k) { r0[k] = dv.x; r1[k] = dv.y; }
#pragma unroll
        for (int l = 0; l < 9; ++l) {
            double2 D = p[1 + l];
            if (MODE == 1) { D.x = __shfl_sync(0xffffffffu, D.x, 0); D.y = __shfl_sync(0xffffffffu, D.y, 0); }
#pragma unroll
            for (int k = 0; k < WPT; ++k) { r0[k] = fma(-a[k][l], D.x, r0[k]); r1[k] = fma(-a[k][l], D.y, r1[k]); }
        }
#pragma unroll
        for (int k = 0; k < WPT; ++k) { chi[k] = fma(r0[k], r0[k], chi[k]); chi[k] = fma(r1[k], r1[k], chi[k]); }
    }
#pragma unroll
    for (int k = 0; k < WPT; ++k) out[threadIdx.x * WPT + k] = chi[k];
}
template __global__ void probe<4, 0>(const double *, double *, int);
template __global__ void probe<8, 0>(const double *, double *, int);
template __global__ void probe<4, 1>(const double *, double *, int);
