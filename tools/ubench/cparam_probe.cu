// bin constants passed BY VALUE as a kernel parameter (constant bank 0): does ptxas feed DFMA
// from uniform registers via LDCU with a uniform dynamic index?
#include <cuda_runtime.h>
struct Consts { double2 v[30 * 10]; };   // C4: 30 angle pairs x (1 + 9) rows = 4800 bytes
template <int WPT>
__global__ void __launch_bounds__(128) probe(const __grid_constant__ Consts c, const double *g, double *out, int pairs) {
    double a[WPT][9], chi[WPT];
#pragma unroll
    for (int k = 0; k < WPT; ++k) { chi[k] = 0;
#pragma unroll
        for (int l = 0; l < 9; ++l) a[k][l] = g[(threadIdx.x * WPT + k) * 9 + l]; }
    for (int j = 0; j < pairs; ++j) {
        double r0[WPT], r1[WPT];
        const double2 dv = c.v[j * 10];
#pragma unroll
        for (int k = 0; k < WPT; ++k) { r0[k] = dv.x; r1[k] = dv.y; }
#pragma unroll
        for (int l = 0; l < 9; ++l) {
            const double2 D = c.v[j * 10 + 1 + l];
#pragma unroll
            for (int k = 0; k < WPT; ++k) { r0[k] = fma(-a[k][l], D.x, r0[k]); r1[k] = fma(-a[k][l], D.y, r1[k]); }
        }
#pragma unroll
        for (int k = 0; k < WPT; ++k) { chi[k] = fma(r0[k], r0[k], chi[k]); chi[k] = fma(r1[k], r1[k], chi[k]); }
    }
#pragma unroll
    for (int k = 0; k < WPT; ++k) out[threadIdx.x * WPT + k] = chi[k];
}
template __global__ void probe<2>(const __grid_constant__ Consts, const double *, double *, int);
template __global__ void probe<4>(const __grid_constant__ Consts, const double *, double *, int);
