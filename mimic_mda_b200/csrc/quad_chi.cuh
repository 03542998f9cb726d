// quad_chi.cuh -- chi^2 from a bin's triangular factor, shared by ensemble_quad_kernel.cu and lnpost_quad_kernel.cu.
//
// chi^2(a) = sum_j (d_j - sum_l a_l D_lj)^2 (C/ChiSquare.c:188-210) = |M [a; -1]|^2 with M = [D^T | d], and with
// M = Q T (T upper triangular, chisq_abi.cu: factor_bin) = |T [a; -1]|^2.  T is packed by rows, T[i][i..L].
#pragma once
#include "mda_kernels.cuh"

namespace mda {

// In shared memory the rows are padded to an even number of doubles, so that a row is read with 16-byte loads
// (half the shared-memory instructions of the packed layout; every lane reads the same address: broadcasts).
__host__ __device__ constexpr int quad_row_offset(int n_l, int i) {       // packed: start of row i
    return i * (n_l + 1) - i * (i - 1) / 2;
}
__host__ __device__ constexpr int quad_padded_offset(int n_l, int i) {    // padded: start of row i
    int o = 0;
    for (int r = 0; r < i; ++r) o += ((n_l + 1 - r) + 1) & ~1;
    return o;
}
__host__ __device__ constexpr int quad_padded_doubles(int n_l) { return quad_padded_offset(n_l, n_l + 1); }

// Cooperative copy of a bin's packed factor (global) into the padded shared-memory layout.
template <int L>
__device__ __forceinline__ void load_factor(double *__restrict__ s_tri, const double *__restrict__ packed, int tid, int nthr) {
    for (int k = tid; k < quad_padded_doubles(L); k += nthr) s_tri[k] = 0.0;
    __syncthreads();
#pragma unroll
    for (int i = 0; i <= L; ++i)
        for (int k = tid; k < L + 1 - i; k += nthr) s_tri[quad_padded_offset(L, i) + k] = packed[quad_row_offset(L, i) + k];
}

template <int L>
__device__ __forceinline__ double chi_from_factor(const double *__restrict__ tri, const double (&a)[L]) {
    double chi = 0.0;
#pragma unroll
    for (int i = 0; i <= L; ++i) {
        constexpr int dummy = 0;
        (void)dummy;
        const int len = L + 1 - i;                                            // T[i][i..L]; the last entry is the transformed data
        const double2 *row = reinterpret_cast<const double2 *>(tri + quad_padded_offset(L, i));
        double u = 0.0;
        double data = 0.0;
#pragma unroll
        for (int k2 = 0; k2 < (len + 1) / 2; ++k2) {
            const double2 v = row[k2];
            const int k = 2 * k2;
            if (k < len - 1) u = fma(v.x, a[i + k], u); else if (k == len - 1) data = v.x;
            if (k + 1 < len - 1) u = fma(v.y, a[i + k + 1], u); else if (k + 1 == len - 1) data = v.y;
        }
        u -= data;
        chi = fma(u, u, chi);
    }
    return chi;
}

// The same with T held in registers (indices are compile-time constants after unrolling): for the build that has an SM's
// register file to itself.  t[] in the padded layout of load_factor.  Rows are summed G at a time, column by column, so
// that G independent multiply-add chains are in flight (a lone walker warp pair per scheduler cannot hide the latency of
// one 55-long dependent chain); every row's sum still runs over ascending k and chi^2 over ascending i, so the result is
// bit for bit the one of chi_from_factor.
template <int L, int G = 5>
__device__ __forceinline__ double chi_from_factor_regs(const double (&t)[quad_padded_doubles(L)], const double (&a)[L]) {
    double chi = 0.0;
#pragma unroll
    for (int i0 = 0; i0 <= L; i0 += G) {
        double u[G];
#pragma unroll
        for (int g = 0; g < G; ++g) u[g] = 0.0;
#pragma unroll
        for (int k = i0; k < L; ++k)
#pragma unroll
            for (int g = 0; g < G; ++g) {
                const int i = i0 + g;
                if (i <= k) u[g] = fma(t[quad_padded_offset(L, i < L ? i : L) + (k - i)], a[k], u[g]);
            }
#pragma unroll
        for (int g = 0; g < G; ++g) {
            const int i = i0 + g;
            if (i <= L) {
                u[g] -= t[quad_padded_offset(L, i) + (L - i)];
                chi = fma(u[g], u[g], chi);
            }
        }
    }
    return chi;
}

// Rows 0 .. SPLIT-1 of T from registers (t[] in the padded layout, only those rows filled), rows SPLIT .. L from shared
// memory (padded layout, 16-byte broadcast reads): the long rows where registers pay, the short ones where a handful of
// broadcast loads is cheaper than the registers they would occupy.  Same per-row and chi^2 order: the same bits.
template <int L, int SPLIT, int G = 5>
__device__ __forceinline__ double chi_from_factor_split(const double (&t)[quad_padded_offset(L, SPLIT) > 0 ? quad_padded_offset(L, SPLIT) : 1],
                                                        const double *__restrict__ tri, const double (&a)[L]) {
    double chi = 0.0;
#pragma unroll
    for (int i0 = 0; i0 < SPLIT; i0 += G) {
        double u[G];
#pragma unroll
        for (int g = 0; g < G; ++g) u[g] = 0.0;
#pragma unroll
        for (int k = i0; k < L; ++k)
#pragma unroll
            for (int g = 0; g < G; ++g) {
                const int i = i0 + g;
                if (i <= k && i < SPLIT) u[g] = fma(t[quad_padded_offset(L, i < SPLIT ? i : 0) + (k - i)], a[k], u[g]);
            }
#pragma unroll
        for (int g = 0; g < G; ++g) {
            const int i = i0 + g;
            if (i < SPLIT) {
                u[g] -= t[quad_padded_offset(L, i) + (L - i)];
                chi = fma(u[g], u[g], chi);
            }
        }
    }
#pragma unroll
    for (int i = SPLIT; i <= L; ++i) {
        const int len = L + 1 - i;
        const double2 *row = reinterpret_cast<const double2 *>(tri + quad_padded_offset(L, i));
        double u = 0.0, data = 0.0;
#pragma unroll
        for (int k2 = 0; k2 < (len + 1) / 2; ++k2) {
            const double2 v = row[k2];
            const int k = 2 * k2;
            if (k < len - 1) u = fma(v.x, a[i + k], u); else if (k == len - 1) data = v.x;
            if (k + 1 < len - 1) u = fma(v.y, a[i + k + 1], u); else if (k + 1 == len - 1) data = v.y;
        }
        u -= data;
        chi = fma(u, u, chi);
    }
    return chi;
}

}  // namespace mda
