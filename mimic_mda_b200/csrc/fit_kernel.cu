// fit_kernel.cu -- batched initial fits.
//
// The reference finds walker generators by running scipy's L-BFGS-B with finite-difference
// gradients on calculateChi from every point of a Cartesian start grid, several rounds per
// start (mda.py:1114, 1664-1709, 1912-1957; 6000 starts x 4 rounds in mda_config0.py), i.e.
// millions of scalar chi^2 calls per Ex bin.  chi^2(a) = |d - D^T a|^2 is a convex
// quadratic and the constraints are a box, so every start has the same constrained
// minimiser; here each (bin, start) pair is solved exactly by one thread with a
// bounded-variable least-squares active-set method (Stark & Parker's BVLS, warm-started at
// the clipped start point) on the bin's normal equations:
//   gram_kernel   G = D D^T, b = D d per bin, accumulated in ascending angle order
//   bvls_kernel   one thread per (bin, start): free-set Cholesky solves + feasible steps
// The chi^2 of every solution is then evaluated by the batched likelihood kernel in kChi
// mode (not through G: the Gram form cancels catastrophically).
#include "mda_kernels.cuh"

namespace mda {

namespace {

// G[bin][L][L], b[bin][L] from the store layout consts[bin][n_pad/2][1+L] double2.
__global__ void __launch_bounds__(256) gram_kernel(const double *__restrict__ consts, const int *__restrict__ n_pts,
                                                   int n_l, int n_pad, double *__restrict__ gram, double *__restrict__ rhs) {
    const int bin = blockIdx.x;
    const int row = n_l + 1;
    const double2 *c = reinterpret_cast<const double2 *>(consts) + (size_t)bin * (n_pad / 2) * row;
    const int n = n_pts[bin];
    const int pairs = n >> 1;
    // entries (l, m) with m <= l of G, and (l, -1) for b
    const int n_entries = n_l * (n_l + 1) / 2 + n_l;
    for (int e = threadIdx.x; e < n_entries; e += blockDim.x) {
        int l, m;
        if (e < n_l) { l = e; m = -1; }
        else {
            int t = e - n_l;
            l = 0;
            while (t >= l + 1) { t -= l + 1; ++l; }
            m = t;
        }
        double acc = 0.0;
        for (int j2 = 0; j2 < pairs; ++j2) {
            const double2 x = c[(size_t)j2 * row + 1 + l];
            const double2 y = c[(size_t)j2 * row + 1 + m];      // m == -1 -> the data pair
            acc = fma(x.x, y.x, acc);
            acc = fma(x.y, y.y, acc);
        }
        if (n & 1) acc = fma(c[(size_t)pairs * row + 1 + l].x, c[(size_t)pairs * row + 1 + m].x, acc);
        if (m < 0) rhs[(size_t)bin * n_l + l] = acc;
        else {
            gram[((size_t)bin * n_l + l) * n_l + m] = acc;
            gram[((size_t)bin * n_l + m) * n_l + l] = acc;
        }
    }
}

// Solve M z = r for the nf x nf leading block (M symmetric positive definite, lower part used).
// Returns false if a pivot is not positive.
template <int L>
__device__ bool cholesky_solve(double (&M)[L * L], double (&r)[L], int nf) {
    for (int j = 0; j < nf; ++j) {
        double d = M[j * L + j];
        for (int k = 0; k < j; ++k) d -= M[j * L + k] * M[j * L + k];
        if (!(d > 0.0)) return false;
        d = sqrt(d);
        M[j * L + j] = d;
        for (int i = j + 1; i < nf; ++i) {
            double s = M[i * L + j];
            for (int k = 0; k < j; ++k) s -= M[i * L + k] * M[j * L + k];
            M[i * L + j] = s / d;
        }
    }
    for (int i = 0; i < nf; ++i) {                      // forward
        double s = r[i];
        for (int k = 0; k < i; ++k) s -= M[i * L + k] * r[k];
        r[i] = s / M[i * L + i];
    }
    for (int i = nf - 1; i >= 0; --i) {                 // backward
        double s = r[i];
        for (int k = i + 1; k < nf; ++k) s -= M[k * L + i] * r[k];
        r[i] = s / M[i * L + i];
    }
    return true;
}

template <int L>
__global__ void __launch_bounds__(128) bvls_kernel(const double *__restrict__ gram, const double *__restrict__ rhs,
                                                   const double *__restrict__ lo_g, const double *__restrict__ hi_g,
                                                   const double *__restrict__ starts, int starts_per_bin,
                                                   int shared_starts, double *__restrict__ out, int *__restrict__ iters) {
    __shared__ double sG[L * L], sB[L], sLo[L], sHi[L];
    const int bin = blockIdx.y;
    for (int i = threadIdx.x; i < L * L; i += blockDim.x) sG[i] = gram[(size_t)bin * L * L + i];
    if (threadIdx.x < L) {
        sB[threadIdx.x] = rhs[(size_t)bin * L + threadIdx.x];
        sLo[threadIdx.x] = lo_g[threadIdx.x];
        sHi[threadIdx.x] = hi_g[threadIdx.x];
    }
    __syncthreads();
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= starts_per_bin) return;
    const double *a0 = starts + ((size_t)(shared_starts ? 0 : bin) * starts_per_bin + s) * L;

    double a[L];
    unsigned free_mask = 0;
    double scale = 0.0;
    for (int i = 0; i < L; ++i) {
        double v = a0[i];
        v = v < sLo[i] ? sLo[i] : (v > sHi[i] ? sHi[i] : v);       // NaN start -> falls to the lower bound below
        if (!(v == v)) v = sLo[i];
        a[i] = v;
        if (v > sLo[i] && v < sHi[i]) free_mask |= 1u << i;
        scale = fmax(scale, sG[i * L + i]);
    }
    const double ridge = scale * 1e-15;                 // only matters for rank-deficient bins (N < L)
    const double tol = 1e-12;

    int it = 0;
    const int max_it = 6 * L + 12;
    unsigned tabu = 0;                  // variables that, once freed, wanted straight back out through their bound
    int just = -1;                      // variable freed by the current outer iteration (-1: warm start)
    for (;;) {
        // ---- inner loop: minimise over the free set, stepping only as far as feasibility allows
        bool bounced = false;
        for (int inner = 0; inner <= L && free_mask; ++inner) {
            int idx[L];
            int nf = 0;
            for (int i = 0; i < L; ++i)
                if (free_mask >> i & 1u) idx[nf++] = i;
            double M[L * L], r[L];
            for (int p = 0; p < nf; ++p) {
                const int i = idx[p];
                double v = sB[i];
                for (int j = 0; j < L; ++j)
                    if (!(free_mask >> j & 1u)) v -= sG[i * L + j] * a[j];
                r[p] = v;
                for (int q = 0; q <= p; ++q) M[p * L + q] = sG[i * L + idx[q]] + (p == q ? ridge : 0.0);
            }
            if (!cholesky_solve<L>(M, r, nf)) { free_mask = 0; break; }
            double alpha = 1.0;
            int hit = -1;
            for (int p = 0; p < nf; ++p) {
                const int i = idx[p];
                if (r[p] < sLo[i] || r[p] > sHi[i]) {
                    const double bound = r[p] < sLo[i] ? sLo[i] : sHi[i];
                    const double den = r[p] - a[i];
                    const double t = den != 0.0 ? (bound - a[i]) / den : 0.0;
                    if (t < alpha) { alpha = t; hit = p; }
                }
            }
            if (hit < 0) {                              // the free-set minimiser is feasible
                for (int p = 0; p < nf; ++p) a[idx[p]] = r[p];
                break;
            }
            if (inner == 0 && idx[hit] == just && alpha <= 0.0) {
                free_mask &= ~(1u << just);             // round-off: put it back and do not pick it again
                tabu |= 1u << just;
                bounced = true;
                continue;
            }
            if (alpha < 0.0) alpha = 0.0;
            for (int p = 0; p < nf; ++p) {
                const int i = idx[p];
                a[i] += alpha * (r[p] - a[i]);
            }
            {                                           // the blocking variable lands exactly on its bound
                const int i = idx[hit];
                a[i] = r[hit] < sLo[i] ? sLo[i] : sHi[i];
                free_mask &= ~(1u << i);
            }
            for (int p = 0; p < nf; ++p) {              // anything else that reached a bound
                const int i = idx[p];
                if (a[i] <= sLo[i]) { a[i] = sLo[i]; free_mask &= ~(1u << i); }
                else if (a[i] >= sHi[i]) { a[i] = sHi[i]; free_mask &= ~(1u << i); }
            }
        }
        if (just >= 0 && !bounced) tabu = 0;            // progress was made: forget old bounces
        if (++it > max_it) break;
        // ---- outer loop: KKT test on the bound variables, free the most violating one
        int best = -1;
        double best_w = 0.0;
        for (int i = 0; i < L; ++i) {
            if ((free_mask >> i & 1u) || (tabu >> i & 1u) || !(sHi[i] > sLo[i])) continue;
            double w = sB[i];
            for (int j = 0; j < L; ++j) w -= sG[i * L + j] * a[j];      // minus half the gradient
            const double thr = tol * (fabs(sB[i]) + scale);
            const bool wants_in = (a[i] <= sLo[i] && w > thr) || (a[i] >= sHi[i] && w < -thr);
            if (wants_in && fabs(w) > best_w) { best_w = fabs(w); best = i; }
        }
        if (best < 0) break;
        free_mask |= 1u << best;
        just = best;
    }
    double *o = out + ((size_t)bin * starts_per_bin + s) * L;
    for (int i = 0; i < L; ++i) o[i] = a[i];
    if (iters) iters[(size_t)bin * starts_per_bin + s] = it;
}

using BvlsKernel = void (*)(const double *, const double *, const double *, const double *, const double *, int, int,
                            double *, int *);

BvlsKernel pick_bvls(int n_l) {
    switch (n_l) {
#define MDA_CASE(l) case l: return bvls_kernel<l>;
        MDA_CASE(1) MDA_CASE(2) MDA_CASE(3) MDA_CASE(4) MDA_CASE(5) MDA_CASE(6) MDA_CASE(7) MDA_CASE(8)
        MDA_CASE(9) MDA_CASE(10) MDA_CASE(11) MDA_CASE(12) MDA_CASE(13) MDA_CASE(14) MDA_CASE(15) MDA_CASE(16)
#undef MDA_CASE
        default: return nullptr;
    }
}

}  // namespace

cudaError_t launch_gram(const double *consts, const int *n_pts, int bins, int n_l, int n_pad, double *gram, double *rhs,
                        cudaStream_t stream) {
    if (bins <= 0) return cudaSuccess;
    gram_kernel<<<bins, 256, 0, stream>>>(consts, n_pts, n_l, n_pad, gram, rhs);
    return cudaGetLastError();
}

cudaError_t launch_bvls(const double *gram, const double *rhs, const double *lo, const double *hi, const double *starts,
                        int bins, int n_l, int starts_per_bin, bool shared_starts, double *out, int *iters,
                        cudaStream_t stream) {
    if (bins <= 0 || starts_per_bin <= 0) return cudaSuccess;
    BvlsKernel k = pick_bvls(n_l);
    if (!k || bins > 65535) return cudaErrorInvalidValue;
    dim3 grid((unsigned)((starts_per_bin + 127) / 128), (unsigned)bins);
    k<<<grid, 128, 0, stream>>>(gram, rhs, lo, hi, starts, starts_per_bin, shared_starts ? 1 : 0, out, iters);
    return cudaGetLastError();
}

}  // namespace mda
