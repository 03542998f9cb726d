// fit_kernel.cu -- batched initial fits.
//
// The reference finds walker generators by running scipy's L-BFGS-B with finite-difference
// gradients on calculateChi from every point of a Cartesian start grid, several rounds per
// start (mda.py:1114, 1664-1709, 1912-1957; 6000 starts x 4 rounds in mda_config0.py), i.e.
// millions of scalar chi^2 calls per Ex bin.  chi^2(a) = |d - D^T a|^2 is a convex
// quadratic and the constraints are a box, so every start has the same constrained
// minimiser; here each (bin, start) pair is solved exactly by one thread with a
// bounded-variable least-squares active-set method (Stark & Parker's BVLS, warm-started at
// the clipped start point).
//
// The linear algebra runs on the bin's TRIANGULAR FACTOR, the one the sampler uses (chisq_abi.cu:
// factor_bin; [D^T | d] = Q [R t; 0 tau], computed in long double when the store is built):
// chi^2(a) = |R a - t|^2 + tau^2.  A free set's minimiser is the least-squares solution of
// R[:, free] x = t - R[:, bound] a_bound, obtained by Householder reflections on those columns -- never
// through the normal equations D D^T, whose condition number is the square of D's (small-angle DWBA
// multipoles are close to collinear).  The multipliers of the bound variables are R^T (t - R a).
// Columns that are numerically null (a bin with fewer angles than distributions) keep their current
// value: every value is a minimiser along them.
// The chi^2 of every solution is then evaluated by the batched likelihood kernel in kChi
// mode, in the reference's operation order.
#include "quad_chi.cuh"

namespace mda {

namespace {

// Least-squares solution over the free set.  A is the L x nf matrix of the free columns of R (row-major, stride L), y the
// right-hand side; both are overwritten.  x[p] = the minimiser's component for free variable p, or keep[p] where the
// column carries no information (pivot below `tiny`).  Returns the number of such null columns.
template <int L>
__device__ int householder_solve(double *A, double *y, int nf, const double *keep, double tiny, double *x) {
    unsigned null_cols = 0;
    for (int p = 0; p < nf; ++p) {
        double norm2 = 0.0;
        for (int k = p; k < L; ++k) norm2 = fma(A[k * L + p], A[k * L + p], norm2);
        if (!(norm2 > tiny)) { null_cols |= 1u << p; continue; }
        const double norm = sqrt(norm2), x0 = A[p * L + p];
        const double alpha = x0 > 0.0 ? -norm : norm;
        const double v0 = x0 - alpha;
        const double vv = norm2 - x0 * x0 + v0 * v0;               // |v|^2 with v = column - alpha e_p
        if (!(vv > 0.0)) { A[p * L + p] = alpha; continue; }
        for (int q = p + 1; q <= nf; ++q) {                        // q == nf: the right-hand side
            double dot = v0 * (q < nf ? A[p * L + q] : y[p]);
            for (int k = p + 1; k < L; ++k) dot = fma(A[k * L + p], q < nf ? A[k * L + q] : y[k], dot);
            const double f = 2.0 * dot / vv;
            if (q < nf) {
                A[p * L + q] -= f * v0;
                for (int k = p + 1; k < L; ++k) A[k * L + q] -= f * A[k * L + p];
            } else {
                y[p] -= f * v0;
                for (int k = p + 1; k < L; ++k) y[k] -= f * A[k * L + p];
            }
        }
        A[p * L + p] = alpha;
    }
    int n_null = 0;
    for (int p = nf - 1; p >= 0; --p) {                            // back substitution
        if (null_cols >> p & 1u) { x[p] = keep[p]; ++n_null; continue; }
        double s = y[p];
        for (int q = p + 1; q < nf; ++q) s -= A[p * L + q] * x[q];
        x[p] = s / A[p * L + p];
    }
    return n_null;
}

template <int L>
__global__ void __launch_bounds__(128) bvls_kernel(const double *__restrict__ quad, const double *__restrict__ lo_g,
                                                   const double *__restrict__ hi_g, const double *__restrict__ starts,
                                                   int starts_per_bin, int shared_starts, double *__restrict__ out,
                                                   int *__restrict__ iters) {
    __shared__ double sR[L * L], sT[L], sLo[L], sHi[L];            // R dense (zeros below the diagonal), t
    const int bin = blockIdx.y;
    const double *packed = quad + (size_t)bin * quad_stride(L);
    for (int e = threadIdx.x; e < L * L; e += blockDim.x) {
        const int i = e / L, c = e - i * L;
        sR[e] = c >= i ? packed[quad_row_offset(L, i) + (c - i)] : 0.0;
    }
    if (threadIdx.x < L) {
        sT[threadIdx.x] = packed[quad_row_offset(L, threadIdx.x) + (L - threadIdx.x)];
        sLo[threadIdx.x] = lo_g[threadIdx.x];
        sHi[threadIdx.x] = hi_g[threadIdx.x];
    }
    __syncthreads();
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= starts_per_bin) return;
    const double *a0 = starts + ((size_t)(shared_starts ? 0 : bin) * starts_per_bin + s) * L;

    double a[L];
    unsigned free_mask = 0;
    double scale = 0.0;                                            // largest squared column norm of R = largest diagonal of D D^T
    for (int i = 0; i < L; ++i) {
        double v = a0[i];
        v = v < sLo[i] ? sLo[i] : (v > sHi[i] ? sHi[i] : v);       // NaN start -> falls to the lower bound below
        if (!(v == v)) v = sLo[i];
        a[i] = v;
        if (v > sLo[i] && v < sHi[i]) free_mask |= 1u << i;
        double col = 0.0;
        for (int k = 0; k <= i; ++k) col = fma(sR[k * L + i], sR[k * L + i], col);
        scale = fmax(scale, col);
    }
    const double tiny = scale * 1e-28;                             // squared pivot below (1e-14)^2 of the largest column: null
    const double tol = 1e-12;

    int it = 0;
    int null_seen = 0;
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
            double A[L * L], y[L], r[L], keep[L];
            for (int k = 0; k < L; ++k) {
                double v = sT[k];
                for (int j = k; j < L; ++j)
                    if (!(free_mask >> j & 1u)) v -= sR[k * L + j] * a[j];
                y[k] = v;
                for (int p = 0; p < nf; ++p) A[k * L + p] = sR[k * L + idx[p]];
            }
            for (int p = 0; p < nf; ++p) keep[p] = a[idx[p]];
            null_seen |= householder_solve<L>(A, y, nf, keep, tiny, r) > 0 ? 1 : 0;
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
        // ---- outer loop: KKT test on the bound variables, free the most violating one.  w = R^T (t - R a) is minus half
        // the gradient of chi^2.
        double resid[L];
        for (int k = 0; k < L; ++k) {
            double v = sT[k];
            for (int j = k; j < L; ++j) v -= sR[k * L + j] * a[j];
            resid[k] = v;
        }
        int best = -1;
        double best_w = 0.0;
        for (int i = 0; i < L; ++i) {
            if ((free_mask >> i & 1u) || (tabu >> i & 1u) || !(sHi[i] > sLo[i])) continue;
            double w = 0.0, b = 0.0;
            for (int k = 0; k <= i; ++k) { w = fma(sR[k * L + i], resid[k], w); b = fma(sR[k * L + i], sT[k], b); }
            const double thr = tol * (fabs(b) + scale);
            const bool wants_in = (a[i] <= sLo[i] && w > thr) || (a[i] >= sHi[i] && w < -thr);
            if (wants_in && fabs(w) > best_w) { best_w = fabs(w); best = i; }
        }
        if (best < 0) break;
        free_mask |= 1u << best;
        just = best;
    }
    double *o = out + ((size_t)bin * starts_per_bin + s) * L;
    for (int i = 0; i < L; ++i) o[i] = a[i];
    // iterations taken; negative: the iteration limit was hit (the point is feasible but need not be the minimiser);
    // bit 16 set: a free set met a numerically null column (a rank-deficient bin: any value along it is a minimiser)
    if (iters) iters[(size_t)bin * starts_per_bin + s] = (it > max_it ? -it : it) | (null_seen && it <= max_it ? 1 << 16 : 0);
}

using BvlsKernel = void (*)(const double *, const double *, const double *, const double *, int, int, double *, int *);

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

cudaError_t launch_bvls(const double *quad, const double *lo, const double *hi, const double *starts,
                        int bins, int n_l, int starts_per_bin, bool shared_starts, double *out, int *iters,
                        cudaStream_t stream) {
    if (bins <= 0 || starts_per_bin <= 0) return cudaSuccess;
    BvlsKernel k = pick_bvls(n_l);
    if (!k || bins > 65535) return cudaErrorInvalidValue;
    dim3 grid((unsigned)((starts_per_bin + 127) / 128), (unsigned)bins);
    k<<<grid, 128, 0, stream>>>(quad, lo, hi, starts, starts_per_bin, shared_starts ? 1 : 0, out, iters);
    return cudaGetLastError();
}

}  // namespace mda
