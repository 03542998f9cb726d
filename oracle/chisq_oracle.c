/*
 * chisq_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU restatement of the reference's chi-square log-likelihood
 * (reference: C/ChiSquare.c) used as the checker for the CUDA path.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this library; nothing under
 * mimic_mda_b200/ links, imports or calls it.
 *
 * Parity status: PINNED.  The restatement is checked (tests/test_oracle.py)
 * against (i) the reference's single known-answer vector
 * (C/ChiSqTester.c:20,43 -> chi 0.456, lnL -0.228) and (ii) outputs of the
 * unmodified reference library compiled here from /root/reference/C/ChiSquare.c
 * (oracle/Makefile -> oracle/_ref/libChiSq.so), frozen in
 * tests/golden/golden_lnl.npz by tests/golden/make_golden.py.
 *
 * Arithmetic restated (all IEEE fp64, the order matters for the last bits):
 *   residual init      r_j  = d_j - a_0 * D_0j              C/ChiSquare.c:188-191
 *   residual updates   r_j -= a_l * D_lj, l = 1..L-1 asc.   C/ChiSquare.c:194-202
 *   accumulation       chi  = sum_j r_j^2, j ascending,
 *                      starting from 0                      C/ChiSquare.c:204-210
 *   lnL                chi / (-2.0)                         C/ChiSquare.c:212
 *   chi^2 objective    same without the division            C/ChiSquare.c:114-172
 *   caller scratch     residuals left in caller's array     C/ChiSquare.c:216-252
 *   row layout         dists[l*numPts + j]                  C/ChiSquare.c:97-102
 *   index guard        distIndex >= numLs is ignored        C/ChiSquare.c:88-89
 *   box prior          a_l < lo_l or a_l > hi_l -> -inf     mda.py:1574-1576
 *
 * The seven reference symbols are also exported (same signatures as
 * C/ChiSquare.h:31-49) so one ctypes loader can drive the reference library,
 * this oracle and the CUDA library interchangeably in the tests.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

typedef struct {
    int n_pts;       /* angles in this Ex bin            */
    int n_dists;     /* fitted distributions (numLs)     */
    double *d;       /* [n_pts]  data / error            */
    double *D;       /* [n_dists][n_pts] dist / error    */
    double *r;       /* [n_pts] scratch                  */
} orc_bin;

/* ---- core arithmetic --------------------------------------------------- */

/* chi^2 of one parameter vector; residuals are left in r (length n). */
static double orc_chi_core(int n, int L, const double *d, const double *D,
                           const double *a, double *r)
{
    const double a0 = a[0];                 /* L >= 1 is assumed, as in the reference */
    for (int j = 0; j < n; ++j)
        r[j] = d[j] - a0 * D[j];
    for (int l = 1; l < L; ++l) {
        const double al = a[l];
        const double *row = D + (size_t)l * (size_t)n;
        for (int j = 0; j < n; ++j)
            r[j] -= al * row[j];
    }
    double chi = 0.0;
    for (int j = 0; j < n; ++j)
        chi += r[j] * r[j];
    return chi;
}

/* ---- batched oracle entry points (used by the parity tests) ------------- */

/* lnL for W parameter vectors of one bin.  params is [W][L], out is [W]. */
void orc_lnlike_batch(int n, int L, const double *d, const double *D,
                      const double *params, int W, double *out)
{
    double *r = (double *)malloc(sizeof(double) * (size_t)(n > 0 ? n : 1));
    for (int w = 0; w < W; ++w)
        out[w] = orc_chi_core(n, L, d, D, params + (size_t)w * L, r) / (-2.0);
    free(r);
}

/* chi^2 for W parameter vectors of one bin. */
void orc_chi_batch(int n, int L, const double *d, const double *D,
                   const double *params, int W, double *out)
{
    double *r = (double *)malloc(sizeof(double) * (size_t)(n > 0 ? n : 1));
    for (int w = 0; w < W; ++w)
        out[w] = orc_chi_core(n, L, d, D, params + (size_t)w * L, r);
    free(r);
}

/*
 * Log-posterior with the flat box prior of mda.py:1540-1578 for W vectors:
 * any a_l < lo_l or a_l > hi_l gives -inf (bounds inclusive, NaN falls
 * through to the likelihood exactly as the Python comparisons do).
 */
void orc_lnpost_batch(int n, int L, const double *d, const double *D,
                      const double *lo, const double *hi,
                      const double *params, int W, double *out)
{
    double *r = (double *)malloc(sizeof(double) * (size_t)(n > 0 ? n : 1));
    for (int w = 0; w < W; ++w) {
        const double *a = params + (size_t)w * L;
        int outside = 0;
        for (int l = 0; l < L; ++l)
            if (a[l] < lo[l] || a[l] > hi[l]) { outside = 1; break; }
        out[w] = outside ? -INFINITY
                         : orc_chi_core(n, L, d, D, a, r) / (-2.0);
    }
    free(r);
}

/*
 * Multi-bin form matching the batched CUDA entry point: consts is
 * [B][(1+L)][n_pad] with row 0 = data, rows 1..L = distributions, each
 * row zero padded from n_pts[b] to n_pad; params [B][W][L]; out [B][W].
 */
void orc_lnpost_bins(int B, int L, int n_pad, const int *n_pts,
                     const double *consts, const double *lo, const double *hi,
                     const double *params, int W, double *out)
{
    const size_t bin_stride = (size_t)(L + 1) * (size_t)n_pad;
    double *D = (double *)malloc(sizeof(double) * (size_t)L * (size_t)(n_pad > 0 ? n_pad : 1));
    for (int b = 0; b < B; ++b) {
        const int n = n_pts[b];
        const double *blk = consts + (size_t)b * bin_stride;
        for (int l = 0; l < L; ++l)   /* repack to the reference's dense [L][n] */
            memcpy(D + (size_t)l * n, blk + (size_t)(l + 1) * n_pad, sizeof(double) * (size_t)n);
        orc_lnpost_batch(n, L, blk, D, lo, hi,
                         params + (size_t)b * W * L, W, out + (size_t)b * W);
    }
    free(D);
}

/* ---- the reference's seven symbols (C/ChiSquare.h:31-49) ---------------- */

void *makeMdaStruct(int numPts, int numL)
{
    orc_bin *s = (orc_bin *)calloc(1, sizeof(orc_bin));
    s->n_pts = numPts;
    s->n_dists = numL;
    s->d = (double *)calloc((size_t)(numPts > 0 ? numPts : 1), sizeof(double));
    s->D = (double *)calloc((size_t)(numPts > 0 ? numPts : 1) * (size_t)(numL > 0 ? numL : 1), sizeof(double));
    s->r = (double *)calloc((size_t)(numPts > 0 ? numPts : 1), sizeof(double));
    return s;
}

void freeMdaStruct(void *p)
{
    orc_bin *s = (orc_bin *)p;
    if (!s) return;
    free(s->r); free(s->D); free(s->d); free(s);
}

void setMdaData(void *p, double *divData)
{
    orc_bin *s = (orc_bin *)p;
    memcpy(s->d, divData, sizeof(double) * (size_t)s->n_pts);
}

void setMdaDist(void *p, int distIndex, double *dist)
{
    orc_bin *s = (orc_bin *)p;
    if (distIndex >= s->n_dists) return;          /* C/ChiSquare.c:88-89 */
    memcpy(s->D + (size_t)distIndex * s->n_pts, dist, sizeof(double) * (size_t)s->n_pts);
}

double calculateChi(void *p, double *params)
{
    orc_bin *s = (orc_bin *)p;
    return orc_chi_core(s->n_pts, s->n_dists, s->d, s->D, params, s->r);
}

double calculateLnLiklihood(void *p, double *params)
{
    orc_bin *s = (orc_bin *)p;
    return orc_chi_core(s->n_pts, s->n_dists, s->d, s->D, params, s->r) / (-2.0);
}

double calculateLnLiklihoodResids(void *p, double *params, double *residArray)
{
    orc_bin *s = (orc_bin *)p;
    return orc_chi_core(s->n_pts, s->n_dists, s->d, s->D, params, residArray) / (-2.0);
}
