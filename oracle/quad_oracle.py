"""CPU restatement of the triangular-factor evaluation of chi^2 -- TEST INFRASTRUCTURE ONLY.

chi^2(a) = sum_j (d_j - sum_l a_l D_lj)^2 (reference C/ChiSquare.c:188-210) = |M [a; -1]|^2 with M = [D^T | d]
(N x (L+1)); with M = Q T, T upper triangular, chi^2(a) = |T [a; -1]|^2.  ``factor`` follows
mimic_mda_b200/csrc/chisq_abi.cu::factor_bin (Householder reflections in extended precision, rounded to double
once); ``chi`` follows csrc/quad_chi.cuh::chi_from_factor.  Pinned to the reference by tests/test_oracle.py: the
golden outputs of the unmodified reference library are reproduced to 1e-12 relative.
"""
import numpy as np


def factor(data, dists):
    """data [N], dists [L][N] (both already divided by the errors, mda.py:176) -> T [(L+1), (L+1)] upper triangular."""
    dists = np.asarray(dists, dtype=np.longdouble)
    n_l, n = dists.shape
    c = n_l + 1
    rows = max(n, c)
    m = np.zeros((rows, c), dtype=np.longdouble)
    m[:n, :n_l] = dists.T
    m[:n, n_l] = np.asarray(data, dtype=np.longdouble)
    for k in range(c):
        x = m[k:, k].copy()
        norm = np.sqrt(np.sum(x * x))
        if norm == 0:
            continue
        x[0] -= -norm if x[0] > 0 else norm
        vv = np.sum(x * x)
        if vv == 0:
            continue
        m[k:, k:] -= np.outer(x, (2 * (x @ m[k:, k:])) / vv)
    return np.triu(m[:c]).astype(np.float64)


KAPPA_MAX = 1e-13 / (4.0 * 2.0 ** -53)      # chisq_abi.cu: kQuadKappaMax


def kappa(t):
    """|d| / sqrt(min chi^2) of a factor (chisq_abi.cu::factor_bin): chi^2 from T and chi^2 in the reference's order
    differ by about eps * kappa relative; bins beyond KAPPA_MAX are not evaluated from T by the library."""
    n_l = t.shape[0] - 1
    norm_d = float(np.sqrt(np.sum(t[:, n_l].astype(np.longdouble) ** 2)))
    if norm_d == 0.0:
        return 1.0
    return norm_d / abs(t[n_l, n_l]) if t[n_l, n_l] != 0.0 else np.inf


def chi(t, params):
    """t [(L+1), (L+1)], params [..., L] -> chi^2 [...], accumulated as the kernels do (row sums, minus the data, squared)."""
    n_l = t.shape[0] - 1
    params = np.asarray(params, dtype=np.float64)
    out = np.zeros(params.shape[:-1])
    for i in range(n_l + 1):
        u = np.zeros(params.shape[:-1])
        for l in range(i, n_l):
            u = u + t[i, l] * params[..., l]
        u = u - t[i, n_l]
        out = out + u * u
    return out
