"""numpy restatement of the reference likelihood -- TEST INFRASTRUCTURE ONLY.

Independent of both the C oracle and the CUDA path so that parity can be
triangulated: (reference .so) vs (C restatement) vs (this) vs (long double).

Reference arithmetic followed here:
  * residuals  r_j = d_j - a_0 D_0j, then r_j -= a_l D_lj for l ascending
    (C/ChiSquare.c:188-202)
  * chi^2      sum of r_j^2 in ascending j from 0 (C/ChiSquare.c:204-210)
  * lnL        chi^2 / (-2.0) (C/ChiSquare.c:212)
  * box prior  any a_l < lo_l or a_l > hi_l  ->  -inf, bounds inclusive, NaN
    compares false and falls through (mda.py:1574-1576)
The reference's own numpy cross-check has the same shape
(C/so_py_tester.py:90-103).
"""
import numpy as np


def chi_batch(data, dists, params, dtype=np.float64):
    """chi^2 for params [W, L] against data [N], dists [L, N]; returns [W].

    Vectorised over walkers, sequential over distributions and over angles so
    that each walker sees the reference's operation order.
    """
    d = np.asarray(data, dtype=dtype)
    D = np.asarray(dists, dtype=dtype)
    a = np.atleast_2d(np.asarray(params, dtype=dtype))
    n_l, n_pts = D.shape
    assert a.shape[1] == n_l and d.shape[0] == n_pts
    res = d[None, :] - a[:, 0:1] * D[0][None, :]
    for l in range(1, n_l):
        res -= a[:, l:l + 1] * D[l][None, :]
    chi = np.zeros(a.shape[0], dtype=dtype)
    for j in range(n_pts):
        chi += res[:, j] * res[:, j]
    return chi


def lnlike_batch(data, dists, params, dtype=np.float64):
    """lnL = chi^2 / (-2) for params [W, L]; returns [W]."""
    return chi_batch(data, dists, params, dtype) / dtype(-2.0)


def outside_box(params, lo, hi):
    """Boolean [W]: True where the flat prior of mda.py:1574-1576 is zero."""
    a = np.atleast_2d(np.asarray(params, dtype=np.float64))
    lo = np.asarray(lo, dtype=np.float64)
    hi = np.asarray(hi, dtype=np.float64)
    return np.any((a < lo[None, :]) | (a > hi[None, :]), axis=1)


def lnpost_batch(data, dists, lo, hi, params, dtype=np.float64):
    """Log-posterior (flat box prior) for params [W, L]; returns [W] float64."""
    out = np.asarray(lnlike_batch(data, dists, params, dtype), dtype=np.float64).copy()
    out[outside_box(params, lo, hi)] = -np.inf
    return out


def lnpost_bins(consts, n_pts, lo, hi, params, dtype=np.float64):
    """Multi-bin form: consts [B, 1+L, n_pad] (row 0 data, rows 1..L dists,
    zero padded past n_pts[b]); params [B, W, L]; returns [B, W]."""
    consts = np.asarray(consts)
    params = np.asarray(params)
    n_bins = consts.shape[0]
    out = np.empty(params.shape[:2], dtype=np.float64)
    for b in range(n_bins):
        n = int(n_pts[b])
        out[b] = lnpost_batch(consts[b, 0, :n], consts[b, 1:, :n], lo, hi, params[b], dtype)
    return out


def lnlike_exact(data, dists, params):
    """Extended-precision (x87 80-bit long double) value, for conditioning checks."""
    return np.asarray(lnlike_batch(data, dists, params, dtype=np.longdouble))
