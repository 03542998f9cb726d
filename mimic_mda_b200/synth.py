"""Seeded synthetic inputs of the shapes BASELINE.json names.

What populates the store mirrors what the reference puts into its struct:
``data = sigma/dsigma`` (mda.py:176) and ``dists[l] = D_l(theta)/dsigma``
(mda.py:2069), i.e. all division by the experimental error is done before the
likelihood ever runs (C/ChiSquare.h:4-13).  Bounds are the reference's
``(0, 1/EWSR_l)`` with EWSR = 1 (mda.py:1124, mda_config0.py:42).

Host-side numpy only; used by bench.py, __graft_entry__.smoke() and the tests.
"""
import numpy as np

# name -> (bins, dists L, angles N, walkers W) -- BASELINE.json "configs"
CONFIGS = {
    "C1": dict(bins=1, n_l=4, n_pts=30, walkers=64),     # mda_config0.py-style single bin (CPU case)
    "C2": dict(bins=100, n_l=5, n_pts=30, walkers=256),  # full spectrum, L=0-3 + IVGDR
    "C3": dict(bins=100, n_l=9, n_pts=60, walkers=512),  # high-multipole fit
    "C4": dict(bins=200, n_l=9, n_pts=60, walkers=512),  # the metric's config, sharded over GPUs
}


def pad_even(n):
    """Rows are padded to an even count of doubles (16-byte rows for bulk copies)."""
    return int(n) + (int(n) & 1)


def model_dists(theta_deg, n_l):
    """Smooth, positive, decaying-oscillatory angular distributions [L, N] (mb/sr-like)."""
    theta = np.asarray(theta_deg, dtype=np.float64)
    ell = np.arange(n_l, dtype=np.float64)[:, None]
    return 10.0 ** (1.0 - 0.15 * theta[None, :]) * (1.1 + np.cos(theta[None, :] * (0.3 + 0.12 * ell) + ell))


def make_bin(bin_index, n_l, n_pts, rel_err=0.05, seed_base=1000):
    """One Ex bin: returns (data [N], dists [L, N], a_true [L]), error-normalised."""
    rng = np.random.default_rng(seed_base + int(bin_index))
    theta = np.linspace(0.5, 10.0, n_pts)
    shapes = model_dists(theta, n_l)
    a_true = rng.uniform(0.0, 0.3, size=n_l)
    y = a_true @ shapes
    err = rel_err * y
    y_obs = y + err * rng.standard_normal(n_pts)
    return y_obs / err, shapes / err[None, :], a_true


def make_bins(n_bins, n_l, n_pts, rel_err=0.05, seed_base=1000, ragged=False, first_bin=0, bin_step=1):
    """A spectrum of Ex bins in the device-store layout.

    Returns ``consts [B, 1+L, n_pad]`` (row 0 data, rows 1..L distributions,
    zero padded), ``n_pts [B]`` int32 and ``a_true [B, L]``.  With
    ``ragged=True`` bin b keeps ``n_pts - (b % 5)`` angles (the reference's
    per-row Max Theta cut makes N differ per bin, mda.py:2387-2394).
    ``first_bin``/``bin_step`` select the global bin indices b = first_bin +
    k*bin_step, so a rank of a round-robin partition generates exactly its bins.
    """
    n_pad = pad_even(n_pts)
    consts = np.zeros((n_bins, 1 + n_l, n_pad), dtype=np.float64)
    counts = np.empty(n_bins, dtype=np.int32)
    a_true = np.empty((n_bins, n_l), dtype=np.float64)
    for k in range(n_bins):
        gbin = first_bin + k * bin_step
        n = n_pts - (gbin % 5) if ragged else n_pts
        n = max(n, 1)
        data, dists, a = make_bin(gbin, n_l, n, rel_err, seed_base)
        consts[k, 0, :n] = data
        consts[k, 1:, :n] = dists
        counts[k] = n
        a_true[k] = a
    return consts, counts, a_true


def default_bounds(n_l, ewsr=None):
    """(lo, hi) = (0, 1/EWSR_l) per dimension (mda.py:1124)."""
    ewsr = np.ones(n_l) if ewsr is None else np.asarray(ewsr, dtype=np.float64)
    return np.zeros(n_l, dtype=np.float64), 1.0 / ewsr


def make_walkers(a_true, n_walkers, seed=7, scatter=0.02, oob_frac=0.01, lo=None, hi=None):
    """Walker blocks ``[B, W, L]`` near a_true, reflected into the box, with
    about ``oob_frac`` of the rows pushed outside it to exercise the -inf prior."""
    a_true = np.asarray(a_true, dtype=np.float64)
    n_bins, n_l = a_true.shape
    if lo is None or hi is None:
        lo, hi = default_bounds(n_l)
    rng = np.random.default_rng(seed)
    p = a_true[:, None, :] + scatter * rng.standard_normal((n_bins, n_walkers, n_l))
    p = np.abs(p - lo) + lo                       # reflect at the lower edge
    p = hi - np.abs(hi - p)                       # and at the upper edge
    p = np.clip(p, lo, hi)
    if oob_frac > 0.0:
        bad = rng.random((n_bins, n_walkers)) < oob_frac
        which = rng.integers(0, n_l, size=(n_bins, n_walkers))
        sign = rng.random((n_bins, n_walkers)) < 0.5
        b_idx, w_idx = np.nonzero(bad)
        l_idx = which[b_idx, w_idx]
        p[b_idx, w_idx, l_idx] = np.where(sign[b_idx, w_idx], lo[l_idx] - 0.01, hi[l_idx] + 0.01)
    return np.ascontiguousarray(p)
