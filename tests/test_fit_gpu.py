"""Batched initial fits: mdaBatchFit against (i) the reference procedure
restated with scipy's L-BFGS-B on the oracle's calculateChi (mda.py:1664-1709) and (ii)
scipy's own bounded-variable least squares on the same design matrix."""
import numpy as np
import pytest
from scipy import optimize

from mimic_mda_b200 import chisq, fits, synth
from oracle import init_fit_oracle, reflib

pytestmark = pytest.mark.gpu


def _design(consts, n_pts, b):
    n = int(n_pts[b])
    return consts[b, 1:, :n].T, consts[b, 0, :n]          # A [N, L], y [N]


@pytest.mark.parametrize("shape", [(3, 4, 30), (2, 9, 60), (2, 8, 41), (2, 5, 6), (1, 16, 40)])
def test_fit_is_the_constrained_least_squares_minimiser(cs_lib, oracle_lib, shape):
    n_bins, n_l, n_pts = shape
    consts, counts, a_true = synth.make_bins(n_bins, n_l, n_pts, ragged=True)
    lo, hi = synth.default_bounds(n_l)
    hi = hi * 0.25                                         # make some bounds active
    store = chisq.DeviceStore(consts, counts, (lo, hi), cs_lib=cs_lib)
    rng = np.random.default_rng(5)
    starts = np.vstack([rng.uniform(lo, hi, size=(20, n_l)), lo[None], hi[None], -np.ones((1, n_l)), 5 * np.ones((1, n_l))])
    params, chi = store.fit(starts)
    assert np.all(params >= lo) and np.all(params <= hi)
    for b in range(n_bins):
        A, y = _design(consts, counts, b)
        ref = optimize.lsq_linear(A, y, bounds=(lo, hi), method="bvls", tol=1e-15, max_iter=2000)
        chi_ref = float(np.sum((A @ ref.x - y) ** 2))
        # never worse than scipy's own BVLS, and every start reaches the same minimum (convex problem)
        assert np.all(chi[b] <= chi_ref * (1 + 1e-9) + 1e-12), (b, chi[b].max(), chi_ref)
        assert chi[b].max() - chi[b].min() <= 1e-9 * chi[b].min() + 1e-12
        # Karush-Kuhn-Tucker conditions of the box-constrained problem, checked directly
        for a in params[b]:
            grad = 2.0 * A.T @ (A @ a - y)
            gtol = 1e-7 * (np.abs(A.T @ y).max() + 1.0)
            inside = (a > lo) & (a < hi)
            assert np.all(np.abs(grad[inside]) <= gtol)
            assert np.all(grad[a <= lo] >= -gtol) and np.all(grad[a >= hi] <= gtol)
        if ref.status > 0 and chi_ref <= chi[b].min() * (1 + 1e-9):
            np.testing.assert_allclose(params[b], np.broadcast_to(ref.x, params[b].shape), atol=1e-5 * max(1.0, np.abs(ref.x).max()))
        # chi is the batched calculateChi of the returned parameters
        np.testing.assert_allclose(chi[b], oracle_lib.chi_batch(y, A.T.copy(), params[b]), rtol=1e-12)
    store.close()


def test_fit_matches_the_reference_procedure(cs_lib):
    """do_init_fit restated (L-BFGS-B, finite-difference gradients, 3 forced extra fits) on the
    oracle's calculateChi: same minimum; the reference stops within its own tolerance."""
    consts, counts, _ = synth.make_bins(2, 4, 30)
    lo, hi = synth.default_bounds(4)
    starts = fits.calc_start_params([[0.05, 0.45], [0.05, 0.25], [0.05, 0.45], [0.05, 0.2]])
    assert starts.shape == (16, 4) and starts[1].tolist() == [0.05, 0.05, 0.05, 0.2]      # last index fastest
    lib = reflib.prepare_shared_object(reflib.reference_lib_path() or reflib.ORACLE_LIB)
    par_ref, chi_ref = init_fit_oracle.fit_bins(lib, consts, counts, lo, hi, starts, extra_fits=3)
    store = chisq.DeviceStore(consts, counts, (lo, hi), cs_lib=cs_lib)
    par, chi = store.fit(starts)
    assert np.all(chi <= chi_ref * (1 + 1e-9))             # never worse than L-BFGS-B
    np.testing.assert_allclose(chi, chi_ref, rtol=1e-6)    # and L-BFGS-B gets this close (factr=10, eps=1e-5)
    np.testing.assert_allclose(par, par_ref, atol=2e-4)
    gens, gchi = fits.initial_fits(store, starts, 5)
    assert gens.shape == (2, 5, 4) and np.all(np.diff(gchi, axis=1) >= 0)
    walkers = fits.gen_walker_starts(gens[0], 32, 2.0, 0.04, 0.02, np.ones(4), np.random.RandomState(3))
    assert walkers.shape == (32, 4) and np.all(walkers > 0) and np.all(walkers <= 1.0)
    store.close()


def test_fit_handles_many_starts_and_rank_deficiency(cs_lib):
    # the reference grid of mda_config0.py:138-145 (6000 starts, L = 8) for a few bins, and a bin
    # with fewer angles than distributions (singular normal equations)
    grid = fits.calc_start_params([[0.05, 0.15, 0.25, 0.35, 0.45]] * 3 + [[0.05, 0.25, 0.45]] + [[0.05, 0.2]] * 4)
    assert grid.shape == (6000, 8)
    consts, counts, _ = synth.make_bins(3, 8, 41)
    store = chisq.DeviceStore(consts, counts, cs_lib=cs_lib)
    par, chi = store.fit(grid)
    assert par.shape == (3, 6000, 8) and np.all(np.isfinite(chi))
    assert np.all(chi.max(axis=1) - chi.min(axis=1) <= 1e-9 * chi.min(axis=1) + 1e-12)
    store.close()
    consts, counts, _ = synth.make_bins(1, 6, 4)
    lo, hi = synth.default_bounds(6)
    store = chisq.DeviceStore(consts, counts, (lo, hi), cs_lib=cs_lib)
    par, chi = store.fit(np.full((3, 6), 0.2))
    A, y = _design(consts, counts, 0)
    ref = optimize.lsq_linear(A, y, bounds=(lo, hi), method="bvls", tol=1e-15, max_iter=2000)
    assert np.all(chi <= np.sum((A @ ref.x - y) ** 2) + 1e-9)
    store.close()


def test_fit_on_nearly_collinear_distributions(cs_lib):
    """Small-angle DWBA multipoles are close to collinear.  The active-set solves run on the bin's triangular factor
    (Householder on the free columns), not on the normal equations D D^T whose condition number is the square: with
    cond(D) ~ 1e7 the normal equations have none of fp64's digits left (cond^2 eps ~ 1e-2), the factor keeps nine."""
    rng = np.random.default_rng(11)
    n_l, n = 6, 40
    theta = np.linspace(0.5, 10.0, n)
    base = 10.0 ** (1.0 - 0.15 * theta)
    dists = np.stack([base * (1.0 + 1e-6 * (l + 1) * np.cos(0.3 * theta * (l + 1))) for l in range(n_l)])
    a_true = np.array([0.2, 0.1, 0.3, 0.05, 0.15, 0.25])
    y = a_true @ dists
    err = 0.05 * y
    data = (y + err * rng.standard_normal(n)) / err
    dists = dists / err
    assert np.linalg.cond(dists.T) > 1e6
    consts = np.zeros((1, 1 + n_l, n))
    consts[0, 0], consts[0, 1:] = data, dists
    lo, hi = np.zeros(n_l), np.full(n_l, 1.0)
    store = chisq.DeviceStore(consts, np.array([n], dtype=np.int32), (lo, hi), cs_lib=cs_lib)
    starts = np.vstack([rng.uniform(lo, hi, size=(30, n_l)), lo[None], hi[None]])
    params, chi = store.fit(starts)
    unconverged, null_col, max_it = store.fit_stats()
    assert unconverged == 0 and max_it > 0
    A = dists.T
    ref = optimize.lsq_linear(A, data, bounds=(lo, hi), method="bvls", tol=1e-15, max_iter=5000)
    chi_ref = float(np.sum((A @ ref.x - data) ** 2))
    assert np.all(chi[0] <= chi_ref * (1 + 1e-9)), (chi[0].max(), chi_ref)          # every start reaches the minimum
    assert chi[0].max() - chi[0].min() <= 1e-9 * chi[0].min()
    for a in params[0]:                                                               # KKT, with the design's own scale
        grad = 2.0 * A.T @ (A @ a - data)
        gtol = 1e-6 * (np.abs(A.T @ data).max() + 1.0)
        inside = (a > lo) & (a < hi)
        assert np.all(np.abs(grad[inside]) <= gtol) and np.all(grad[a <= lo] >= -gtol) and np.all(grad[a >= hi] <= gtol)
    store.close()
    # a rank-deficient bin reports its null columns instead of failing silently
    consts, counts, _ = synth.make_bins(1, 6, 4)
    store = chisq.DeviceStore(consts, counts, synth.default_bounds(6), cs_lib=cs_lib)
    store.fit(np.full((3, 6), 0.2))
    unconverged, null_col, _ = store.fit_stats()
    assert unconverged == 0 and null_col == 3
    store.close()
