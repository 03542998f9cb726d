"""chi^2 from the bins' triangular factors (quad_chi.cuh; lnpost_quad_kernel, ensemble_quad_kernel) against the
reference's outputs and the oracle.

chi^2(a) = |d - D^T a|^2 (C/ChiSquare.c:188-210) = |T [a; -1]|^2 with [D^T | d] = Q T: the same function evaluated
with (L+1)(L+2)/2 multiply-adds.  Bar: the tolerance north_star states for fp64 -- 1e-12 relative to the reference
library's log-likelihood -- with -inf / NaN exactly where the reference has them.
"""
import numpy as np
import pytest

from conftest import rel_err, same_special
from mimic_mda_b200 import chisq, synth

pytestmark = pytest.mark.gpu

TOL = 1e-12                      # fp64, relative -- the tolerance north_star states


@pytest.fixture
def quad(monkeypatch):
    monkeypatch.setenv("MDA_BATCH_IMPL", "quad")


def test_factor_evaluation_matches_reference_outputs(cs_lib, golden, quad):
    """Every golden case: outputs of the unmodified reference library (tests/golden/make_golden.py)."""
    for name in golden.cases:
        g = golden.case(name)
        store = chisq.DeviceStore(g["consts"], g["n_pts"], (g["lo"], g["hi"]), cs_lib=cs_lib)
        lnp, chi = store.lnpost(g["params"]), store.chi(g["params"])
        assert same_special(lnp, g["lnpost"]), name
        assert rel_err(lnp, g["lnpost"]) <= TOL, (name, rel_err(lnp, g["lnpost"]))
        assert same_special(chi, g["chi"]), name
        assert rel_err(chi, g["chi"]) <= TOL, (name, rel_err(chi, g["chi"]))
        store.close()


@pytest.mark.parametrize("shape", [(3, 4, 16, 33), (2, 9, 60, 1), (5, 12, 256, 200), (2, 16, 24, 95), (1, 7, 41, 8192),
                                   (150, 5, 30, 128), (4, 8, 60, 2500), (3, 1, 9, 40),
                                   (3, 6, 4, 64),          # fewer angles than parameters: a rank-deficient factor
                                   (2, 9, 9, 64)])         # exactly as many
def test_factor_evaluation_against_oracle_and_direct_kernel(cs_lib, oracle_lib, monkeypatch, shape):
    n_bins, n_l, n_pts, walkers = shape
    consts, n_pts_arr, a_true = synth.make_bins(n_bins, n_l, n_pts, ragged=n_pts > 8)
    lo, hi = synth.default_bounds(n_l)
    params = synth.make_walkers(a_true, walkers, seed=9)
    if walkers > 4:
        params[0, 3, 0] = np.nan                          # a NaN passes the box prior and poisons the sum (mda.py:1574-1576)
        params[-1, 1, n_l - 1] = hi[n_l - 1]              # on the inclusive bound: inside
    want = oracle_lib.lnpost_bins(consts, n_pts_arr, lo, hi, params)
    store = chisq.DeviceStore(consts, n_pts_arr, (lo, hi), cs_lib=cs_lib)
    monkeypatch.setenv("MDA_BATCH_IMPL", "dfma")
    direct = store.lnpost(params)
    monkeypatch.setenv("MDA_BATCH_IMPL", "quad")
    got = store.lnpost(params)
    assert same_special(got, want) and same_special(got, direct)
    assert rel_err(got, want) <= TOL, rel_err(got, want)
    assert rel_err(got, direct) <= TOL
    # a walker's value does not depend on the batch it is evaluated in
    if walkers > 8:
        part = np.ascontiguousarray(params[:, 3:8])
        np.testing.assert_array_equal(store.lnpost(part), got[:, 3:8])
    store.close()


def test_factor_evaluation_far_from_the_minimum_and_with_small_errors(cs_lib, oracle_lib, quad):
    """1 % cross-section errors (|d|^2 ~ 6e5 against chi^2 ~ N) and parameter vectors all over the box: the
    cancellation inside a row of T [a; -1] is that of a residual, so the bound holds there too."""
    consts, n_pts, a_true = synth.make_bins(6, 9, 60, rel_err=0.01, ragged=True)
    lo, hi = synth.default_bounds(9)
    rng = np.random.default_rng(4)
    near = synth.make_walkers(a_true, 256, seed=3, scatter=0.002, oob_frac=0.0)
    far = rng.uniform(0.0, 1.0, size=(6, 256, 9))
    store = chisq.DeviceStore(consts, n_pts, (lo, hi), cs_lib=cs_lib)
    for params in (near, far):
        want = oracle_lib.lnpost_bins(consts, n_pts, lo, hi, params)
        got = store.lnpost(params)
        assert same_special(got, want)
        assert rel_err(got, want) <= TOL, rel_err(got, want)
    store.close()


def test_sampler_log_probabilities_equal_the_oracle_on_the_visited_positions(cs_lib, oracle_lib):
    """The default stepper's own log-probabilities, on C4-shaped bins, against the oracle on the same positions."""
    from mimic_mda_b200.ensemble import DeviceEnsemble
    consts, n_pts, a_true = synth.make_bins(8, 9, 60, ragged=True)
    lo, hi = synth.default_bounds(9)
    store = chisq.DeviceStore(consts, n_pts, (lo, hi), cs_lib=cs_lib)
    ens = DeviceEnsemble(store, 512, seed=77)
    assert ens.describe(40).startswith(("ensemble_qws_kernel", "ensemble_quad_kernel"))      # the triangular-factor steppers
    ens.run_mcmc(synth.make_walkers(a_true, 512, seed=1, oob_frac=0.0), 40, thin=4)
    chain, lnp = ens.chain, ens.lnprobability                                  # [B, W, kept, L], [B, W, kept]
    for k in range(chain.shape[2]):
        want = oracle_lib.lnpost_bins(consts, n_pts, lo, hi, np.ascontiguousarray(chain[:, :, k]))
        assert rel_err(lnp[:, :, k], want) <= TOL
    ens.close()
    store.close()


@pytest.mark.parametrize("rel", [1e-2, 1e-3, 3e-4, 1e-4])
def test_sampler_stays_within_the_bound_whatever_the_conditioning(cs_lib, oracle_lib, rel):
    """Near the minimum, chi^2 from T differs from the reference's arithmetic by ~eps |d| / sqrt(chi^2): the library
    measures kappa per bin when the store is built and steps a store with a bin beyond the limit in the reference's
    operation order (C/ChiSquare.c:188-210).  Whatever it picks, the sampler's log-probabilities are within 1e-12."""
    from mimic_mda_b200.ensemble import DeviceEnsemble
    consts, n_pts, a_true = synth.make_bins(6, 9, 60, rel_err=rel, ragged=True)
    lo, hi = synth.default_bounds(9)
    store = chisq.DeviceStore(consts, n_pts, (lo, hi), cs_lib=cs_lib)
    kappa, limit, safe = store.conditioning()
    assert kappa.shape == (6,) and np.all(kappa > 0.5 / rel) and np.all(kappa < 3.0 / rel)
    assert safe == bool(np.all(kappa <= limit)) and safe == (rel >= 1e-2)
    ens = DeviceEnsemble(store, 256, seed=5)
    text = ens.describe(24)
    if safe:
        assert text.startswith(("ensemble_qws_kernel", "ensemble_quad_kernel")), text
    else:
        assert text.startswith("ensemble_kernel") and "reference-order stepper" in text, text
    ens.run_mcmc(synth.make_walkers(a_true, 256, seed=1, scatter=0.04 * rel, oob_frac=0.0), 24, thin=4)
    chain, lnp = ens.chain, ens.lnprobability
    for k in range(chain.shape[2]):
        want = oracle_lib.lnpost_bins(consts, n_pts, lo, hi, np.ascontiguousarray(chain[:, :, k]))
        assert rel_err(lnp[:, :, k], want) <= TOL, (rel, rel_err(lnp[:, :, k], want))
    assert 0.05 < ens.acceptance_fraction.mean() < 0.9
    ens.close()
    store.close()


def test_one_ill_conditioned_bin_routes_the_store(cs_lib):
    consts, n_pts, a_true = synth.make_bins(4, 5, 30)
    bad, _, _ = synth.make_bins(1, 5, 30, rel_err=1e-4, seed_base=77)
    store = chisq.DeviceStore(consts, n_pts, synth.default_bounds(5), cs_lib=cs_lib)
    assert store.conditioning()[2]
    consts[2] = bad[0]
    mixed = chisq.DeviceStore(consts, n_pts, synth.default_bounds(5), cs_lib=cs_lib)
    kappa, limit, safe = mixed.conditioning()
    assert not safe and int(np.argmax(kappa)) == 2 and np.sum(kappa > limit) == 1
    store.close()
    mixed.close()
