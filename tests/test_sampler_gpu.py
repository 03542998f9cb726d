"""Posterior parity: the same stretch-move driver with the same seed, once on the oracle
and once on the CUDA path.  lnL agrees to ~1e-15, so accept/reject decisions -- and hence
whole chains -- are expected to be identical; percentile summaries are the fallback bar
(mda.py:1429-1431)."""
import numpy as np
import pytest

from mimic_mda_b200 import chisq, synth
from mimic_mda_b200.sampler import BatchedEnsembleSampler, GpuLogProb, summarize_chain

pytestmark = pytest.mark.gpu


def _run(log_prob, n_bins, walkers, n_l, p0, steps, seed):
    smp = BatchedEnsembleSampler(n_bins, walkers, n_l, log_prob, seeds=seed)
    smp.run_mcmc(p0, steps)
    return smp


@pytest.mark.parametrize("shape", [("C1", 1, 4, 30, 64, 400), ("multi", 6, 9, 60, 96, 120)])
def test_chains_match_oracle_driven_chains(cs_lib, oracle_lib, shape):
    _, n_bins, n_l, n_pts, walkers, steps = shape
    consts, counts, a_true = synth.make_bins(n_bins, n_l, n_pts, ragged=n_bins > 1)
    lo, hi = synth.default_bounds(n_l)
    p0 = synth.make_walkers(a_true, walkers, seed=17, oob_frac=0.0)
    ref = _run(lambda p: oracle_lib.lnpost_bins(consts, counts, lo, hi, p), n_bins, walkers, n_l, p0, steps, 1234)
    store = chisq.DeviceStore(consts, counts, (lo, hi), cs_lib=cs_lib)
    gpu = _run(GpuLogProb(store), n_bins, walkers, n_l, p0, steps, 1234)
    # identical decisions except for measure-zero ties: allow a vanishing fraction of steps to differ
    same = np.all(gpu.chain == ref.chain, axis=(1, 3))            # [bins, steps]
    assert same.mean() > 0.999
    s_ref = summarize_chain(ref.chain, burn_in=steps // 4)
    s_gpu = summarize_chain(gpu.chain, burn_in=steps // 4)
    for key in ("p16", "p50", "p84", "mean"):
        np.testing.assert_allclose(s_gpu[key], s_ref[key], rtol=0, atol=0.05 * np.max(s_ref["std"]) + 1e-12)
    np.testing.assert_allclose(gpu.acceptance_fraction, ref.acceptance_fraction, atol=0.01)
    store.close()


def test_shared_rng_mode_samples_the_same_posterior(cs_lib):
    consts, counts, a_true = synth.make_bins(4, 3, 40)
    store = chisq.DeviceStore(consts, counts, cs_lib=cs_lib)
    p0 = synth.make_walkers(a_true, 64, seed=3, oob_frac=0.0)
    a = BatchedEnsembleSampler(4, 64, 3, GpuLogProb(store), seeds=7, rng_mode="shared")
    b = BatchedEnsembleSampler(4, 64, 3, GpuLogProb(store), seeds=8, rng_mode="per_bin")
    a.run_mcmc(p0, 400)
    b.run_mcmc(p0, 400)
    sa, sb = summarize_chain(a.chain, 100), summarize_chain(b.chain, 100)
    assert np.all(np.abs(sa["mean"] - sb["mean"]) < 0.5 * sb["std"] + 1e-4)
    store.close()
