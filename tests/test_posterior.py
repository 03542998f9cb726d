"""Posterior summaries: the restatement against the reference's own functions
(golden fixture), and the device reductions against the restatement on real chains."""
import os

import numpy as np
import pytest

from conftest import ROOT
from oracle import posterior_oracle


def _golden():
    z = np.load(os.path.join(ROOT, "tests", "golden", "golden_posterior.npz"))
    keys = sorted({k.rsplit("/", 1)[0] for k in z.files})
    return z, keys


def test_restatement_matches_reference_functions():
    z, keys = _golden()
    for key in keys:
        samples, ci, nb = z[key + "/samples"], float(z[key + "/ci"]), int(key.split("/")[1])
        ql = [(0.5 - ci / 2.0), 0.5, (0.5 + ci / 2.0)]
        points, peaks = posterior_oracle.calc_param_values(samples, ql, samples.shape[1], nb, ci)
        np.testing.assert_array_equal(np.array(points), z[key + "/points"])
        np.testing.assert_array_equal(np.array(peaks), z[key + "/peaks"])


@pytest.mark.gpu
def test_device_summaries_equal_numpy_on_the_chain(cs_lib):
    """Exact order statistics + numpy's interpolation => bit-identical to running the reference's
    summarisation on the chain pulled to the host."""
    from mimic_mda_b200 import chisq, synth
    from mimic_mda_b200.ensemble import DeviceEnsemble
    consts, counts, a_true = synth.make_bins(5, 4, 30, ragged=True)
    lo, hi = synth.default_bounds(4)
    store = chisq.DeviceStore(consts, counts, (lo, hi), cs_lib=cs_lib)
    ens = DeviceEnsemble(store, 64, seed=12)
    p0 = synth.make_walkers(a_true, 64, seed=1, oob_frac=0.0)
    ens.run_mcmc(p0, 240, thin=2)
    chain = ens.chain                                           # [B, W, kept, L]
    for burn, ci, nb in ((20, 0.682689492, 1000), (0, 0.9, 37), (100, 0.5, 5)):
        got = ens.summarize(burn_in=burn, confidence_interval=ci, num_bins=nb)
        points, peaks = posterior_oracle.summarize(chain, burn, ci, nb)
        np.testing.assert_array_equal(got["points"], points)
        np.testing.assert_array_equal(got["peaks"], peaks)
        np.testing.assert_allclose(got["acceptance_fraction"], ens.acceptance_fraction.mean(axis=1), rtol=1e-14)
    ens.close()
    store.close()


@pytest.mark.gpu
def test_device_summaries_on_golden_sample_sets(cs_lib):
    """The fixture's sample sets (repeats, a constant dimension, a peak in the first bin) pushed
    through the device reductions via a hand-built chain."""
    import ctypes as ct
    from mimic_mda_b200 import chisq, synth
    from mimic_mda_b200.ensemble import DeviceEnsemble
    z, keys = _golden()
    for key in keys:
        samples, ci, nb = z[key + "/samples"], float(z[key + "/ci"]), int(key.split("/")[1])
        n, n_l = samples.shape
        walkers = 2 * n_l + 2 + (n_l % 2) * 0
        walkers += walkers % 2
        kept = n // walkers
        used = samples[: kept * walkers]
        # a one-bin ensemble whose device chain we overwrite with the sample set: chain[kept][1][W][L]
        consts, counts, a_true = synth.make_bins(1, n_l, 12)
        store = chisq.DeviceStore(consts, counts, cs_lib=cs_lib)
        ens = DeviceEnsemble(store, walkers, seed=1)
        ens.set_state(synth.make_walkers(a_true, walkers, seed=1, oob_frac=0.0))
        ens.reserve_chain(kept)
        ens.advance(kept)
        d_chain = cs_lib.mdaEnsembleDeviceChain(ens.handle)
        block = np.ascontiguousarray(used.reshape(kept, 1, walkers, n_l))
        chisq.check(cs_lib, cs_lib.mdaMemcpyH2D(d_chain, block.ctypes.data, block.nbytes))
        got = ens.summarize(0, ci, nb)
        ql = [(0.5 - ci / 2.0), 0.5, (0.5 + ci / 2.0)]
        points, peaks = posterior_oracle.calc_param_values(used, ql, n_l, nb, ci)
        np.testing.assert_array_equal(got["points"][0], np.array(points))
        np.testing.assert_array_equal(got["peaks"][0], np.array(peaks))
        ens.close()
        store.close()
