"""Integrated autocorrelation time (mda.py:1304-1308 -> emcee 2.x get_autocorr_time).

emcee is a third-party dependency absent from the reference tree and from this image, so the oracle
(oracle/autocorr_oracle.py) restates its published algorithm -- parity unpinned, see its header.  CPU tests
check the restatement against an independent direct-sum evaluation and a closed-form case; GPU tests check the
device reductions (ensemble mean, autocorrelation sums, window search behind the C ABI) against the restatement
run on the same chain pulled to the host.  Tolerance: the device adds the lag products directly, the oracle goes
through a zero-padded FFT -- both are fp64 evaluations of the same sums, 1e-9 relative on tau is ample; the
accepted window M is an integer and must agree exactly.
"""
import numpy as np
import pytest

from oracle import autocorr_oracle as ao

TAU_RTOL = 1e-9


def _ar1(phi, n, dim, seed):
    rng = np.random.default_rng(seed)
    x = np.zeros((n, dim))
    e = rng.standard_normal((n, dim))
    for t in range(1, n):
        x[t] = phi * x[t - 1] + e[t]
    return x


def test_fft_function_is_the_linear_autocorrelation():
    x = _ar1(0.8, 700, 3, 1)
    np.testing.assert_allclose(ao.function(x)[:120], ao.acf_direct(x, 120), rtol=0, atol=1e-12)
    # fast: the first 512 samples only
    np.testing.assert_allclose(ao.function(x, fast=True)[:120], ao.acf_direct(x[:512], 120), rtol=0, atol=1e-12)


def test_integrated_time_of_ar1_series():
    phi = 0.7                                            # tau = (1 + phi) / (1 - phi) = 5.67
    x = _ar1(phi, 200000, 2, 2)
    tau, M = ao.integrated_time(x, c=10, full_output=True)
    assert tau.shape == (2,) and M > 10 * tau.max()
    np.testing.assert_allclose(tau, (1 + phi) / (1 - phi), rtol=0.08)
    # one-dimensional input takes the other branch of the same loop
    np.testing.assert_allclose(ao.integrated_time(x[:, 0], c=10), tau[0], rtol=1e-12)


def test_short_chain_raises():
    x = _ar1(0.9, 150, 2, 3)
    with pytest.raises(ao.AutocorrError):
        ao.integrated_time(x, c=10)                      # c * low >= size
    with pytest.raises(ao.AutocorrError):
        ao.integrated_time(_ar1(0.999, 4000, 2, 4), c=10)   # never converges inside the window range


def _run_ensemble(cs_lib, bins, n_l, walkers, steps, thin=1, seed=5):
    from mimic_mda_b200 import chisq, synth
    from mimic_mda_b200.ensemble import DeviceEnsemble
    consts, counts, a_true = synth.make_bins(bins, n_l, 30, ragged=True)
    store = chisq.DeviceStore(consts, counts, synth.default_bounds(n_l), cs_lib=cs_lib)
    ens = DeviceEnsemble(store, walkers, seed=seed)
    ens.run_mcmc(synth.make_walkers(a_true, walkers, seed=1, oob_frac=0.0), steps, thin=thin)
    return store, ens


def _oracle_tau(chain_b, **kw):
    try:
        return ao.get_autocorr_time(chain_b, full_output=True, **kw)
    except ao.AutocorrError:
        return None, 0


@pytest.mark.gpu
@pytest.mark.parametrize("n_l,walkers,steps,thin", [(4, 64, 1500, 1), (9, 32, 2049, 1), (5, 48, 1200, 2)])
def test_device_autocorr_time_equals_restatement(cs_lib, n_l, walkers, steps, thin):
    store, ens = _run_ensemble(cs_lib, 5, n_l, walkers, steps, thin)
    chain = ens.chain                                    # [B, W, kept, L]
    for kw in (dict(c=3.0, fast=True), dict(c=3.0, fast=False), dict(c=2.0, fast=False, low=5, step=3),
               dict(c=1.5, fast=True, high=200), dict(c=10, fast=False)):
        got = ens.get_autocorr_time(quiet=True, **kw)
        win = ens.autocorr_window.copy()
        for b in range(chain.shape[0]):
            tau, M = _oracle_tau(chain[b], **kw)
            assert win[b] == M, (kw, b, win[b], M)
            if M:
                np.testing.assert_allclose(got[b], tau, rtol=TAU_RTOL)
            else:
                assert np.all(np.isnan(got[b]))
    ens.close()
    store.close()


@pytest.mark.gpu
def test_device_autocorr_raises_like_emcee(cs_lib):
    from mimic_mda_b200.ensemble import AutocorrError
    store, ens = _run_ensemble(cs_lib, 2, 4, 32, 150)
    with pytest.raises(AutocorrError):
        ens.get_autocorr_time(c=10)                      # c * low >= kept / 2
    assert np.all(np.isnan(ens.get_autocorr_time(c=10, quiet=True)))
    with pytest.raises(Exception):
        ens.get_autocorr_time(c=-1.0)
    ens.close()
    store.close()


@pytest.mark.gpu
def test_device_autocorr_on_a_known_series(cs_lib):
    """An AR(1) chain written into the device chain: the device estimate is the closed-form time."""
    from mimic_mda_b200 import chisq
    store, ens = _run_ensemble(cs_lib, 3, 4, 16, 40000)
    kept, B, W, L = ens.kept, 3, 16, 4
    phi = 0.6
    series = _ar1(phi, kept, B * L, 9).reshape(kept, B, 1, L)
    noise = np.random.default_rng(3).standard_normal((kept, B, W, L))
    block = np.ascontiguousarray(series + noise - noise.mean(axis=2, keepdims=True))    # walkers scatter around the series
    chisq.check(cs_lib, cs_lib.mdaMemcpyH2D(cs_lib.mdaEnsembleDeviceChain(ens.handle), block.ctypes.data, block.nbytes))
    ens._chain = None
    got = ens.get_autocorr_time(c=10, fast=True)
    np.testing.assert_allclose(got, (1 + phi) / (1 - phi), rtol=0.15)
    for b in range(B):
        tau, M = ao.get_autocorr_time(block[:, b].transpose(1, 0, 2), c=10, fast=True, full_output=True)
        assert ens.autocorr_window[b] == M
        np.testing.assert_allclose(got[b], tau, rtol=TAU_RTOL)
    ens.close()
    store.close()


@pytest.mark.gpu
def test_pipeline_returns_autocorr_times_and_dumps_chains(cs_lib, tmp_path):
    from mimic_mda_b200 import chisq, fits, pipeline, synth
    consts, counts, a_true = synth.make_bins(3, 4, 30)
    store = chisq.DeviceStore(consts, counts, synth.default_bounds(4), cs_lib=cs_lib)
    starts = fits.calc_start_params([[0.1, 0.4]] * 4)
    cfg = {"Number of Walkers": 64, "Sample Points": 2048, "Burn-in Points": 256, "Number Walker Generators": 4,
           "Calc AutoCorr": True, "ACorr WindSize": 3.0, "ACorr Use FFT": True,
           "Save Chain Data": True, "Chain Directory": str(tmp_path / "chains"), "Target A": 116}
    energies = [8.5, 9.5, 10.5]
    results, ens = pipeline.fit_and_mcmc_all(store, starts, cfg, seed=3, energies=energies)
    chain = ens.chain
    # the reference's per-bin dump (mda.py:1244-1249): arr_0 = sampler.chain, [walker, step, dim]
    for b, energy in enumerate(energies):
        with np.load(str(tmp_path / "chains" / ("A116_chain_E%5.2f.npz" % energy))) as z:
            np.testing.assert_array_equal(z["arr_0"], chain[b])
    for b, (values, (acor, chis, acc)) in enumerate(results):
        np.testing.assert_allclose(acor, ao.get_autocorr_time(chain[b], c=3.0, fast=True), rtol=TAU_RTOL)
        assert len(acor) == 4 and max(acor) > 1.0        # what write_fit_diagnostics divides by (mda.py:269)
    ens.close()
    store.close()
