"""Device-resident ensemble stepping against its CPU restatement and
against the host-driven emcee-2-style sampler."""
import numpy as np
import pytest

from mimic_mda_b200 import chisq, synth
from mimic_mda_b200.ensemble import DeviceEnsemble
from mimic_mda_b200.sampler import BatchedEnsembleSampler, summarize_chain
from oracle import device_sampler_oracle as dso

pytestmark = pytest.mark.gpu


def _problem(n_bins, n_l, n_pts, walkers, ragged=True, seed=17):
    consts, counts, a_true = synth.make_bins(n_bins, n_l, n_pts, ragged=ragged)
    lo, hi = synth.default_bounds(n_l)
    p0 = synth.make_walkers(a_true, walkers, seed=seed, oob_frac=0.0)
    return consts, counts, lo, hi, p0, a_true


@pytest.mark.parametrize("shape", [(1, 4, 30, 64, 150), (5, 9, 60, 96, 60), (3, 5, 31, 512, 25), (2, 8, 41, 34, 80),
                                   (2, 1, 7, 8, 50), (1, 16, 24, 40, 30),
                                   # more than 512 walkers: the DFMA stepper with the state in shared memory
                                   (2, 6, 33, 1100, 8),
                                   # ensembles larger than shared memory (state in HBM/L2): the reference's own
                                   # 2500 walkers x 8 parameters (mda_config0.py:93,105), and a wider one
                                   (1, 8, 41, 2500, 12), (2, 9, 60, 3000, 6)])
# the triangular-factor stepper (the default) and the DMMA steppers (shapes none of those takes run on the DFMA stepper)
# "qws": the triangular-factor stepper warp-specialised (the default below two bins per SM), T in registers or in shared memory
@pytest.mark.parametrize("impl", ["quad", "ws", "mma", "mma-2cta", "qws-treg1", "qws-treg0"])
def test_device_chain_equals_cpu_restatement(cs_lib, oracle_lib, monkeypatch, shape, impl):
    monkeypatch.setenv("MDA_ENS_IMPL", impl.split("-")[0])
    if "-treg" in impl:
        monkeypatch.setenv("MDA_ENS_TREG", impl[-1])
    monkeypatch.setenv("MDA_ENS_DRAWW", "2" if impl.endswith("2cta") else "0")     # 2: without draw warps (two CTAs per SM)
    n_bins, n_l, n_pts, walkers, steps = shape
    consts, counts, lo, hi, p0, _ = _problem(n_bins, n_l, n_pts, walkers)
    want = dso.run(lambda p: oracle_lib.lnpost_bins(consts, counts, lo, hi, p), p0, steps, seed=0xC0FFEE1234)
    store = chisq.DeviceStore(consts, counts, (lo, hi), cs_lib=cs_lib)
    ens = DeviceEnsemble(store, walkers, seed=0xC0FFEE1234)
    pos, lnp = ens.run_mcmc(p0, steps)
    got = ens.chain.transpose(0, 2, 1, 3)                     # -> [B, kept, W, L]
    assert got.shape == want["chain"].shape == (n_bins, steps, walkers, n_l)
    np.testing.assert_allclose(got[:, :5], want["chain"][:, :5], rtol=1e-13, atol=0)   # same trajectory at all
    same = np.all(got == want["chain"], axis=(2, 3))
    assert same.mean() > 0.999, same.mean()                   # identical decisions (ties are measure zero)
    if same.all():
        np.testing.assert_array_equal(pos, want["pos"])
        np.testing.assert_array_equal(ens.naccepted, want["nacc"])
        np.testing.assert_allclose(lnp, want["lnp"], rtol=1e-12)
        np.testing.assert_allclose(ens.lnprobability.transpose(0, 2, 1), want["lnp_chain"], rtol=1e-12)
    assert np.all(got >= lo) and np.all(got <= hi)
    ens.close()
    store.close()


def test_split_runs_thinning_and_partition_invariance(cs_lib):
    consts, counts, lo, hi, p0, _ = _problem(6, 5, 30, 64)
    store = chisq.DeviceStore(consts, counts, (lo, hi), cs_lib=cs_lib)
    one = DeviceEnsemble(store, 64, seed=99)
    one.run_mcmc(p0, 120)
    full = one.chain.copy()
    # 50 + 70 steps == 120 steps
    two = DeviceEnsemble(store, 64, seed=99)
    two.set_state(p0)
    two.reserve_chain(120)
    two.advance(50)
    two.advance(70)
    np.testing.assert_array_equal(two.chain, full)
    assert two.iterations == 120 and two.kept == 120
    # thin = 3 keeps steps 3, 6, 9, ...
    thin = DeviceEnsemble(store, 64, seed=99)
    thin.run_mcmc(p0, 120, thin=3)
    np.testing.assert_array_equal(thin.chain, full[:, :, 2::3])
    # a different seed gives a different chain
    other = DeviceEnsemble(store, 64, seed=100)
    other.run_mcmc(p0, 120)
    assert not np.array_equal(other.chain, full)
    # bins sharded round-robin over two "ranks" reproduce the same chains via bin ids
    for r in range(2):
        idx = np.arange(r, 6, 2)
        st = chisq.DeviceStore(consts[idx], counts[idx], (lo, hi), cs_lib=cs_lib)
        part = DeviceEnsemble(st, 64, seed=99, bin_ids=idx)
        part.run_mcmc(p0[idx], 120)
        np.testing.assert_array_equal(part.chain, full[idx])
        part.close()
        st.close()
    for e in (one, two, thin, other):
        e.close()
    store.close()


@pytest.mark.parametrize("env", [{"MDA_ENS_GRID": "2", "MDA_ENS_CHUNK": "7"},        # 5 bins on 2 CTAs, 9 chunks: hand-over through HBM
                                 {"MDA_ENS_GRID": "3", "MDA_ENS_CHUNK": "1"},
                                 {"MDA_ENS_GRID": "4", "MDA_ENS_CHUNK": "16"},
                                 {"MDA_ENS_WT": "2"}, {"MDA_ENS_WT": "4"},
                                 {"MDA_ENS_IMPL": "mma"}, {"MDA_ENS_IMPL": "mma", "MDA_ENS_WT": "4"},
                                 {"MDA_ENS_IMPL": "mma", "MDA_ENS_DRAWW": "2"}, {"MDA_ENS_IMPL": "mma", "MDA_ENS_DRAWW": "2", "MDA_ENS_WT": "2"},
                                 {"MDA_ENS_IMPL": "mma", "MDA_ENS_GRID": "3", "MDA_ENS_CHUNK": "5"},
                                 {"MDA_ENS_IMPL": "dfma"}, {"MDA_ENS_IMPL": "quad"},
                                 {"MDA_ENS_IMPL": "quad", "MDA_ENS_GRID": "2", "MDA_ENS_CHUNK": "7"},    # units handed over through HBM
                                 {"MDA_ENS_IMPL": "quad", "MDA_ENS_GRID": "3", "MDA_ENS_CHUNK": "1"},
                                 # the warp-specialised stepper: a CTA per bin, and 5 bins' steps cut evenly over 2 / 3 / 4 CTAs
                                 {"MDA_ENS_IMPL": "qws"}, {"MDA_ENS_IMPL": "qws", "MDA_ENS_TREG": "0"},
                                 {"MDA_ENS_IMPL": "qws", "MDA_ENS_GRID": "2"}, {"MDA_ENS_IMPL": "qws", "MDA_ENS_GRID": "3", "MDA_ENS_TREG": "0"},
                                 {"MDA_ENS_IMPL": "qws", "MDA_ENS_GRID": "4"}])
def test_chain_does_not_depend_on_the_schedule(cs_lib, monkeypatch, env):
    """Units of (chunk of steps, bin) walked by fewer CTAs than bins, other tilings of the chi^2 work and the
    DFMA stepper: the chains are the same (identical for every schedule of the same kernel; the DFMA
    stepper sums chi^2 in another order, so it may differ in a measure-zero set of near-tie decisions)."""
    consts, counts, lo, hi, p0, _ = _problem(5, 9, 60, 128)
    store = chisq.DeviceStore(consts, counts, (lo, hi), cs_lib=cs_lib)
    monkeypatch.setenv("MDA_ENS_IMPL", "ws")
    ref = DeviceEnsemble(store, 128, seed=11)
    ref.run_mcmc(p0, 60, thin=2)
    want, want_state = ref.chain.copy(), ref.state()
    for k, v in env.items():
        monkeypatch.setenv(k, v)
    ens = DeviceEnsemble(store, 128, seed=11)
    ens.set_state(p0)
    ens.reserve_chain(30)
    ens.advance(25, thin=2)
    ens.advance(35, thin=2)
    if env.get("MDA_ENS_IMPL") == "dfma":
        same = np.all(ens.chain == want, axis=(1, 3))
        assert same.mean() > 0.999
        np.testing.assert_allclose(ens.chain[:, :, :3], want[:, :, :3], rtol=1e-13)
    elif env.get("MDA_ENS_IMPL") in ("quad", "qws"):
        # chi^2 from the triangular factor: log-probabilities agree to rounding, positions and decisions exactly
        np.testing.assert_array_equal(ens.chain, want)
        np.testing.assert_array_equal(ens.state()[0], want_state[0])
        np.testing.assert_allclose(ens.state()[1], want_state[1], rtol=1e-12)
        np.testing.assert_array_equal(ens.state()[2], want_state[2])
    else:
        np.testing.assert_array_equal(ens.chain, want)
        for a, b in zip(ens.state(), want_state):
            np.testing.assert_array_equal(a, b)
    ens.close()
    ref.close()
    store.close()


@pytest.mark.parametrize("shape", [(200, 9, 60, 512, 24),      # BASELINE.json configs[3] (C4): 148 CTAs walk 200 bins in chunks
                                   (200, 9, 60, 512, 130),     # the same with 16 hand-overs per bin
                                   (100, 9, 60, 512, 16),      # configs[2] (C3): one bin per SM
                                   (100, 5, 30, 256, 40),      # configs[1] (C2)
                                   (400, 9, 60, 512, 16)])     # more than two bins per SM: 296 CTAs of 8 warps walk
def test_full_size_steppers_agree_and_follow_the_restatement(cs_lib, oracle_lib, monkeypatch, shape):
    """BASELINE.json shapes at full size: every stepper gives bit-identical chains (size-independent property: a
    chain depends on seed, bin id and start only -- not on the kernel, the unit schedule or the hand-overs), three
    bins spot-checked against the CPU restatement, every sample inside the box prior, and the chain on 2 'ranks'
    (bins round-robin) is the chain on one."""
    n_bins, n_l, n_pts, walkers, steps = shape
    consts, counts, lo, hi, p0, _ = _problem(n_bins, n_l, n_pts, walkers)
    store = chisq.DeviceStore(consts, counts, (lo, hi), cs_lib=cs_lib)
    def run(impl, draww, grid="", chunk=""):
        monkeypatch.setenv("MDA_ENS_IMPL", impl)
        monkeypatch.setenv("MDA_ENS_DRAWW", draww)
        monkeypatch.setenv("MDA_ENS_GRID", grid)       # "" = the library's own schedule
        monkeypatch.setenv("MDA_ENS_CHUNK", chunk)
        ens = DeviceEnsemble(store, walkers, seed=20240229)
        ens.run_mcmc(p0, steps)
        out = (ens.chain.copy(), ens.lnprobability.copy(), ens.naccepted.copy())
        ens.close()
        return out

    ref = run("ws", "0")
    for impl, draww in (("ws", "0"), ("mma", "1"), ("mma", "2")):      # the first: run to run
        got = run(impl, draww)
        for a, b in zip(got, ref):
            np.testing.assert_array_equal(a, b, err_msg=impl + " " + draww)
        del got
    if steps <= 24:                        # many units per CTA, hand-over after every 3 steps
        for impl, draww in (("ws", "0"), ("mma", "1"), ("mma", "2")):
            got = run(impl, draww, grid="37", chunk="3")
            np.testing.assert_array_equal(got[0], ref[0], err_msg=impl + " " + draww + " 37 CTAs")
            del got
    got = run("quad", "0")                 # chi^2 from the triangular factor: same decisions, log-probabilities to rounding
    assert np.all(got[0] == ref[0], axis=(1, 3)).mean() > 0.999
    if np.array_equal(got[0], ref[0]):
        np.testing.assert_allclose(got[1], ref[1], rtol=1e-12)
        np.testing.assert_array_equal(got[2], ref[2])
    # the warp-specialised build of the same stepper (the default at these shapes: the library's own schedule -- C4: 148 CTAs
    # carry 200 bins' steps, some bins stepped by two CTAs in turn), and a coarser cut over 37 CTAs: identical to the plain one
    for grid in ("", "37"):
        qws = run("qws", "0", grid=grid)
        for a, b in zip(qws, got):
            np.testing.assert_array_equal(a, b, err_msg="qws grid=" + grid)
        del qws
    del got
    got = run("dfma", "0")                 # another summation order: may differ in a measure-zero set of near ties
    assert np.all(got[0] == ref[0], axis=(1, 3)).mean() > 0.999
    del got
    assert np.all(ref[0] >= lo) and np.all(ref[0] <= hi)
    acc = ref[2].sum() / (n_bins * walkers * steps)
    assert 0.15 < acc < 0.75, acc
    pick = np.array([0, 77, n_bins - 1])
    spot = min(steps, 24)
    want = dso.run(lambda p: oracle_lib.lnpost_bins(consts[pick], counts[pick], lo, hi, p), p0[pick], spot,
                   seed=20240229, bin_ids=pick)
    got = ref[0][pick].transpose(0, 2, 1, 3)[:, :spot]
    np.testing.assert_allclose(got[:, :3], want["chain"][:, :3], rtol=1e-13, atol=0)
    assert np.all(got == want["chain"], axis=(2, 3)).mean() > 0.99
    monkeypatch.setenv("MDA_ENS_IMPL", "ws")
    monkeypatch.setenv("MDA_ENS_DRAWW", "0")
    for r in range(2):
        idx = np.arange(r, n_bins, 2)
        st = chisq.DeviceStore(consts[idx], counts[idx], (lo, hi), cs_lib=cs_lib)
        part = DeviceEnsemble(st, walkers, seed=20240229, bin_ids=idx)
        part.run_mcmc(p0[idx], steps)
        np.testing.assert_array_equal(part.chain, ref[0][idx])
        part.close()
        st.close()
    store.close()


def test_posterior_agrees_with_host_driven_emcee_style_sampler(cs_lib, oracle_lib):
    """Different generators (Philox vs Mersenne Twister) => statistical parity:
    posterior means and quantiles within MCMC noise (mda.py:1429-1431)."""
    consts, counts, lo, hi, p0, a_true = _problem(4, 4, 40, 128, ragged=False)
    host = BatchedEnsembleSampler(4, 128, 4, lambda p: oracle_lib.lnpost_bins(consts, counts, lo, hi, p), seeds=5)
    host.run_mcmc(p0, 600)
    store = chisq.DeviceStore(consts, counts, (lo, hi), cs_lib=cs_lib)
    ens = DeviceEnsemble(store, 128, seed=5)
    ens.run_mcmc(p0, 600)
    s_h, s_d = summarize_chain(host.chain, 200), summarize_chain(ens.chain, 200)
    for key in ("mean", "p16", "p50", "p84"):
        assert np.all(np.abs(s_h[key] - s_d[key]) < 0.25 * s_h["std"] + 1e-5), key
    np.testing.assert_allclose(s_d["std"], s_h["std"], rtol=0.2)
    assert abs(ens.acceptance_fraction.mean() - host.acceptance_fraction.mean()) < 0.03
    ens.close()
    store.close()


def test_ensemble_argument_checks(cs_lib):
    consts, counts, lo, hi, p0, _ = _problem(2, 3, 10, 16)
    store = chisq.DeviceStore(consts, counts, (lo, hi), cs_lib=cs_lib)
    with pytest.raises(ValueError):
        DeviceEnsemble(store, 15)                  # odd
    with pytest.raises(ValueError):
        DeviceEnsemble(store, 4)                   # fewer than 2 * dim
    ens = DeviceEnsemble(store, 16, seed=1)
    with pytest.raises(chisq.MdaError):
        ens.advance(1)                             # no state yet
    bad = p0.copy()
    bad[1, 3, 0] = np.nan
    with pytest.raises(ValueError):
        ens.set_state(bad)
    ens.set_state(p0)
    ens.advance(5)                                 # no chain reserved: state only
    assert ens.kept == 0 and ens.iterations == 5
    pos, lnp, acc = ens.state()
    assert np.all(np.isfinite(lnp)) and acc.max() <= 5
    ens.close()
    store.close()


def test_run_to_host_streams_the_same_chain(cs_lib):
    """Chunked, double-buffered delivery to host memory == one resident run."""
    consts, counts, lo, hi, p0, _ = _problem(5, 5, 30, 64)
    store = chisq.DeviceStore(consts, counts, (lo, hi), cs_lib=cs_lib)
    ref = DeviceEnsemble(store, 64, seed=3)
    ref.run_mcmc(p0, 90, thin=3)
    want_chain, want_lnp = ref.chain.transpose(2, 0, 1, 3), ref.lnprobability.transpose(2, 0, 1)   # step-major
    for chunk in (0, 3, 12, 33, 90, 500):
        ens = DeviceEnsemble(store, 64, seed=3)
        ens.set_state(p0)
        chain, lnp = ens.run_to_host(90, thin=3, chunk_steps=chunk)
        np.testing.assert_array_equal(chain, want_chain)
        np.testing.assert_array_equal(lnp, want_lnp)
        np.testing.assert_array_equal(ens.state()[0], ref.state()[0])
        ens.close()
    ref.close()
    store.close()


def test_balanced_schedule_of_the_default_stepper_with_thinning_and_split_runs(cs_lib, monkeypatch):
    """Just above one bin per SM the default stepper cuts the bins' step ranges evenly over one CTA per SM (some bins are
    stepped by two CTAs, one after the other): the chain, thinned and advanced in two calls, equals the one a CTA per bin
    produces (the plain triangular-factor kernel, MDA_ENS_IMPL=quad with its own walk switched off)."""
    n_bins = 200
    consts, counts, lo, hi, p0, _ = _problem(n_bins, 4, 16, 32)
    store = chisq.DeviceStore(consts, counts, (lo, hi), cs_lib=cs_lib)
    monkeypatch.setenv("MDA_ENS_IMPL", "quad")
    monkeypatch.setenv("MDA_ENS_QUAD_WALK", "0")
    ref = DeviceEnsemble(store, 32, seed=31)
    assert "200 CTAs" in ref.describe(90)
    ref.run_mcmc(p0, 90, thin=3)
    want, want_state = ref.chain.copy(), ref.state()
    monkeypatch.delenv("MDA_ENS_IMPL")
    ens = DeviceEnsemble(store, 32, seed=31)
    text = ens.describe(90)
    assert text.startswith("ensemble_qws_kernel") and "CTAs for 200 bins" in text and "200 CTAs for" not in text, text
    ens.set_state(p0)
    ens.reserve_chain(30)
    ens.advance(40, thin=3)
    ens.advance(50, thin=3)
    np.testing.assert_array_equal(ens.chain, want)
    for a, b in zip(ens.state(), want_state):
        np.testing.assert_array_equal(a, b)
    ens.close()
    ref.close()
    store.close()


def test_restart_on_the_same_ensemble_draws_a_fresh_stream_and_store_changes_are_noticed(cs_lib, oracle_lib):
    """burn-in, then run_mcmc again on the same ensemble: the second run must not replay the first run's draws (its Philox
    key is the seed advanced by the restart count); a store modified after set_state invalidates the stored
    log-probabilities; a store freed before its ensemble stays alive until the ensemble goes."""
    consts, counts, lo, hi, p0, _ = _problem(3, 4, 12, 32)
    store = chisq.DeviceStore(consts, counts, (lo, hi), cs_lib=cs_lib)
    ens = DeviceEnsemble(store, 32, seed=11)
    ens.run_mcmc(p0, 12)
    first = ens.chain.copy()
    ens.run_mcmc(p0, 12)
    second = ens.chain.copy()
    assert not np.array_equal(first, second)
    fn = lambda p: oracle_lib.lnpost_bins(consts, counts, lo, hi, p)
    np.testing.assert_allclose(first.transpose(0, 2, 1, 3), dso.run(fn, p0, 12, seed=11)["chain"], rtol=1e-12, atol=0)
    np.testing.assert_allclose(second.transpose(0, 2, 1, 3), dso.run(fn, p0, 12, seed=11, epoch=1)["chain"], rtol=1e-12, atol=0)
    # the store changes under the ensemble
    store.set_bounds((lo, hi * 0.9))
    with pytest.raises(chisq.MdaError):
        ens.advance(1)
    ens.set_state(p0 * 0.9)
    ens.advance(2)
    # freeing the store first is safe
    store.close()
    pos, lnp, _ = ens.state()
    assert np.all(np.isfinite(pos))
    ens.close()
