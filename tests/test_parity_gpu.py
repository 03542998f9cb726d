"""Parity of the CUDA path with the oracle, through the C ABI, on a B200.

Bar (BASELINE.json north_star): log-likelihoods within 1e-12 relative of the reference
C library on the same inputs; -inf / NaN positions identical.  The oracle here is
(i) the frozen outputs of the unmodified reference library (tests/golden) and (ii) the
C restatement evaluated on seeded inputs, up to BASELINE's full sizes.
"""
import ctypes as ct

import numpy as np
import pytest

from conftest import rel_err, same_special
from mimic_mda_b200 import chisq, synth

pytestmark = pytest.mark.gpu

TOL = 1e-12                      # fp64, relative -- the tolerance north_star states
_DP = ct.POINTER(ct.c_double)


def _ptr(a):
    return a.ctypes.data_as(_DP)


# ---- the reference's seven symbols ---------------------------------------------------------
def test_reference_tester_sequence(cs_lib, golden):
    """C/ChiSqTester.c:10-89 / so_py_tester.py:28-86, call for call, against the CUDA library."""
    st = cs_lib.makeMdaStruct(10, 3)
    assert st
    cs_lib.freeMdaStruct(st)
    st = cs_lib.makeMdaStruct(10, 3)
    x = np.array([0.0, 0.1, 0.2, 0.3, 0.4, 0.5, 0.6, 0.7, 0.8, 0.9])
    cs_lib.setMdaData(st, _ptr(x))
    for i in range(3):
        cs_lib.setMdaDist(st, i, _ptr(x))
    a = np.array([0.2, 0.2, 0.2])
    kat = golden.case("kat")
    chi = [cs_lib.calculateChi(st, _ptr(a)) for _ in range(31)]
    assert all(abs(c - 0.456) < 1e-15 for c in chi) and abs(chi[0] - float(kat["chi"])) < 1e-15
    lnl = [cs_lib.calculateLnLiklihood(st, _ptr(a)) for _ in range(31)]
    assert all(abs(v + 0.228) < 1e-15 for v in lnl)
    resid = np.arange(10.0)
    for _ in range(31):
        v = cs_lib.calculateLnLiklihoodResids(st, _ptr(a), _ptr(resid))
    assert abs(v + 0.228) < 1e-15
    np.testing.assert_allclose(resid, kat["resid"], rtol=0, atol=3e-16)
    cs_lib.freeMdaStruct(st)
    assert cs_lib.mdaLastError() == 0, cs_lib.mdaLastErrorString()


def test_scalar_entry_points_match_reference_outputs(cs_lib, golden):
    """make_calc_struct + ln_post_prob per walker (mda.py:1742-1780, 1540-1578)."""
    for name in golden.cases:
        g = golden.case(name)
        n_l = g["params"].shape[2]
        bounds = [(float(g["lo"][i]), float(g["hi"][i])) for i in range(n_l)]
        for b in range(g["consts"].shape[0]):
            n = int(g["n_pts"][b])
            st = chisq.make_calc_struct(cs_lib, g["consts"][b, 0, :n], [g["consts"][b, 1 + l, :n] for l in range(n_l)])
            n_w = min(12, g["params"].shape[1])
            lnp = np.array([chisq.ln_post_prob(g["params"][b, w], cs_lib, st, bounds) for w in range(n_w)])
            chi = np.array([chisq.call_chi_sq(g["params"][b, w], cs_lib, st) for w in range(n_w)])
            assert same_special(lnp, g["lnpost"][b, :n_w]) and rel_err(lnp, g["lnpost"][b, :n_w]) <= TOL, name
            assert same_special(chi, g["chi"][b, :n_w]) and rel_err(chi, g["chi"][b, :n_w]) <= TOL, name
            scratch = np.zeros(max(n, 1))
            v = cs_lib.calculateLnLiklihoodResids(st, _ptr(g["params"][b, 0]), _ptr(scratch))
            assert rel_err(np.array([v]), g["lnl_resid"][b, :1]) <= TOL, name
            np.testing.assert_allclose(scratch[:n], g["resid"][b, 0, :n], rtol=1e-12, atol=1e-12)
            cs_lib.freeMdaStruct(st)
    assert cs_lib.mdaLastError() == 0, cs_lib.mdaLastErrorString()


def test_dist_index_guard_and_overwrite(cs_lib):
    x = np.linspace(0.1, 1.0, 12)
    st = chisq.make_calc_struct(cs_lib, x, [x, 2 * x])
    a = np.array([0.25, 0.125])
    before = cs_lib.calculateChi(st, _ptr(a))
    junk = np.full(12, 1e9)
    cs_lib.setMdaDist(st, 2, _ptr(junk))                 # >= numLs: ignored (C/ChiSquare.c:88-89)
    cs_lib.setMdaDist(st, 99, _ptr(junk))
    assert cs_lib.calculateChi(st, _ptr(a)) == before
    cs_lib.setMdaDist(st, 1, _ptr(x))                    # a later set replaces the row
    assert abs(cs_lib.calculateChi(st, _ptr(a)) - np.sum((x - 0.375 * x) ** 2)) < 1e-14
    cs_lib.freeMdaStruct(st)


# ---- the batched entry point --------------------------------------------------------------
def test_batched_matches_reference_outputs(cs_lib, golden):
    for name in golden.cases:
        g = golden.case(name)
        store = chisq.DeviceStore(g["consts"], g["n_pts"], (g["lo"], g["hi"]), cs_lib=cs_lib)
        lnp = store.lnpost(g["params"])
        chi = store.chi(g["params"])
        assert same_special(lnp, g["lnpost"]), name
        assert rel_err(lnp, g["lnpost"]) <= TOL, (name, rel_err(lnp, g["lnpost"]))
        assert same_special(chi, g["chi"]), name
        assert rel_err(chi, g["chi"]) <= TOL, (name, rel_err(chi, g["chi"]))
        store.close()


@pytest.mark.parametrize("name", ["C2", "C3", "C4"])
def test_full_size_configs_against_oracle(cs_lib, oracle_lib, name):
    """BASELINE.json's configs at full size (C4: 200 bins x 512 walkers x 9 x 60)."""
    cfg = synth.CONFIGS[name]
    consts, n_pts, a_true = synth.make_bins(cfg["bins"], cfg["n_l"], cfg["n_pts"], ragged=(name == "C4"))
    lo, hi = synth.default_bounds(cfg["n_l"])
    params = synth.make_walkers(a_true, cfg["walkers"], seed=21)
    want = oracle_lib.lnpost_bins(consts, n_pts, lo, hi, params)
    store = chisq.DeviceStore(consts, n_pts, (lo, hi), cs_lib=cs_lib)
    got = store.lnpost(params)
    assert same_special(got, want) and np.isneginf(want).sum() > 0
    assert rel_err(got, want) <= TOL, rel_err(got, want)
    # size-independent property: a (bin, walker) result does not depend on the batch it is in
    half = np.ascontiguousarray(params[:, : cfg["walkers"] // 2])
    np.testing.assert_array_equal(store.lnpost(half), got[:, : cfg["walkers"] // 2])
    one = np.ascontiguousarray(params[:, 5:6])
    np.testing.assert_array_equal(store.lnpost(one), got[:, 5:6])
    store.close()


def test_every_tiling_is_bit_identical(cs_lib, oracle_lib, monkeypatch):
    """The register-tiled DFMA kernel (a forced tiling selects it): the reference's operation order in every tiling.
    The DMMA kernel sums in another order: within 1e-14 of it, and still inside the oracle tolerance."""
    monkeypatch.setenv("MDA_BATCH_IMPL", "mma")
    consts, n_pts, a_true = synth.make_bins(7, 9, 60, ragged=True)
    lo, hi = synth.default_bounds(9)
    params = synth.make_walkers(a_true, 301, seed=5)     # not a multiple of any tile
    want = oracle_lib.lnpost_bins(consts, n_pts, lo, hi, params)
    store = chisq.DeviceStore(consts, n_pts, (lo, hi), cs_lib=cs_lib)
    assert store.get_tile(301)[:2] == (128, 0)           # the DMMA kernel
    default = store.lnpost(params)
    assert same_special(default, want) and rel_err(default, want) <= TOL
    store.set_tile(64, 2)
    base = store.lnpost(params)
    assert same_special(base, want) and rel_err(base, want) <= TOL
    assert rel_err(default, base) <= 1e-14
    for threads in (32, 64, 128, 256):
        for wpt in (1, 2, 4):
            store.set_tile(threads, wpt)
            assert store.get_tile(301)[:2] == (threads, wpt)
            np.testing.assert_array_equal(store.lnpost(params), base)
    store.set_tile(0, 0)
    np.testing.assert_array_equal(store.lnpost(params), default)
    monkeypatch.setenv("MDA_BATCH_IMPL", "")
    assert store.get_tile(301)[1] != 0                   # 60 angles: the DFMA kernel is the default
    np.testing.assert_array_equal(store.lnpost(params), base)
    store.close()


@pytest.mark.parametrize("shape", [(3, 4, 18, 33), (2, 9, 60, 1), (5, 12, 256, 200), (2, 16, 24, 95), (1, 7, 41, 700),
                                   (150, 5, 30, 128), (4, 8, 61, 77), (200, 9, 60, 256), (3, 1, 17, 40), (2, 3, 100, 64), (2, 8, 350, 9)])
def test_angle_split_kernel_is_bit_identical_to_the_tile_kernel(cs_lib, oracle_lib, monkeypatch, shape):
    """lnpost_split_kernel (G lanes per walker, MDA_BATCH_IMPL=split) hands chi^2 from lane to lane in the
    reference's order: the same bits as the thread-per-walker kernel whatever the shape -- ragged and odd angle counts,
    NaN and infinite parameters, chi^2 and log-posterior modes -- and within the oracle tolerance."""
    n_bins, n_l, n_pts, walkers = shape
    consts, n_pts_arr, a_true = synth.make_bins(n_bins, n_l, n_pts, ragged=n_pts > 8)
    lo, hi = synth.default_bounds(n_l)
    params = synth.make_walkers(a_true, walkers, seed=12)
    if walkers > 4:
        params[0, 1, 0] = np.nan
        params[-1, 2, n_l - 1] = np.inf                    # chi^2 = inf, not NaN: pads are never multiplied
        params[0, 3, 0] = -np.inf
    store = chisq.DeviceStore(consts, n_pts_arr, (lo, hi), cs_lib=cs_lib)
    monkeypatch.setenv("MDA_BATCH_IMPL", "split")
    threads, wpt, grid = store.get_tile(walkers)
    assert threads == 256 and wpt < 0 and (-wpt) & (-wpt - 1) == 0 and -wpt * 16 >= n_pts     # -G lanes per walker, 16 angles each
    got_ln, got_chi = store.lnpost(params), store.chi(params)
    monkeypatch.setenv("MDA_BATCH_IMPL", "dfma")
    assert store.get_tile(walkers)[1] > 0
    tile_ln, tile_chi = store.lnpost(params), store.chi(params)
    np.testing.assert_array_equal(got_ln, tile_ln)
    np.testing.assert_array_equal(got_chi, tile_chi)
    want = oracle_lib.lnpost_bins(consts, n_pts_arr, lo, hi, params)
    assert same_special(got_ln, want) and rel_err(got_ln, want) <= TOL
    # it is opt-in: measured slower than the tile kernel at every size (profiles/experiments/r02_split_kernel.md)
    monkeypatch.delenv("MDA_BATCH_IMPL")
    assert store.get_tile(walkers)[1] >= 0
    store.close()


@pytest.mark.parametrize("shape", [(3, 4, 16, 33), (2, 9, 60, 1), (5, 12, 256, 200), (2, 16, 24, 95), (1, 7, 41, 8192),
                                   (150, 5, 30, 128), (4, 8, 60, 2500)])
@pytest.mark.parametrize("mode", ["lnpost", "chi"])
def test_dmma_batched_kernel_against_oracle_and_dfma_kernel(cs_lib, oracle_lib, monkeypatch, shape, mode):
    """lnpost_mma_kernel at ragged walker counts, every numL % 4, many angle tiles, several passes per warp: within
    the oracle tolerance, -inf / NaN exactly where the reference has them, within 1e-14 of the DFMA kernel, and a
    walker's value does not depend on the batch it is evaluated in."""
    monkeypatch.setenv("MDA_BATCH_IMPL", "mma")
    n_bins, n_l, n_pts, walkers = shape
    consts, n_pts_arr, a_true = synth.make_bins(n_bins, n_l, n_pts, ragged=True)
    lo, hi = synth.default_bounds(n_l)
    params = synth.make_walkers(a_true, walkers, seed=9)
    if walkers > 4:
        params[0, 3, 0] = np.nan                          # a NaN passes the box prior and poisons the sum (mda.py:1574-1576)
        params[-1, 1, n_l - 1] = hi[n_l - 1]              # on the inclusive bound: inside
    store = chisq.DeviceStore(consts, n_pts_arr, (lo, hi), cs_lib=cs_lib)
    fn = store.lnpost if mode == "lnpost" else store.chi
    if mode == "lnpost":
        want = oracle_lib.lnpost_bins(consts, n_pts_arr, lo, hi, params)
    else:                                                 # chi^2 has no prior (C/ChiSquare.c:114-172)
        want = np.stack([oracle_lib.chi_batch(consts[b, 0, :n_pts_arr[b]], consts[b, 1:, :n_pts_arr[b]], params[b])
                         for b in range(n_bins)])
    got = fn(params)
    assert same_special(got, want), mode
    assert rel_err(got, want) <= TOL, rel_err(got, want)
    if walkers > 40:
        part = np.ascontiguousarray(params[:, 7:40])
        np.testing.assert_array_equal(fn(part), got[:, 7:40])
    store.set_tile(64, 2)
    ref = fn(params)
    assert same_special(got, ref) and rel_err(got, ref) <= 1e-14
    store.close()


@pytest.mark.parametrize("n_l", list(range(1, 17)) + [20, 33])
def test_every_dist_count(cs_lib, oracle_lib, n_l):
    """Every register-tiled instantiation (L = 1..16) and the generic kernel beyond."""
    consts, n_pts, a_true = synth.make_bins(3, n_l, 23, ragged=True)
    lo, hi = synth.default_bounds(n_l)
    params = synth.make_walkers(a_true, 70, seed=n_l)
    want = oracle_lib.lnpost_bins(consts, n_pts, lo, hi, params)
    store = chisq.DeviceStore(consts, n_pts, (lo, hi), cs_lib=cs_lib)
    for wpt in (1, 2, 4):
        store.set_tile(64, wpt)
        got = store.lnpost(params)
        assert same_special(got, want) and rel_err(got, want) <= TOL, (n_l, wpt)
    store.close()


def test_edge_shapes(cs_lib, oracle_lib):
    # empty bin (numPts = 0 -> lnL = -0.0), single angle, single walker, odd angle counts
    n_l = 3
    consts = np.zeros((4, 1 + n_l, 8))
    n_pts = np.array([0, 1, 7, 8], dtype=np.int32)
    rng = np.random.default_rng(3)
    for b in range(4):
        consts[b, :, : n_pts[b]] = rng.uniform(0.5, 2.0, size=(1 + n_l, n_pts[b]))
    lo, hi = synth.default_bounds(n_l)
    for walkers in (1, 2, 33):
        params = rng.uniform(0.0, 1.0, size=(4, walkers, n_l))
        want = oracle_lib.lnpost_bins(consts, n_pts, lo, hi, params)
        store = chisq.DeviceStore(consts, n_pts, (lo, hi), cs_lib=cs_lib)
        got = store.lnpost(params)
        assert rel_err(got, want) <= TOL
        assert np.all(got[0] == 0.0) and np.all(np.signbit(got[0]))
        store.close()
    # zero walkers / zero bins are no-ops
    store = chisq.DeviceStore(consts, n_pts, (lo, hi), cs_lib=cs_lib)
    assert store.lnpost(np.zeros((4, 0, n_l))).shape == (4, 0)
    store.close()


def test_angle_chunking_for_large_bins(cs_lib, oracle_lib):
    """A bin whose constants exceed shared memory is streamed in chunks of angles."""
    n_l, n = 6, 20001
    consts, n_pts, a_true = synth.make_bins(2, n_l, n)
    lo, hi = synth.default_bounds(n_l)
    params = synth.make_walkers(a_true, 40, seed=8)
    want = oracle_lib.lnpost_bins(consts, n_pts, lo, hi, params)
    store = chisq.DeviceStore(consts, n_pts, (lo, hi), cs_lib=cs_lib)
    got = store.lnpost(params)
    assert same_special(got, want) and rel_err(got, want) <= TOL
    store.close()


def test_custom_bounds_and_rebinding(cs_lib, oracle_lib):
    consts, n_pts, a_true = synth.make_bins(3, 4, 30)
    lo = np.array([0.0, 0.05, 0.0, 0.1])
    hi = np.array([0.5, 2.0, 0.25, 1.0 / 0.7])           # 1/EWSR_l (mda.py:1124)
    params = synth.make_walkers(a_true, 64, seed=4, oob_frac=0.0)
    store = chisq.DeviceStore(consts, n_pts, [(l, h) for l, h in zip(lo, hi)], cs_lib=cs_lib)
    want = oracle_lib.lnpost_bins(consts, n_pts, lo, hi, params)
    got = store.lnpost(params)
    assert same_special(got, want) and np.isneginf(want).any() and rel_err(got, want) <= TOL
    # staged (graph) path must see a later change of bounds
    buf, out = store.staging(64)
    buf[...] = params
    store.lnpost_staged(64)
    np.testing.assert_array_equal(out, got)
    lo2, hi2 = synth.default_bounds(4)
    store.set_bounds((lo2, hi2))
    store.lnpost_staged(64)
    want2 = oracle_lib.lnpost_bins(consts, n_pts, lo2, hi2, params)
    assert same_special(out, want2) and rel_err(out, want2) <= TOL
    assert not np.array_equal(np.isneginf(want2), np.isneginf(want))
    np.testing.assert_array_equal(store.lnpost(params), out)
    store.close()


def test_staged_graph_path_equals_host_path(cs_lib):
    consts, n_pts, a_true = synth.make_bins(12, 5, 30)
    store = chisq.DeviceStore(consts, n_pts, cs_lib=cs_lib)
    for walkers in (128, 32, 256):                        # growing re-allocates the pinned staging
        params = synth.make_walkers(a_true, walkers, seed=walkers)
        direct = store.lnpost(params)
        buf, out = store.staging(walkers)
        for rep in range(3):
            buf[...] = params
            out[...] = 0.0
            store.lnpost_staged(walkers)
            np.testing.assert_array_equal(out, direct)
        buf[...] = params
        store.chi_staged(walkers)
        np.testing.assert_array_equal(out, store.chi(params))
    store.close()


def test_device_pointer_path(cs_lib):
    consts, n_pts, a_true = synth.make_bins(9, 9, 60)
    store = chisq.DeviceStore(consts, n_pts, cs_lib=cs_lib)
    params = synth.make_walkers(a_true, 256, seed=1)
    want = store.lnpost(params)
    d_p = cs_lib.mdaDeviceAlloc(params.nbytes)
    d_o = cs_lib.mdaDeviceAlloc(want.nbytes)
    assert d_p and d_o
    chisq.check(cs_lib, cs_lib.mdaMemcpyH2D(d_p, params.ctypes.data, params.nbytes))
    before = cs_lib.mdaKernelLaunchCount()
    store.lnpost_device(256, d_p, d_o)
    store.lnpost_device_many(256, [d_p], [d_o], 3)
    store.synchronize()
    assert cs_lib.mdaKernelLaunchCount() - before == 4
    got = np.empty_like(want)
    chisq.check(cs_lib, cs_lib.mdaMemcpyD2H(got.ctypes.data, d_o, got.nbytes))
    np.testing.assert_array_equal(got, want)
    cs_lib.mdaDeviceFree(d_p)
    cs_lib.mdaDeviceFree(d_o)
    store.close()


def test_store_from_legacy_structs(cs_lib, golden):
    """Handles filled the reference way can be gathered into one device store."""
    g = golden.case("C4")
    n_l = g["params"].shape[2]
    structs = [chisq.make_calc_struct(cs_lib, g["consts"][b, 0, : g["n_pts"][b]],
                                      [g["consts"][b, 1 + l, : g["n_pts"][b]] for l in range(n_l)])
               for b in range(g["consts"].shape[0])]
    store = chisq.DeviceStore.from_structs(cs_lib, structs, n_l, int(g["n_pts"].max()), (g["lo"], g["hi"]))
    got = store.lnpost(g["params"])
    assert same_special(got, g["lnpost"]) and rel_err(got, g["lnpost"]) <= TOL
    for st in structs:
        cs_lib.freeMdaStruct(st)
    store.close()


def test_error_channel_reports_misuse(cs_lib):
    cs_lib.mdaClearError()
    consts, n_pts, _ = synth.make_bins(2, 3, 10)
    store = chisq.DeviceStore(consts, n_pts, cs_lib=cs_lib)
    with pytest.raises(ValueError):
        store.lnpost(np.zeros((3, 4, 3)))
    assert cs_lib.mdaBatchLnPost(store.handle, 4, None, None) == 3
    assert b"NULL" in cs_lib.mdaLastErrorString()
    assert cs_lib.mdaStoreSetBin(store.handle, 5, 10, None, None) == 3
    assert cs_lib.mdaBatchLnPostStaged(store.handle, 10 ** 6) == 5        # staging never requested
    st = cs_lib.makeMdaStruct(4, 2)
    assert cs_lib.mdaBatchLnPost(st, 1, None, None) == 3                  # wrong handle type
    assert np.isnan(cs_lib.calculateChi(store.handle, _ptr(np.zeros(3))))
    cs_lib.freeMdaStruct(st)
    store.close()
    cs_lib.mdaClearError()
    assert cs_lib.mdaLastError() == 0


def test_roofline_denominators_are_measurable(cs_lib):
    tf = ct.c_double()
    chisq.check(cs_lib, cs_lib.mdaMeasureFp64Peak(2, ct.byref(tf)))
    assert 5.0 < tf.value < 100.0            # B200 fp64 FMA pipe: tens of TFLOP/s
    gbs = ct.c_double()
    chisq.check(cs_lib, cs_lib.mdaMeasureHbmCopy(1 << 28, 3, ct.byref(gbs)))
    assert 1000.0 < gbs.value < 10000.0
