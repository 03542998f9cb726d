"""The oracle is pinned before it is trusted: C restatement, numpy restatement and
long-double evaluation against (i) the reference's known-answer vector and (ii) the
outputs of the unmodified reference library frozen in tests/golden/."""
import os

import numpy as np
import pytest

from conftest import rel_err, same_special
from oracle import numpy_oracle, reflib

TOL = 1e-13   # oracle vs reference library: both are fp64 CPU code, only FMA contraction differs


def test_known_answer_vector(golden, oracle_lib):
    # C/ChiSqTester.c:20,43 and so_py_tester.py:34,52: chi^2 = 0.456, lnL = -0.228
    kat = golden.case("kat")
    x, a = kat["data"], kat["params"]
    assert abs(float(kat["chi"]) - 0.456) < 1e-15
    assert abs(float(kat["lnl"]) + 0.228) < 1e-15
    assert abs(float(kat["lnl_resids"]) + 0.228) < 1e-15
    np.testing.assert_allclose(kat["resid"], 0.4 * x, rtol=0, atol=3e-16)
    D = np.stack([x, x, x])
    assert abs(oracle_lib.chi_batch(x, D, a)[0] - 0.456) < 1e-15
    assert abs(oracle_lib.lnlike_batch(x, D, a)[0] + 0.228) < 1e-15
    assert abs(numpy_oracle.chi_batch(x, D, a)[0] - 0.456) < 1e-15
    assert abs(numpy_oracle.lnlike_batch(x, D, a)[0] + 0.228) < 1e-15


def test_c_oracle_matches_reference_outputs(golden, oracle_lib):
    for name in golden.cases:
        g = golden.case(name)
        got = oracle_lib.lnpost_bins(g["consts"], g["n_pts"], g["lo"], g["hi"], g["params"])
        assert same_special(got, g["lnpost"]), name
        assert rel_err(got, g["lnpost"]) <= TOL, (name, rel_err(got, g["lnpost"]))
        for b in range(g["consts"].shape[0]):
            n = int(g["n_pts"][b])
            chi = oracle_lib.chi_batch(g["consts"][b, 0, :n], g["consts"][b, 1:, :n], g["params"][b])
            assert same_special(chi, g["chi"][b]), name
            assert rel_err(chi, g["chi"][b]) <= TOL, name


def test_numpy_oracle_matches_reference_outputs(golden):
    for name in golden.cases:
        g = golden.case(name)
        got = numpy_oracle.lnpost_bins(g["consts"], g["n_pts"], g["lo"], g["hi"], g["params"])
        assert same_special(got, g["lnpost"]), name
        assert rel_err(got, g["lnpost"]) <= TOL, (name, rel_err(got, g["lnpost"]))


def test_reference_is_well_conditioned_against_long_double(golden):
    # headroom under the 1e-12 gate: the reference itself is ~1e-15..1e-14 from exact
    for name in ("C2", "C3", "C5_large"):
        g = golden.case(name)
        b = 0
        n = int(g["n_pts"][b])
        exact = numpy_oracle.lnlike_exact(g["consts"][b, 0, :n], g["consts"][b, 1:, :n], g["params"][b])
        inside = np.isfinite(g["lnpost"][b])
        assert rel_err(g["lnpost"][b][inside], np.asarray(exact, dtype=np.float64)[inside]) < 1e-13


def test_prior_semantics(golden, oracle_lib):
    g = golden.case("C2")
    lo, hi = g["lo"], g["hi"]
    p = g["params"]
    # inclusive bounds evaluate, NaN falls through to the likelihood (mda.py:1575)
    assert np.isfinite(g["lnpost"][-1, 1]) and p[-1, 1, 0] == lo[0]
    assert np.isfinite(g["lnpost"][-1, 2]) and p[-1, 2, -1] == hi[-1]
    assert np.isnan(g["lnpost"][-1, 3]) and np.isnan(p[-1, 3, 0])
    outside = numpy_oracle.outside_box(p.reshape(-1, p.shape[2]), lo, hi).reshape(p.shape[:2])
    assert np.array_equal(outside, np.isneginf(g["lnpost"]))
    assert outside.sum() > 0


def test_residuals_and_index_guard(golden, oracle_lib):
    g = golden.case("odd_N")
    lib = reflib.prepare_shared_object(reflib.ORACLE_LIB)
    b = 1
    n = int(g["n_pts"][b])
    st = reflib.make_calc_struct(lib, g["consts"][b, 0, :n], [g["consts"][b, 1 + l, :n] for l in range(g["params"].shape[2])])
    scratch = np.zeros(n)
    val = lib.calculateLnLiklihoodResids(st, g["params"][b, 0].ctypes.data_as(reflib._DP), scratch.ctypes.data_as(reflib._DP))
    assert abs(val - g["lnl_resid"][b, 0]) <= TOL * abs(val)
    np.testing.assert_allclose(scratch, g["resid"][b, 0, :n], rtol=1e-12, atol=1e-12)
    # distIndex >= numLs is silently ignored (C/ChiSquare.c:88-89)
    before = lib.calculateChi(st, g["params"][b, 0].ctypes.data_as(reflib._DP))
    junk = np.full(n, 1e9)
    lib.setMdaDist(st, g["params"].shape[2], junk.ctypes.data_as(reflib._DP))
    lib.setMdaDist(st, g["params"].shape[2] + 5, junk.ctypes.data_as(reflib._DP))
    assert lib.calculateChi(st, g["params"][b, 0].ctypes.data_as(reflib._DP)) == before
    lib.freeMdaStruct(st)


def test_empty_bin_and_single_point(oracle_lib):
    # numPts = 0: every loop is empty, chi = 0, lnL = -0.0 (C/ChiSquare.c:204-212)
    lib = reflib.prepare_shared_object(reflib.ORACLE_LIB)
    st = lib.makeMdaStruct(0, 2)
    a = np.array([0.3, 0.4])
    assert lib.calculateChi(st, a.ctypes.data_as(reflib._DP)) == 0.0
    assert lib.calculateLnLiklihood(st, a.ctypes.data_as(reflib._DP)) == 0.0
    lib.freeMdaStruct(st)


@pytest.mark.skipif(reflib.reference_lib_path() is None, reason="no runnable build of the reference library in oracle/_ref")
def test_live_reference_library_agrees_with_fixture_and_oracle(golden, oracle_lib):
    """Where oracle/_ref exists (build container, or shipped to the GPU box) the
    unmodified reference is driven live, scalar call by scalar call."""
    cs_lib = reflib.prepare_shared_object(reflib.reference_lib_path())
    loop = reflib.RefLoop(reflib.reference_lib_path())
    for name in ("C1", "C3", "odd_N", "L20"):
        g = golden.case(name)
        live = reflib.lnpost_bins_scalar(cs_lib, g["consts"], g["n_pts"], g["lo"], g["hi"], g["params"])
        assert same_special(live, g["lnpost"]), name
        assert rel_err(live, g["lnpost"]) <= TOL, name          # native vs portable -march builds
        fast, _ = loop.run(g["consts"], g["n_pts"], g["lo"], g["hi"], g["params"], n_threads=2)
        assert same_special(fast, live) and rel_err(fast, live) == 0.0, name
        orc = oracle_lib.lnpost_bins(g["consts"], g["n_pts"], g["lo"], g["hi"], g["params"])
        assert rel_err(orc, live) <= TOL, name


@pytest.mark.skipif(not os.path.exists(reflib.REF_TESTER), reason="reference tester not built")
def test_reference_tester_prints_known_answer():
    import subprocess
    if reflib.reference_lib_path() is None:
        pytest.skip("reference build does not run on this CPU")
    out = subprocess.run([reflib.REF_TESTER], capture_output=True, text=True, timeout=60).stdout
    assert "the chi was: 0.456000" in out and out.count("the chi was: -0.228000") == 2


def test_triangular_factor_evaluation_reproduces_reference_outputs(golden):
    """chi^2(a) = |T [a; -1]|^2 with [D^T | d] = Q T (oracle/quad_oracle.py, the restatement of the device's default
    sampler arithmetic) against the golden outputs of the unmodified reference library: 1e-12 relative, the
    tolerance north_star states."""
    from oracle import quad_oracle
    worst = 0.0
    for name in golden.cases:
        g = golden.case(name)
        n_l = g["params"].shape[2]
        for b in range(g["consts"].shape[0]):
            n = int(g["n_pts"][b])
            t = quad_oracle.factor(g["consts"][b, 0, :n], g["consts"][b, 1:1 + n_l, :n])
            params = g["params"][b]
            ok = np.all(np.isfinite(params), axis=1)
            got = quad_oracle.chi(t, params[ok])
            want = g["chi"][b][ok]
            scale = np.maximum(np.abs(want), 1e-300)
            worst = max(worst, float(np.max(np.abs(got - want) / scale))) if want.size else worst
    assert worst <= 1e-12, worst


@pytest.mark.parametrize("rel", [5e-2, 1e-2, 1e-3, 3e-4, 1e-4])
def test_conditioning_guard_of_the_triangular_factor(rel):
    """Near the minimum the triangular-factor chi^2 and the reference-order chi^2 differ by about eps * kappa,
    kappa = |d| / sqrt(min chi^2) ~ 1 / (relative error of the cross-sections): inside the 1e-12 bound for kappa up to
    KAPPA_MAX (which is what the library evaluates from T), beyond it for 0.01 % errors -- those bins must be flagged."""
    from mimic_mda_b200 import synth
    from oracle import quad_oracle
    consts, n_pts, a_true = synth.make_bins(6, 9, 60, rel_err=rel, ragged=True)
    lo, hi = synth.default_bounds(9)
    params = synth.make_walkers(a_true, 256, seed=3, scatter=0.04 * rel, oob_frac=0.0)
    want = numpy_oracle.lnpost_bins(consts, n_pts, lo, hi, params)        # the reference's order (pinned above)
    worst, kappas = 0.0, []
    for b in range(6):
        n = int(n_pts[b])
        t = quad_oracle.factor(consts[b, 0, :n], consts[b, 1:, :n])
        kappas.append(quad_oracle.kappa(t))
        got = -0.5 * quad_oracle.chi(t, params[b])
        worst = max(worst, float(np.max(np.abs(got - want[b]) / np.abs(want[b]))))
    k = max(kappas)
    assert 0.5 / rel < k < 3.0 / rel                                       # kappa tracks 1 / relative error
    assert worst <= 8.0 * 2.0 ** -53 * k, (worst, k)                       # measured: ~2 eps kappa
    if k <= quad_oracle.KAPPA_MAX:
        assert worst <= 1e-13, (worst, k)                                  # what the library evaluates from T: 10x inside the bound
    assert (k <= quad_oracle.KAPPA_MAX) == (rel >= 1e-2)                   # 1 % and above from T, below in the reference's order
