"""Store builder against outputs of the reference's own preprocessing
functions (tests/golden/golden_builder.npz, made by make_golden_builder.py by executing
mda.py:1989-2401 on synthetic files)."""
import os

import numpy as np
import pytest

from conftest import ROOT, rel_err
from mimic_mda_b200 import builder


@pytest.fixture(scope="module")
def case(tmp_path_factory):
    z = np.load(os.path.join(ROOT, "tests", "golden", "golden_builder.npz"))
    root = tmp_path_factory.mktemp("mda_inputs")
    os.makedirs(root / "dists")
    (root / "exp.csv").write_text(str(z["exp_text"]))
    for name, text in zip(z["file_names"], z["file_texts"]):
        (root / "dists" / str(name)).write_text(str(text))
    config = {"Input File Path": str(root / "exp.csv"), "Distribution Directory": str(root / "dists"), "Target A": 116,
              "Start Energy": 10.0, "Final Energy": 12.0, "Max Theta": 10.0, "Maximum L": 3,
              "EWSR Fractions": [1.0, 1.0, 0.8, 1.0], "Subtract IVGDR": False, "IVGDR Height": 280.0,
              "IVGDR Center": 15.3, "IVGDR Width": 4.9, "IVGDR CS Integral": 2100.0}
    return z, config


@pytest.mark.parametrize("tag", ["noivgdr", "ivgdr"])
def test_builder_reproduces_reference_preprocessing(case, tag):
    z, config = case
    config = dict(config, **{"Subtract IVGDR": tag == "ivgdr"})
    got = builder.build_inputs(config)
    np.testing.assert_array_equal(got["energies"], z[tag + "/energies"])           # Start/Final Energy window
    np.testing.assert_array_equal(got["n_pts"], z[tag + "/n_pts"])                 # per-row Max Theta cut
    np.testing.assert_array_equal([len(b.points) for b in got["bins"]], z[tag + "/n_plot"])    # every angle is kept for plots
    assert len(set(got["n_pts"].tolist())) > 1                                      # ragged, as in real data
    for b in range(len(got["energies"])):
        n = int(got["n_pts"][b])
        # the packed store block holds exactly the reference's values (bit for bit), zero padded
        np.testing.assert_array_equal(got["consts"][b, 0, :n], z["%s/fit_data_%d" % (tag, b)])
        np.testing.assert_array_equal(got["consts"][b, 1:, :n], z["%s/interp_dists_%d" % (tag, b)])
        assert np.all(got["consts"][b, :, n:] == 0.0)
    if tag == "ivgdr":
        np.testing.assert_array_equal(got["ivgdr_fractions"], z["ivgdr/ewsrs"])
        assert not np.array_equal(got["consts"][0, 0, :int(got["n_pts"][0])], z["noivgdr/fit_data_0"])     # the subtraction happened
    else:
        assert got["ivgdr_fractions"] is None
    assert got["bounds"] == [(0.0, 1.0), (0.0, 1.0), (0.0, 1.25), (0.0, 1.0)]       # (0, 1/EWSR), mda.py:1124


def test_ivgdr_fraction_formula():
    e = np.array([15.3, 10.0, 20.0])                   # at the centroid the Lorentzian is at its height
    cs = builder.lorentzian_cross_section(e, 280.0, 15.3, 4.9)
    assert abs(cs[0] - 280.0) < 1e-9
    want = 280.0 / (1.0 + (e ** 2 - 15.3 ** 2) ** 2 / (e ** 2 * 4.9 ** 2))      # the textbook form
    np.testing.assert_allclose(cs, want, rtol=1e-13)
    frac = builder.ivgdr_sum_rule_fraction(e, 280.0, 15.3, 4.9, 2100.0)
    assert abs(frac[0] - 280.0 / 2100.0) < 1e-12


@pytest.mark.gpu
def test_built_store_evaluates_like_the_oracle(case, cs_lib, oracle_lib):
    """Files -> store -> batched log-posterior, against the oracle on the builder's arrays."""
    from mimic_mda_b200 import shard, synth
    z, config = case
    store, inputs = builder.build_store(dict(config, **{"Subtract IVGDR": True}), cs_lib=cs_lib)
    lo = np.array([b[0] for b in inputs["bounds"]])
    hi = np.array([b[1] for b in inputs["bounds"]])
    params = synth.make_walkers(np.full((store.n_bins, 4), 0.2), 40, seed=2, lo=lo, hi=hi)
    want = oracle_lib.lnpost_bins(inputs["consts"], inputs["n_pts"], lo, hi, params)
    got = store.lnpost(params)
    assert np.array_equal(np.isneginf(got), np.isneginf(want)) and rel_err(got, want) <= 1e-12
    # a rank of a 2-way partition builds only its bins
    mine = shard.bins_for_rank(store.n_bins, 1, 2)
    part, _ = builder.build_store(dict(config, **{"Subtract IVGDR": True}), cs_lib=cs_lib, bins=mine)
    np.testing.assert_array_equal(part.lnpost(params[mine]), got[mine])
    part.close()
    store.close()
