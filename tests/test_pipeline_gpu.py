"""End to end: files -> store -> fits -> walker starts -> device sampler -> device summaries,
with the per-bin result tuples of the reference's worker (mda.py:1318), sharded over two
"ranks" and merged like Pool.map's return (mda.py:188)."""
import os

import numpy as np
import pytest

from conftest import ROOT
from mimic_mda_b200 import builder, fits, pipeline, shard, synth
from oracle import posterior_oracle

pytestmark = pytest.mark.gpu


def _config(tmp_path):
    z = np.load(os.path.join(ROOT, "tests", "golden", "golden_builder.npz"))
    os.makedirs(tmp_path / "dists")
    (tmp_path / "exp.csv").write_text(str(z["exp_text"]))
    for name, text in zip(z["file_names"], z["file_texts"]):
        (tmp_path / "dists" / str(name)).write_text(str(text))
    return {"Input File Path": str(tmp_path / "exp.csv"), "Distribution Directory": str(tmp_path / "dists"), "Target A": 116,
            "Start Energy": 10.0, "Final Energy": 12.0, "Max Theta": 10.0, "Maximum L": 3,
            "EWSR Fractions": [1.0, 1.0, 0.8, 1.0], "Subtract IVGDR": True, "IVGDR Height": 280.0, "IVGDR Center": 15.3,
            "IVGDR Width": 4.9, "IVGDR CS Integral": 2100.0,
            "Number of Walkers": 48, "Sample Points": 300, "Burn-in Points": 60, "Number Walker Generators": 4,
            "Num Bins": 200}


def test_whole_path_and_sharding(tmp_path, cs_lib):
    config = _config(tmp_path)
    starts = fits.calc_start_params([[0.05, 0.25], [0.05, 0.25], [0.05, 0.25], [0.05, 0.2]])
    store, inputs = builder.build_store(config, cs_lib=cs_lib)
    n_bins = store.n_bins
    results, ens = pipeline.fit_and_mcmc_all(store, starts, config, ewsr=config["EWSR Fractions"], seed=5,
                                             bin_ids=np.arange(n_bins), rng=np.random.RandomState(1))
    assert len(results) == n_bins == 6
    chain = ens.chain
    assert chain.shape == (n_bins, 48, 300, 4)
    hi = 1.0 / np.array(config["EWSR Fractions"])
    assert np.all(chain >= 0.0) and np.all(chain <= hi)                      # (0, 1/EWSR) prior, mda.py:1124
    points, peaks = posterior_oracle.summarize(chain, 60, 0.682689492, 200)
    for b, (values, (acor, chis, acc)) in enumerate(results):
        np.testing.assert_array_equal(np.array(values[0]), points[b])
        np.testing.assert_array_equal(np.array(values[1]), peaks[b])
        assert acor == [] and 0.05 < acc < 0.95
        n = int(inputs["n_pts"][b])
        want = [np.sum((inputs["consts"][b, 0, :n] - np.array([v[0] for v in vals]) @ inputs["consts"][b, 1:, :n]) ** 2)
                for vals in values]
        np.testing.assert_allclose(chis, want, rtol=1e-12)
    ens.close()
    store.close()
    # two ranks, bins round-robin; per-rank RNG for the (host-side) walker starts is rank-local, so
    # compare what must be partition independent: the fits, and the sampler given the same starts
    per_rank = []
    for r in range(2):
        mine = shard.bins_for_rank(n_bins, r, 2)
        st, _ = builder.build_store(config, cs_lib=cs_lib, bins=mine)
        res, e = pipeline.fit_and_mcmc_all(st, starts, config, ewsr=config["EWSR Fractions"], seed=5, bin_ids=mine,
                                           rng=np.random.RandomState(1))
        per_rank.append(res)
        e.close()
        st.close()
    merged = shard.merge_by_bin(n_bins, 2, per_rank)
    assert len(merged) == n_bins and all(len(m[0][0]) == 4 for m in merged)
    # rank 0's first bin is global bin 0 with the same walker-start draws => identical result
    np.testing.assert_array_equal(np.array(merged[0][0][0]), np.array(results[0][0][0]))


def test_config_driver_runs_a_reference_style_config(tmp_path, cs_lib):
    """tools/run_config.py (a convenience driver outside the package): a configuration file in the reference's
    format (a Python file filling CONFIG) drives the whole path; keys unrelated to the likelihood path are ignored."""
    import importlib.util
    import json
    spec = importlib.util.spec_from_file_location("run_config", os.path.join(ROOT, "tools", "run_config.py"))
    run = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(run)
    cfg = _config(tmp_path)
    cfg.update({"Start Pts a0": [0.05, 0.25], "Start Pts a1": [0.05, 0.25], "Start Pts a2": [0.1], "Start Pts a3": [0.05, 0.2],
                "Shared Lib Path": "./C/libChiSq.so", "Generate Fit Plots": True, "Plot DPI": 300, "Number of Threads": 4})
    path = tmp_path / "mda_config_test.py"
    path.write_text("CONFIG = {}\n" + "".join("CONFIG[%r] = %r\n" % kv for kv in cfg.items()))
    out = tmp_path / "res.json"
    assert run.main([str(path), "--seed", "3", "--out", str(out)]) == 0
    table = json.loads(out.read_text())
    assert len(table) == 6 and abs(table[0]["Ex"] - 10.1) < 1e-12
    for row in table:
        assert 0.05 < row["acceptance_fraction"] < 0.95 and row["chi2_percentile"] > 0
        assert all(0.0 <= row["a%d" % l]["percentile"][0] <= 1.25 for l in range(4))


def test_grouped_pipeline_equals_the_one_group_run(tmp_path, cs_lib):
    """Bins worked through in groups sized to the free memory (the reference's own sizes are ~3 GB of chain per bin): the
    results are those of all bins at once, because the random streams are keyed by the global bin index."""
    cfg = _config(tmp_path)
    cfg.update({"Number of Walkers": 32, "Sample Points": 60, "Burn-in Points": 10, "Number Walker Generators": 3, "Num Bins": 50})
    inputs = builder.build_inputs(cfg)
    starts = fits.calc_start_params([[0.05, 0.25], [0.05, 0.25], [0.05, 0.25], [0.05, 0.2]])
    ewsr = cfg["EWSR Fractions"][:4]
    whole = pipeline.fit_and_mcmc_grouped(inputs["consts"], inputs["n_pts"], inputs["bounds"], starts, cfg, ewsr=ewsr, seed=4,
                                          rng=np.random.RandomState(9), group_size=100, cs_lib=cs_lib)
    parts = pipeline.fit_and_mcmc_grouped(inputs["consts"], inputs["n_pts"], inputs["bounds"], starts, cfg, ewsr=ewsr, seed=4,
                                          rng=np.random.RandomState(9), group_size=4, cs_lib=cs_lib)
    assert len(whole) == len(parts) == len(inputs["energies"])
    for a, b in zip(whole, parts):
        np.testing.assert_array_equal(np.array(a[0][0]), np.array(b[0][0]))
        np.testing.assert_array_equal(np.array(a[0][1]), np.array(b[0][1]))
        assert a[1][1] == b[1][1] and a[1][2] == b[1][2]
    # sizing: the reference's defaults are 2.95 GB per bin; a 180 GB part takes a few dozen at a time, never zero
    per_bin = pipeline.chain_bytes_per_bin({}, 8)
    assert 2.9e9 < per_bin < 3.0e9
    assert pipeline.bins_per_group({}, 8, 116, free_bytes=170e9) == 34
    assert pipeline.bins_per_group({}, 8, 116, free_bytes=1e9) == 1
    assert 1 <= pipeline.bins_per_group(cfg, 4, 6, cs_lib=cs_lib) <= 6


def test_an_ill_conditioned_bin_is_sampled_in_a_group_of_its_own(tmp_path, cs_lib):
    """One bin with 0.01 % errors in a spectrum: only that bin goes to the reference-order stepper, and every bin's result
    is the one it has when the whole spectrum is sampled at once."""
    cfg = {"Number of Walkers": 32, "Sample Points": 40, "Burn-in Points": 8, "Number Walker Generators": 3, "Num Bins": 40}
    consts, n_pts, _ = synth.make_bins(5, 4, 30)
    bad, _, _ = synth.make_bins(1, 4, 30, rel_err=1e-4, seed_base=77)
    consts[3] = bad[0]
    bounds = synth.default_bounds(4)
    starts = fits.calc_start_params([[0.05, 0.25]] * 4)
    split = pipeline.fit_and_mcmc_grouped(consts, n_pts, bounds, starts, cfg, seed=2, rng=np.random.RandomState(5), cs_lib=cs_lib)
    whole = pipeline.fit_and_mcmc_grouped(consts, n_pts, bounds, starts, cfg, seed=2, rng=np.random.RandomState(5), cs_lib=cs_lib,
                                          split_by_conditioning=False)
    assert len(split) == 5 and all(r is not None for r in split)
    for b, (a, w) in enumerate(zip(split, whole)):
        if b == 3:
            continue        # sampled by the same (reference-order) stepper in both runs, but next to different bins: same chain
        # the well-conditioned bins ran on the triangular-factor stepper in one run and on the reference-order stepper in the
        # other: same decisions, so the same samples and summaries
        np.testing.assert_array_equal(np.array(a[0][0]), np.array(w[0][0]))
        assert a[1][2] == w[1][2]
    np.testing.assert_array_equal(np.array(split[3][0][0]), np.array(whole[3][0][0]))
