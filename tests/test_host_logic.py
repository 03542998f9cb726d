"""Host-side logic that needs no GPU: synthetic inputs, bin sharding (including a
world_size-2 gloo run), and the batched stretch-move sampler driven by the oracle."""
import os
import subprocess
import sys

import numpy as np
import pytest

from conftest import ROOT
from mimic_mda_b200 import shard, synth
from mimic_mda_b200.sampler import BatchedEnsembleSampler, summarize_chain
from oracle import numpy_oracle


def test_synth_shapes_and_padding():
    consts, n_pts, a_true = synth.make_bins(7, 5, 31, ragged=True)
    assert consts.shape == (7, 6, 32) and n_pts.tolist() == [31, 30, 29, 28, 27, 31, 30]
    for b in range(7):
        assert np.all(consts[b, :, n_pts[b]:] == 0.0) and np.all(consts[b, 1:, :n_pts[b]] > 0.0)
    assert np.all((a_true >= 0) & (a_true <= 0.3))
    lo, hi = synth.default_bounds(5)
    p = synth.make_walkers(a_true, 200, oob_frac=0.05)
    outside = numpy_oracle.outside_box(p.reshape(-1, 5), lo, hi)
    assert 0 < outside.mean() < 0.15
    # a rank of a round-robin partition generates exactly its own bins
    mine, _, _ = synth.make_bins(4, 5, 31, ragged=True, first_bin=1, bin_step=2)
    np.testing.assert_array_equal(mine[1], consts[3])


def test_round_robin_partition_and_merge():
    for n_bins, world in ((200, 8), (7, 2), (3, 4), (0, 2)):
        seen = np.concatenate([shard.bins_for_rank(n_bins, r, world) for r in range(world)]) if world else []
        assert sorted(seen.tolist()) == list(range(n_bins))
        per_rank = [[("bin", int(b)) for b in shard.bins_for_rank(n_bins, r, world)] for r in range(world)]
        merged = shard.merge_by_bin(n_bins, world, per_rank)
        assert merged == [("bin", b) for b in range(n_bins)]
    assert shard.owner_of_bin(13, 8) == 5
    with pytest.raises(ValueError):
        shard.bins_for_rank(10, 3, 2)


def _oracle_logprob(consts, n_pts, lo, hi):
    return lambda params: numpy_oracle.lnpost_bins(consts, n_pts, lo, hi, params)


def test_sampler_bins_are_independent_streams():
    """Running bins together must give each bin the chain of a single-bin run with the
    same seed: only the likelihood call is batched."""
    consts, n_pts, a_true = synth.make_bins(3, 4, 20)
    lo, hi = synth.default_bounds(4)
    p0 = synth.make_walkers(a_true, 16, seed=3, oob_frac=0.0)
    both = BatchedEnsembleSampler(3, 16, 4, _oracle_logprob(consts, n_pts, lo, hi), seeds=[5, 6, 7])
    both.run_mcmc(p0, 12)
    for b in range(3):
        solo = BatchedEnsembleSampler(1, 16, 4, _oracle_logprob(consts[b:b + 1], n_pts[b:b + 1], lo, hi), seeds=[5 + b])
        solo.run_mcmc(p0[b:b + 1], 12)
        np.testing.assert_array_equal(solo.chain[0], both.chain[b])
        np.testing.assert_array_equal(solo.naccepted[0], both.naccepted[b])


def test_sampler_recovers_truth_and_respects_prior():
    consts, n_pts, a_true = synth.make_bins(2, 3, 40)
    lo, hi = synth.default_bounds(3)
    p0 = synth.make_walkers(a_true, 32, seed=9, oob_frac=0.0)
    smp = BatchedEnsembleSampler(2, 32, 3, _oracle_logprob(consts, n_pts, lo, hi), seeds=100)
    smp.run_mcmc(p0, 300)
    assert smp.chain.shape == (2, 32, 300, 3)
    assert np.all(smp.chain >= lo) and np.all(smp.chain <= hi)       # -inf proposals are never accepted
    acc = smp.acceptance_fraction
    assert 0.2 < acc.mean() < 0.9
    summ = summarize_chain(smp.chain, burn_in=100)
    assert np.all(np.abs(summ["p50"] - a_true) < 6 * summ["std"] + 1e-3)
    assert np.all(np.isfinite(smp.lnprobability))


def test_sampler_argument_checks():
    fn = lambda p: np.zeros(p.shape[:2])
    with pytest.raises(ValueError):
        BatchedEnsembleSampler(1, 7, 2, fn)
    with pytest.raises(ValueError):
        BatchedEnsembleSampler(1, 4, 3, fn)
    smp = BatchedEnsembleSampler(1, 8, 2, lambda p: np.full(p.shape[:2], np.nan), seeds=1)
    with pytest.raises(ValueError):
        smp.run_mcmc(np.random.default_rng(0).random((1, 8, 2)), 1)


def test_two_rank_gloo_sharded_run_matches_single_process(tmp_path):
    """world_size 2 over gloo on CPU: each rank samples its round-robin share of the
    bins with the oracle likelihood; the gathered chains equal the 1-process run."""
    script = os.path.join(ROOT, "tests", "_gloo_worker.py")
    out = str(tmp_path / "gloo")
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29631", PYTHONPATH=ROOT)
    procs = [subprocess.Popen([sys.executable, script, out], env=dict(env, RANK=str(r), LOCAL_RANK=str(r), WORLD_SIZE="2"))
             for r in range(2)]
    for p in procs:
        assert p.wait(timeout=300) == 0
    merged = np.load(out + "_rank0.npy")
    consts, n_pts, a_true = synth.make_bins(5, 3, 16)
    lo, hi = synth.default_bounds(3)
    p0 = synth.make_walkers(a_true, 12, seed=2, oob_frac=0.0)
    ref = BatchedEnsembleSampler(5, 12, 3, _oracle_logprob(consts, n_pts, lo, hi), seeds=[40 + b for b in range(5)])
    ref.run_mcmc(p0, 8)
    np.testing.assert_array_equal(merged, ref.chain)


def test_philox_known_answers(cs_lib):
    """Random123's published vectors for Philox4x32-10, on the numpy restatement and on the
    host build of the very source the kernels compile (csrc/philox.cuh)."""
    import ctypes as ct
    from oracle.device_sampler_oracle import philox4x32_10
    kat = [((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
           ((0xffffffff,) * 4, (0xffffffff,) * 2, (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
           ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0),
            (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1))]
    for ctr, key, want in kat:
        assert tuple(int(x) for x in philox4x32_10(*ctr, *key)) == want
        c, k, o = (ct.c_uint * 4)(*ctr), (ct.c_uint * 2)(*key), (ct.c_uint * 4)()
        cs_lib.mdaPhilox4x32(c, k, o)
        assert tuple(o) == want


def test_device_sampler_restatement_samples_the_same_posterior():
    """The CPU restatement of the device stepper (Philox streams) and the emcee-2-style
    sampler (Mersenne Twister streams) agree statistically."""
    from oracle import device_sampler_oracle as dso
    consts, n_pts, a_true = synth.make_bins(2, 3, 30)
    lo, hi = synth.default_bounds(3)
    p0 = synth.make_walkers(a_true, 48, seed=4, oob_frac=0.0)
    fn = _oracle_logprob(consts, n_pts, lo, hi)
    dev = dso.run(fn, p0, 500, seed=77)
    host = BatchedEnsembleSampler(2, 48, 3, fn, seeds=11)
    host.run_mcmc(p0, 500)
    s_d = summarize_chain(dev["chain"].transpose(0, 2, 1, 3), 150)
    s_h = summarize_chain(host.chain, 150)
    assert np.all(np.abs(s_d["mean"] - s_h["mean"]) < 0.3 * s_h["std"] + 1e-5)
    assert np.all(dev["chain"] >= lo) and np.all(dev["chain"] <= hi)
    assert 0.2 < dev["nacc"].mean() / 500 < 0.9


def test_walker_starts_obey_the_box_rules():
    """mda.py:1613-1661: negative values mirrored, tiny ones become 0.001, values above 1/EWSR become 1/EWSR - 0.001."""
    from mimic_mda_b200 import fits

    class Fixed:                                       # a "generator" with chosen normal draws
        def __init__(self, draws):
            self.draws = draws

        def standard_normal(self, shape):
            return np.broadcast_to(self.draws, shape).copy()
    gens = np.array([[0.2, 0.5, 0.0, 0.9], [0.1, 0.1, 0.1, 0.1]])
    ewsr = np.array([1.0, 1.0, 1.0, 2.0])                # upper bounds 1, 1, 1, 0.5
    # spread 1, scale draw 3 -> factor (1 - 3) = -2: values turn negative and are mirrored
    got = fits.gen_walker_starts(gens, 3, 1.0, 0.0, 0.0, ewsr, Fixed(np.array([[3.0], [0.0]])[None]))
    np.testing.assert_allclose(got[0], [0.4, 1.0 - 0.001 if 2 * 0.5 > 1.0 else 1.0, 0.001, 0.5 - 0.001])
    np.testing.assert_allclose(got[1], [0.2, 0.2, 0.2, 0.2])                    # walker 1 starts from generator 1
    np.testing.assert_allclose(got[2], got[0])                                   # walker 2 from generator 0 again
    rnd = fits.gen_walker_starts(gens, 500, 2.0, 0.04, 0.02, ewsr, np.random.RandomState(1))
    assert rnd.shape == (500, 4) and np.all(rnd > 0) and np.all(rnd <= 1.0 / ewsr)


def _ribbon_units(bins, n_steps, grid):
    """The stepper's schedule, restated (csrc/ensemble_qws_kernel.cu: qws_next_unit and the segment bounds of the kernel): the
    bins' step ranges laid end to end, cut into `grid` segments; per CTA the list of (bin, s0, s1, waits, signals)."""
    total = bins * n_steps
    out = []
    for cta in range(grid):
        p, r1 = cta * total // grid, (cta + 1) * total // grid
        units = []
        while p < r1:
            b, o = divmod(p, n_steps)
            length = min(n_steps - o, r1 - p)
            if o == 0 and length == n_steps:
                units.append((b, 0, n_steps, False, False))
            elif o == 0:                                   # the end of the segment: the bin's LAST steps, behind another CTA's piece
                units.append((b, n_steps - length, n_steps, True, False))
            else:                                          # the start of the segment: its FIRST steps
                units.append((b, 0, length, False, True))
            p += length
        out.append(units)
    return out


@pytest.mark.parametrize("bins,n_steps,grid", [(200, 20, 148), (200, 2000, 148), (200, 1, 148), (592, 20, 148), (149, 7, 148),
                                               (5, 3, 5), (7, 1, 3), (1000, 13, 148), (25, 20, 25)])
def test_ribbon_schedule_invariants(bins, n_steps, grid):
    """Every bin-step exactly once and in order; equal shares; a bin in at most two pieces, the later one starting (at
    equal speed) no earlier than the earlier one ends; at most one waiting and one signalling piece per CTA, and only at
    the end / the start of its segment."""
    sched = _ribbon_units(bins, n_steps, grid)
    covered = {}
    loads = []
    for cta, units in enumerate(sched):
        t = 0
        loads.append(sum(s1 - s0 for _, s0, s1, _, _ in units))
        for k, (b, s0, s1, waits, signals) in enumerate(units):
            assert 0 <= s0 < s1 <= n_steps
            assert not waits or k == len(units) - 1                # only the last piece of a segment continues another CTA's bin
            assert not signals or k == 0                           # only the first one is continued elsewhere
            covered.setdefault(b, []).append((s0, s1, cta, t, t + (s1 - s0), waits, signals))
            t += s1 - s0
    assert max(loads) - min(loads) <= 1 and sum(loads) == bins * n_steps
    assert sorted(covered) == list(range(bins))
    for b, pieces in covered.items():
        pieces.sort()
        assert len(pieces) <= 2 and pieces[0][0] == 0 and pieces[-1][1] == n_steps
        if len(pieces) == 2:
            first, second = pieces
            assert first[1] == second[0] and first[6] and second[5] and first[2] == second[2] + 1
            assert second[3] >= first[4]                           # the waiting piece starts after the other one has ended
        else:
            assert not pieces[0][5] and not pieces[0][6]
