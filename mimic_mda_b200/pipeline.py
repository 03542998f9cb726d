"""The per-bin worker of the reference, for every Ex bin of a rank at once.

``fit_and_mcmc`` (mda.py:1072-1139) does, per bin inside a pool worker: load the library,
fill the struct, multi-start fits, pick generators, randomise walker starts, run emcee,
summarise the samples, free the struct.  ``fit_and_mcmc_all`` runs the same stages for all
bins of a ``DeviceStore`` in a handful of launches:

    initial fits        mdaBatchFit              (mda.py:1114-1120)
    walker starts       fits.gen_walker_starts   (mda.py:1122, host numpy, once per bin)
    MCMC                mdaEnsembleAdvance       (mda.py:1128-1131; chain stays in HBM)
    summaries           mdaEnsembleSummarize     (mda.py:1253-1259, 1399-1491, 1316)
    autocorrelation     mdaEnsembleAutocorrTime  (mda.py:1303-1310, if "Calc AutoCorr")
    chi^2 of estimates  mdaBatchChi              (calculate_fit_chis, mda.py:1321-1354)

and returns, per bin, the tuple the reference's worker returns:
``((points, peaks), (acor_time, (perc_chi, peak_chi), accept_frac))``.  Only the walker
starts go to the device and only the summaries come back.
"""
import os

import numpy as np

from . import fits
from .ensemble import DeviceEnsemble

# the CONFIG keys this path reads, with the values of mda_config0.py
DEFAULTS = {
    "Number of Walkers": 2500, "Sample Points": 16385, "Burn-in Points": 512,
    "Number Walker Generators": 50, "Sample Spread": 2.0, "Sample Offset Centroid": 0.04,
    "Sample Offset Width": 0.02, "Confidence Interval": 0.682689492, "Num Bins": 1000,
    "Calc AutoCorr": False, "ACorr WindSize": 3.0, "ACorr Use FFT": True,
    "Save Chain Data": False, "Chain Directory": "chains", "Target A": 0,
}


def chain_file_name(config, energy):
    """The reference's per-bin chain dump (mda.py:1244-1249)."""
    return os.path.join(config["Chain Directory"], "A{0:d}_chain_E{1:5.2f}.npz".format(config["Target A"], energy))


def save_chains(ens, config, energies):
    """``np.savez_compressed(name, sampler.chain)`` per bin, as perform_sample_manips does when
    "Save Chain Data" is set (mda.py:1243-1250): ``arr_0`` is emcee's [walker, step, dim] array.  The chain is
    pulled from HBM once (the fetch the ensemble caches)."""
    chain = ens.chain                                     # [B, W, kept, L]
    os.makedirs(config["Chain Directory"], exist_ok=True)
    names = []
    for b, energy in enumerate(energies):
        names.append(chain_file_name(config, energy))
        np.savez_compressed(names[-1], np.ascontiguousarray(chain[b]))
    return names


def fit_and_mcmc_all(store, start_params, config, ewsr=None, seed=0, bin_ids=None, rng=None, thin=1, energies=None):
    """Returns ``(results, ensemble)``; ``results[b]`` has the shape of the reference worker's
    return value for bin b (mda.py:1318).  ``config`` holds the reference's CONFIG keys (missing
    ones take mda_config0.py's values); ``ewsr`` defaults to 1 for every multipole.  With "Save Chain Data" and the
    bins' ``energies`` given, every bin's chain is also written in the reference's file format."""
    cfg = dict(DEFAULTS)
    cfg.update(config)
    n_l = store.n_l
    ewsr = np.ones(n_l) if ewsr is None else np.asarray(ewsr, dtype=np.float64)
    rng = np.random.RandomState(seed) if rng is None else rng
    # initial fits from every start, best "Number Walker Generators" per bin
    gens, _ = fits.initial_fits(store, start_params, cfg["Number Walker Generators"])
    # walker start positions (mda.py:1122)
    walkers = cfg["Number of Walkers"]
    starts = np.stack([fits.gen_walker_starts(gens[b], walkers, cfg["Sample Spread"], cfg["Sample Offset Centroid"],
                                              cfg["Sample Offset Width"], ewsr, rng) for b in range(store.n_bins)])
    # the ensemble (mda.py:1128-1131): bounds (0, 1/EWSR) are the store's
    ens = DeviceEnsemble(store, walkers, seed=seed, bin_ids=bin_ids)
    ens.set_state(starts)
    steps = cfg["Sample Points"]
    ens.reserve_chain(steps // thin)
    ens.advance(steps, thin)
    if cfg["Save Chain Data"] and energies is not None:
        save_chains(ens, cfg, energies)
    # summaries on the device chain (perform_sample_manips, mda.py:1193-1318)
    summ = ens.summarize(cfg["Burn-in Points"] // thin, cfg["Confidence Interval"], cfg["Num Bins"])
    points, peaks = summ["points"], summ["peaks"]
    # autocorrelation times of the whole chain (mda.py:1303-1310).  Where the chain is too short emcee raises inside
    # the bin's worker; with every bin in one call the others' results are worth keeping: NaN times for that bin
    acor = (ens.get_autocorr_time(c=cfg["ACorr WindSize"], fast=cfg["ACorr Use FFT"], quiet=True)
            if cfg["Calc AutoCorr"] else None)
    # chi^2 of the percentile and peak parameter sets (calculate_fit_chis, mda.py:1350-1353)
    chis = store.chi(np.stack([points[:, :, 0], peaks[:, :, 0]], axis=1))             # [B, 2]
    results = []
    for b in range(store.n_bins):
        values = ([tuple(points[b, l]) for l in range(n_l)], [tuple(peaks[b, l]) for l in range(n_l)])
        acor_time = [] if acor is None else acor[b].copy()
        results.append((values, (acor_time, (float(chis[b, 0]), float(chis[b, 1])), float(summ["acceptance_fraction"][b]))))
    return results, ens


def chain_bytes_per_bin(config, n_l, thin=1):
    """HBM a bin's kept chain and log-probabilities take (plus its state): what limits how many bins run at once."""
    cfg = dict(DEFAULTS)
    cfg.update(config)
    kept = cfg["Sample Points"] // thin
    return 8 * cfg["Number of Walkers"] * (n_l + 1) * (kept + 2)


def bins_per_group(config, n_l, n_bins, thin=1, free_bytes=None, cs_lib=None, headroom=0.6):
    """How many bins' chains fit the device at once: ``headroom`` of the free HBM (the posterior reductions need
    scratch of the order of a histogram per bin and parameter, the chain itself is the bulk).  The reference's own
    defaults (mda_config0.py: 2500 walkers x 16385 samples x 8 parameters) are 2.95 GB per bin -- its full energy
    range, ~116 bins, does not fit one 180 GB part at once, so bins are worked through in groups."""
    if free_bytes is None:
        import ctypes as ct
        from . import chisq
        lib = chisq.prepare_shared_object() if cs_lib is None else cs_lib
        free, total = ct.c_size_t(), ct.c_size_t()
        chisq.check(lib, lib.mdaDeviceMemInfo(ct.byref(free), ct.byref(total)), "mdaDeviceMemInfo")
        free_bytes = free.value
    per_bin = chain_bytes_per_bin(config, n_l, thin)
    return int(max(1, min(n_bins, (headroom * free_bytes) // per_bin)))


class _DrawnNoise:
    """Hands out normal draws made beforehand, bin by bin: the walker starts of a bin must not depend on which group the
    bin is sampled in."""

    def __init__(self, blocks):
        self.blocks = list(blocks)

    def standard_normal(self, shape):
        block = self.blocks.pop(0)
        assert block.shape == tuple(shape)
        return block


def fit_and_mcmc_grouped(consts, n_pts, bounds, start_params, config, ewsr=None, seed=0, bin_ids=None, rng=None, thin=1,
                         energies=None, group_size=None, cs_lib=None, split_by_conditioning=True):
    """``fit_and_mcmc_all`` over groups of bins sized to the device's free memory (``group_size`` bins at a time;
    default: ``bins_per_group``).  Every group gets its own store and ensemble, closed before the next one starts; chain
    files ("Save Chain Data") are written per group.  With ``split_by_conditioning`` the bins whose triangular-factor
    likelihood is not safe (``DeviceStore.conditioning``: cross-section errors below ~0.5 %) are sampled in groups of
    their own, so that one such bin sends only itself -- not the spectrum -- to the slower reference-order stepper.
    Results do not depend on the grouping: the sampler's random streams are keyed by the global bin index and the
    walker-start draws are made bin by bin, in order, before any group runs."""
    from . import chisq
    consts = np.asarray(consts)
    n_pts = np.asarray(n_pts)
    n_bins, n_l = consts.shape[0], consts.shape[1] - 1
    cfg = dict(DEFAULTS)
    cfg.update(config)
    bin_ids = np.arange(n_bins) if bin_ids is None else np.asarray(bin_ids)
    rng = np.random.RandomState(seed) if rng is None else rng
    noise = [rng.standard_normal((cfg["Number of Walkers"], 2, n_l)) for _ in range(n_bins)]
    if group_size is None:
        group_size = bins_per_group(config, n_l, n_bins, thin, cs_lib=cs_lib)
    order = [np.arange(n_bins)]
    if split_by_conditioning and n_bins:
        probe = chisq.DeviceStore(consts, n_pts, bounds, cs_lib=cs_lib)
        kappa, limit, safe = probe.conditioning()
        probe.close()
        if not safe:
            order = [np.nonzero(kappa <= limit)[0], np.nonzero(~(kappa <= limit))[0]]
    results = [None] * n_bins
    for members in order:
        for first in range(0, len(members), group_size):
            sel = members[first:first + group_size]
            store = chisq.DeviceStore(consts[sel], n_pts[sel], bounds, cs_lib=cs_lib)
            part, ens = fit_and_mcmc_all(store, start_params, config, ewsr=ewsr, seed=seed, bin_ids=bin_ids[sel],
                                         rng=_DrawnNoise(noise[b] for b in sel), thin=thin,
                                         energies=None if energies is None else [list(energies)[b] for b in sel])
            for b, res in zip(sel, part):
                results[b] = res
            ens.close()
            store.close()
    return results
