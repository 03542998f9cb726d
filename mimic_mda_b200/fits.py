"""Initial fits and walker start positions: the stage between the store and the sampler.

Mirrors, with the reference's names and argument meaning:

    calc_start_params      mda.py:1912-1957  Cartesian grid of start points ("Start Pts aN")
    initial_fits           mda.py:1114-1120  fit from every start, sort by chi^2, keep the best
                                             "Number Walker Generators"  (do_init_fit 1664-1709
                                             runs on the GPU: DeviceStore.fit -> mdaBatchFit)
    gen_walker_starts      mda.py:1581-1661  (with its per-walker helper folded in, vectorised)

The fits are done for all Ex bins and all starts in one library call.
"""
import itertools

import numpy as np

FLOAT_EPSILON = 1.0e-7          # mda.py:55


def calc_start_params(start_lists):
    """``start_lists[i]`` = the reference's CONFIG["Start Pts a%d" % i].  Returns ``[S, L]``
    in the reference's order (last index fastest, mda.py:1960-1986)."""
    return np.array(list(itertools.product(*start_lists)), dtype=np.float64)


def initial_fits(store, starts, n_generators):
    """Fit every bin from every start; per bin the ``n_generators`` lowest-chi^2 fits
    (stable sort on chi^2, as ``final_points.sort(key=lambda x: x[0])``, mda.py:1117-1120).
    Returns ``(generators [B, n_generators, L], chi2 [B, n_generators])``."""
    params, chi = store.fit(starts)
    order = np.argsort(chi, axis=1, kind="stable")[:, :n_generators]
    gens = np.take_along_axis(params, order[:, :, None], axis=1)
    return gens, np.take_along_axis(chi, order, axis=1)


def gen_walker_starts(gens, n_walkers, spread, offset_centroid, offset_width, ewsr, rng):
    """Walker start positions for one bin, ``[n_walkers, L]``: walker i starts from generator ``i mod n_gen``, shifted by
    a Gaussian offset (centroid, width) and scaled by ``1 - spread * N(0, 1)`` per parameter, then brought inside the
    box: negative values are mirrored, values below ``FLOAT_EPSILON`` become 0.001 and values above ``1 / EWSR_l``
    become ``1 / EWSR_l - 0.001`` (the rule of mda.py:1581-1661, applied to all walkers at once).  ``rng`` supplies
    ``standard_normal`` (the reference draws from numpy's global, unseeded generator, so no draw order is defined)."""
    gens = np.asarray(gens, dtype=np.float64)
    upper = 1.0 / np.asarray(ewsr, dtype=np.float64)
    noise = rng.standard_normal((n_walkers, 2, gens.shape[1]))
    base = gens[np.arange(n_walkers) % gens.shape[0]]
    pos = np.abs((1.0 - spread * noise[:, 0]) * (base + offset_centroid + offset_width * noise[:, 1]))
    pos = np.where(pos < FLOAT_EPSILON, 0.001, pos)
    return np.where(pos > upper, upper - 0.001, pos)
