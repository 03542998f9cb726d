"""Initial fits and walker start positions: the stage between the store and the sampler.

Mirrors, with the reference's names and argument meaning:

    calc_start_params      mda.py:1912-1957  Cartesian grid of start points ("Start Pts aN")
    initial_fits           mda.py:1114-1120  fit from every start, sort by chi^2, keep the best
                                             "Number Walker Generators"  (do_init_fit 1664-1709
                                             runs on the GPU: DeviceStore.fit -> mdaBatchFit)
    gen_walker_starts      mda.py:1581-1610
    randomize_position     mda.py:1613-1661

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


def randomize_position(gen, ndims, spread, offset_centroid, offset_width, ewsr, rng):
    """One randomised walker start from a generator (mda.py:1613-1661); ``rng`` supplies
    ``standard_normal`` (the reference uses the global, unseeded numpy generator)."""
    rands = spread * rng.standard_normal(ndims)
    randomizer = np.ones(ndims, dtype=np.float64) - rands
    base_offsets = offset_width * rng.standard_normal(ndims)
    offsets = offset_centroid * np.ones(ndims, dtype=np.float64) + base_offsets
    position = randomizer * (gen + offsets)
    for i in range(ndims):
        if position[i] < 0.0:
            position[i] *= -1.0
        if position[i] < FLOAT_EPSILON:
            position[i] = 0.001
        if position[i] > (1.0 / ewsr[i]):
            position[i] = (1.0 / ewsr[i]) - 0.001
    return position


def gen_walker_starts(gens, n_walkers, spread, offset_centroid, offset_width, ewsr, rng):
    """Walker start positions for one bin from its generators (mda.py:1581-1610)."""
    ndims = gens.shape[1]
    n_gen = gens.shape[0]
    return np.array([randomize_position(gens[i % n_gen], ndims, spread, offset_centroid, offset_width, ewsr, rng)
                     for i in range(n_walkers)])
