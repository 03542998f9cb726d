"""CPU restatement of the reference's initial-fit loop -- TEST INFRASTRUCTURE ONLY.

``do_init_fit`` follows mda.py:1664-1709 call for call: scipy's ``fmin_l_bfgs_b`` with
``approx_grad=True, epsilon=1e-05, factr=10.0`` on ``call_chi_sq`` (mda.py:1712-1739) inside
the bounds ``(0, 1/EWSR)``, then ``Forced Extra Fits`` restarts from the previous result.
scipy is a third-party dependency of the reference (unpinned); 1.18 is what is here, which no
longer accepts the `iprint=0` the reference passes (it only silenced output).
"""
import numpy as np
from scipy import optimize

from . import reflib


def do_init_fit(start, struct, cs_lib, bounds, extra_fits):
    init_params = np.array(start, dtype=np.float64)
    ret_vals = optimize.fmin_l_bfgs_b(reflib.call_chi_sq, init_params, bounds=bounds, epsilon=1e-05,
                                      approx_grad=True, args=(cs_lib, struct), factr=10.0)
    for _ in range(extra_fits):
        ret_vals = optimize.fmin_l_bfgs_b(reflib.call_chi_sq, ret_vals[0], bounds=bounds, epsilon=1e-05,
                                          approx_grad=True, args=(cs_lib, struct), factr=10.0)
    return ret_vals[1], ret_vals[0]


def fit_bins(cs_lib, consts, n_pts, lo, hi, starts, extra_fits=3):
    """[B, S] chi^2 and [B, S, L] parameters from the reference procedure."""
    n_bins, rows, _ = consts.shape
    n_l = rows - 1
    bounds = [(float(lo[i]), float(hi[i])) for i in range(n_l)]
    chi = np.empty((n_bins, len(starts)))
    par = np.empty((n_bins, len(starts), n_l))
    for b in range(n_bins):
        n = int(n_pts[b])
        struct = reflib.make_calc_struct(cs_lib, consts[b, 0, :n], [consts[b, 1 + l, :n] for l in range(n_l)])
        for s, start in enumerate(starts):
            chi[b, s], par[b, s] = do_init_fit(start, struct, cs_lib, bounds, extra_fits)
        cs_lib.freeMdaStruct(struct)
    return par, chi
