#!/usr/bin/env python
"""Generates tests/golden/golden_lnl.npz from the UNMODIFIED reference library.

Run in the build container, where /root/reference exists:

    make -C oracle all && python tests/golden/make_golden.py

Every value in the fixture is produced by oracle/_ref/libChiSq.so -- the
reference's own C/ChiSquare.c compiled in place with its makefile's flags
(portable -march, see oracle/Makefile) -- driven call-for-call the way mda.py
drives it: makeMdaStruct / setMdaData / setMdaDist (mda.py:1742-1780), then
per walker ``ln_post_prob`` (mda.py:1540-1578), ``calculateChi``
(mda.py:1712-1739) and, for a few walkers, ``calculateLnLiklihoodResids``.
The inputs are stored next to the outputs so the fixture does not depend on
numpy's random streams.  /root/reference is NOT needed to *use* the fixture.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from mimic_mda_b200 import synth  # noqa: E402
from oracle import reflib  # noqa: E402

_DP = reflib._DP

# name -> (bins, L, N, walkers, ragged)
CASES = {
    "C1": (1, 4, 30, 64, False),
    "C2": (3, 5, 30, 48, False),
    "C3": (3, 9, 60, 48, False),
    "C4": (3, 9, 60, 48, True),          # ragged angle counts, C4's L/N
    "C5_small": (2, 4, 16, 32, False),
    "C5_large": (2, 12, 256, 32, False),
    "odd_N": (2, 7, 33, 32, True),
    "L1_N1": (2, 1, 1, 8, False),
    "L8_ref": (2, 8, 41, 40, False),     # the reference's own Maximum L = 7 (mda_config0.py:93)
    "L16": (1, 16, 24, 16, False),
    "L20": (1, 20, 18, 16, False),       # beyond the register-tiled kernels -> generic kernel
}


def reference_outputs(cs_lib, consts, n_pts, lo, hi, params):
    n_bins, n_walk, n_l = params.shape
    bounds = [(float(lo[i]), float(hi[i])) for i in range(n_l)]
    lnpost = np.empty((n_bins, n_walk))
    chi = np.empty((n_bins, n_walk))
    n_res = min(4, n_walk)
    resid = np.zeros((n_bins, n_res, consts.shape[2]))
    lnl_resid = np.empty((n_bins, n_res))
    for b in range(n_bins):
        n = int(n_pts[b])
        struct = reflib.make_calc_struct(cs_lib, consts[b, 0, :n], [consts[b, 1 + l, :n] for l in range(n_l)])
        for w in range(n_walk):
            lnpost[b, w] = reflib.ln_post_prob(params[b, w], cs_lib, struct, bounds)
            chi[b, w] = reflib.call_chi_sq(params[b, w], cs_lib, struct)
        for w in range(n_res):
            scratch = np.zeros(max(n, 1))
            lnl_resid[b, w] = cs_lib.calculateLnLiklihoodResids(
                struct, params[b, w].ctypes.data_as(_DP), scratch.ctypes.data_as(_DP))
            resid[b, w, :n] = scratch[:n]
        cs_lib.freeMdaStruct(struct)
    return lnpost, chi, lnl_resid, resid


def main():
    if not os.path.exists(reflib.REF_LIB_PORTABLE):
        raise SystemExit("build the reference first: make -C oracle all")
    cs_lib = reflib.prepare_shared_object(reflib.REF_LIB_PORTABLE)
    out = {}
    # the reference's only known-answer vector (C/ChiSqTester.c:20,43; so_py_tester.py:34,52)
    x = np.array([0.0, 0.1, 0.2, 0.3, 0.4, 0.5, 0.6, 0.7, 0.8, 0.9])
    a = np.array([0.2, 0.2, 0.2])
    st = reflib.make_calc_struct(cs_lib, x, [x, x, x])
    scratch = np.zeros(10)
    out["kat/data"] = x
    out["kat/params"] = a
    out["kat/chi"] = np.array(cs_lib.calculateChi(st, a.ctypes.data_as(_DP)))
    out["kat/lnl"] = np.array(cs_lib.calculateLnLiklihood(st, a.ctypes.data_as(_DP)))
    out["kat/lnl_resids"] = np.array(cs_lib.calculateLnLiklihoodResids(st, a.ctypes.data_as(_DP), scratch.ctypes.data_as(_DP)))
    out["kat/resid"] = scratch
    cs_lib.freeMdaStruct(st)

    for name, (bins, n_l, n_pts, walkers, ragged) in CASES.items():
        consts, counts, a_true = synth.make_bins(bins, n_l, n_pts, ragged=ragged)
        lo, hi = synth.default_bounds(n_l)
        params = synth.make_walkers(a_true, walkers, seed=11, oob_frac=0.08)
        params[0, 0, :] = a_true[0]                      # the truth itself
        if walkers > 4:
            params[-1, 1:4, :] = a_true[-1]
            params[-1, 1, 0] = lo[0]                     # exactly on the inclusive bounds
            params[-1, 2, n_l - 1] = hi[n_l - 1]
            params[-1, 3, 0] = np.nan                    # NaN falls through the prior (mda.py:1575)
        lnpost, chi, lnl_resid, resid = reference_outputs(cs_lib, consts, counts, lo, hi, params)
        for key, val in (("consts", consts), ("n_pts", counts), ("lo", lo), ("hi", hi), ("params", params),
                         ("lnpost", lnpost), ("chi", chi), ("lnl_resid", lnl_resid), ("resid", resid)):
            out["%s/%s" % (name, key)] = val
        n_inf = int(np.isneginf(lnpost).sum())
        print("%-9s bins=%d L=%d N=%d W=%d  -inf rows: %d  lnL range [%.4g, %.4g]" % (
            name, bins, n_l, n_pts, walkers, n_inf, np.nanmin(lnpost[np.isfinite(lnpost)]),
            np.nanmax(lnpost[np.isfinite(lnpost)])))
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden_lnl.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
