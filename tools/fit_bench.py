#!/usr/bin/env python
"""Times the batched initial fits (mdaBatchFit) next to the reference procedure
(scipy L-BFGS-B with finite-difference gradients on calculateChi, mda.py:1664-1709).
    python tools/fit_bench.py [--bins 200 --n-l 9 --n-pts 60]
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mimic_mda_b200 import chisq, fits, synth  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--bins", type=int, default=200)
    ap.add_argument("--n-l", type=int, default=9)
    ap.add_argument("--n-pts", type=int, default=60)
    ap.add_argument("--ref-fits", type=int, default=24)
    args = ap.parse_args()
    lists = ([[0.05, 0.15, 0.25, 0.35, 0.45]] * 3 + [[0.05, 0.25, 0.45]] + [[0.05, 0.2]] * 12)[:args.n_l]
    starts = fits.calc_start_params(lists)                 # mda_config0.py:138-145 extended to n_l dims
    consts, counts, _ = synth.make_bins(args.bins, args.n_l, args.n_pts)
    lo, hi = synth.default_bounds(args.n_l)
    store = chisq.DeviceStore(consts, counts, (lo, hi))
    store.fit(starts[:8])
    t0 = time.perf_counter()
    par, chi = store.fit(starts)
    gpu_s = time.perf_counter() - t0
    out = {"bins": args.bins, "L": args.n_l, "N": args.n_pts, "starts_per_bin": len(starts),
           "fits": args.bins * len(starts), "gpu_seconds_incl_copies": gpu_s,
           "gpu_fits_per_s": args.bins * len(starts) / gpu_s,
           "chi_spread_over_starts": float(np.max(chi.max(axis=1) / chi.min(axis=1) - 1.0))}
    try:
        from oracle import init_fit_oracle, reflib
        lib = reflib.prepare_shared_object(reflib.reference_lib_path() or reflib.ORACLE_LIB)
        t0 = time.perf_counter()
        rpar, rchi = init_fit_oracle.fit_bins(lib, consts[:1], counts[:1], lo, hi, starts[:args.ref_fits], extra_fits=3)
        ref_s = time.perf_counter() - t0
        out.update({"reference_seconds_per_fit_1core": ref_s / args.ref_fits,
                    "reference_fits_per_s_1core": args.ref_fits / ref_s,
                    "reference_chi_minus_gpu_chi_max_rel": float(np.max(rchi[0] / chi[0, :args.ref_fits] - 1.0))})
    except Exception as exc:                               # oracle not present: GPU numbers only
        out["reference"] = "unavailable: %s" % exc
    print(json.dumps(out))


if __name__ == "__main__":
    main()
