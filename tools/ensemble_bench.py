#!/usr/bin/env python
"""Times the device-resident ensemble stepper (one persistent launch per run).
    python tools/ensemble_bench.py [--bins 200 --walkers 512 --n-l 9 --n-pts 60 --steps 500]
"""
import argparse
import ctypes as ct
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mimic_mda_b200 import chisq, synth  # noqa: E402
from mimic_mda_b200.ensemble import DeviceEnsemble  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--bins", type=int, default=200)
    ap.add_argument("--walkers", type=int, default=512)
    ap.add_argument("--n-l", type=int, default=9)
    ap.add_argument("--n-pts", type=int, default=60)
    ap.add_argument("--steps", type=int, default=500)
    ap.add_argument("--chain", type=int, default=1)
    ap.add_argument("--profile", type=int, default=0)
    ap.add_argument("--spacer", type=int, default=100, help="untimed steps enqueued right before the timed launch (0: none)")
    args = ap.parse_args()
    lib = chisq.prepare_shared_object()
    peak = ct.c_double()
    chisq.check(lib, lib.mdaMeasureFp64Peak(3, ct.byref(peak)))
    consts, counts, a_true = synth.make_bins(args.bins, args.n_l, args.n_pts)
    store = chisq.DeviceStore(consts, counts, cs_lib=lib)
    p0 = synth.make_walkers(a_true, args.walkers, seed=3, oob_frac=0.0)
    ens = DeviceEnsemble(store, args.walkers, seed=1)
    ens.set_state(p0)
    if args.chain:
        ens.reserve_chain(args.steps)
    if args.profile:
        ens.profile()
    ens.advance(20)                                   # warm-up (also fills 20 chain slots)
    ev0, ev1 = lib.mdaEventCreate(), lib.mdaEventCreate()
    if args.chain:
        ens.reserve_chain(args.steps)
    if args.spacer:                                   # keeps the GPU busy while the host enqueues the timed launch
        ens.advance(args.spacer, thin=10 ** 6, wait=False)
    lib.mdaEventRecord(ev0, store.handle)
    ens.advance(args.steps, wait=False)
    lib.mdaEventRecord(ev1, store.handle)
    lib.mdaEventSynchronize(ev1)
    ms = ct.c_float()
    lib.mdaEventElapsedMs(ev0, ev1, ct.byref(ms))
    ens.synchronize()
    evals = args.bins * args.walkers * args.steps
    flops = 2.0 * args.n_pts * (args.n_l + 1) * evals
    t0 = time.perf_counter()
    acc = ens.acceptance_fraction.mean()
    out = {"bins": args.bins, "walkers": args.walkers, "L": args.n_l, "N": args.n_pts, "steps": args.steps,
           "chain": bool(args.chain), "ms": ms.value, "us_per_step": ms.value * 1e3 / args.steps,
           "steps_per_s": args.steps / (ms.value * 1e-3), "evals_per_s": evals / (ms.value * 1e-3),
           "tflops": flops / (ms.value * 1e-3) / 1e12, "frac_fp64": flops / (ms.value * 1e-3) / 1e12 / peak.value,
           "fp64_peak": peak.value, "acceptance": float(acc)}
    if args.profile:
        cyc = ens.profile().astype(float)
        if ens.describe(args.steps).startswith("ensemble_qws"):
            ctas = min(args.bins, 148)
            wall = cyc[:ctas, 124:128]                     # ns (globaltimer): -, before the last wait, kernel entry, kernel exit
            t0 = wall[:, 2].min()
            extra = cyc[:ctas, 117:120]                     # ns: last piece done, first piece ready to step, first draws made
            out["wall_ns_median_since_own_entry"] = {"first_draws_made": float(np.median(extra[:, 2] - wall[:, 2])),
                                                     "first_piece_ready": float(np.median(extra[:, 1] - wall[:, 2])),
                                                     "last_piece_done": float(np.median(extra[:, 0] - wall[:, 2])),
                                                     "before_final_wait": float(np.median(wall[:, 1] - wall[:, 2])),
                                                     "exit": float(np.median(wall[:, 3] - wall[:, 2]))}
            out["wall_ns"] = {"entry_spread": float(wall[:, 2].max() - t0), "last_cta_before_final_wait": float(wall[:, 1].max() - t0), "last_cta_exit": float(wall[:, 3].max() - t0),
                              "first_cta_exit": float(wall[:, 3].min() - t0)}
            for name, o in (("walker_t0", 0), ("walker_t255", 8)):
                v = cyc[:ctas, o:o + 5].sum(axis=0)
                out[name + "_cycles_per_half_step"] = dict(zip(["gather+propose", "chi2", "accept+store+prefetch", "barrier"], [round(x / max(v[4], 1), 1) for x in v[:4]]))
            per_half = cyc[:, :8].mean(axis=0) / (2 * args.steps)
            u = cyc[:, 16:112].reshape(args.bins, 12, 8)      # per CTA and unit: stamps start, (U), before wait, after wait, (A), loop end, steps, (D)
            rows = []
            for cta in (0, 1, 2, 147):
                if cta < args.bins:
                    t0 = u[cta, 0, 0]
                    rows.append([[int(v - t0) for v in u[cta, k, [0, 1, 2, 3, 4, 5, 7]]] + [int(u[cta, k, 6])] for k in range(4) if u[cta, k, 0] > 0])
            out["unit_timeline_cycles"] = rows
            for name, o in (("draw0", 112), ("draw1", 120)):
                v = cyc[:, o:o + 5].mean(axis=0)
                out[name + "_cycles_per_half_step"] = dict(zip(["draws", "wait_read", "barrier", "post"], [round(x / max(v[4], 1), 1) for x in v[:4]]))
        elif os.environ.get("MDA_ENS_IMPL") == "dfma" or args.n_l < 4:
            per_half = cyc[:, :8].mean(axis=0) / (2 * args.steps)
            out["cycles_per_half_step"] = dict(zip(["propose+bar", "bar_after_A", "accept+bar", "chi_loop", "draw"], [round(x, 1) for x in per_half[:5]]))
        else:                                           # timeline of one half-step: [bin, warp, stamp], relative to the CTA's first stamp
            tl = cyc.reshape(args.bins, 16, 8)
            used = tl[:, :, 0] > 0
            t0 = np.where(used, tl[:, :, 0], np.inf).min(axis=1)[:, None, None]
            rel = np.where(tl > 0, tl - t0, np.nan)
            out["timeline_cycles_median_over_bins"] = [[None if np.isnan(v) else round(float(v)) for v in row] for row in np.nanmedian(rel, axis=0)]
    print(json.dumps(out))


if __name__ == "__main__":
    main()
