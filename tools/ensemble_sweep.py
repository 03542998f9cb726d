#!/usr/bin/env python
"""Device-resident sampler over the BASELINE.json sweep (configs[4]: walkers x angles x dists), 296 bins per GPU,
one persistent launch of --steps steps each.  One JSON line per shape.
    python tools/ensemble_sweep.py > profiles/rNN_sampler_sweep.jsonl"""
import argparse
import ctypes as ct
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mimic_mda_b200 import chisq, synth  # noqa: E402
from mimic_mda_b200.ensemble import DeviceEnsemble  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--bins", type=int, default=296)
    ap.add_argument("--steps", type=int, default=300)
    args = ap.parse_args()
    lib = chisq.prepare_shared_object()
    peak = ct.c_double()
    chisq.check(lib, lib.mdaMeasureFp64Peak(3, ct.byref(peak)))
    shapes = [(w, n, l) for w in (32, 128, 512) for n in (16, 60, 256) for l in (4, 9, 12)]
    shapes += [(2048, 60, 9), (2500, 41, 8)]          # beyond 512 walkers: the DFMA stepper (the second: mda_config0.py:93,105)
    for walkers, n_pts, n_l in shapes:
        bins = args.bins if walkers <= 512 else 148
        consts, counts, a_true = synth.make_bins(bins, n_l, n_pts)
        store = chisq.DeviceStore(consts, counts, cs_lib=lib)
        p0 = synth.make_walkers(a_true, walkers, seed=3, oob_frac=0.0)
        ens = DeviceEnsemble(store, walkers, seed=1)
        ens.set_state(p0)
        ens.advance(10)
        text = ct.create_string_buffer(512)
        lib.mdaEnsembleDescribe(ens.handle, args.steps, text, 512)
        ev0, ev1 = lib.mdaEventCreate(), lib.mdaEventCreate()
        lib.mdaEventRecord(ev0, store.handle)
        ens.advance(args.steps, wait=False)
        lib.mdaEventRecord(ev1, store.handle)
        lib.mdaEventSynchronize(ev1)
        ms = ct.c_float()
        lib.mdaEventElapsedMs(ev0, ev1, ct.byref(ms))
        ens.synchronize()
        evals = bins * walkers * args.steps
        flops = 2.0 * n_pts * (n_l + 1) * evals
        print(json.dumps({"bins": bins, "walkers": walkers, "N": n_pts, "L": n_l, "steps": args.steps, "us_per_step": ms.value * 1e3 / args.steps,
                          "evals_per_s": evals / (ms.value * 1e-3), "tflops": flops / (ms.value * 1e-3) / 1e12,
                          "frac_fp64": flops / (ms.value * 1e-3) / 1e12 / peak.value, "acceptance": float(ens.acceptance_fraction.mean()),
                          "kernel": text.value.decode().split(":")[0]}), flush=True)
        ens.close()
        store.close()


if __name__ == "__main__":
    main()
