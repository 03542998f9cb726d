#!/usr/bin/env python
"""Runs every stepper (ws, mma with / without draw warps, dfma) on the full C4 shape and reports, per pair, the bins whose
chains differ, plus four bins against the CPU restatement.  All must agree bit for bit (it found the prologue race of r01)."""
import os, sys, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mimic_mda_b200 import chisq, synth
from mimic_mda_b200.ensemble import DeviceEnsemble
from oracle import device_sampler_oracle as dso, reflib
lib = chisq.prepare_shared_object()
oracle = reflib.OracleLib()
n_bins, n_l, n_pts, walkers, steps = 200, 9, 60, 512, 24
consts, counts, a_true = synth.make_bins(n_bins, n_l, n_pts, ragged=True)
lo, hi = synth.default_bounds(n_l)
p0 = synth.make_walkers(a_true, walkers, seed=17, oob_frac=0.0)
store = chisq.DeviceStore(consts, counts, (lo, hi), cs_lib=lib)
pick = np.array([0, 77, 148, 199])
want = dso.run(lambda p: oracle.lnpost_bins(consts[pick], counts[pick], lo, hi, p), p0[pick], steps, seed=20240229, bin_ids=pick)["chain"]
def run(impl, dw, extra=None):
    os.environ["MDA_ENS_IMPL"] = impl; os.environ["MDA_ENS_DRAWW"] = dw
    for k in ("MDA_ENS_GRID", "MDA_ENS_CHUNK"): os.environ.pop(k, None)
    if extra: os.environ.update(extra)
    ens = DeviceEnsemble(store, walkers, seed=20240229)
    ens.run_mcmc(p0, steps)
    c = ens.chain.copy(); ens.close(); return c
res = {}
for name, impl, dw, extra in (("ws-a", "ws", "0", None), ("ws-b", "ws", "0", None), ("ws-nochunk?grid200", "ws", "0", {"MDA_ENS_CHUNK": "1000"}),
                              ("mma-dw-a", "mma", "1", None), ("mma-dw-b", "mma", "1", None), ("mma-2cta", "mma", "2", None), ("dfma", "dfma", "0", None)):
    c = run(impl, dw, extra); res[name] = c
    got = c[pick].transpose(0, 2, 1, 3)
    same = np.all(got == want, axis=(2, 3))
    print(name, "bins-vs-oracle identical steps per picked bin:", same.sum(axis=1), "of", steps, " first mismatch step per bin:", [int(np.argmin(s)) if not s.all() else -1 for s in same])
for a in res:
    for b in res:
        if a < b:
            diff = np.any(res[a] != res[b], axis=(1, 2, 3))
            print(a, b, "bins differing:", int(diff.sum()), np.nonzero(diff)[0][:12])
