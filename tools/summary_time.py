"""Wall time of the device posterior reductions on a C4-sized chain (200 bins x 512 walkers x 9, `steps` kept)."""
import sys, time; sys.path.insert(0, '/root/repo')
import numpy as np
from mimic_mda_b200 import chisq, synth
from mimic_mda_b200.ensemble import DeviceEnsemble
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 500
consts, counts, a_true = synth.make_bins(200, 9, 60)
store = chisq.DeviceStore(consts, counts)
ens = DeviceEnsemble(store, 512, seed=1)
ens.set_state(synth.make_walkers(a_true, 512, seed=3, oob_frac=0.0))
ens.reserve_chain(steps); ens.advance(steps)
ens.summarize(steps // 4)
t = time.perf_counter(); ens.summarize(steps // 4); print("summarize total ms", (time.perf_counter() - t) * 1e3)
for kw in (dict(c=3.0, fast=True), dict(c=3.0, fast=False), dict(c=1.0, fast=False)):
    ens.get_autocorr_time(quiet=True, **kw)
    t = time.perf_counter(); tau = ens.get_autocorr_time(quiet=True, **kw); dt = time.perf_counter() - t
    gb = steps * 200 * 512 * 9 * 8 / 1e9
    print("autocorr", kw, "ms %.3f" % (dt * 1e3), "chain GB %.2f -> %.0f GB/s" % (gb, gb / dt),
          "windows", np.bincount(ens.autocorr_window)[:1], "median tau %.2f" % np.nanmedian(tau))
