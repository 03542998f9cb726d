"""Wall time of host p0 in -> 1024 steps -> device summaries out (bench.py's e2e_summaries), stage by stage."""
import sys, time; sys.path.insert(0, '/root/repo')
import numpy as np
from mimic_mda_b200 import chisq, synth
from mimic_mda_b200.ensemble import DeviceEnsemble
consts, counts, a_true = synth.make_bins(200, 9, 60)
store = chisq.DeviceStore(consts, counts)
ens = DeviceEnsemble(store, 512, seed=1)
p0 = synth.make_walkers(a_true, 512, seed=3, oob_frac=0.0)
for steps in (16, 1024, 1024, 1024, 300, 1024):
    t0 = time.perf_counter(); ens.set_state(p0); t1 = time.perf_counter()
    ens.reserve_chain(steps); t2 = time.perf_counter()
    ens.advance(steps, 1); t3 = time.perf_counter()
    ens.summarize(steps // 4); t4 = time.perf_counter()
    print(steps, "set_state %.2f reserve %.2f advance %.2f summarize %.2f total %.2f ms" % tuple(1e3 * x for x in (t1 - t0, t2 - t1, t3 - t2, t4 - t3, t4 - t0)), ens.describe(steps)[-60:])
