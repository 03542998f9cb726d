import sys, time; sys.path.insert(0,'/root/repo')
import numpy as np
from mimic_mda_b200 import chisq, synth
from mimic_mda_b200.ensemble import DeviceEnsemble
consts, counts, a_true = synth.make_bins(200, 9, 60)
store = chisq.DeviceStore(consts, counts)
ens = DeviceEnsemble(store, 512, seed=1)
ens.set_state(synth.make_walkers(a_true, 512, seed=3, oob_frac=0.0))
ens.reserve_chain(500); ens.advance(500)
ens.summarize(125)
t=time.perf_counter(); ens.summarize(125); print("summarize total ms", (time.perf_counter()-t)*1e3)
