"""The triangular-factor stepper (the default) against the angle-by-angle steppers (MDA_ENS_IMPL=direct) on the same seeds: positions identical, log-probabilities
to 1e-12 relative (and against the numpy oracle on the final state)."""
import os, sys; sys.path.insert(0, '/root/repo')
import numpy as np
from mimic_mda_b200 import chisq, synth
from mimic_mda_b200.ensemble import DeviceEnsemble
from oracle import numpy_oracle

def run(impl, bins, n_l, n_pts, walkers, steps, thin, ragged):
    if impl: os.environ["MDA_ENS_IMPL"] = impl
    else: os.environ.pop("MDA_ENS_IMPL", None)
    consts, counts, a_true = synth.make_bins(bins, n_l, n_pts, ragged=ragged)
    store = chisq.DeviceStore(consts, counts, synth.default_bounds(n_l))
    ens = DeviceEnsemble(store, walkers, seed=11)
    ens.run_mcmc(synth.make_walkers(a_true, walkers, seed=2, oob_frac=0.02), steps, thin=thin)
    out = (ens.chain.copy(), ens.lnprobability.copy(), ens.naccepted.copy(), ens.describe(steps))
    ens.close(); store.close()
    return out + (consts, counts)

worst = 0.0
for (bins, n_l, n_pts, walkers, steps, thin, ragged) in [(5, 4, 30, 64, 200, 1, True), (3, 9, 60, 512, 120, 1, False), (4, 5, 30, 256, 150, 3, True),
                                                         (2, 8, 41, 2500, 40, 1, False), (3, 12, 256, 96, 100, 2, False), (2, 3, 16, 10, 300, 1, False),
                                                         (2, 1, 20, 32, 100, 1, False), (2, 6, 4, 40, 100, 1, False)]:
    a = run("quad", bins, n_l, n_pts, walkers, steps, thin, ragged)
    b = run("direct", bins, n_l, n_pts, walkers, steps, thin, ragged)
    same_pos = np.array_equal(a[0], b[0])
    fin = np.isfinite(b[1])
    rel = np.max(np.abs(a[1][fin] - b[1][fin]) / np.abs(b[1][fin]))
    worst = max(worst, rel)
    print((bins, n_l, n_pts, walkers, steps, thin), "positions identical:", same_pos, "inf pattern:", np.array_equal(np.isfinite(a[1]), fin),
          "lnp max rel %.2e" % rel, "acc equal:", np.array_equal(a[2], b[2]))
    print("   ", a[3][:90], "|", b[3][:60])
print("worst rel", worst)
