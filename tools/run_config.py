"""Run a reference configuration file end to end on the GPU(s).

    python tools/run_config.py config.py [--seed N] [--out results.json]
    torchrun --nproc-per-node G tools/run_config.py config.py      # bins round-robin over G GPUs

A convenience driver, outside the hot-path scope and therefore not part of the package.

``config.py`` is the reference's configuration format: a Python file that fills a ``CONFIG``
dictionary (mda.py:50-53, config_example.py).  The keys on the likelihood path are honoured
(input files, energy window, Max Theta, Maximum L, EWSR Fractions, IVGDR subtraction, start
grid, walkers, sample points, burn-in, confidence interval, Num Bins, Calc AutoCorr / ACorr WindSize /
ACorr Use FFT); plotting, directory
and CSV-format keys are ignored -- this driver prints / saves the per-bin estimates that
``fit_and_mcmc`` returns (mda.py:1318) instead of the reference's plot and CSV tree.
"""
import argparse
import json
import sys

import os

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mimic_mda_b200 import builder, chisq, fits, pipeline, shard  # noqa: E402


def load_config(path):
    scope = {}
    with open(path) as fh:
        exec(compile(fh.read(), path, "exec"), scope)          # mda.py:50-53
    return scope["CONFIG"]


def main(argv=None):
    ap = argparse.ArgumentParser(description=__doc__.split("\n")[0])
    ap.add_argument("config")
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--out", default="")
    ap.add_argument("--thin", type=int, default=1)
    ap.add_argument("--group-size", type=int, default=0, help="bins sampled at once (default: what the free HBM holds)")
    args = ap.parse_args(argv)
    config = load_config(args.config)
    rank, _, world = shard.rank_and_world()
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("gloo")                        # host-side gather only; no data-path collective
    inputs = builder.build_inputs(config)
    n_bins = len(inputs["energies"])
    mine = shard.bins_for_rank(n_bins, rank, world)
    n_l = config["Maximum L"] + 1
    starts = fits.calc_start_params([config["Start Pts a%d" % i] for i in range(n_l)])
    # bins in groups sized to the free HBM: the reference's own sizes (2500 walkers x 16385 samples) are ~3 GB of chain per bin
    results = pipeline.fit_and_mcmc_grouped(inputs["consts"][mine], inputs["n_pts"][mine], inputs["bounds"], starts, config,
                                            ewsr=config["EWSR Fractions"][:n_l], seed=args.seed, bin_ids=mine,
                                            rng=np.random.RandomState(args.seed + 7919 * rank), thin=args.thin,
                                            energies=[inputs["energies"][b] for b in mine], group_size=args.group_size or None)
    merged = shard.gather_results(results, n_bins)
    if rank == 0:
        table = []
        for b, (values, (acor, chis, acc)) in enumerate(merged):
            row = {"Ex": inputs["energies"][b], "n_points": int(inputs["n_pts"][b]), "acceptance_fraction": acc,
                   "chi2_percentile": chis[0], "chi2_peak": chis[1]}
            if len(acor):                                      # the diagnostics columns of mda.py:266-272
                samples = config["Number of Walkers"] * (config["Sample Points"] - config["Burn-in Points"])
                row["autocorr_times"] = [float(t) for t in acor]
                row["num_ind_samples"] = float(samples) / max(acor)
                row["dist_error"] = 1.0 / np.sqrt(row["num_ind_samples"])
            for l in range(n_l):
                row["a%d" % l] = {"percentile": list(values[0][l]), "peak": list(values[1][l])}
            table.append(row)
            print("Ex %6.2f MeV  N=%3d  acc=%.3f  chi2=%.3f  " % (row["Ex"], row["n_points"], acc, chis[0]) +
                  "  ".join("a%d=%.4f+%.4f-%.4f" % ((l,) + tuple(values[0][l])) for l in range(n_l)))
        if args.out:
            with open(args.out, "w") as fh:
                json.dump(table, fh, indent=1)
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
