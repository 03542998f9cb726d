"""Helper for test_two_rank_gloo_sharded_run_matches_single_process (CPU, gloo)."""
import sys

import numpy as np
import torch.distributed as dist

from mimic_mda_b200 import shard, synth
from mimic_mda_b200.sampler import BatchedEnsembleSampler
from oracle import numpy_oracle


def main():
    dist.init_process_group("gloo")
    rank, _, world = shard.rank_and_world()
    n_bins = 5
    mine = shard.bins_for_rank(n_bins, rank, world)
    consts, n_pts, a_true = synth.make_bins(len(mine), 3, 16, first_bin=rank, bin_step=world)
    lo, hi = synth.default_bounds(3)
    all_consts, _, all_true = synth.make_bins(n_bins, 3, 16)
    p0 = synth.make_walkers(all_true, 12, seed=2, oob_frac=0.0)[mine]
    smp = BatchedEnsembleSampler(len(mine), 12, 3,
                                 lambda p: numpy_oracle.lnpost_bins(consts, n_pts, lo, hi, p),
                                 seeds=[40 + int(b) for b in mine])
    smp.run_mcmc(p0, 8)
    chains = shard.gather_results([smp.chain[k] for k in range(len(mine))], n_bins)
    if rank == 0:
        np.save(sys.argv[1] + "_rank0.npy", np.stack(chains))
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
