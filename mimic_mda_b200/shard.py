"""Partition of Ex bins over GPUs: one process per GPU, no data-path collective.

The reference farms Ex bins to OS processes with ``multiprocessing.Pool.map``
(mda.py:182-188): bins are independent MCMC problems and the only "gather" is
the pickled per-bin result tuples coming back to the parent.  Here bin ``i``
goes to rank ``i mod G`` (round-robin, so neighbouring bins -- which have
similar angle counts -- spread evenly), each rank owns its bins' device store,
stream, graphs and sampler state, and results are gathered on the host.
Walkers of one bin are never split across GPUs: they interact every half-step
through the stretch move.
"""
import os

import numpy as np


def rank_and_world():
    """(rank, local_rank, world_size) from the torchrun environment (1 process = 1 GPU)."""
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", str(rank)))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    return rank, local_rank, world


def bind_to_gpu_numa_node(device_index):
    """Best effort, before any pinned allocation: run this process on the CPUs of the NUMA node the GPU hangs off, so that
    first-touch places the pinned chain buffers there (with one process per GPU all chains otherwise cross whichever socket
    the processes happen to start on).  Returns a short description of what was done; never raises."""
    try:
        import subprocess
        bus = subprocess.run(["nvidia-smi", "-i", str(device_index), "--query-gpu=pci.bus_id", "--format=csv,noheader"],
                             capture_output=True, text=True, timeout=20).stdout.strip().lower()
        if not bus:
            return "no pci bus id"
        bus = bus[-12:]                                    # 00000000:1B:00.0 -> 0000:1b:00.0
        with open("/sys/bus/pci/devices/%s/numa_node" % bus) as fh:
            node = int(fh.read().strip())
        if node < 0:
            return "gpu %s: numa node unknown (-1)" % bus
        with open("/sys/devices/system/node/node%d/cpulist" % node) as fh:
            cpus = set()
            for part in fh.read().strip().split(","):
                lo, _, hi = part.partition("-")
                cpus.update(range(int(lo), int(hi or lo) + 1))
        allowed = os.sched_getaffinity(0) & cpus
        if not allowed:
            return "gpu %s: node %d has no allowed cpus" % (bus, node)
        os.sched_setaffinity(0, allowed)
        return "gpu %s: bound to %d cpus of numa node %d" % (bus, len(allowed), node)
    except Exception as exc:                               # an unusual sysfs / no nvidia-smi: stay where we are
        return "not bound (%s)" % type(exc).__name__


def bins_for_rank(n_bins, rank, world):
    """Global bin indices owned by ``rank``: rank, rank + world, rank + 2*world, ..."""
    if world < 1 or not 0 <= rank < world:
        raise ValueError("bad rank/world: %r/%r" % (rank, world))
    return np.arange(rank, n_bins, world, dtype=np.int64)


def owner_of_bin(bin_index, world):
    return int(bin_index) % int(world)


def merge_by_bin(n_bins, world, per_rank_results):
    """Inverse of ``bins_for_rank``: ``per_rank_results[r][k]`` is the result for
    global bin ``r + k*world``; returns the list ordered by global bin, like the
    list ``Pool.map`` returns at mda.py:188."""
    merged = [None] * n_bins
    for r in range(world):
        mine = bins_for_rank(n_bins, r, world)
        if len(per_rank_results[r]) != len(mine):
            raise ValueError("rank %d returned %d results for %d bins" % (r, len(per_rank_results[r]), len(mine)))
        for k, b in enumerate(mine):
            merged[int(b)] = per_rank_results[r][k]
    return merged


def gather_results(local_results, n_bins, group=None):
    """Gather per-bin Python results from all ranks to every rank over
    torch.distributed (host-side object gather; gloo or nccl world).  With no
    initialised process group this is the single-process identity."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return list(local_results)
    world = dist.get_world_size(group)
    gathered = [None] * world
    dist.all_gather_object(gathered, list(local_results), group=group)
    return merge_by_bin(n_bins, world, gathered)
