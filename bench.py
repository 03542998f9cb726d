#!/usr/bin/env python
"""bench.py -- likelihood evals/s (walker x Ex-bin) of the batched chi-square
log-posterior path, with its roofline and the reference's CPU path beside it.

    python bench.py --gpus N --steps K --warmup W [--impl reference]

Launched by the driver as ``python -m torch.distributed.run --nproc-per-node N ...
bench.py --gpus N ...`` for N > 1: one process per GPU, Ex bins partitioned
round-robin over ranks, no data-path collective; torch.distributed is
used for the barrier and the max-over-ranks of the device timings only.

Workload (BASELINE.json configs[3], the one the metric is quoted on): per GPU
200 Ex bins x 512 walkers x 9 distributions x 60 angles, fp64, synthetic seeded inputs
(mimic_mda_b200/synth.py).  A *step* is one ensemble step of every bin: two dependent
half-ensemble launches of [bins][256][9] parameter vectors (the two halves of the
stretch move, mda.py:1128-1131), i.e. bins x 512 likelihood evaluations.

  value  evals/s of the device-resident ensemble sampler (mdaEnsembleAdvance): K steps of
         every bin's 512 walkers in ONE persistent launch -- proposals, box prior,
         likelihood, accept/reject and the chain all stay in HBM/shared memory; CUDA events
         on the launching stream.
  e2e    the same sampler through the C-ABI call with HOST buffers (mdaEnsembleSetState +
         mdaEnsembleRunToHost): start positions come from pinned host memory and the whole
         chain + log-probabilities are delivered to pinned host memory, chunk c+1 computed
         while chunk c crosses PCIe.  This is what emcee's run_mcmc leaves on the host.
  batched_half_step  the host-driven alternative (one launch per half-ensemble, parameters
         from the host each time): resident ring > 2x L2, and staged-graph e2e.
  roofline       the sampler kernel against the fp64 pipe (measured live by a DFMA
                 microbenchmark) and against HBM (MEASURED_PEAKS.json); the kernel and its
                 schedule are named by the library (mdaEnsembleDescribe).  The default stepper
                 takes chi^2 from the bin's triangular factor (55 instead of 600 multiply-adds
                 at C4): its headline fraction is the HBM one (the chain going out); the
                 ALGORITHMIC flop rate of SURVEY.md 8d and the flops it executes are sub-objects.
                 direct_evaluation is the same run on the angle-by-angle DMMA stepper
                 (MDA_ENS_IMPL=direct), whose roofline is the fp64 pipe.
  cpu_baseline   the unmodified reference library (oracle/_ref) through the Python
                 ctypes path of mda.py:1540-1578 on all host cores (N = 1, rank 0).
"""
import argparse
import ctypes as ct
import json
import multiprocessing
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from mimic_mda_b200 import shard, synth  # noqa: E402

METRIC = "likelihood evals/sec (walker x Ex-bin)"
UNIT = "evals/s"
L2_BYTES = 126 * 1024 * 1024


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2000)
    ap.add_argument("--warmup", type=int, default=50)
    ap.add_argument("--impl", choices=["b200", "reference"], default="b200")
    ap.add_argument("--workload", default="C4", help="C2 | C3 | C4 (default: the metric's config)")
    ap.add_argument("--bins", type=int, default=0, help="override Ex bins per GPU")
    ap.add_argument("--walkers", type=int, default=0, help="override walkers per bin")
    ap.add_argument("--n-l", type=int, default=0)
    ap.add_argument("--n-pts", type=int, default=0)
    ap.add_argument("--tile", default="", help="force kernel tiling THREADSxWPT, e.g. 64x2")
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="CPU work budget of the baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-strong", action="store_true", help="skip the 200-bins-total strong-scaling companion run")
    ap.add_argument("--no-configs", action="store_true", help="skip the C2 / C3 / C5-corner companion measurements")
    return ap.parse_args()


def workload_shape(args):
    cfg = dict(synth.CONFIGS[args.workload])
    if args.bins:
        cfg["bins"] = args.bins
    if args.walkers:
        cfg["walkers"] = args.walkers
    if args.n_l:
        cfg["n_l"] = args.n_l
    if args.n_pts:
        cfg["n_pts"] = args.n_pts
    return cfg


# ------------------------------------------------------------------------------------------
# clocks during the timed region (B200_PROFILING.md, "clocks line")
# ------------------------------------------------------------------------------------------
class ClockSampler:
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows = []
        self.proc = None
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(index), "--query-gpu=" + self.FIELDS, "--format=csv,noheader,nounits", "-lms", "50"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"], "samples": 0}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, power, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for row in self.rows:
            parts = [p.strip() for p in row.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0]))
                mx.append(float(parts[1]))
                power.append(float(parts[2]))
            except ValueError:
                continue
            for name, flag in zip(names, parts[3:7]):
                if flag.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(power) if power else None, "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------
# CPU arm: the reference's own implementation of the path on the host cores
# ------------------------------------------------------------------------------------------
_CPU_STATE = {}


def _cpu_worker_init(lib_path, consts, n_pts, lo, hi, params):
    """Runs inside each pool worker: the library is loaded in the worker, as the
    reference does (mda.py:1109), and one struct per bin is built on first use."""
    from oracle import reflib
    _CPU_STATE.update(lib=reflib.prepare_shared_object(lib_path), reflib=reflib, consts=consts, n_pts=n_pts,
                      bounds=[(float(l), float(h)) for l, h in zip(lo, hi)], params=params)


def _cpu_worker_bin(task):
    """One Ex bin: emcee's per-walker calls of ln_post_prob (mda.py:1540-1578)."""
    b, reps = task
    st = _CPU_STATE
    reflib, lib = st["reflib"], st["lib"]
    n = int(st["n_pts"][b])
    n_l = st["params"].shape[2]
    struct = reflib.make_calc_struct(lib, st["consts"][b, 0, :n], [st["consts"][b, 1 + l, :n] for l in range(n_l)])
    rows = [st["params"][b, w] for w in range(st["params"].shape[1])]
    bounds = st["bounds"]
    t0 = time.perf_counter()
    acc = 0.0
    for _ in range(reps):
        for row in rows:
            v = reflib.ln_post_prob(row, lib, struct, bounds)
            if v > -1e300:
                acc += v
    dt = time.perf_counter() - t0
    lib.freeMdaStruct(struct)
    return dt, acc


def cpu_reference_arm(cfg, cpu_seconds, consts=None, n_pts=None, params=None):
    """Times the reference path on a bounded sample of the workload.  Returns a dict."""
    from oracle import reflib
    if not os.path.exists(reflib.ORACLE_LIB) or not os.path.exists(reflib.REF_LOOP_LIB):
        reflib.build()
    lib_path = reflib.reference_lib_path()
    kind = "reference"
    if lib_path is None:                      # the reference build did not travel / does not run here
        lib_path, kind = reflib.ORACLE_LIB, "port"
    cores = os.cpu_count() or 1
    n_l, n_ang, walkers = cfg["n_l"], cfg["n_pts"], cfg["walkers"]
    sample_bins = min(cfg["bins"], max(cores, 8))
    if consts is None:
        consts, n_pts, a_true = synth.make_bins(sample_bins, n_l, n_ang)
        params = synth.make_walkers(a_true, walkers, seed=7)
    else:
        consts, n_pts, params = consts[:sample_bins], n_pts[:sample_bins], params[:sample_bins]
    lo, hi = synth.default_bounds(n_l)
    # calibrate on one bin in this process, then size the sample to ~cpu_seconds of CPU work
    _cpu_worker_init(lib_path, consts, n_pts, lo, hi, params)
    dt1, _ = _cpu_worker_bin((0, 1))
    per_eval = dt1 / walkers
    reps = max(1, int(cpu_seconds / max(per_eval * walkers * sample_bins, 1e-9)))
    tasks = [(b, reps) for b in range(sample_bins)]
    ctx = multiprocessing.get_context("fork")
    with ctx.Pool(processes=min(cores, sample_bins), initializer=_cpu_worker_init,
                  initargs=(lib_path, consts, n_pts, lo, hi, params)) as pool:
        pool.map(_cpu_worker_bin, [(b, 1) for b in range(sample_bins)])          # warm-up
        t0 = time.perf_counter()
        pool.map(_cpu_worker_bin, tasks)
        wall = time.perf_counter() - t0
    evals = sample_bins * walkers * reps
    # secondary honesty line: the same library in a tight C loop (no Python in the way)
    loop = reflib.RefLoop(lib_path)
    _, s1 = loop.run(consts, n_pts, lo, hi, params, n_threads=1, repeats=1)
    c_reps = max(1, int(1.0 / max(s1, 1e-6)))
    _, s_one = loop.run(consts, n_pts, lo, hi, params, n_threads=1, repeats=c_reps)
    _, s_all = loop.run(consts, n_pts, lo, hi, params, n_threads=min(cores, sample_bins), repeats=c_reps * min(cores, sample_bins))
    n_c = sample_bins * walkers
    return {
        "value": evals / wall, "unit": UNIT, "cores": min(cores, sample_bins), "kind": kind,
        "sample": "%d of %d bins x %d walkers x %d passes, L=%d N=%d; per-walker ln_post_prob -> ctypes -> "
                  "calculateLnLiklihood (mda.py:1540-1578), multiprocessing.Pool over bins (mda.py:186-188)"
                  % (sample_bins, cfg["bins"], walkers, reps, n_l, n_ang),
        "library": os.path.relpath(lib_path, ROOT),
        "seconds": wall, "evals": evals,
        "per_core_python_evals_per_s": 1.0 / per_eval,
        "pure_c_evals_per_s_1core": n_c * c_reps / s_one,
        "pure_c_evals_per_s_allcores": n_c * c_reps * min(cores, sample_bins) / s_all,
        "host_cpu": _cpu_model(), "host_cores_total": cores,
    }


def _cpu_model():
    try:
        with open("/proc/cpuinfo") as fh:
            for line in fh:
                if line.startswith("model name"):
                    return line.split(":", 1)[1].strip()
    except OSError:
        pass
    return "unknown"


def run_reference_impl(args, cfg, rank, world):
    """--impl reference: the reference's CPU implementation, same metric/config/JSON line."""
    if rank != 0:
        return
    base = cpu_reference_arm(cfg, args.cpu_seconds)
    evals_step = cfg["bins"] * cfg["walkers"] * world
    line = {
        "impl": "reference", "metric": METRIC, "value": base["value"], "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * evals_step / base["value"],
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(cfg, world),
        "note": "reference CPU path on host cores; bounded sample, see cpu_baseline.sample",
        "cpu_baseline": base,
        "e2e": {"value": base["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def workload_config(cfg, world):
    """Identical in both arms (the driver compares them); what an arm does with it goes into the line's "note"."""
    return {
        "workload": "%s: %d Ex bins/GPU x %d walkers x %d dists x %d angles (BASELINE.json configs[%s])"
                    % (cfg.get("name", "C4"), cfg["bins"], cfg["walkers"], cfg["n_l"], cfg["n_pts"],
                       {"C1": 0, "C2": 1, "C3": 2, "C4": 3}.get(cfg.get("name", "C4"), "?")),
        "bins_per_gpu": cfg["bins"], "bins_total": cfg["bins"] * world, "walkers": cfg["walkers"],
        "dists": cfg["n_l"], "angles": cfg["n_pts"], "walkers_per_launch": cfg["walkers"] // 2,
        "launches_per_step": 2, "partition": "bins round-robin over ranks, no collective",
        "l2_policy": "sampler: state lives in shared memory, chain streams out (nothing to re-read); "
                     "batched_half_step: ring of distinct parameter blocks > 2x126 MB",
    }


# ------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------
class Dist:
    """torch.distributed used as plumbing only: barrier + max over ranks."""

    def __init__(self, world, local_rank):
        self.world = world
        self.torch = None
        if world > 1:
            import torch
            import torch.distributed as dist
            self.torch, self.dist = torch, dist
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
                os.environ["NCCL_DEBUG"] = "WARN"          # keep stdout to the one JSON line:
            os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")   # NCCL's own lines (its version banner) go to stderr
            try:
                torch.cuda.set_device(local_rank)
                dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
                self.device = torch.device("cuda", local_rank)
            except Exception:
                dist.init_process_group("gloo")
                self.device = torch.device("cpu")

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()

    def max(self, value):
        if self.world == 1:
            return value
        t = self.torch.tensor([value], dtype=self.torch.float64, device=self.device)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def sum(self, value):
        if self.world == 1:
            return value
        t = self.torch.tensor([value], dtype=self.torch.float64, device=self.device)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM)
        return float(t.item())

    def close(self):
        if self.world > 1:
            self.dist.barrier()
            self.dist.destroy_process_group()


class GpuWorkload:
    """One rank's share: store in HBM, ring of resident parameter blocks, pinned staging."""

    def __init__(self, lib, cfg, rank, world, n_bins_local, first_bin, bin_step, tile):
        from mimic_mda_b200 import chisq
        self.lib, self.chisq, self.cfg = lib, chisq, cfg
        self.bins, self.n_l, self.n_ang = n_bins_local, cfg["n_l"], cfg["n_pts"]
        self.walkers, self.half = cfg["walkers"], cfg["walkers"] // 2
        self.consts, self.n_pts, self.a_true = synth.make_bins(self.bins, self.n_l, self.n_ang,
                                                               first_bin=first_bin, bin_step=bin_step)
        self.lo, self.hi = synth.default_bounds(self.n_l)
        self.store = chisq.DeviceStore(self.consts, self.n_pts, (self.lo, self.hi), cs_lib=lib)
        if tile:
            t, w = tile.lower().split("x")
            self.store.set_tile(int(t), int(w))
        self.params = synth.make_walkers(self.a_true, self.walkers, seed=7 + rank)     # [B, W, L]
        self.block_bytes = self.bins * self.half * self.n_l * 8
        self.out_bytes = self.bins * self.half * 8
        n_ring = max(4, int(2 * L2_BYTES / max(self.block_bytes, 1)) + 1)
        self.n_ring = n_ring + (n_ring & 1)
        self.d_params, self.d_out = [], []
        rng = np.random.default_rng(99 + rank)
        halves = [np.ascontiguousarray(self.params[:, :self.half]), np.ascontiguousarray(self.params[:, self.half:])]
        for i in range(self.n_ring):
            blk = halves[i & 1]
            if i >= 2:      # distinct contents per block: jitter inside the box
                blk = np.clip(blk + 1e-3 * rng.standard_normal(blk.shape), self.lo, self.hi)
            dp = lib.mdaDeviceAlloc(self.block_bytes)
            do = lib.mdaDeviceAlloc(self.out_bytes)
            if not dp or not do:
                raise chisq.MdaError(lib.mdaLastError(), lib.mdaLastErrorString().decode())
            chisq.check(lib, lib.mdaMemcpyH2D(dp, blk.ctypes.data, self.block_bytes))
            self.d_params.append(dp)
            self.d_out.append(do)
        self.ev0, self.ev1 = lib.mdaEventCreate(), lib.mdaEventCreate()
        self.stage_p, self.stage_o = self.store.staging(self.half)
        self.halves = halves
        # device-resident sampler: global bin ids make the chains independent of the partition
        from mimic_mda_b200.ensemble import DeviceEnsemble
        self.bin_ids = first_bin + bin_step * np.arange(self.bins)
        self.ens = DeviceEnsemble(self.store, self.walkers, seed=20261019, bin_ids=self.bin_ids)
        self.p0 = synth.make_walkers(self.a_true, self.walkers, seed=7 + rank, oob_frac=0.0)
        self.pinned = []

    def pinned_array(self, shape):
        n = int(np.prod(shape))
        ptr = self.lib.mdaHostAlloc(max(n, 1) * 8)
        if not ptr:
            raise self.chisq.MdaError(self.lib.mdaLastError(), self.lib.mdaLastErrorString().decode())
        self.pinned.append(ptr)
        buf = (ct.c_double * max(n, 1)).from_address(ptr)
        return np.frombuffer(buf, dtype=np.float64, count=n).reshape(shape)

    def sampler_resident(self, steps, warmup):
        """K ensemble steps in one persistent launch, chain kept in HBM; elapsed ms (events)."""
        lib = self.lib
        self.ens.set_state(self.p0)
        thin = max(1, -(-steps // 2048))                   # keep at most 2048 samples per walker in HBM
        self.ens.reserve_chain(max(warmup, 1) // thin + steps // thin + 2)
        if warmup > 0:
            self.ens.advance(warmup, thin)
        # More untimed steps, enqueued without waiting: the device is still busy with them while the host enqueues the start
        # event, the timed launch and the stop event, so the events bracket the K steps on the device and not the host's
        # launch latency (~7 us of a 70 us region at K = 20).  Nothing of them is kept (thin beyond their length).
        # (150 steps ~ 350 us: the clock-sampling thread may hold the interpreter while this thread wants to enqueue)
        self.ens.advance(max(warmup, 150), 10 ** 6, wait=False)
        self.chisq.check(lib, lib.mdaEventRecord(self.ev0, self.store.handle))
        self.ens.advance(steps, thin, wait=False)
        self.chisq.check(lib, lib.mdaEventRecord(self.ev1, self.store.handle))
        self.chisq.check(lib, lib.mdaEventSynchronize(self.ev1))
        self.ens.synchronize()
        ms = ct.c_float()
        self.chisq.check(lib, lib.mdaEventElapsedMs(self.ev0, self.ev1, ct.byref(ms)))
        return float(ms.value), thin

    def sampler_e2e(self, steps, host_p0, host_chain, host_lnp):
        """Host buffers in, host buffers out; wall seconds."""
        self.store.synchronize()
        t0 = time.perf_counter()
        self.ens.set_state(host_p0)
        self.ens.run_to_host(steps, 1, 0, host_chain, host_lnp)
        return time.perf_counter() - t0

    def sampler_summary_e2e(self, steps, host_p0, burn_frac=0.25):
        """Host start positions in, posterior summaries out (percentile and histogram-peak
        estimates per bin and parameter, acceptance fractions); the chain never leaves HBM."""
        self.store.synchronize()
        t0 = time.perf_counter()
        self.ens.set_state(host_p0)
        self.ens.reserve_chain(steps)
        self.ens.advance(steps, 1)
        summ = self.ens.summarize(int(steps * burn_frac), 0.682689492, 1000)
        secs = time.perf_counter() - t0
        return secs, summ

    def sampler_parity(self, oracle, steps=3):
        """The sampler's own log-probabilities against the oracle on the positions it visited."""
        self.ens.run_mcmc(self.p0, steps)
        chain = np.ascontiguousarray(self.ens.chain[:, :, -1, :])          # [B, W, L] after the last step
        got = self.ens.lnprobability[:, :, -1]
        want = oracle.lnpost_bins(self.consts, self.n_pts, self.lo, self.hi, chain)
        fin = np.isfinite(want)
        return float(np.max(np.abs(got[fin] - want[fin]) / np.abs(want[fin])))

    def evals_per_step(self):
        return self.bins * self.walkers

    def resident(self, steps):
        """Enqueue 2*steps launches over the ring; returns elapsed ms (CUDA events)."""
        lib = self.lib
        self.chisq.check(lib, lib.mdaEventRecord(self.ev0, self.store.handle))
        self.store.lnpost_device_many(self.half, self.d_params, self.d_out, 2 * steps)
        self.chisq.check(lib, lib.mdaEventRecord(self.ev1, self.store.handle))
        self.chisq.check(lib, lib.mdaEventSynchronize(self.ev1))
        ms = ct.c_float()
        self.chisq.check(lib, lib.mdaEventElapsedMs(self.ev0, self.ev1, ct.byref(ms)))
        return float(ms.value)

    def e2e(self, steps):
        """steps x 2 staged calls (H2D from pinned host, kernel, D2H, sync); wall seconds."""
        self.store.synchronize()
        t0 = time.perf_counter()
        acc = 0.0
        for _ in range(steps):
            for h in (0, 1):
                self.store.lnpost_staged(self.half)
                acc += self.stage_o[0, 0]                 # the step's result is read on the host
        self.store.synchronize()
        return time.perf_counter() - t0, acc

    def check_against(self, oracle):
        """Parity gate in the same run: resident block 0/1 and the staged path vs the oracle."""
        want = [oracle.lnpost_bins(self.consts, self.n_pts, self.lo, self.hi, h) for h in self.halves]
        worst = 0.0
        for i in (0, 1):
            got = np.empty((self.bins, self.half))
            self.chisq.check(self.lib, self.lib.mdaMemcpyD2H(got.ctypes.data, self.d_out[i], got.nbytes))
            if not np.array_equal(np.isneginf(got), np.isneginf(want[i])):
                raise AssertionError("-inf positions differ from the oracle")
            fin = np.isfinite(want[i])
            worst = max(worst, float(np.max(np.abs(got[fin] - want[i][fin]) / np.abs(want[i][fin]))))
        return worst

    def close(self):
        self.ens.close()
        for p in self.pinned:
            self.lib.mdaHostFree(p)
        for p in self.d_params + self.d_out:
            self.lib.mdaDeviceFree(p)
        self.lib.mdaEventDestroy(self.ev0)
        self.lib.mdaEventDestroy(self.ev1)
        self.store.close()


def pinned_block(lib, chisq, shape, keep):
    n = int(np.prod(shape))
    ptr = lib.mdaHostAlloc(max(n, 1) * 8)
    if not ptr:
        raise chisq.MdaError(lib.mdaLastError(), lib.mdaLastErrorString().decode())
    keep.append(ptr)
    return np.frombuffer((ct.c_double * max(n, 1)).from_address(ptr), dtype=np.float64, count=n).reshape(shape)


def pcie_peaks(lib, chisq, mbytes=64):
    """Measured roof of the e2e paths: one cudaMemcpy of `mbytes` MB between pinned host memory and HBM, each way
    (best of 4, wall clock around the synchronous copy)."""
    n = mbytes << 20
    keep = []
    host = pinned_block(lib, chisq, (n // 8,), keep)
    host[...] = 1.0
    dev = lib.mdaDeviceAlloc(n)
    out = {}
    for name, fn, a, b in (("d2h_gbs", lib.mdaMemcpyD2H, host.ctypes.data, dev), ("h2d_gbs", lib.mdaMemcpyH2D, dev, host.ctypes.data)):
        chisq.check(lib, fn(a, b, n))
        best = 1e30
        for _ in range(4):
            t0 = time.perf_counter()
            chisq.check(lib, fn(a, b, n))
            best = min(best, time.perf_counter() - t0)
        out[name] = n / best / 1e9
    lib.mdaDeviceFree(dev)
    lib.mdaHostFree(keep[0])
    out["how"] = "one cudaMemcpy of %d MB, pinned host <-> HBM, best of 4" % mbytes
    return out


def scalar_call_latency(lib, chisq, calls=300):
    """The seven-symbol path: microseconds per calculateLnLiklihood call through ctypes (a launch + a sync each), next
    to the same call on the CPU library the reference ships (oracle/_ref, or the oracle port when that did not travel)."""
    from oracle import reflib
    consts, n_pts, a_true = synth.make_bins(1, 9, 60)
    data, dists, a = consts[0, 0, :60], [consts[0, 1 + l, :60] for l in range(9)], np.ascontiguousarray(a_true[0])
    out = {}
    cpu_path = reflib.reference_lib_path() or reflib.ORACLE_LIB
    for name, handle, mod in (("gpu_us", lib, chisq), ("cpu_us", reflib.prepare_shared_object(cpu_path), reflib)):
        st = mod.make_calc_struct(handle, data, dists)
        ptr = a.ctypes.data_as(ct.POINTER(ct.c_double))
        for _ in range(20):
            handle.calculateLnLiklihood(st, ptr)
        t0 = time.perf_counter()
        for _ in range(calls):
            handle.calculateLnLiklihood(st, ptr)
        out[name] = (time.perf_counter() - t0) / calls * 1e6
        handle.freeMdaStruct(st)
    out["what"] = ("calculateLnLiklihood(L=9, N=60) per call through ctypes: the GPU library pays a kernel launch per call (parameters in "
                   "the argument block, result and completion ticket written to mapped host memory, no stream synchronisation), so "
                   "unmodified mda.py pointed at it is slower per call than the CPU library -- the batched entry points exist for that")
    return out


def batched_point(lib, chisq, orc, bins, walkers, n_l, n_pts, launches, peak_tf, hbm_peak):
    """The batched likelihood kernel at one shape: ring of distinct parameter blocks larger than 2 x L2 (at least two),
    CUDA events around `launches` launches; roofline per SURVEY.md 8d; parity of a sample against the oracle."""
    consts, counts, a_true = synth.make_bins(bins, n_l, n_pts)
    lo, hi = synth.default_bounds(n_l)
    store = chisq.DeviceStore(consts, counts, (lo, hi), cs_lib=lib)
    blk = synth.make_walkers(a_true, walkers, seed=3)
    n_ring = max(2, 2 * L2_BYTES // blk.nbytes + 1)
    dps, dos = [], []
    rng = np.random.default_rng(5)
    for i in range(n_ring):
        dp, do = lib.mdaDeviceAlloc(blk.nbytes), lib.mdaDeviceAlloc(bins * walkers * 8)
        src = blk if i < 2 else np.clip(blk + 1e-3 * rng.standard_normal((1, 1, n_l)), lo, hi)
        chisq.check(lib, lib.mdaMemcpyH2D(dp, src.ctypes.data, blk.nbytes))
        dps.append(dp)
        dos.append(do)
    ev0, ev1 = lib.mdaEventCreate(), lib.mdaEventCreate()
    store.lnpost_device_many(walkers, dps, dos, max(3, min(launches, 2 * n_ring)))
    store.synchronize()
    lib.mdaEventRecord(ev0, store.handle)
    store.lnpost_device_many(walkers, dps, dos, launches)
    lib.mdaEventRecord(ev1, store.handle)
    lib.mdaEventSynchronize(ev1)
    ms = ct.c_float()
    lib.mdaEventElapsedMs(ev0, ev1, ct.byref(ms))
    sec = ms.value * 1e-3 / launches
    # parity: block 0's results for the first bins / walkers against the oracle
    got = np.empty((bins, walkers))
    chisq.check(lib, lib.mdaMemcpyD2H(got.ctypes.data, dos[0], got.nbytes))
    nb, nw = min(bins, 4), min(walkers, 256)
    want = orc.lnpost_bins(consts[:nb], counts[:nb], lo, hi, np.ascontiguousarray(blk[:nb, :nw]))
    fin = np.isfinite(want)
    same_inf = bool(np.array_equal(np.isneginf(got[:nb, :nw]), np.isneginf(want)))
    err = float(np.max(np.abs(got[:nb, :nw][fin] - want[fin]) / np.abs(want[fin])))
    t, w, g = store.get_tile(walkers)
    for ptr in dps + dos:
        lib.mdaDeviceFree(ptr)
    lib.mdaEventDestroy(ev0)
    lib.mdaEventDestroy(ev1)
    store.close()
    flops = 2.0 * n_pts * (n_l + 1) * bins * walkers
    nbytes = 8.0 * (n_l + 1) * bins * (walkers + n_pts)
    f64, fhbm = flops / sec / 1e12 / peak_tf, nbytes / sec / 1e9 / hbm_peak
    kernel = ("lnpost_tile_kernel %dx%d" % (t, w) if w > 0 else ("lnpost_mma_kernel (fp64 DMMA)" if w == 0 else "lnpost_split_kernel G=%d" % -w))
    return {"shape": "%d bins x %d walkers x %d dists x %d angles" % (bins, walkers, n_l, n_pts), "kernel": kernel + ", %d CTAs" % g,
            "us_per_launch": sec * 1e6, "value": bins * walkers / sec, "unit": UNIT, "launches": launches, "ring_blocks": n_ring,
            "roofline": {"bound": "fp64" if f64 >= fhbm else "hbm", "frac": max(f64, fhbm), "frac_fp64": f64, "frac_hbm": fhbm,
                         "achieved": flops / sec / 1e12 if f64 >= fhbm else nbytes / sec / 1e9,
                         "peak": peak_tf if f64 >= fhbm else hbm_peak, "unit": "TFLOP/s" if f64 >= fhbm else "GB/s"},
            "parity_max_rel_err": err, "inf_positions_equal": same_inf}


def sampler_point(lib, chisq, orc, name, steps, hbm_peak, e2e_steps=48):
    """The device-resident sampler on another BASELINE config (one GPU's worth): resident value, e2e to pinned host
    memory, the chain's HBM fraction, parity of the visited positions."""
    from mimic_mda_b200.ensemble import DeviceEnsemble
    cfg = synth.CONFIGS[name]
    bins, walkers, n_l, n_pts = cfg["bins"], cfg["walkers"], cfg["n_l"], cfg["n_pts"]
    consts, counts, a_true = synth.make_bins(bins, n_l, n_pts)
    lo, hi = synth.default_bounds(n_l)
    store = chisq.DeviceStore(consts, counts, (lo, hi), cs_lib=lib)
    ens = DeviceEnsemble(store, walkers, seed=20261020)
    p0 = synth.make_walkers(a_true, walkers, seed=7, oob_frac=0.0)
    ens.set_state(p0)
    thin = max(1, -(-steps // 2048))
    ens.reserve_chain(steps // thin + 8)
    ens.advance(8, thin)
    ev0, ev1 = lib.mdaEventCreate(), lib.mdaEventCreate()
    ens.advance(40, 10 ** 6, wait=False)                  # untimed, keeps the device busy while the timed launch is enqueued
    lib.mdaEventRecord(ev0, store.handle)
    ens.advance(steps, thin, wait=False)
    lib.mdaEventRecord(ev1, store.handle)
    lib.mdaEventSynchronize(ev1)
    ens.synchronize()
    ms = ct.c_float()
    lib.mdaEventElapsedMs(ev0, ev1, ct.byref(ms))
    text = ens.describe(steps)
    # parity on the last kept sample
    chain = np.ascontiguousarray(ens.chain[:, :, -1, :])
    want = orc.lnpost_bins(consts, counts, lo, hi, chain)
    fin = np.isfinite(want)
    err = float(np.max(np.abs(ens.lnprobability[:, :, -1][fin] - want[fin]) / np.abs(want[fin])))
    # e2e: host start positions in, whole chain + log-probabilities out (pinned)
    keep = []
    e2e_steps = max(1, min(steps, e2e_steps))
    h_p0 = pinned_block(lib, chisq, p0.shape, keep)
    h_p0[...] = p0
    h_chain = pinned_block(lib, chisq, (e2e_steps, bins, walkers, n_l), keep)
    h_lnp = pinned_block(lib, chisq, (e2e_steps, bins, walkers), keep)
    ens.set_state(h_p0)
    ens.run_to_host(min(e2e_steps, 4), 1, 0, h_chain[:min(e2e_steps, 4)], h_lnp[:min(e2e_steps, 4)])
    store.synchronize()
    t0 = time.perf_counter()
    ens.set_state(h_p0)
    ens.run_to_host(e2e_steps, 1, 0, h_chain, h_lnp)
    secs = time.perf_counter() - t0
    ens.close()
    for ptr in keep:
        lib.mdaHostFree(ptr)
    lib.mdaEventDestroy(ev0)
    lib.mdaEventDestroy(ev1)
    store.close()
    sec = ms.value * 1e-3
    chain_bytes = 8.0 * (n_l + 1) * bins * walkers * (steps // thin)
    return {"shape": "%s: %d bins x %d walkers x %d dists x %d angles" % (name, bins, walkers, n_l, n_pts), "kernel": text.split(":")[0],
            "value": bins * walkers * steps / sec, "unit": UNIT, "steps": steps, "ms_per_step": ms.value / steps,
            "roofline": {"bound": "hbm", "frac": chain_bytes / sec / 1e9 / hbm_peak, "achieved": chain_bytes / sec / 1e9, "peak": hbm_peak,
                         "unit": "GB/s", "what": "the chain going out, 8 (L+1) bytes per evaluation kept (thin %d)" % thin},
            "e2e": {"value": bins * walkers * e2e_steps / secs, "unit": UNIT, "steps": e2e_steps,
                    "d2h_bytes_per_step": bins * walkers * (n_l + 1) * 8, "pcie_gbs": bins * walkers * (n_l + 1) * 8 * e2e_steps / secs / 1e9},
            "parity_max_rel_err": err}


def other_configs(lib, chisq, orc, steps, peak_tf, hbm_peak):
    """BASELINE.json's other configs in the driver's record: C2 and C3 (sampler resident + e2e, batched half-step) and four
    corners of the C5 sweep of the batched kernel at 200 bins x 8192 walkers."""
    out = {}
    for name in ("C2", "C3"):
        cfg = synth.CONFIGS[name]
        out[name] = {"sampler": sampler_point(lib, chisq, orc, name, steps, hbm_peak),
                     "batched_half_step": batched_point(lib, chisq, orc, cfg["bins"], cfg["walkers"] // 2, cfg["n_l"], cfg["n_pts"], 200, peak_tf, hbm_peak)}
    out["C5_batched_200x8192"] = [batched_point(lib, chisq, orc, 200, 8192, l, n, 12, peak_tf, hbm_peak)
                                  for (l, n) in ((4, 16), (5, 30), (9, 60), (12, 256))]
    return out


def measure_sampler(work, dist, steps, warmup, e2e_steps=192):
    """Device-resident sampler: resident (events) and host-buffer e2e (wall), max over ranks."""
    work.store.synchronize()
    dist.barrier()
    ms, thin = work.sampler_resident(steps, max(warmup, 3))
    work.store.synchronize()
    dist.barrier()
    ms = dist.max(ms)
    # The e2e run is always 192 steps (1.6 GB of chain per GPU at C4), whatever K is: one run_mcmc call of the reference is
    # 16 385 steps (mda_config0.py:103), so the per-call costs (start positions H2D, first log-probabilities, a pass over p0)
    # belong to a run, not to a step; at K = 20 they would be a quarter of a 4 ms region.
    host_p0 = work.pinned_array(work.p0.shape)
    host_p0[...] = work.p0
    host_chain = work.pinned_array((e2e_steps, work.bins, work.walkers, work.n_l))
    host_lnp = work.pinned_array((e2e_steps, work.bins, work.walkers))
    work.sampler_e2e(min(e2e_steps, 8), host_p0, host_chain[:min(e2e_steps, 8)], host_lnp[:min(e2e_steps, 8)])
    dist.barrier()
    secs = work.sampler_e2e(e2e_steps, host_p0, host_chain, host_lnp)
    dist.barrier()
    secs = dist.max(secs)
    work.sampler_summary_e2e(min(e2e_steps, 16), host_p0)                     # warm-up
    dist.barrier()
    sum_steps = 1024                                        # a fixed-length run, like e2e above (8.4 GB of chain in HBM at C4)
    sum_secs = []
    for _ in range(2):                                      # wall clock of a ~15 ms region: a host hiccup (seen once: 650 ms) is not the path
        secs_i, _ = work.sampler_summary_e2e(sum_steps, host_p0)
        dist.barrier()
        sum_secs.append(dist.max(secs_i))
    work.summary_e2e = (min(sum_secs), sum_steps)
    return ms, thin, secs, e2e_steps


def measure(work, dist, steps, warmup):
    """Barrier + sync on both sides, exactly `steps` timed steps, max over ranks."""
    work.resident(max(warmup, 3))
    work.store.synchronize()
    dist.barrier()
    ms = work.resident(steps)
    work.store.synchronize()
    dist.barrier()
    ms = dist.max(ms)
    # end to end: fill staging once per half (what the sampler writes proposals into)
    work.stage_p[...] = work.halves[0]
    work.e2e(max(min(warmup, 20), 3))
    dist.barrier()
    e2e_steps = steps if steps <= 200 else max(200, steps // 10)
    secs, _ = work.e2e(e2e_steps)
    dist.barrier()
    secs = dist.max(secs)
    return ms, secs, e2e_steps


def main():
    args = parse_args()
    rank, local_rank, world = shard.rank_and_world()
    cfg = workload_shape(args)
    cfg["name"] = args.workload
    if cfg["walkers"] % 2:
        raise SystemExit("walkers must be even (two half-ensembles)")
    if args.impl == "reference":
        run_reference_impl(args, cfg, rank, world)
        return

    os.environ.setdefault("LOCAL_RANK", str(local_rank))
    all_cpus = os.sched_getaffinity(0)
    numa = shard.bind_to_gpu_numa_node(local_rank) if os.environ.get("MDA_BENCH_NUMA", "1") != "0" else "off"
    from mimic_mda_b200 import chisq
    lib = chisq.prepare_shared_object()               # raises if the CUDA library is missing
    if lib.mdaDeviceCount() < 1:
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback")
    dist = Dist(world, local_rank)
    name = ct.create_string_buffer(128)
    sms, cc = ct.c_int(), ct.c_int()
    chisq.check(lib, lib.mdaDeviceInfo(name, 128, ct.byref(sms), ct.byref(cc)))

    # weak scaling: every GPU owns cfg["bins"] bins (global bin index = rank + k*world)
    work = GpuWorkload(lib, cfg, rank, world, cfg["bins"], rank, world, args.tile)
    tile = work.store.get_tile(work.half)

    peak_tf = ct.c_double()
    chisq.check(lib, lib.mdaMeasureFp64Peak(5, ct.byref(peak_tf)))
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    hbm_peak, hbm_src = 6650.0, "fallback"
    if os.path.exists(peaks_path):
        try:
            hbm_peak, hbm_src = float(json.load(open(peaks_path))["hbm_gbs"]), "measured"
        except Exception:
            pass

    launches0 = lib.mdaKernelLaunchCount()
    sampler = ClockSampler(lib.mdaGetDevice())
    # headline: the device-resident sampler (K steps = one persistent launch)
    s_ms, s_thin, s_secs, s_e2e_steps = measure_sampler(work, dist, args.steps, args.warmup)
    # the host-driven alternative: one launch per half-ensemble
    ms, e2e_secs, e2e_steps = measure(work, dist, args.steps, args.warmup)
    timed_launches = 2 * args.steps
    # keep the same load running while nvidia-smi samples (the timed regions are milliseconds long)
    t_end = time.perf_counter() + 1.0
    while time.perf_counter() < t_end:
        work.ens.advance(200, 10 ** 6)            # thin beyond the run: nothing is kept
    clocks = sampler.stop()
    total_launches = lib.mdaKernelLaunchCount() - launches0

    # parity gate on the very blocks that were timed
    from oracle import reflib
    if not os.path.exists(reflib.ORACLE_LIB):
        reflib.build()
    orc = reflib.OracleLib()
    worst = max(work.check_against(orc), work.sampler_parity(orc))
    if worst > 1e-12:
        raise SystemExit("bench.py: parity gate failed, max rel err %.3e" % worst)

    n_l, n_ang, bins, half = cfg["n_l"], cfg["n_pts"], cfg["bins"], work.half
    evals_step = work.evals_per_step() * world
    value = evals_step * args.steps / (s_ms * 1e-3)
    e2e_value = evals_step * s_e2e_steps / s_secs
    traffic = {}
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath):
        try:
            traffic = json.load(open(tpath))
        except Exception:
            traffic = {}

    def traffic_of(key):                     # DRAM bytes per launch from the ncu capture (per step x steps for the sampler)
        exact = "%s_%dstep_launch" % (key, args.steps)
        if exact in traffic:                 # a capture of a launch of exactly this length exists
            return traffic[exact]
        if key + "_per_step" in traffic:
            return traffic[key + "_per_step"] * args.steps
        return traffic.get(key)

    def traffic_source_of(key):
        if "%s_%dstep_launch" % (key, args.steps) in traffic:
            return ("dram__bytes_read.sum + dram__bytes_write.sum of an ncu --set full capture of a launch of this very length "
                    "(profiles/traffic.json names the capture)")
        return ("dram__bytes_read.sum + dram__bytes_write.sum of the ncu --set full capture named in profiles/traffic.json, per step x this "
                "launch's steps for the sampler (a capture of another launch length, scaled -- not a measurement of the timed launch)")

    def roofline_of(kernel, seconds_per_launch, flops_launch, bytes_launch, traffic_key, tensor_note, executed_flops=None):
        tf = flops_launch / seconds_per_launch / 1e12
        gbs = bytes_launch / seconds_per_launch / 1e9
        f64, fhbm = tf / peak_tf.value, gbs / hbm_peak
        extra = {}
        if executed_flops is not None:
            etf = executed_flops / seconds_per_launch / 1e12
            extra = {"fp64_executed": {"achieved_tflops": etf, "frac": etf / peak_tf.value, "flops_per_launch": executed_flops},
                     "note": "this kernel obtains chi^2 from the bin's triangular factor with (L+1)(L+2) + 5 L flops per evaluation "
                             "(fp64_executed) instead of the 2 N (L+1) of SURVEY.md 8d (fp64: the ALGORITHMIC rate, which the pipe peak "
                             "does not cap), so the roofline that can bind it is HBM: the chain, 8 (L+1) bytes per evaluation, going "
                             "out (achieved / frac).  What limits it before that: a bin's half-steps are sequential, so an SM steps "
                             "its bins one after the other at a half-step's latency (~1500 cycles at 512 x 9), most of it the "
                             "shared-memory pipe -- the stretch move gathers a RANDOM partner row per walker and half-step, ~3 "
                             "wavefronts per half-warp access instead of 1 (profiles/r02_ncu_full_*: 725 wavefronts per bin and "
                             "half-step, 43 % of them bank-collision replays) -- and the fp64 dependency chains of chi^2"}
        # the triangular-factor kernel executes a tenth of the algorithmic flops: the roofline that can bind it is HBM
        # (the chain going out); its algorithmic flop rate is reported in "fp64" below, not as the headline fraction
        by_flops = f64 >= fhbm and executed_flops is None
        return {**extra, **{
            "bound": "fp64" if by_flops else "hbm",
            "achieved": tf if by_flops else gbs,
            "peak": peak_tf.value if by_flops else hbm_peak,
            "unit": "TFLOP/s" if by_flops else "GB/s",
            "frac": f64 if by_flops else fhbm, "traffic": traffic_of(traffic_key),
            "traffic_source": traffic_source_of(traffic_key),
            "kernel": kernel, "launch_us": seconds_per_launch * 1e6,
            "flops_per_launch": flops_launch, "bytes_per_launch": bytes_launch,
            "fp64": {"achieved_tflops": tf, "peak_tflops": peak_tf.value, "frac": f64,
                     "peak_source": "DFMA microbenchmark measured in this run (mdaMeasureFp64Peak); MEASURED_PEAKS.json has no fp64 figure"},
            "hbm": {"achieved_gbs": gbs, "peak_gbs": hbm_peak, "frac": fhbm, "peak_source": hbm_src + " (MEASURED_PEAKS.json hbm_gbs)"},
            "tensor_cores": tensor_note,
        }}

    # sampler launch: 2 N (L+1) flop per evaluation; HBM: kept chain + lnprob out,
    # constants + state in/out once
    s_flops = 2.0 * n_ang * (n_l + 1) * bins * cfg["walkers"] * args.steps
    # 8 (L+1) bytes per evaluation (here: per kept sample, the chain and its log-probability going out) plus the
    # bins' constants once per launch.  The state that comes in and goes back (2 x 8 (L+1) bytes per walker per launch) is real
    # traffic too but not part of 8d's figure: left out of the fraction, so that it is the one a reader recomputes.
    s_bytes = 8.0 * (n_l + 1) * bins * (cfg["walkers"] * (args.steps // s_thin) + n_ang)
    dmma_note = ("fp64 tensor path: mma.sync m8n8k4 f64 (DMMA.8x8x4) for 8 of the 9 distributions, DFMA for the rest; it shares the "
                 "DFMA pipe and its peak (measured: 36.8 vs 37.1 TFLOP/s) and holds the scheduler's issue port while it runs, so the "
                 "denominator is the same fp64 peak; tcgen05 has no fp64 kind")
    kernel_text = work.ens.describe(args.steps)
    is_quad = kernel_text.startswith(("ensemble_quad_kernel", "ensemble_qws_kernel"))
    roofline = roofline_of(kernel_text + "; one persistent launch of %d steps" % args.steps,
                           s_ms * 1e-3, s_flops, s_bytes, args.workload + ("_sampler_quad" if is_quad else "_sampler"),
                           "not used by this kernel: (L+1)(L+2)/2 DFMA per evaluation from the triangular factor; the DMMA steppers "
                           "(direct_evaluation below) remain selectable with MDA_ENS_IMPL=direct" if is_quad else dmma_note,
                           executed_flops=((n_l + 1) * (n_l + 2) + 5.0 * n_l) * bins * cfg["walkers"] * args.steps if is_quad else None)
    direct = None
    if is_quad:
        # the same steps on the angle-by-angle stepper (chi^2 as C/ChiSquare.c:188-210 sums it, on the fp64 tensor path)
        os.environ["MDA_ENS_IMPL"] = "direct"
        try:
            work.store.synchronize()
            dist.barrier()
            d_ms, _ = work.sampler_resident(args.steps, max(args.warmup, 3))
            dist.barrier()
            d_ms = dist.max(d_ms)
            d_text = work.ens.describe(args.steps)
        finally:
            os.environ.pop("MDA_ENS_IMPL", None)
        direct = {"value": evals_step * args.steps / (d_ms * 1e-3), "unit": UNIT, "ms_per_step": d_ms / args.steps,
                  "roofline": roofline_of(d_text + "; one persistent launch of %d steps" % args.steps, d_ms * 1e-3, s_flops, s_bytes,
                                          args.workload + "_sampler", dmma_note)}
    dur_s = ms * 1e-3 / timed_launches
    b_kernel = ("lnpost_tile_kernel<L=%d,WPT=%d> %d threads x %d CTAs" % (n_l, tile[1], tile[0], tile[2]) if tile[1] > 0 else
                "lnpost_mma_kernel<L=%d> (fp64 DMMA) %d threads x %d CTAs" % (n_l, tile[0], tile[2]))
    b_roofline = roofline_of(b_kernel, dur_s,
                             2.0 * n_ang * (n_l + 1) * bins * half, 8.0 * (n_l + 1) * bins * (half + n_ang), args.workload,
                             ("not used by this kernel: register-tiled DFMA (K = L = %d); the DMMA variant is the default "
                              "from 12 angle tiles on" % n_l) if tile[1] > 0 else "fp64 tensor path: mma.sync m8n8k4 f64 (DMMA.8x8x4)")
    batched = {
        "value": evals_step * args.steps / (ms * 1e-3), "unit": UNIT, "ms_per_step": ms / args.steps,
        "gpu_launches": timed_launches, "roofline": b_roofline,
        "e2e": {"value": evals_step * e2e_steps / e2e_secs, "unit": UNIT, "steps": e2e_steps,
                "h2d_bytes_per_step": bins * cfg["walkers"] * n_l * 8, "d2h_bytes_per_step": bins * cfg["walkers"] * 8,
                "path": "mdaBatchLnPostStaged: pinned host block -> H2D -> kernel -> D2H -> sync, one CUDA graph per call"},
        "l2_policy": "inputs larger than L2: ring of distinct parameter blocks > 2x126 MB",
    }

    strong = None
    if world > 1 and not args.no_strong:
        # companion: BASELINE configs[3] literally -- 200 bins in total, sharded round-robin
        mine = shard.bins_for_rank(cfg["bins"], rank, world)
        sw = GpuWorkload(lib, cfg, rank, world, len(mine), rank, world, args.tile)
        q_ms, _, q_secs, q_steps = measure_sampler(sw, dist, args.steps, args.warmup)
        tot = cfg["bins"] * cfg["walkers"]
        strong = {"bins_total": cfg["bins"], "bins_per_gpu": len(mine), "value": tot * args.steps / (q_ms * 1e-3),
                  "unit": UNIT, "e2e": tot * q_steps / q_secs, "ms_per_step": q_ms / args.steps}
        sw.close()

    # the copy roofs of the e2e paths (rank 0's GPU only; the other ranks wait at the closing barrier): measured here, next to
    # the e2e run, and once more after the host-side measurements below -- the roof is the larger of the two (a single
    # reading taken while the host was still busy came out below the e2e path's own rate once)
    pcie = pcie_peaks(lib, chisq) if rank == 0 else None
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        os.sched_setaffinity(0, all_cpus)                  # the reference arm gets every host core again
        cpu = cpu_reference_arm(cfg, args.cpu_seconds, work.consts, work.n_pts, work.params)
    # rank 0's GPU only: the second reading of the copy roofs, the scalar-call latency, and BASELINE.json's other configs
    scalar = configs = None
    if rank == 0:
        again = pcie_peaks(lib, chisq)
        for key in ("d2h_gbs", "h2d_gbs"):
            pcie[key] = max(pcie[key], again[key])
        pcie["how"] += ", the larger of two such readings"
        scalar = scalar_call_latency(lib, chisq)
        if not args.no_configs:
            configs = other_configs(lib, chisq, orc, args.steps, peak_tf.value, hbm_peak)

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": s_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": workload_config(cfg, world),
            "note": "device-resident stretch-move sampler: proposal, box prior, fp64 chi-square log-likelihood and accept/reject in "
                    "one persistent kernel; chain kept in HBM (thin %d).  scaling = weak: every GPU owns %d bins; the literal "
                    "200-bins-in-total split of BASELINE configs[3] is value_strong / e2e_strong" % (s_thin, cfg["bins"]),
            "e2e": {"value": e2e_value, "unit": UNIT, "steps": s_e2e_steps,
                    "h2d_bytes_per_step": bins * cfg["walkers"] * n_l * 8 // s_e2e_steps,
                    "d2h_bytes_per_step": bins * cfg["walkers"] * (n_l + 1) * 8,
                    "ensemble_steps_per_s": s_e2e_steps / s_secs,
                    # the roof of this path is the host link: bytes delivered per second against one big pinned copy
                    "pcie_gbs": world * bins * cfg["walkers"] * (n_l + 1) * 8 * s_e2e_steps / s_secs / 1e9,
                    "pcie_gbs_per_gpu": bins * cfg["walkers"] * (n_l + 1) * 8 * s_e2e_steps / s_secs / 1e9,
                    "numa": numa,
                    "d2h_peak_gbs": pcie["d2h_gbs"], "h2d_peak_gbs": pcie["h2d_gbs"], "peak_how": pcie["how"] + " (rank 0's GPU alone)",
                    "frac_of_d2h_peak": bins * cfg["walkers"] * (n_l + 1) * 8 * s_e2e_steps / s_secs / 1e9 / pcie["d2h_gbs"],
                    "limit": "PCIe: 8 (L+1) = %d bytes per evaluation cross the host link (Save Chain Data, mda.py:1244-1250: the reference "
                             "keeps the whole chain on the host)%s" % (8 * (n_l + 1), "" if world == 1 else
                             "; with several GPUs the VM's host bridge is shared -- frac_of_d2h_peak is per GPU against a lone GPU's copy rate"),
                    "path": "mdaEnsembleSetState(pinned host p0) + mdaEnsembleRunToHost: full chain + lnprob delivered to pinned "
                            "host memory, chunked D2H overlapped with stepping (PCIe bound: 80 B per evaluation); one call of "
                            "%d steps whatever --steps is (a run_mcmc call of the reference is 16 385 steps)" % s_e2e_steps},
            "e2e_summaries": {"value": evals_step * work.summary_e2e[1] / work.summary_e2e[0], "unit": UNIT,
                              "steps": work.summary_e2e[1], "seconds": work.summary_e2e[0],
                              "h2d_bytes": bins * cfg["walkers"] * n_l * 8, "d2h_bytes": bins * (n_l * 6 + 1) * 8,
                              "runs": "the faster of two runs (wall clock, max over ranks each)",
                              "path": "mdaEnsembleSetState(host p0) + mdaEnsembleAdvance + mdaEnsembleSummarize: the chain stays in "
                                      "HBM; percentile / histogram-peak estimates (mda.py:1399-1491) and acceptance fractions "
                                      "come back -- what the reference's worker returns (mda.py:1318)"},
            "gpu_launches": 1,
            "gpu_launches_total_in_run": int(total_launches),
            "ensemble_steps_per_s": args.steps / (s_ms * 1e-3),
            "roofline": roofline, "direct_evaluation": direct, "batched_half_step": batched, "cpu_baseline": cpu, "clocks": clocks,
            "scalar_call_us": scalar, "configs": configs,
            "parity_max_rel_err": worst,
            "device": {"name": name.value.decode(), "sms": sms.value, "cc": cc.value},
        }
        if strong is not None:                     # BASELINE configs[3] read literally, first class: 200 bins in total over the GPUs
            line["strong_200_bins_total"] = strong
            line["value_strong"] = strong["value"]
            line["e2e_strong"] = strong["e2e"]
            line["scaling_strong"] = "strong: %d bins in total, %d per GPU" % (strong["bins_total"], strong["bins_per_gpu"])
        print(json.dumps(line))
    work.close()
    dist.close()


if __name__ == "__main__":
    main()
