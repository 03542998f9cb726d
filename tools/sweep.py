#!/usr/bin/env python
"""Throughput sweep of the batched likelihood kernel (BASELINE.json configs[4]):
walkers x angles x dists (and kernel tilings), device-resident parameter blocks, CUDA
events on the launching stream.  Prints one JSON line per point.

    python tools/sweep.py [--tiles] [--shapes c4|sweep|big] [--launches 200]
"""
import argparse
import ctypes as ct
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mimic_mda_b200 import chisq, synth  # noqa: E402

L2 = 126 << 20


def time_point(lib, bins, walkers, n_l, n_pts, tile, launches, peak_tf):
    consts, counts, a_true = synth.make_bins(bins, n_l, n_pts)
    store = chisq.DeviceStore(consts, counts, cs_lib=lib)
    if tile:
        store.set_tile(*tile)
    blk = synth.make_walkers(a_true, walkers, seed=3)
    nbytes = blk.nbytes
    n_ring = min(max(2, 2 * L2 // nbytes + 1), 64)
    dps, dos = [], []
    for i in range(n_ring):
        dp, do = lib.mdaDeviceAlloc(nbytes), lib.mdaDeviceAlloc(bins * walkers * 8)
        chisq.check(lib, lib.mdaMemcpyH2D(dp, blk.ctypes.data, nbytes))
        dps.append(dp)
        dos.append(do)
    ev0, ev1 = lib.mdaEventCreate(), lib.mdaEventCreate()
    store.lnpost_device_many(walkers, dps, dos, 5)
    store.synchronize()
    best = 1e30
    for _ in range(3):
        lib.mdaEventRecord(ev0, store.handle)
        store.lnpost_device_many(walkers, dps, dos, launches)
        lib.mdaEventRecord(ev1, store.handle)
        lib.mdaEventSynchronize(ev1)
        ms = ct.c_float()
        lib.mdaEventElapsedMs(ev0, ev1, ct.byref(ms))
        best = min(best, ms.value)
    us = best * 1e3 / launches
    evals = bins * walkers
    flops = 2.0 * n_pts * (n_l + 1) * evals
    nbytes_alg = 8.0 * (n_l + 1) * bins * (walkers + n_pts)
    t, w, g = store.get_tile(walkers)
    for p in dps + dos:
        lib.mdaDeviceFree(p)
    lib.mdaEventDestroy(ev0)
    lib.mdaEventDestroy(ev1)
    store.close()
    tf = flops / (us * 1e-6) / 1e12
    return {"bins": bins, "walkers": walkers, "L": n_l, "N": n_pts, "tile": "%dx%d" % (t, w), "grid": g,
            "us_per_launch": round(us, 3), "evals_per_s": evals / (us * 1e-6), "tflops": round(tf, 3),
            "frac_fp64": round(tf / peak_tf, 4), "gbs": round(nbytes_alg / (us * 1e-6) / 1e9, 1),
            "frac_hbm": round(nbytes_alg / (us * 1e-6) / 1e9 / 6551.7, 4)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--shapes", default="c4")
    ap.add_argument("--tiles", action="store_true")
    ap.add_argument("--launches", type=int, default=200)
    args = ap.parse_args()
    lib = chisq.prepare_shared_object()
    peak = ct.c_double()
    chisq.check(lib, lib.mdaMeasureFp64Peak(5, ct.byref(peak)))
    print(json.dumps({"fp64_peak_tflops": peak.value}))
    if args.shapes == "c4":
        shapes = [(200, 256, 9, 60), (100, 256, 9, 60), (100, 128, 5, 30), (25, 256, 9, 60), (1, 32, 4, 30)]
    elif args.shapes == "big":
        shapes = [(200, 8192, 9, 60), (200, 2048, 9, 60), (200, 8192, 12, 256), (200, 8192, 4, 16), (200, 8192, 5, 30),
                  (200, 8192, 8, 40)]
    elif args.shapes == "cross":           # where the angle-split kernel stops paying: 200 bins, growing ensembles
        shapes = [(200, w, 9, 60) for w in (64, 128, 256, 512, 1024, 2048, 4096)] + [(100, w, 5, 30) for w in (128, 512, 2048)] + \
                 [(100, w, 12, 256) for w in (128, 512, 2048)]
    else:
        shapes = [(100, w, l, n) for w in (32, 128, 512, 2048, 8192) for n in (16, 64, 256) for l in (4, 8, 12)]
    tiles = [None]
    if args.tiles:
        tiles = [(t, w) for w in (1, 2, 4) for t in (32, 64, 128, 256)]
    for (b, w, l, n) in shapes:
        for tile in tiles:
            print(json.dumps(time_point(lib, b, w, l, n, tile, args.launches, peak.value)), flush=True)


if __name__ == "__main__":
    main()
