#!/usr/bin/env python
"""Times the triangular-factor steppers over the named shapes: a CTA per bin (ensemble_quad_kernel) against a bin per
thread-block cluster (ensemble_cluster_kernel) at every cluster size.  One JSON line per run (tools/ensemble_bench.py).
    python tools/cluster_sweep.py [--quick] > gpurun_out/cluster_sweep.jsonl
"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SHAPES = [  # name, bins, walkers, L, N
    ("C4", 200, 512, 9, 60), ("C4/8gpu", 25, 512, 9, 60), ("C4/4gpu", 50, 512, 9, 60), ("C4/2gpu", 100, 512, 9, 60),
    ("C2", 100, 256, 5, 30), ("one-per-sm", 148, 512, 9, 60), ("two-per-sm", 296, 512, 9, 60),
]
VARIANTS = [("quad", {"MDA_ENS_IMPL": "quad"})] + [("cl%d%s" % (cl, "r" if tr == "1" else "s"), {"MDA_ENS_CL": str(cl), "MDA_ENS_TREG": tr})
                                                    for cl in (1, 2, 4, 8) for tr in ("1", "0")]


def main():
    quick = "--quick" in sys.argv
    steps_list = (20, 500) if not quick else (200,)
    for name, bins, walkers, n_l, n_pts in SHAPES:
        for vname, env in VARIANTS:
            for steps in steps_list:
                e = dict(os.environ)
                e.update(env)
                cmd = [sys.executable, os.path.join(ROOT, "tools", "ensemble_bench.py"), "--bins", str(bins), "--walkers", str(walkers),
                       "--n-l", str(n_l), "--n-pts", str(n_pts), "--steps", str(steps)]
                res = subprocess.run(cmd, env=e, capture_output=True, text=True, timeout=300)
                line = res.stdout.strip().splitlines()[-1] if res.stdout.strip() else ""
                try:
                    out = json.loads(line)
                except ValueError:
                    out = {"error": (res.stderr or res.stdout)[-300:]}
                out.update(shape=name, variant=vname)
                print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
