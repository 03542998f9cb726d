#!/usr/bin/env python
"""Generates tests/golden/golden_posterior.npz from the REFERENCE'S OWN calc_param_values /
find_most_likely_values (mda.py:1399-1491), lifted from the source text and executed on
seeded sample sets (build container only: needs /root/reference)."""
import os
import re
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
REF = "/root/reference/mda.py"


def lift(names):
    lines = open(REF).read().split("\n")
    tops = [i for i, ln in enumerate(lines) if re.match(r"^(def|class) \w+", ln)] + [len(lines)]
    out = []
    for name in names:
        start = next(i for i in tops[:-1] if re.match(r"^def %s\b" % name, lines[i]))
        out.append("\n".join(lines[start:tops[tops.index(start) + 1]]))
    return "\n\n".join(out)


def main():
    rng = np.random.default_rng(77)
    out = {}
    cases = {
        "gauss": np.abs(0.2 + 0.03 * rng.standard_normal((6000, 3))),
        "skewed": np.clip(rng.gamma(2.0, 0.05, size=(4096, 4)), 0.0, 1.0),
        "repeats": np.repeat(rng.uniform(0.0, 0.4, size=(700, 2)), 5, axis=0),      # MCMC chains repeat values
        "flat_dim": np.column_stack([rng.uniform(0.1, 0.3, 3000), np.full(3000, 0.001)]),
        "edge_peak": np.clip(rng.exponential(0.02, size=(5000, 2)), 0.0, 1.0),      # peak in the first bins -> lo_p clipped to 0
    }
    for name, samples in cases.items():
        for ci, nb in ((0.682689492, 1000), (0.9, 50)):
            ns = {"CONFIG": {"Num Bins": nb, "Confidence Interval": ci}, "FLOAT_EPSILON": 1.0e-7, "np": np}
            exec(lift(["calc_param_values", "find_most_likely_values"]), ns)
            ql = [(0.5 - ci / 2.0), 0.5, (0.5 + ci / 2.0)]
            points, peaks = ns["calc_param_values"](samples, ql, samples.shape[1])
            key = "%s/%d" % (name, nb)
            out[key + "/samples"] = samples
            out[key + "/ci"] = np.array(ci)
            out[key + "/points"] = np.array(points)
            out[key + "/peaks"] = np.array(peaks)
            print(key, np.array(points)[0], np.array(peaks)[0])
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden_posterior.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path))


if __name__ == "__main__":
    main()
