#!/usr/bin/env python
"""Generates tests/golden/golden_builder.npz by running the REFERENCE'S OWN preprocessing
functions on synthetic input files (build container only: needs /root/reference).

mda.py is Python 2 and cannot be imported, so the functions on this path --
read_row_cs_data_file, read_dist, read_ivgdr_dists, IVGDRFraction,
interpolate_and_scale_ivgdr, handle_ivgdr, interpolate_dist, interp_all_dists
(mda.py:1989-2401) -- are lifted from the source text where it lies, unmodified except
that `print "..."` statements become `pass` and `xrange` is bound to `range`, and executed
against a CONFIG dictionary.  The fixture stores the input files' text and the reference's
outputs; tests/test_builder.py replays them through mimic_mda_b200/builder.py.
"""
import copy
import math
import os
import re
import sys
import tempfile

import numpy as np
from scipy import interpolate

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
REF = "/root/reference/mda.py"
WANTED = ["read_row_cs_data_file", "read_dist", "read_ivgdr_dists", "IVGDRFraction",
          "interpolate_and_scale_ivgdr", "handle_ivgdr", "interpolate_dist", "interp_all_dists"]


def lift(names):
    lines = open(REF).read().split("\n")
    tops = [i for i, ln in enumerate(lines) if re.match(r"^(def|class) \w+", ln)] + [len(lines)]
    src = []
    for name in names:
        start = next(i for i in tops[:-1] if re.match(r"^(def|class) %s\b" % name, lines[i]))
        end = tops[tops.index(start) + 1]
        block = [re.sub(r"^(\s*)print .*$", r"\1pass", ln) for ln in lines[start:end]]
        src.append("\n".join(block))
    return "\n\n".join(src)


def write_inputs(root):
    rng = np.random.default_rng(2024)
    energies = [10.1, 10.3, 10.5, 10.7, 10.9, 11.1, 25.0]           # the last one is outside the window
    dist_dir = os.path.join(root, "dists")
    os.makedirs(dist_dir)
    rows = []
    files = {}
    for k, en in enumerate(energies):
        n = 9 + (k % 3)
        theta = np.sort(rng.uniform(0.6, 11.5, size=n))
        sigma = 40.0 * np.exp(-0.25 * theta) * (1.2 + np.cos(0.9 * theta + 0.3 * k)) + 0.5
        err = sigma * rng.uniform(0.03, 0.08, size=n)
        row = "%.2f" % en + "".join(", %.6f, %.6f, %.6f" % (t, s, e) for t, s, e in zip(theta, sigma, err))
        rows.append(row)
        grid = np.arange(0.0, 15.01, 0.25)
        for l in range(4):
            cs = 10.0 ** (1.3 - 0.12 * grid) * (1.1 + np.cos(grid * (0.35 + 0.1 * l) + l + 0.05 * k))
            name = "A116_Ex%4.2f_L%02d_T0_F%03d.csv" % (en, l, [100, 100, 80, 100][l])
            files[name] = "\n".join("%.4f, %.8e" % (g, c) for g, c in zip(grid, cs)) + "\n"
        cs = 10.0 ** (0.8 - 0.1 * grid) * (1.05 + np.sin(0.5 * grid + 0.1 * k))
        files["A116_Ex%4.2f_L01_T1_F100.csv" % en] = "\n".join("%.4f,%.8e" % (g, c) for g, c in zip(grid, cs)) + "\n"
    exp_text = "\n".join(rows) + "\n"
    with open(os.path.join(root, "exp.csv"), "w") as fh:
        fh.write(exp_text)
    for name, text in files.items():
        with open(os.path.join(dist_dir, name), "w") as fh:
            fh.write(text)
    return exp_text, files


def main():
    with tempfile.TemporaryDirectory() as root:
        exp_text, files = write_inputs(root)
        out = {"exp_text": np.array(exp_text), "file_names": np.array(sorted(files)),
               "file_texts": np.array([files[k] for k in sorted(files)])}
        for tag, subtract in (("noivgdr", False), ("ivgdr", True)):
            config = {"Input File Path": os.path.join(root, "exp.csv"), "Distribution Directory": os.path.join(root, "dists"),
                      "Target A": 116, "Start Energy": 10.0, "Final Energy": 12.0, "Max Theta": 10.0, "Maximum L": 3,
                      "EWSR Fractions": [1.0, 1.0, 0.8, 1.0], "Subtract IVGDR": subtract, "IVGDR Height": 280.0,
                      "IVGDR Center": 15.3, "IVGDR Width": 4.9, "IVGDR CS Integral": 2100.0}
            ns = {"CONFIG": config, "np": np, "interpolate": interpolate, "os": os, "math": math, "copy": copy, "xrange": range}
            exec(lift(WANTED), ns)
            # the sequence of initialize_mda, mda.py:163-176
            exp_data, plot_data = ns["read_row_cs_data_file"]()
            ivgdr_info = ns["handle_ivgdr"](exp_data)
            dists = [[ns["read_dist"](elem[0], i) for i in range(config["Maximum L"] + 1)] for elem in exp_data]
            interp_dists = ns["interp_all_dists"](dists, exp_data)
            fit_data = [(exp_en[1][:, 1] / exp_en[1][:, 2]) for exp_en in ivgdr_info[2]]
            out[tag + "/energies"] = np.array([e[0] for e in exp_data])
            out[tag + "/n_pts"] = np.array([len(d) for d in fit_data])
            out[tag + "/n_plot"] = np.array([len(p[1]) for p in plot_data])
            for b in range(len(fit_data)):
                out["%s/fit_data_%d" % (tag, b)] = fit_data[b]
                out["%s/interp_dists_%d" % (tag, b)] = np.array(interp_dists[b])
            if subtract:
                out[tag + "/ewsrs"] = np.array(ivgdr_info[1])
            print(tag, "bins", len(fit_data), "angles per bin", [len(d) for d in fit_data])
        path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden_builder.npz")
        np.savez_compressed(path, **out)
        print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
