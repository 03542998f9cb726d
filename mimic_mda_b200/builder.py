"""Store builder: from the reference's input files to the device-resident store
(SURVEY.md 8f-3).

Python-3 mirror of the reference's one-off preprocessing, same names and argument
meaning, with the CONFIG keys it reads passed explicitly:

    read_row_cs_data_file       mda.py:2331-2401   exp data rows "Ex, (theta, sigma, dsigma) x N",
                                                   Start/Final Energy window, per-row Max Theta cut
    read_dist                   mda.py:2072-2113   DWBA csv  A%d_Ex%4.2f_L%02d_T0_F%03d.csv
    read_ivgdr_dists            mda.py:2211-2245   IVGDR csv A%d_Ex%4.2f_L01_T1_F100.csv
    IVGDRFraction               mda.py:2249-2328   Lorentzian EWSR fraction
    interpolate_and_scale_ivgdr mda.py:2180-2208
    handle_ivgdr                mda.py:2116-2177   IVGDR subtracted from the data (not fitted)
    interpolate_dist            mda.py:2040-2069   interp1d(kind="cubic") at exp angles / error
    interp_all_dists            mda.py:1989-2037
    build_store                 mda.py:160-183     the sequence initialize_mda runs, ending in a
                                                   DeviceStore instead of per-bin tuples

The interpolation is scipy's (the reference's own third-party dependency): the store
contents are defined as ``dists[l][j] = cubic_l(theta_j) / dsigma_j`` and
``data[j] = (sigma_j - ivgdr_j) / dsigma_j`` (mda.py:176, 2069).
"""
import copy
import math
import os

import numpy as np
from scipy import interpolate


def read_row_cs_data_file(path, start_energy, final_energy, max_theta):
    """Returns (fit_output, plot_output): lists of [energy, array(N, 3)] with columns
    (angle, cross-section, error); fit_output keeps only angles <= max_theta."""
    fit_output, plot_output = [], []
    with open(path, "r") as input_file:
        for line in input_file:
            data_list = line.split(",")
            energy = float(data_list[0].strip())
            if energy < final_energy and energy > start_energy:
                distribution_data = [float(x.strip()) for x in data_list[1:]]
                fit_distribution, plot_distribution = [], []
                for i in range(0, len(distribution_data), 3):
                    point = (distribution_data[i], distribution_data[i + 1], distribution_data[i + 2])
                    if distribution_data[i] <= max_theta:
                        fit_distribution.append(point)
                    plot_distribution.append(point)
                fit_output.append([energy, np.array(fit_distribution, dtype=np.float64)])
                plot_output.append([energy, np.array(plot_distribution, dtype=np.float64)])
    return fit_output, plot_output


def dist_file_name(directory, target_a, energy, l_value, ewsr_fraction):
    fmt_str = "A{0:d}_Ex{1:4.2f}_L{2:02d}_T0_F{3:03d}.csv"
    return os.path.join(directory, fmt_str.format(target_a, energy, l_value, int(100.0 * ewsr_fraction)))


def read_dist(directory, target_a, energy, l_value, ewsr_fractions):
    """DWBA angular distribution (angle, cross-section) of multipole l_value at this energy."""
    output = []
    with open(dist_file_name(directory, target_a, energy, l_value, ewsr_fractions[l_value]), "r") as dist_file:
        for line in dist_file:
            vals = [float(x.strip()) for x in line.strip().split(",")]
            output.append((vals[0], vals[1]))
    return np.array(output, dtype=np.float64)


def read_ivgdr_dists(directory, target_a, en_list):
    fmt_str = "A{0:d}_Ex{1:4.2f}_L01_T1_F100.csv"
    dists_list = []
    for energy in en_list:
        with open(os.path.join(directory, fmt_str.format(target_a, energy)), "r") as dist_file:
            dist = [[float(elem) for elem in line.strip().split(",")] for line in dist_file]
        dists_list.append(np.array(dist, dtype=np.float64))
    return dists_list


class IVGDRFraction(object):
    """Lorentzian normalised by its integral (mda.py:2249-2328)."""

    def __init__(self, max_sigma, centroid, width, total_sigma):
        self.sig_ratio = (max_sigma / total_sigma)
        self.total = total_sigma
        self.denom_const = (1.0 - 2.0 * math.pow(centroid / width, 2.0))
        self.width_sq = math.pow(width, 2.0)
        self.denom_inv = math.pow(centroid, 4.0) / self.width_sq

    def get_ivgdr_fraction(self, excitation_energy):
        en_sq = math.pow(excitation_energy, 2.0)
        denom = self.denom_const + self.denom_inv / en_sq + en_sq / self.width_sq
        return self.sig_ratio / denom

    def get_ivgdr_cs(self, excitation_energy):
        en_sq = math.pow(excitation_energy, 2.0)
        denom = self.denom_const + self.denom_inv / en_sq + en_sq / self.width_sq
        return (self.total * self.sig_ratio) / denom


def interpolate_and_scale_ivgdr(dist, ewsr, angle_list):
    interp = interpolate.interp1d(dist[:, 0], dist[:, 1], kind="cubic")
    return ewsr * interp(angle_list)


def handle_ivgdr(data, subtract, directory=None, target_a=None, height=None, center=None, width=None, cs_integral=None):
    """Returns (dists, ewsrs, sub_data) exactly as mda.py:2116-2177."""
    if not subtract:
        return None, None, copy.deepcopy(data)
    frac = IVGDRFraction(height, center, width, cs_integral)
    ewsrs = [frac.get_ivgdr_fraction(elem[0]) for elem in data]
    dists = read_ivgdr_dists(directory, target_a, [elem[0] for elem in data])
    angle_lists = [en_set[1][:, 0] for en_set in data]
    ivgdr_values = [interpolate_and_scale_ivgdr(dist, ewsr, angle_list)
                    for (dist, ewsr, angle_list) in zip(dists, ewsrs, angle_lists)]
    sub_data = copy.deepcopy(data)
    for i, sub_set in enumerate(sub_data):
        sub_set[1][:, 1] = sub_set[1][:, 1] - ivgdr_values[i]
    return dists, ewsrs, sub_data


def interpolate_dist(dist, angles, errors):
    interp = interpolate.interp1d(dist[:, 0], dist[:, 1], kind="cubic")
    return (interp(angles) / errors).astype(np.float64)


def interp_all_dists(dists, data):
    output = []
    for i, dis_set in enumerate(dists):
        angle_list = data[i][1][:, 0]
        error_list = data[i][1][:, 2]
        output.append([interpolate_dist(dist, angle_list, error_list).astype(np.float64) for dist in dis_set])
    return output


def pack_bins(fit_data, interp_dists):
    """Per-bin arrays of initialize_mda (mda.py:173-176) -> ``consts [B, 1+L, n_pad]``
    (row 0 data/error, rows 1..L distributions/error, zero padded) and ``n_pts [B]``."""
    n_bins = len(fit_data)
    n_l = len(interp_dists[0]) if n_bins else 0
    n_max = max((len(d) for d in fit_data), default=0)
    n_pad = n_max + (n_max & 1)
    consts = np.zeros((n_bins, 1 + n_l, max(n_pad, 2)), dtype=np.float64)
    n_pts = np.zeros(n_bins, dtype=np.int32)
    for b in range(n_bins):
        n = len(fit_data[b])
        n_pts[b] = n
        consts[b, 0, :n] = fit_data[b]
        for l in range(n_l):
            consts[b, 1 + l, :n] = interp_dists[b][l]
    return consts, n_pts


def build_inputs(config):
    """The sequence initialize_mda runs up to the per-bin tuples (mda.py:160-183), from a
    CONFIG dictionary with the reference's keys.  Returns a dict with ``energies``,
    ``consts``, ``n_pts``, ``bounds`` and the intermediate products."""
    exp_data, plot_data = read_row_cs_data_file(config["Input File Path"], config["Start Energy"],
                                                config["Final Energy"], config["Max Theta"])
    ivgdr_info = handle_ivgdr(exp_data, config.get("Subtract IVGDR", False), config["Distribution Directory"],
                              config["Target A"], config.get("IVGDR Height"), config.get("IVGDR Center"),
                              config.get("IVGDR Width"), config.get("IVGDR CS Integral"))
    n_l = config["Maximum L"] + 1
    dists = [[read_dist(config["Distribution Directory"], config["Target A"], elem[0], i, config["EWSR Fractions"])
              for i in range(n_l)] for elem in exp_data]
    interp_dists = interp_all_dists(dists, exp_data)
    fit_data = [(exp_en[1][:, 1] / exp_en[1][:, 2]) for exp_en in ivgdr_info[2]]
    consts, n_pts = pack_bins(fit_data, interp_dists)
    bounds = [(0.0, 1.0 / x) for x in config["EWSR Fractions"][:n_l]]              # mda.py:1124
    return {"energies": [elem[0] for elem in exp_data], "consts": consts, "n_pts": n_pts, "bounds": bounds,
            "exp_data": exp_data, "plot_data": plot_data, "ivgdr_info": ivgdr_info, "dists": dists,
            "interp_dists": interp_dists, "fit_data": fit_data}


def build_store(config, cs_lib=None, bins=None):
    """Everything above, then the device-resident store.  ``bins`` selects a subset of the
    spectrum (e.g. ``shard.bins_for_rank``) so every rank uploads only its own bins."""
    from . import chisq
    inputs = build_inputs(config)
    sel = np.arange(len(inputs["energies"])) if bins is None else np.asarray(bins, dtype=np.int64)
    store = chisq.DeviceStore(inputs["consts"][sel], inputs["n_pts"][sel], inputs["bounds"], cs_lib=cs_lib)
    return store, inputs
