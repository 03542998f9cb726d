"""Store builder: from the reference's input files to the device-resident store
(SURVEY.md 8f-3).

What the store must hold is fixed by the reference (mda.py:173-176, 2069):

    dists[l][j] = cubic_l(theta_j) / dsigma_j          cubic_l: scipy interp1d(kind="cubic") through the
                                                       DWBA table of multipole l at this bin's energy
    data[j]     = (sigma_j - ivgdr_j) / dsigma_j       ivgdr_j: the isovector dipole's table, interpolated
                                                       the same way and scaled by its Lorentzian
                                                       sum-rule fraction (only if "Subtract IVGDR")

and so are the on-disk formats it is computed from (README.md:11-17):

    experimental data     one row per Ex bin: "Ex, theta, sigma, dsigma, theta, sigma, dsigma, ..."
                          (mda.py:2331-2401: bins strictly inside (Start Energy, Final Energy), angles
                          up to Max Theta are fitted, every angle is kept for plots)
    DWBA tables           A<A>_Ex<E>_L<LL>_T0_F<FFF>.csv, rows "theta, sigma"  (mda.py:2072-2113)
    IVGDR tables          A<A>_Ex<E>_L01_T1_F100.csv                            (mda.py:2211-2245)

How it is computed here is this module's own: a row of the data file is parsed into one array
and split by slicing, tables are parsed whole, the Lorentzian is one vectorised expression over
the spectrum's energies, and bins are packed straight into the store's [B, 1+L, n_pad] block.
The arithmetic per value -- one interpolation, one multiply, one subtract, one divide -- is the
reference's, so the results are bit-identical to its preprocessing
(tests/golden/golden_builder.npz holds the outputs of the reference's own functions).
"""
import os
from dataclasses import dataclass

import numpy as np
from scipy.interpolate import interp1d

DWBA_NAME = "A{a:d}_Ex{e:4.2f}_L{l:02d}_T0_F{f:03d}.csv"
IVGDR_NAME = "A{a:d}_Ex{e:4.2f}_L01_T1_F100.csv"


@dataclass
class ExBin:
    """One excitation-energy bin of the measured spectrum."""
    energy: float
    points: np.ndarray        # [n, 3] every (theta, sigma, dsigma) of the row, in file order
    fitted: np.ndarray        # [n] bool: theta <= Max Theta

    @property
    def theta(self):
        return self.points[self.fitted, 0]

    @property
    def sigma(self):
        return self.points[self.fitted, 1]

    @property
    def dsigma(self):
        return self.points[self.fitted, 2]


def _numbers(text):
    """Comma-separated decimal numbers -> float64 (Python's correctly rounded parser)."""
    return np.array(text.split(","), dtype=np.float64)


def load_table(path):
    """A two-column csv table (theta, sigma) as [n, 2]."""
    with open(path, "r") as fh:
        rows = [ln for ln in fh.read().splitlines() if ln.strip()]
    return _numbers(",".join(rows)).reshape(len(rows), -1)[:, :2]


def load_spectrum(path, start_energy, final_energy, max_theta):
    """The experimental data file -> [ExBin] for the rows with start < Ex < final."""
    bins = []
    with open(path, "r") as fh:
        for row in fh:
            if not row.strip():
                continue
            values = _numbers(row)
            energy = float(values[0])
            if not (start_energy < energy < final_energy):
                continue
            points = values[1:].reshape(-1, 3)
            bins.append(ExBin(energy, points, points[:, 0] <= max_theta))
    return bins


def dwba_table_path(directory, target_a, energy, l_value, ewsr_fraction):
    return os.path.join(directory, DWBA_NAME.format(a=target_a, e=energy, l=l_value, f=int(100.0 * ewsr_fraction)))


def ivgdr_table_path(directory, target_a, energy):
    return os.path.join(directory, IVGDR_NAME.format(a=target_a, e=energy))


def lorentzian_cross_section(energies, height, centroid, width):
    """The IVGDR photo-absorption Lorentzian  s(E) = h / (1 + (E^2 - c^2)^2 / (E^2 w^2))  at ``energies``, written as
    the reference evaluates it (mda.py:2249-2328): h / ((1 - 2 c^2/w^2) + (c^4/w^2)/E^2 + E^2/w^2)."""
    e_sq = np.square(np.asarray(energies, dtype=np.float64))
    w_sq = width * width
    ratio = centroid / width
    return height / ((1.0 - 2.0 * (ratio * ratio)) + (centroid ** 4.0 / w_sq) / e_sq + e_sq / w_sq)


def ivgdr_sum_rule_fraction(energies, height, centroid, width, cs_integral):
    """Fraction of the IVGDR sum rule per MeV at each energy: Lorentzian / its integral (the peak-to-integral ratio is
    formed first, as the reference does, so the values agree bit for bit)."""
    return lorentzian_cross_section(energies, height / cs_integral, centroid, width)


def spline_at(table, theta):
    """scipy's cubic interpolant (the reference's own third-party dependency) through a table, at ``theta``."""
    return interp1d(table[:, 0], table[:, 1], kind="cubic")(theta)


def store_block(bins, dwba_tables, ivgdr=None):
    """``consts [B, 1+L, n_pad]`` (row 0: data / error, rows 1..L: distributions / error, zero padded to an even
    n_pad) and ``n_pts [B]`` for bins of ragged length.  ``dwba_tables[b][l]`` is the [n, 2] table of multipole l at
    bin b's energy; ``ivgdr`` = (tables, fractions) to subtract the dipole from the data first."""
    n_bins = len(bins)
    n_l = len(dwba_tables[0]) if n_bins else 0
    n_pts = np.array([int(b.fitted.sum()) for b in bins], dtype=np.int32)
    n_max = int(n_pts.max()) if n_bins else 0
    consts = np.zeros((n_bins, 1 + n_l, max(n_max + (n_max & 1), 2)), dtype=np.float64)
    for b, ex in enumerate(bins):
        n, theta, err = n_pts[b], ex.theta, ex.dsigma
        sigma = ex.sigma
        if ivgdr is not None:
            sigma = sigma - ivgdr[1][b] * spline_at(ivgdr[0][b], theta)
        consts[b, 0, :n] = sigma / err
        for l in range(n_l):
            consts[b, 1 + l, :n] = spline_at(dwba_tables[b][l], theta) / err
    return consts, n_pts


def build_inputs(config):
    """Everything ``initialize_mda`` prepares per bin (mda.py:160-183), from a CONFIG dictionary with the reference's
    keys.  Returns a dict: ``bins`` ([ExBin]), ``energies``, ``consts``, ``n_pts``, ``bounds`` (0, 1/EWSR_l:
    mda.py:1124) and ``ivgdr_fractions`` (None unless "Subtract IVGDR")."""
    bins = load_spectrum(config["Input File Path"], config["Start Energy"], config["Final Energy"], config["Max Theta"])
    directory, target_a, ewsr = config["Distribution Directory"], config["Target A"], config["EWSR Fractions"]
    n_l = config["Maximum L"] + 1
    energies = [b.energy for b in bins]
    dwba = [[load_table(dwba_table_path(directory, target_a, e, l, ewsr[l])) for l in range(n_l)] for e in energies]
    ivgdr = fractions = None
    if config.get("Subtract IVGDR", False):
        fractions = ivgdr_sum_rule_fraction(energies, config["IVGDR Height"], config["IVGDR Center"], config["IVGDR Width"],
                                            config["IVGDR CS Integral"])
        ivgdr = ([load_table(ivgdr_table_path(directory, target_a, e)) for e in energies], fractions)
    consts, n_pts = store_block(bins, dwba, ivgdr)
    return {"bins": bins, "energies": energies, "consts": consts, "n_pts": n_pts, "bounds": [(0.0, 1.0 / x) for x in ewsr[:n_l]],
            "ivgdr_fractions": fractions}


def build_store(config, cs_lib=None, bins=None):
    """``build_inputs``, then the device-resident store.  ``bins`` selects a subset of the spectrum (e.g.
    ``shard.bins_for_rank``) so every rank uploads only its own bins."""
    from . import chisq
    inputs = build_inputs(config)
    sel = np.arange(len(inputs["energies"])) if bins is None else np.asarray(bins, dtype=np.int64)
    store = chisq.DeviceStore(inputs["consts"][sel], inputs["n_pts"][sel], inputs["bounds"], cs_lib=cs_lib)
    return store, inputs
