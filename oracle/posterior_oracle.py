"""CPU restatement of the reference's posterior summaries -- TEST INFRASTRUCTURE ONLY.

Follows mda.py:1253-1259 (burn-in cut, walkers merged), calc_param_values (mda.py:1399-1433)
and find_most_likely_values (mda.py:1436-1491) line for line with numpy; pinned against the
reference's own functions by tests/golden/golden_posterior.npz (make_golden_posterior.py).
"""
import numpy as np

FLOAT_EPSILON = 1.0e-7          # mda.py:55


def find_most_likely_values(samples, ndims, num_bins, confidence_interval):
    output = []
    last_index = (len(samples[:, 0]) - 1)
    for i in range(ndims):
        values = np.sort(samples[:, i])
        if FLOAT_EPSILON > abs(values[last_index] - values[0]):
            output.append((values[0], 0.0, 0.0))
            continue
        hist = np.histogram(values, bins=num_bins)
        ind = np.argmax(hist[0])
        quantile = 0.0
        for j in range(ind + 1):
            quantile += float(hist[0][j])
        quantile /= float(len(values))
        peak_percentile = (100.0 * quantile)
        lo_p = peak_percentile - 100.0 * (confidence_interval / 2.0)
        if lo_p < 0.0:
            lo_p = 0.0
        hi_p = peak_percentile + 100.0 * (confidence_interval / 2.0)
        if hi_p > 100.0:
            hi_p = 100.0
        vals = np.percentile(values, (lo_p, peak_percentile, hi_p))
        output.append((vals[1], vals[2] - vals[1], vals[1] - vals[0]))
    return output


def calc_param_values(samples, quantile_list, ndims, num_bins, confidence_interval):
    percentile_list = [100.0 * q for q in quantile_list]
    points = [(v[1], v[2] - v[1], v[1] - v[0]) for v in zip(*np.percentile(samples, percentile_list, axis=0))]
    peaks = find_most_likely_values(samples, ndims, num_bins, confidence_interval)
    return (points, peaks)


def summarize(chain, burn_in, confidence_interval, num_bins):
    """``chain [B, W, kept, L]`` (emcee layout per bin) -> points [B, L, 3], peaks [B, L, 3]."""
    n_bins, n_walk, kept, n_l = chain.shape
    quantile_list = [(0.5 - confidence_interval / 2.0), 0.5, (0.5 + confidence_interval / 2.0)]   # mda.py:1257-1258
    points = np.empty((n_bins, n_l, 3))
    peaks = np.empty((n_bins, n_l, 3))
    for b in range(n_bins):
        samples = chain[b][:, burn_in:, :].reshape((n_walk * (kept - burn_in), n_l))               # mda.py:1253-1254
        pts, pks = calc_param_values(samples, quantile_list, n_l, num_bins, confidence_interval)
        points[b] = np.array(pts)
        peaks[b] = np.array(pks)
    return points, peaks
