"""CPU restatement of emcee 2.x's integrated autocorrelation time -- TEST INFRASTRUCTURE ONLY.

The reference calls ``sampler.get_autocorr_time(c=CONFIG["ACorr WindSize"], fast=CONFIG["ACorr Use FFT"])``
(mda.py:1304-1308) and writes the result into its diagnostics file (mda.py:253-272).  The arithmetic lives in
a third-party dependency that is NOT under /root/reference and not installed here: ``emcee`` (no version pin
in the reference; 2.2.x inferred from the ``get_autocorr_time(c=, fast=)`` signature and Python 2, SURVEY.md
8c).  **Parity unpinned**: this file restates the published algorithm of emcee 2.2's ``autocorr.function`` /
``autocorr.integrated_time`` / ``EnsembleSampler.get_autocorr_time`` (zero-padded FFT autocorrelation of the
ensemble-mean series, Sokal's self-consistent window); there is no golden vector to pin it to, so it is
additionally cross-checked against an independent direct-sum evaluation and against AR(1) series whose
autocorrelation time is known in closed form (tests/test_autocorr.py).

One deliberate reading: with ``fast=True`` emcee documents that the series is cropped to the largest power of
two ("restricted to 2^n samples", as the reference's own config comment says, config_example.py:255-259); the
restatement crops, while ``size`` -- hence the default ``high`` and the give-up test -- keeps using the full
length, as in emcee.
"""
import numpy as np


class AutocorrError(Exception):
    pass


def function(x, axis=0, fast=False):
    """Normalised autocorrelation function of a series along ``axis`` (FFT, zero padded to 2n: linear, not circular)."""
    x = np.atleast_1d(x)
    m = [slice(None)] * x.ndim
    if fast:
        n = int(2 ** np.floor(np.log2(x.shape[axis])))
        m[axis] = slice(0, n)
        x = x[tuple(m)]
    else:
        n = x.shape[axis]
    f = np.fft.fft(x - np.mean(x, axis=axis), n=2 * n, axis=axis)
    m[axis] = slice(0, n)
    acf = np.fft.ifft(f * np.conjugate(f), axis=axis)[tuple(m)].real
    m[axis] = 0
    return acf / acf[tuple(m)]


def integrated_time(x, low=10, high=None, step=1, c=10, full_output=False, axis=0, fast=False):
    """Sokal-window estimate: the smallest M in arange(low, high, step) with every tau > 1 and M > c max(tau)."""
    size = 0.5 * x.shape[axis]
    if c * low >= size:
        raise AutocorrError("The chain is too short")
    f = function(x, axis=axis, fast=fast)
    oned = f.ndim == 1
    m = [slice(None)] * f.ndim
    if high is None:
        high = int(size / c)
    for M in np.arange(low, high, step).astype(int):
        if oned:
            tau = 1 + 2 * np.sum(f[1:M])
        else:
            m[axis] = slice(1, M)
            tau = 1 + 2 * np.sum(f[tuple(m)], axis=axis)
        if np.all(tau > 1.0) and M > c * np.max(tau):
            if full_output:
                return tau, M
            return tau
        if c * np.max(tau) >= size:
            break
    raise AutocorrError("The chain is too short to reliably estimate the autocorrelation time")


def get_autocorr_time(chain, low=10, high=None, step=1, c=10, fast=False, full_output=False):
    """``chain [walkers, steps, dim]`` (emcee's sampler.chain of one bin) -> tau [dim]."""
    return integrated_time(np.mean(chain, axis=0), axis=0, low=low, high=high, step=step, c=c, fast=fast,
                           full_output=full_output)


def acf_direct(x, n_lags):
    """Independent check of ``function``: the linear autocorrelation sums, one lag at a time."""
    y = np.asarray(x, dtype=np.float64) - np.mean(x, axis=0)
    out = np.stack([np.sum(y[: len(y) - k] * y[k:], axis=0) for k in range(n_lags)])
    return out / out[0]
