"""Device-resident ensemble sampling: the whole emcee loop of mda.py:1128-1131 in one launch.

``DeviceEnsemble`` wraps section 6 of include/mda_chisq.h.  Interface follows the
emcee 2.x attributes the reference reads (mda.py:1253, 1316, 1393):

    ens = DeviceEnsemble(store, n_walkers, seed=1234)
    ens.run_mcmc(p0, n_steps)            # p0 [B, W, L]
    ens.chain                            # [B, W, kept, L]   (emcee: sampler.chain per bin)
    ens.lnprobability                    # [B, W, kept]
    ens.acceptance_fraction              # [B, W]

Random numbers are Philox4x32-10 on the device (counter = walker, bin id, half-step,
draw), so chains are statistically -- not seed-for-seed -- comparable with emcee's
Mersenne-Twister runs, but identical across launch shapes and GPU counts.
"""
import ctypes as ct

import numpy as np

from . import chisq

_DP = ct.POINTER(ct.c_double)


def _declare(lib):
    if getattr(lib, "_mda_ensemble_declared", False):
        return
    vp, ll = ct.c_void_p, ct.c_longlong
    sig = {
        "mdaEnsembleCreate": (vp, [vp, ct.c_int, ct.c_ulonglong, ct.c_double]),
        "mdaEnsembleFree": (None, [vp]),
        "mdaEnsembleSetBinIds": (ct.c_int, [vp, ct.POINTER(ct.c_int)]),
        "mdaEnsembleSetState": (ct.c_int, [vp, vp]),
        "mdaEnsembleReserveChain": (ct.c_int, [vp, ll]),
        "mdaEnsembleAdvance": (ct.c_int, [vp, ct.c_int, ct.c_int]),
        "mdaEnsembleSynchronize": (ct.c_int, [vp]),
        "mdaEnsembleSteps": (ll, [vp]),
        "mdaEnsembleKept": (ll, [vp]),
        "mdaEnsembleGetState": (ct.c_int, [vp, vp, vp, vp]),
        "mdaEnsembleGetChain": (ct.c_int, [vp, ll, ll, vp, vp]),
        "mdaEnsembleProfile": (ct.c_int, [vp, vp]),
        "mdaEnsembleDescribe": (ct.c_int, [vp, ct.c_int, ct.c_char_p, ct.c_int]),
        "mdaEnsembleSummarize": (ct.c_int, [vp, ll, ct.c_double, ct.c_int, vp, vp, vp]),
        "mdaEnsembleRunToHost": (ct.c_int, [vp, ct.c_int, ct.c_int, ct.c_int, vp, vp]),
        "mdaEnsembleAutocorrTime": (ct.c_int, [vp, ct.c_int, ct.c_int, ct.c_int, ct.c_double, ct.c_int, vp, vp]),
        "mdaEnsembleDeviceChain": (vp, [vp]),
        "mdaEnsembleDevicePositions": (vp, [vp]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(lib, name)
        fn.restype, fn.argtypes = res, args
    lib._mda_ensemble_declared = True


class AutocorrError(Exception):
    """The chain is too short to estimate the autocorrelation time (emcee.autocorr.AutocorrError)."""


class DeviceEnsemble:
    def __init__(self, store, n_walkers, seed=0, a=2.0, bin_ids=None):
        self.store, self.lib = store, store.lib
        _declare(self.lib)
        self.n_bins, self.k, self.dim = store.n_bins, int(n_walkers), store.n_l
        self.handle = self.lib.mdaEnsembleCreate(store.handle, self.k, int(seed) & (2 ** 64 - 1), float(a))
        if not self.handle:
            msg = self.lib.mdaLastErrorString().decode()
            if self.lib.mdaLastError() == 3:
                raise ValueError(msg)
            raise chisq.MdaError(self.lib.mdaLastError(), msg)
        if bin_ids is not None:
            ids = np.ascontiguousarray(bin_ids, dtype=np.int32)
            if ids.shape != (self.n_bins,):
                raise ValueError("bin_ids must have one entry per local bin")
            chisq.check(self.lib, self.lib.mdaEnsembleSetBinIds(self.handle, ids.ctypes.data_as(ct.POINTER(ct.c_int))))
        self._chain = None

    # -- stepping ---------------------------------------------------------------------
    def set_state(self, p0):
        p0 = np.ascontiguousarray(p0, dtype=np.float64)
        if p0.shape != (self.n_bins, self.k, self.dim):
            raise ValueError("p0 must be [bins, walkers, dim]")
        rc = self.lib.mdaEnsembleSetState(self.handle, p0.ctypes.data)
        if rc == 3:
            raise ValueError(self.lib.mdaLastErrorString().decode())
        chisq.check(self.lib, rc, "mdaEnsembleSetState")
        self._chain = None

    def reserve_chain(self, capacity):
        chisq.check(self.lib, self.lib.mdaEnsembleReserveChain(self.handle, int(capacity)), "mdaEnsembleReserveChain")
        self._chain = None

    def advance(self, n_steps, thin=1, wait=True):
        chisq.check(self.lib, self.lib.mdaEnsembleAdvance(self.handle, int(n_steps), int(thin)), "mdaEnsembleAdvance")
        self._chain = None
        if wait:
            self.synchronize()

    def synchronize(self):
        rc = self.lib.mdaEnsembleSynchronize(self.handle)
        if rc == 5:
            raise ValueError(self.lib.mdaLastErrorString().decode())      # "lnprob returned NaN", as emcee
        chisq.check(self.lib, rc, "mdaEnsembleSynchronize")

    def run_mcmc(self, p0, n_steps, thin=1, store_chain=True):
        """emcee-style driver: returns (positions, lnprob) after the last step."""
        self.set_state(p0)
        if store_chain:
            self.reserve_chain(n_steps // thin)
        self.advance(n_steps, thin)
        return self.state()[:2]

    def run_to_host(self, n_steps, thin=1, chunk_steps=0, chain=None, lnprob=None):
        """Advance ``n_steps`` and stream the kept samples into host arrays while the next
        chunk is being computed.  ``chain [n_steps//thin, B, W, L]`` / ``lnprob [kept, B, W]``
        (step-major, the order the run produces them) may be given, e.g. as views of pinned
        memory; they are allocated otherwise.  ``chain.transpose(1, 2, 0, 3)`` is emcee's
        per-bin ``sampler.chain`` layout [B, W, kept, L]."""
        kept = n_steps // thin
        if chain is None:
            chain = np.empty((kept, self.n_bins, self.k, self.dim))
        if lnprob is None:
            lnprob = np.empty((kept, self.n_bins, self.k))
        assert chain.shape == (kept, self.n_bins, self.k, self.dim) and chain.flags.c_contiguous
        assert lnprob.shape == (kept, self.n_bins, self.k) and lnprob.flags.c_contiguous
        rc = self.lib.mdaEnsembleRunToHost(self.handle, int(n_steps), int(thin), int(chunk_steps),
                                           chain.ctypes.data, lnprob.ctypes.data)
        if rc == 5 and b"NaN" in self.lib.mdaLastErrorString():
            raise ValueError(self.lib.mdaLastErrorString().decode())
        chisq.check(self.lib, rc, "mdaEnsembleRunToHost")
        self._chain = None
        return chain, lnprob

    # -- results -----------------------------------------------------------------------
    @property
    def iterations(self):
        return int(self.lib.mdaEnsembleSteps(self.handle))

    @property
    def kept(self):
        return int(self.lib.mdaEnsembleKept(self.handle))

    def state(self):
        pos = np.empty((self.n_bins, self.k, self.dim))
        lnp = np.empty((self.n_bins, self.k))
        acc = np.empty((self.n_bins, self.k), dtype=np.int64)
        chisq.check(self.lib, self.lib.mdaEnsembleGetState(self.handle, pos.ctypes.data, lnp.ctypes.data, acc.ctypes.data),
                    "mdaEnsembleGetState")
        return pos, lnp, acc

    def _fetch(self):
        if self._chain is None:
            kept = self.kept
            chain = np.empty((kept, self.n_bins, self.k, self.dim))
            lnp = np.empty((kept, self.n_bins, self.k))
            chisq.check(self.lib, self.lib.mdaEnsembleGetChain(self.handle, 0, kept, chain.ctypes.data, lnp.ctypes.data),
                        "mdaEnsembleGetChain")
            self._chain = (chain, lnp)
        return self._chain

    @property
    def chain(self):
        """[bins, walkers, kept, dim] view (per bin the layout of emcee's sampler.chain)."""
        return self._fetch()[0].transpose(1, 2, 0, 3)

    @property
    def lnprobability(self):
        return self._fetch()[1].transpose(1, 2, 0)

    @property
    def naccepted(self):
        return self.state()[2]

    @property
    def acceptance_fraction(self):
        return self.naccepted / max(self.iterations, 1)

    def summarize(self, burn_in=0, confidence_interval=0.682689492, num_bins=1000):
        """Posterior summaries computed on the device chain (mda.py:1253-1259, 1399-1491):
        ``points [B, L, 3]`` and ``peaks [B, L, 3]`` as (value, +error, -error) -- the
        reference's percentile and histogram-peak estimates -- and the mean acceptance
        fraction per bin.  ``burn_in`` counts kept samples."""
        points = np.empty((self.n_bins, self.dim, 3))
        peaks = np.empty((self.n_bins, self.dim, 3))
        acc = np.empty(self.n_bins)
        chisq.check(self.lib, self.lib.mdaEnsembleSummarize(self.handle, int(burn_in), float(confidence_interval), int(num_bins),
                                                            points.ctypes.data, peaks.ctypes.data, acc.ctypes.data),
                    "mdaEnsembleSummarize")
        return {"points": points, "peaks": peaks, "acceptance_fraction": acc}

    def get_autocorr_time(self, low=10, high=None, step=1, c=10, fast=False, quiet=False):
        """emcee 2.x's ``EnsembleSampler.get_autocorr_time`` for every bin (mda.py:1307-1308 passes
        ``c=CONFIG["ACorr WindSize"], fast=CONFIG["ACorr Use FFT"]``), computed on the device chain:
        ``tau [B, L]``.  Where emcee raises ``AutocorrError`` this does too, naming the bins; with
        ``quiet=True`` those bins' rows are NaN instead.  ``self.autocorr_window [B]`` holds the accepted
        window sizes (0: none)."""
        tau = np.empty((self.n_bins, self.dim))
        window = np.zeros(self.n_bins, dtype=np.int32)
        chisq.check(self.lib, self.lib.mdaEnsembleAutocorrTime(self.handle, int(low), 0 if high is None else int(high), int(step),
                                                               float(c), int(bool(fast)), tau.ctypes.data, window.ctypes.data),
                    "mdaEnsembleAutocorrTime")
        self.autocorr_window = window
        if not quiet and np.any(window == 0):
            bad = np.flatnonzero(window == 0)
            raise AutocorrError("The chain is too short to reliably estimate the autocorrelation time (bins %s)"
                                % ", ".join(str(b) for b in bad[:8]) + (" ..." if len(bad) > 8 else ""))
        return tau

    def profile(self):
        """Per-bin debug counters [bins, 128] of the last advance (first call enables them; see
        mdaEnsembleProfile in include/mda_chisq.h)."""
        out = np.zeros((self.n_bins, 128), dtype=np.int64)
        chisq.check(self.lib, self.lib.mdaEnsembleProfile(self.handle, out.ctypes.data), "mdaEnsembleProfile")
        return out

    def describe(self, n_steps=1):
        """The stepping kernel and schedule an advance of ``n_steps`` would use (text)."""
        buf = ct.create_string_buffer(1024)
        chisq.check(self.lib, self.lib.mdaEnsembleDescribe(self.handle, int(n_steps), buf, len(buf)), "mdaEnsembleDescribe")
        return buf.value.decode()

    def close(self):
        if getattr(self, "handle", None):
            self.lib.mdaEnsembleFree(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
