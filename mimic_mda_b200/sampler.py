"""Affine-invariant ensemble sampler driving the batched likelihood, all Ex bins
in lock-step.

The reference runs one ``emcee.EnsembleSampler`` per Ex bin inside a worker
process (mda.py:1128-1131) and emcee calls ``ln_post_prob`` once per walker.
emcee is a third-party package that is neither vendored in the reference nor
installed here, so the move is restated from its published algorithm (Goodman
& Weare 2010 stretch move as implemented by emcee 2.x, the version implied by
mda.py:1253,1307-1308): the ensemble is split into a first and a second half;
for each walker k of the active half

    z  = ((a - 1) u + 1)^2 / a,   u ~ U(0,1),  a = 2
    c  = a uniformly chosen walker of the other half
    q  = c - z (c - x_k)
    accept iff (L - 1) ln z + lnp(q) - lnp(x_k) > ln u',  u' ~ U(0,1)

and the second half then moves against the updated first half.  Per bin the
random draws are made in emcee 2.x's order (u for the whole half, partner
indices for the whole half, then u' for the whole half) from a per-bin
``numpy.random.RandomState``, so bin b's chain is what an independent
single-bin run seeded the same way produces; the only thing batched across
bins is the likelihood call: one launch of ``[B, W/2, L]`` per half-step.

``log_prob_fn(params[B, Wc, L]) -> [B, Wc]`` is any vectorised log-posterior;
``GpuLogProb`` is the product one (pinned staging + captured CUDA graph).
"""
import numpy as np


class GpuLogProb:
    """Vectorised log-posterior over a ``DeviceStore``.

    Called with ``[B, Wc, L]`` it returns ``[B, Wc]``.  ``proposal_buffer(Wc)``
    exposes the store's pinned staging block so the sampler can write proposals
    straight into the memory the H2D copy reads from; ``evaluate_staged(Wc)``
    then replays the per-shape CUDA graph {H2D, kernel, D2H}.
    """

    def __init__(self, store):
        self.store = store
        self._views = {}
        self.calls = 0

    def proposal_buffer(self, walkers):
        if walkers not in self._views:
            self._views = {walkers: self.store.staging(walkers)}   # a larger request moves the buffers
        return self._views[walkers][0]

    def evaluate_staged(self, walkers):
        self.proposal_buffer(walkers)
        self.store.lnpost_staged(walkers)
        self.calls += 1
        return self._views[walkers][1]

    def __call__(self, params):
        params = np.asarray(params, dtype=np.float64)
        buf = self.proposal_buffer(params.shape[1])
        if params is not buf:
            buf[...] = params
        return self.evaluate_staged(params.shape[1]).copy()


class BatchedEnsembleSampler:
    """emcee-2-style stretch-move sampler for B independent bins at once.

    Attributes after ``run_mcmc``: ``chain [B, W, steps, L]`` (emcee's
    ``sampler.chain`` per bin, mda.py:1253), ``lnprobability [B, W, steps]``,
    ``naccepted [B, W]``, ``acceptance_fraction [B, W]`` (mda.py:1316).
    """

    def __init__(self, n_bins, n_walkers, n_dim, log_prob_fn, a=2.0, seeds=None, rng_mode="per_bin"):
        if n_walkers % 2 != 0:
            raise ValueError("The number of walkers must be even.")
        if n_walkers < 2 * n_dim:
            raise ValueError("The number of walkers needs to be more than twice the dimension of your parameter space.")
        if rng_mode not in ("per_bin", "shared"):
            raise ValueError("rng_mode must be 'per_bin' or 'shared'")
        self.n_bins, self.k, self.dim, self.a = n_bins, n_walkers, n_dim, float(a)
        self.log_prob_fn = log_prob_fn
        self.rng_mode = rng_mode
        if seeds is None:
            seeds = [None] * n_bins
        elif np.ndim(seeds) == 0:
            seeds = [int(seeds) + b for b in range(n_bins)]
        if rng_mode == "per_bin":
            self.rngs = [np.random.RandomState(s) for s in seeds]
        else:
            self.rng = np.random.default_rng(seeds[0])
        self.reset()

    def reset(self):
        self.iterations = 0
        self.naccepted = np.zeros((self.n_bins, self.k), dtype=np.int64)
        self.chain = np.empty((self.n_bins, self.k, 0, self.dim))
        self.lnprobability = np.empty((self.n_bins, self.k, 0))

    @property
    def acceptance_fraction(self):
        return self.naccepted / max(self.iterations, 1)

    # -- pieces ---------------------------------------------------------------------
    def _lnprob(self, block):
        if np.any(np.isinf(block)):
            raise ValueError("At least one parameter value was infinite.")
        if np.any(np.isnan(block)):
            raise ValueError("At least one parameter value was NaN.")
        if hasattr(self.log_prob_fn, "proposal_buffer") and block is self.log_prob_fn.proposal_buffer(block.shape[1]):
            out = self.log_prob_fn.evaluate_staged(block.shape[1])
        else:
            out = np.asarray(self.log_prob_fn(block), dtype=np.float64)
        if np.any(np.isnan(out)):
            raise ValueError("lnprob returned NaN.")
        return out

    def _draw(self, n_s, n_c):
        """Stretch factors [B, Ns] and partner indices [B, Ns]."""
        if self.rng_mode == "shared":
            zz = ((self.a - 1.0) * self.rng.random((self.n_bins, n_s)) + 1.0) ** 2.0 / self.a
            return zz, self.rng.integers(n_c, size=(self.n_bins, n_s))
        zz = np.empty((self.n_bins, n_s))
        rint = np.empty((self.n_bins, n_s), dtype=np.int64)
        for b, rs in enumerate(self.rngs):
            zz[b] = ((self.a - 1.0) * rs.rand(n_s) + 1.0) ** 2.0 / self.a
            rint[b] = rs.randint(n_c, size=(n_s,))
        return zz, rint

    def _uniform(self, n_s):
        if self.rng_mode == "shared":
            return self.rng.random((self.n_bins, n_s))
        u = np.empty((self.n_bins, n_s))
        for b, rs in enumerate(self.rngs):
            u[b] = rs.rand(n_s)
        return u

    def _propose_stretch(self, s, c, lnprob0):
        n_s, n_c = s.shape[1], c.shape[1]
        zz, rint = self._draw(n_s, n_c)
        partner = np.take_along_axis(c, rint[:, :, None], axis=1)          # c[rint] per bin
        if hasattr(self.log_prob_fn, "proposal_buffer"):
            q = self.log_prob_fn.proposal_buffer(n_s)
            np.subtract(partner, zz[:, :, None] * (partner - s), out=q)
        else:
            q = partner - zz[:, :, None] * (partner - s)
        newlnprob = self._lnprob(q)
        lnpdiff = (self.dim - 1.0) * np.log(zz) + newlnprob - lnprob0
        accept = lnpdiff > np.log(self._uniform(n_s))
        return q, newlnprob, accept

    # -- driver -----------------------------------------------------------------------
    def run_mcmc(self, p0, n_steps, lnprob0=None, store_chain=True, thin=1):
        """Advance every bin's ensemble ``n_steps`` steps from ``p0 [B, W, L]``.
        Returns (positions, lnprob) after the last step."""
        p = np.array(p0, dtype=np.float64)
        if p.shape != (self.n_bins, self.k, self.dim):
            raise ValueError("p0 must be [bins, walkers, dim]")
        halfk = self.k // 2
        lnprob = np.array(self._lnprob(p) if lnprob0 is None else lnprob0, dtype=np.float64)
        if store_chain:
            n_keep = n_steps // thin
            old = self.chain.shape[2]
            self.chain = np.concatenate((self.chain, np.zeros((self.n_bins, self.k, n_keep, self.dim))), axis=2)
            self.lnprobability = np.concatenate((self.lnprobability, np.zeros((self.n_bins, self.k, n_keep))), axis=2)
        first, second = slice(halfk), slice(halfk, self.k)
        for i in range(int(n_steps)):
            self.iterations += 1
            for s0, s1 in ((first, second), (second, first)):
                q, newlnp, acc = self._propose_stretch(p[:, s0], p[:, s1], lnprob[:, s0])
                if np.any(acc):
                    lnprob[:, s0][acc] = newlnp[acc]
                    p[:, s0][acc] = q[acc]
                    self.naccepted[:, s0][acc] += 1
            if store_chain and (i + 1) % thin == 0:
                ind = old + i // thin
                self.chain[:, :, ind, :] = p
                self.lnprobability[:, :, ind] = lnprob
        return p, lnprob


def summarize_chain(chain, burn_in=0):
    """Per bin and dimension the 16/50/84 percentiles and mean of the flattened
    post-burn-in samples -- the quantities mda.py:1253-1259,1429-1431 extract."""
    samples = chain[:, :, burn_in:, :].reshape(chain.shape[0], -1, chain.shape[3])
    pct = np.percentile(samples, [16.0, 50.0, 84.0], axis=1)            # [3, B, L]
    return {"p16": pct[0], "p50": pct[1], "p84": pct[2], "mean": samples.mean(axis=1), "std": samples.std(axis=1)}
