"""CPU restatement of the device-resident ensemble stepper -- TEST INFRASTRUCTURE ONLY.

Follows mimic_mda_b200/csrc/ensemble_kernel.cu step for step with numpy: the same
Philox4x32-10 counters, the same stretch-move arithmetic (emcee 2.x semantics, see
mimic_mda_b200/sampler.py), the likelihood from the C/numpy oracle.  Used to check that
the device sampler does what it says (chains are expected to be identical, because
every decision compares quantities that agree to ~1e-15).
"""
import numpy as np

M0, M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
W0, W1 = np.uint32(0x9E3779B9), np.uint32(0xBB67AE85)
MASK = np.uint64(0xFFFFFFFF)


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    """Vectorised Philox4x32-10 (Salmon et al., SC'11).  All inputs uint32 arrays/scalars."""
    c0, c1, c2, c3 = (np.asarray(c, dtype=np.uint64) & MASK for c in np.broadcast_arrays(c0, c1, c2, c3))
    k0 = np.uint64(k0) & MASK
    k1 = np.uint64(k1) & MASK
    for _ in range(10):
        p0 = M0 * c0
        p1 = M1 * c2
        hi0, lo0 = p0 >> np.uint64(32), p0 & MASK
        hi1, lo1 = p1 >> np.uint64(32), p1 & MASK
        c0, c1, c2, c3 = (hi1 ^ c1 ^ k0) & MASK, lo1, (hi0 ^ c3 ^ k1) & MASK, lo0
        k0 = (k0 + np.uint64(W0)) & MASK
        k1 = (k1 + np.uint64(W1)) & MASK
    return c0, c1, c2, c3


def uniform43(w, x, y):
    """Acceptance uniform from the bits uniform53 leaves unused (word w + 5 + 6 low bits of x, y)."""
    low = ((x & np.uint64(31)) << np.uint64(6)) | (y & np.uint64(63))
    return (w.astype(np.float64) * 2048.0 + low.astype(np.float64)) / 8796093022208.0


def uniform53(a, b):
    return ((a >> np.uint64(5)).astype(np.float64) * 67108864.0 + (b >> np.uint64(6)).astype(np.float64)) / 9007199254740992.0


def run(lnpost_fn, p0, n_steps, seed, bin_ids=None, a=2.0, thin=1, step0=0, lnprob0=None, naccepted0=None, epoch=0):
    """Advance ``p0 [B, W, L]`` by ``n_steps`` steps.  ``lnpost_fn([B, Wc, L]) -> [B, Wc]``.
    Returns dict(pos, lnp, nacc, chain [B, kept, W, L], lnp_chain [B, kept, W])."""
    p = np.array(p0, dtype=np.float64)
    n_bins, n_walk, n_l = p.shape
    bin_ids = np.arange(n_bins, dtype=np.uint64) if bin_ids is None else np.asarray(bin_ids, dtype=np.uint64)
    # the Philox key: the seed for an ensemble's first state, seed + epoch * 2^64/phi for its epoch-th restart
    # (mdaEnsembleSetState called again on the same ensemble; chisq_abi.cu: mdaEnsembleAdvance)
    seed = (seed + epoch * 0x9E3779B97F4A7C15) & 0xFFFFFFFFFFFFFFFF
    k0, k1 = seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF
    lnp = np.array(lnpost_fn(p) if lnprob0 is None else lnprob0, dtype=np.float64)
    nacc = np.zeros((n_bins, n_walk), dtype=np.int64) if naccepted0 is None else np.array(naccepted0)
    halfk = n_walk // 2
    chain, lnp_chain = [], []
    for step in range(step0, step0 + n_steps):
        for half in (0, 1):
            s_sl = slice(halfk, n_walk) if half else slice(0, halfk)
            c_base = 0 if half else halfk
            n_s = n_walk - halfk if half else halfk
            n_c = n_walk - n_s
            gw = np.arange(s_sl.start, s_sl.stop, dtype=np.uint64)[None, :]
            t = np.uint64((2 * step + half) & 0xFFFFFFFF)
            r = philox4x32_10(gw, bin_ids[:, None], t, np.uint64(0), k0, k1)
            u = uniform53(r[0], r[1])
            zr = (a - 1.0) * u + 1.0
            z = zr * zr / a
            partner = c_base + ((r[2] * np.uint64(n_c)) >> np.uint64(32)).astype(np.int64)
            c = np.take_along_axis(p, partner[:, :, None], axis=1)
            s = p[:, s_sl]
            q = c - z[:, :, None] * (c - s)
            newlnp = np.asarray(lnpost_fn(q), dtype=np.float64)
            with np.errstate(divide="ignore", invalid="ignore"):
                lnpdiff = (n_l - 1.0) * np.log(z) + newlnp - lnp[:, s_sl]
                accept = lnpdiff > np.log(uniform43(r[3], r[0], r[1]))
            p[:, s_sl][accept] = q[accept]
            lnp[:, s_sl][accept] = newlnp[accept]
            nacc[:, s_sl][accept] += 1
        if (step + 1) % thin == 0:
            chain.append(p.copy())
            lnp_chain.append(lnp.copy())
    return {"pos": p, "lnp": lnp, "nacc": nacc,
            "chain": np.stack(chain, axis=1) if chain else np.empty((n_bins, 0, n_walk, n_l)),
            "lnp_chain": np.stack(lnp_chain, axis=1) if lnp_chain else np.empty((n_bins, 0, n_walk))}
