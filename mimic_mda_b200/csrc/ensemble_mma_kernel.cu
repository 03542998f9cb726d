// ensemble_mma_kernel.cu -- device-resident ensemble stepping with the chi-square contraction on
// the fp64 tensor path (DMMA.8x8x4), SURVEY.md 8f-1.
//
// Same job and same random streams as ensemble_kernel.cu (the reference's host loop
// emcee.EnsembleSampler.run_mcmc -> ln_post_prob per walker, mda.py:1128-1131, 1540-1578, as ONE
// persistent launch; a CTA owns an Ex bin).  What differs is how a half-step is laid out:
//
//   * A warp owns tiles of 8 walkers.  Lane (g, tig) = (lane / 4, lane % 4) holds, for walker g
//     of the tile, the proposal elements l = R + 4 ks + tig (ks < Q = L / 4) -- exactly the A
//     fragment of mma.m8n8k4 -- and, replicated over the quad, the R = L % 4 leading elements.
//   * residuals of an 8-walker x 8-angle tile: c = d - sum_{l<R} q_l D_l by DFMA, then Q DMMAs
//     against B fragments of the constants staged fragment-major in shared memory; chi^2 is
//     accumulated from the C fragment and reduced over the quad by two shuffles.  Every term of
//     the reference's sum (C/ChiSquare.c:188-210) is evaluated exactly once, only the order of
//     the fp64 additions differs (|error| ~ 1e-16 relative, tested against the oracle).
//     Why: a DFMA with three distinct register operands is register-bandwidth bound at ~72 % of
//     the fp64 pipe on sm_100a; DMMA.8x8x4 runs on the same pipe at its full rate with 1/6 of the
//     operand traffic (tools/ubench/dmma_ubench.cu).
//   * the warp that evaluated a tile also draws its Philox numbers, builds its proposals and
//     accepts / rejects it: no proposal or partial-sum staging and ONE barrier per half-step.
//   * a warp owns the same walker tiles for the whole run: it alone writes their rows (in place, on
//     acceptance) and it copies them to the HBM chain after a kept step -- no staging, no fences.
#include "likelihood.cuh"
#include "philox.cuh"

namespace mda {

namespace {

constexpr int kStatusNaN = 1;

struct EnsembleMmaSmem {
    size_t consts, pos, lnp, acc, total;
    __host__ __device__ EnsembleMmaSmem(int walkers, int n_l, int n_tiles) {
        size_t off = 0;
        consts = off;  off += (size_t)n_tiles * 8 * (size_t)(n_l + 1) * 8;     // [tiles][8 (1+L)] fragment-major
        pos = off;     off += (size_t)walkers * (size_t)n_l * 8;               // [W][L] dense, updated in place
        lnp = off;     off += (size_t)walkers * 8;                             // [W]
        acc = off;     off += (size_t)walkers * 4;                             // [W] accepted proposals
        total = (off + 15) & ~(size_t)15;
    }
};

__device__ __forceinline__ void dmma884(double &c0, double &c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}

__host__ __device__ __forceinline__ double uniform43(uint32_t w, uint32_t x, uint32_t y) {
    return ((double)w * 2048.0 + (double)(((x & 31u) << 6) | (y & 63u))) * (1.0 / 8796093022208.0);
}

// Work of a warp: walker tiles tile_first + k (k < WT) + j * tile_stride of BOTH halves, for the whole
// run.  Only their owner writes the rows of these walkers (in place, on acceptance); the other warps
// read them as partners, and only in the half-step in which they are not being updated.
template <int L, int WT>
__global__ void __launch_bounds__(512) ensemble_mma_kernel(const EnsembleArgs args) {
    constexpr int Q = L / 4, R = L % 4, RR = R > 0 ? R : 1;
    constexpr int TS = 8 * (L + 1);               // doubles per angle tile of the constants

    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ __align__(8) uint64_t mbar;

    const int tid = threadIdx.x;
    const int T = blockDim.x;
    const int lane = tid & 31, warp = tid >> 5, n_warps = T >> 5;
    const int g = lane >> 2, tig = lane & 3;
    const int bin = blockIdx.x;
    const int W = args.walkers;
    const int H = W >> 1;
    const int NT = args.n_tiles;
    const int w_tiles = (H + 7) >> 3;

    const EnsembleMmaSmem lay(W, L, NT);
    const double *sC = reinterpret_cast<const double *>(smem_raw + lay.consts);
    double *sPos = reinterpret_cast<double *>(smem_raw + lay.pos);
    double *sLnp = reinterpret_cast<double *>(smem_raw + lay.lnp);
    unsigned int *sAcc = reinterpret_cast<unsigned int *>(smem_raw + lay.acc);

    if (tid == 0) {
        mbar_init(&mbar, 1);
        fence_mbar_init();
    }
    __syncthreads();
    double *gPos = args.pos + (size_t)bin * W * L;
    double *gLnp = args.lnp + (size_t)bin * W;
    unsigned int *gAcc = args.nacc + (size_t)bin * W;
    if (tid == 0) {
        const uint32_t cb = (uint32_t)NT * TS * 8u;
        mbar_arrive_expect_tx(&mbar, cb);
        bulk_g2s(smem_raw + lay.consts, args.consts_mma + (size_t)bin * NT * TS, cb, &mbar);
    }
    for (int i = tid; i < W * L; i += T) sPos[i] = gPos[i];
    for (int w = tid; w < W; w += T) {
        sLnp[w] = gLnp[w];
        sAcc[w] = gAcc[w];
    }

    const uint32_t bin_id = (uint32_t)args.bin_ids[bin];
    const uint32_t k0 = args.seed_lo, k1 = args.seed_hi;
    const double a_scale = args.a;
    const double dim_m1 = (double)(L - 1);
    const float dim_m1f = (float)(L - 1);

    // box prior of the elements this lane holds (mda.py:1124, 1574-1576)
    double lo_r[RR], hi_r[RR], lo_q[Q], hi_q[Q];
#pragma unroll
    for (int l = 0; l < R; ++l) { lo_r[l] = args.lo[l]; hi_r[l] = args.hi[l]; }
#pragma unroll
    for (int ks = 0; ks < Q; ++ks) { lo_q[ks] = args.lo[R + 4 * ks + tig]; hi_q[ks] = args.hi[R + 4 * ks + tig]; }

    const bool keeping = args.chain != nullptr;
    int until_keep = args.thin - (int)((uint64_t)args.step0 % (uint64_t)args.thin);
    long long slot = args.keep0;
    bool saw_nan = false;

    const bool prof = args.prof != nullptr && tid == 0;
    long long t_prop = 0, t_chi = 0, t_acc = 0, t_bar = 0, t_draw = 0, t_store = 0, t_mark = 0;
#define MDA_SECTION(accum)                        \
    if (prof) {                                   \
        const long long now__ = clock64();        \
        accum += now__ - t_mark;                  \
        t_mark = now__;                           \
    }

    const uint32_t t_begin = 2u * (uint32_t)args.step0;
    const uint32_t t_end = t_begin + 2u * (uint32_t)args.n_steps;

    // The Philox draw of one work item (half-step t, WT walker tiles from tile0): it does not depend
    // on positions, so it is made one item ahead, and by half of the warps before and by the other
    // half after their chi^2 loop -- integer work of some warps under the DMMA stream of the others.
    // Lane j draws for walker 8 tile0 + j; tile k's lanes fetch their walker's values by shuffle.
    struct Draw { double z[WT]; float fast[WT]; int partner[WT]; };
    auto draw = [&](uint32_t t, int tile0) {
        const int c_base = (t & 1u) ? 0 : H;
        const int s_base = (t & 1u) ? H : 0;
        const int i = min(tile0 * 8 + (lane & (8 * WT - 1)), H - 1);
        const Philox4 r = philox4x32_10(Philox4{(uint32_t)(s_base + i), bin_id, t, 0u}, k0, k1);
        const double u = uniform53(r.x, r.y);
        // numpy arithmetic is one rounding per operation: no FMA contraction here
        const double zr = __dadd_rn(__dmul_rn(a_scale - 1.0, u), 1.0);
        const double z2 = __dmul_rn(zr, zr);
        const double z = (a_scale == 2.0) ? z2 * 0.5 : z2 / a_scale;            // ((a-1) u + 1)^2 / a
        const double ua = uniform43(r.w, r.x, r.y);
        // single-precision estimate of (L-1) ln z - ln u'; decides all but near-ties
        const float fast = dim_m1f * __logf((float)z) - __logf((float)ua);
        const int partner = c_base + (int)(((uint64_t)r.z * (uint64_t)H) >> 32);
        Draw d;
#pragma unroll
        for (int k = 0; k < WT; ++k) {
            d.z[k] = __shfl_sync(0xffffffffu, z, 8 * k + g);
            d.fast[k] = __shfl_sync(0xffffffffu, fast, 8 * k + g);
            d.partner[k] = __shfl_sync(0xffffffffu, partner, 8 * k + g);
        }
        return d;
    };
    const bool draw_early = ((warp >> 2) & 1) == 0;       // warps w, w + 4, ... share a scheduler: two of each kind
    const int tile_first = warp * WT, tile_stride = n_warps * WT;
    const int elem0 = R + tig;                            // first DMMA element of this lane within a row

    mbar_wait(&mbar, 0);
    __syncthreads();
    Draw dr = draw(t_begin, min(tile_first, w_tiles - 1));
    if (prof) t_mark = clock64();

    for (uint32_t t = t_begin; t < t_end; ++t) {
        const uint32_t half = t & 1u;
        const int s_base = half ? H : 0;          // the half being updated; partners come from the other one

        for (int tile0 = tile_first; tile0 < w_tiles; tile0 += tile_stride) {
            const bool last_pass = tile0 + tile_stride >= w_tiles;
            const uint32_t t_nxt = last_pass ? t + 1 : t;
            const int tile_nxt = last_pass ? tile_first : tile0 + tile_stride;
            const bool has_next = t_nxt < t_end;
            Draw dn = dr;
            // ---- propose: q = c - z (c - x), as -q in A-fragment order
            double A[WT][Q > 0 ? Q : 1], nq[WT][RR];
            int gw[WT];
            bool outside[WT];
#pragma unroll
            for (int k = 0; k < WT; ++k) {
                gw[k] = s_base + min((tile0 + k) * 8 + g, H - 1);
                const double z = dr.z[k];
                const double *xs = sPos + gw[k] * L;
                const double *xc = sPos + dr.partner[k] * L;
                bool out = false;
#pragma unroll
                for (int l = 0; l < R; ++l) {
                    const double c = xc[l];
                    const double v = __dsub_rn(c, __dmul_rn(z, __dsub_rn(c, xs[l])));
                    nq[k][l] = -v;
                    out = out || (v < lo_r[l]) || (v > hi_r[l]);                    // mda.py:1575
                }
#pragma unroll
                for (int ks = 0; ks < Q; ++ks) {
                    const double c = xc[elem0 + 4 * ks];
                    const double v = __dsub_rn(c, __dmul_rn(z, __dsub_rn(c, xs[elem0 + 4 * ks])));
                    A[k][ks] = -v;
                    out = out || (v < lo_q[ks]) || (v > hi_q[ks]);
                }
                const unsigned int ballot = __ballot_sync(0xffffffffu, out);
                outside[k] = ((ballot >> (lane & ~3)) & 0xFu) != 0u;
            }
            MDA_SECTION(t_prop)
            if (draw_early && has_next) dn = draw(t_nxt, tile_nxt);
            MDA_SECTION(t_draw)

            // ---- chi^2 of WT tiles of 8 walkers against every angle tile
            double chi0[WT], chi1[WT];
#pragma unroll
            for (int k = 0; k < WT; ++k) { chi0[k] = 0.0; chi1[k] = 0.0; }
#pragma unroll 2
            for (int at = 0; at < NT; ++at) {
                const double *tp = sC + at * TS;
                const double2 dv = *reinterpret_cast<const double2 *>(tp + 2 * tig);
                double2 Di[RR];
#pragma unroll
                for (int l = 0; l < R; ++l) Di[l] = *reinterpret_cast<const double2 *>(tp + 8 * (1 + l) + 2 * tig);
                double Bf[Q > 0 ? Q : 1];
#pragma unroll
                for (int ks = 0; ks < Q; ++ks) Bf[ks] = tp[8 * (1 + R) + 32 * ks + lane];
                double c0[WT], c1[WT];
#pragma unroll
                for (int k = 0; k < WT; ++k) {
                    c0[k] = dv.x;
                    c1[k] = dv.y;
#pragma unroll
                    for (int l = 0; l < R; ++l) {
                        c0[k] = fma(nq[k][l], Di[l].x, c0[k]);
                        c1[k] = fma(nq[k][l], Di[l].y, c1[k]);
                    }
                }
#pragma unroll
                for (int ks = 0; ks < Q; ++ks)
#pragma unroll
                    for (int k = 0; k < WT; ++k) dmma884(c0[k], c1[k], A[k][ks], Bf[ks]);
#pragma unroll
                for (int k = 0; k < WT; ++k) {
                    chi0[k] = fma(c0[k], c0[k], chi0[k]);
                    chi1[k] = fma(c1[k], c1[k], chi1[k]);
                }
            }
            MDA_SECTION(t_chi)
            if (!draw_early && has_next) dn = draw(t_nxt, tile_nxt);
            MDA_SECTION(t_draw)

            // ---- accept / reject, in place (only this warp ever writes these rows)
#pragma unroll
            for (int k = 0; k < WT; ++k) {
                double chi = chi0[k] + chi1[k];
                chi += __shfl_xor_sync(0xffffffffu, chi, 1);
                chi += __shfl_xor_sync(0xffffffffu, chi, 2);
                const double newlnp = chi * -0.5;                                   // C/ChiSquare.c:212
                saw_nan = saw_nan || (newlnp != newlnp && !outside[k]);
                const double oldlnp = sLnp[gw[k]];
                const double est = __dsub_rn(newlnp, oldlnp) + (double)dr.fast[k];
                bool accept;
                if (fabs(est) > 1e-3) {
                    accept = est > 0.0;            // far from a tie: the float logs cannot change the sign
                } else {                           // emcee's test in full precision (rare: redo the draw for u')
                    const Philox4 r = philox4x32_10(Philox4{(uint32_t)gw[k], bin_id, t, 0u}, k0, k1);
                    const double zlog = __dmul_rn(dim_m1, log(dr.z[k]));
                    accept = __dsub_rn(__dadd_rn(zlog, newlnp), oldlnp) > log(uniform43(r.w, r.x, r.y));
                }
                // outside the box the log-posterior is -inf (mda.py:1576): never accepted
                accept = accept && !outside[k] && ((tile0 + k) * 8 + g < H);
                if (accept) {
                    double *nx = sPos + gw[k] * L;
#pragma unroll
                    for (int l = 0; l < R; ++l)
                        if (tig == l) nx[l] = -nq[k][l];
#pragma unroll
                    for (int ks = 0; ks < Q; ++ks) nx[elem0 + 4 * ks] = -A[k][ks];
                    if (tig == 0) {
                        sLnp[gw[k]] = newlnp;
                        sAcc[gw[k]] += 1u;
                    }
                }
            }
            dr = dn;
            MDA_SECTION(t_acc)
        }
        __syncthreads();
        MDA_SECTION(t_bar)
        if (half && --until_keep == 0) {           // a full step is complete and kept: every warp copies out ITS rows
            until_keep = args.thin;
            if (keeping && slot < args.keep_cap) {
                double *gc = args.chain + (((size_t)slot * args.bins + (size_t)bin) * W) * L;
                double *gl = args.lnp_chain + ((size_t)slot * args.bins + (size_t)bin) * W;
                for (int tile0 = tile_first; tile0 < w_tiles; tile0 += tile_stride) {
                    const int w0 = tile0 * 8, nw = min(8 * WT, H - w0);
#pragma unroll
                    for (int hb = 0; hb < 2; ++hb) {
                        const int row0 = hb * H + w0;
                        for (int i = lane; i < nw * L; i += 32) gc[row0 * L + i] = sPos[row0 * L + i];
                        if (lane < nw) gl[row0 + lane] = sLnp[row0 + lane];
                    }
                }
                ++slot;
            }
            MDA_SECTION(t_store)
        }
    }
    if (prof) {
        long long *o = args.prof + (size_t)bin * 8;
        o[0] = t_prop; o[1] = t_chi + t_bar; o[2] = t_acc; o[3] = t_chi; o[4] = t_draw; o[5] = t_bar; o[6] = t_store;
    }
#undef MDA_SECTION

    __syncthreads();
    for (int i = tid; i < W * L; i += T) gPos[i] = sPos[i];
    for (int w = tid; w < W; w += T) {
        gLnp[w] = sLnp[w];
        gAcc[w] = sAcc[w];
    }
    if (saw_nan) atomicOr(args.status, kStatusNaN);
}

using EnsKernel = void (*)(const EnsembleArgs);

template <int L>
EnsKernel pick_mma_wt(int wt) {
    switch (wt) {
        case 1: return ensemble_mma_kernel<L, 1>;
        case 2: return ensemble_mma_kernel<L, 2>;
        case 4: return ensemble_mma_kernel<L, 4>;
        default: return nullptr;
    }
}

EnsKernel pick_mma_kernel(int n_l, int wt) {
    switch (n_l) {
        case 4: return pick_mma_wt<4>(wt);
        case 5: return pick_mma_wt<5>(wt);
        case 6: return pick_mma_wt<6>(wt);
        case 7: return pick_mma_wt<7>(wt);
        case 8: return pick_mma_wt<8>(wt);
        case 9: return pick_mma_wt<9>(wt);
        case 10: return pick_mma_wt<10>(wt);
        case 11: return pick_mma_wt<11>(wt);
        case 12: return pick_mma_wt<12>(wt);
        case 13: return pick_mma_wt<13>(wt);
        case 14: return pick_mma_wt<14>(wt);
        case 15: return pick_mma_wt<15>(wt);
        case 16: return pick_mma_wt<16>(wt);
        default: return nullptr;
    }
}

}  // namespace

bool choose_ensemble_mma_shape(int bins, int walkers, int n_l, int n_tiles, int sm_count, size_t smem_optin, int force_wt,
                               EnsembleShape *out) {
    if (n_l < kMinMmaL || n_l > kMaxRegL) return false;
    EnsembleShape s{};
    const int half = walkers / 2;
    const int w_tiles = (half + 7) / 8;
    const int max_warps = bins > sm_count ? 8 : 16;            // two CTAs per SM need 2 x threads x registers <= 64 K
    s.wpt = w_tiles >= 32 ? 4 : (w_tiles >= 16 ? 2 : 1);       // at least 8 warps where the ensemble allows
    if (force_wt == 1 || force_wt == 2 || force_wt == 4) s.wpt = force_wt;
    int warps = (w_tiles + s.wpt - 1) / s.wpt;
    if (warps > max_warps) warps = max_warps;
    s.threads = 32 * warps;
    s.groups = 1;
    s.smem_bytes = EnsembleMmaSmem(walkers, n_l, n_tiles).total;
    s.global_state = false;
    s.mma = true;
    s.fits = s.smem_bytes <= smem_optin && (size_t)n_tiles * 8 * (size_t)(n_l + 1) * 8 < (1u << 20);
    if (!s.fits) return false;
    *out = s;
    return true;
}

cudaError_t configure_ensemble_mma_kernels(size_t smem_optin) {
    for (int l = kMinMmaL; l <= kMaxRegL; ++l)
        for (int wt = 1; wt <= 4; wt *= 2) {
            cudaError_t e = cudaFuncSetAttribute(reinterpret_cast<const void *>(pick_mma_kernel(l, wt)),
                                                 cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_optin);
            if (e != cudaSuccess) return e;
        }
    return cudaSuccess;
}

cudaError_t launch_ensemble_mma(const EnsembleArgs &args, const EnsembleShape &shape, cudaStream_t stream) {
    if (args.bins <= 0 || args.n_steps <= 0) return cudaSuccess;
    EnsKernel k = pick_mma_kernel(args.n_l, shape.wpt);
    if (!k || !shape.fits || !shape.mma || !args.consts_mma) return cudaErrorInvalidValue;
    k<<<args.bins, shape.threads, shape.smem_bytes, stream>>>(args);
    return cudaGetLastError();
}

}  // namespace mda
