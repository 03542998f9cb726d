// ensemble_kernel.cu -- device-resident affine-invariant ensemble stepping.
//
// Replaces the reference's per-step host loop  emcee.EnsembleSampler.run_mcmc ->
// ln_post_prob per walker  (mda.py:1128-1131, 1540-1578)  by ONE persistent launch per
// run: a CTA owns an Ex bin for all steps; the bin's constants, the walker positions and
// their log-probabilities live in shared memory, and per half-step every thread
//   1. draws its walkers' stretch factor z, partner index and acceptance uniform from
//      Philox4x32-10 keyed by (seed; walker, bin id, half-step) -- counter based, so the
//      chain does not depend on launch shape or on how bins are spread over GPUs,
//   2. proposes q = c - z (c - x) against the complementary half of the ensemble
//      (Goodman & Weare stretch move as emcee 2.x implements it, see sampler.py),
//   3. evaluates the box prior and the chi-square log-likelihood with the same
//      register-tiled inner loop as the batched kernel (likelihood.cuh, reference order
//      C/ChiSquare.c:188-212),
//   4. accepts iff (L-1) ln z + lnp(q) - lnp(x) > ln u and updates its walker in place.
// Nothing crosses PCIe per step; kept samples go to an HBM chain [kept][bin][W][L].
#include "likelihood.cuh"
#include "philox.cuh"

#ifndef MDA_ENS_UNROLL
#define MDA_ENS_UNROLL 2
#endif

namespace mda {

namespace {

constexpr int kStatusNaN = 1;          // a log-probability came out NaN (emcee raises on this)

// Shared-memory carve-up, identical on host (size) and device (pointers).
struct EnsembleSmem {
    size_t consts, pos, lnp, q, part, z, uacc, fast, partner, outside, acc, total;
    // state_in_smem = false: positions, log-probabilities and acceptance counters stay in HBM/L2
    // (ensembles too large for shared memory, e.g. the reference's 2500 walkers x 8 parameters)
    __host__ __device__ EnsembleSmem(int walkers, int n_l, int n_pad, int groups, bool state_in_smem) {
        const size_t lp = (size_t)(n_l | 1), h = (size_t)walkers / 2;
        size_t off = 0;
        consts = off;  off += (size_t)n_pad * (size_t)(n_l + 1) * 8;     // [n_pad/2][1+L] double2
        pos = off;     off += state_in_smem ? (size_t)walkers * (size_t)n_l * 8 : 0;   // [W][L] dense: bulk-stored as is
        lnp = off;     off += state_in_smem ? (size_t)walkers * 8 : 0;                 // [W]
        off = (off + 15) & ~(size_t)15;
        q = off;       off += h * lp * 8;                                // [H][LP]  proposals of the half-step
        part = off;    off += (size_t)groups * h * 8;                    // [G][H]   partial chi^2 per angle group
        z = off;       off += 2 * h * 8;                                 // [2][H]   stretch factor, double buffered
        uacc = off;    off += 2 * h * 8;                                 // [2][H]   acceptance uniform, double buffered
        fast = off;    off += 2 * h * 4;                                 // [2][H]   float (L-1) ln z - ln u
        partner = off; off += h * 4;                                     // [H] int
        outside = off; off += h * 4;                                     // [H] int
        acc = off;     off += state_in_smem ? (size_t)walkers * 4 : 0;   // [W] unsigned
        total = (off + 15) & ~(size_t)15;
    }
};

__device__ __forceinline__ void fence_proxy_async_smem() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void bulk_s2g(void *dst, const void *src, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(smem_u32(src)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read_all() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

// Acceptance uniform from the Philox bits the 53-bit stretch uniform leaves unused:
// word w plus the 5 + 6 discarded low bits of x and y -> 43 bits in [0, 1).
__host__ __device__ __forceinline__ double uniform43(uint32_t w, uint32_t x, uint32_t y) {
    return ((double)w * 2048.0 + (double)(((x & 31u) << 6) | (y & 63u))) * (1.0 / 8796093022208.0);
}

// One CTA = one Ex bin for the whole run.  blockDim = T_w * G: thread (tw, g) evaluates the
// chi^2 contribution of angle group g for WPT walkers; the stretch-move bookkeeping of a
// half-step is spread over the threads.  Per half-step t (= 2*step + half):
//   phase A  all threads : partial chi^2 of the staged proposals  +  the Philox draw for t+1
//                          (the integer work fills the issue slots the DFMA stream leaves
//                          free); thread 0 also launches the bulk store of a finished step
//   phase B  H threads   : sum partials in ascending group order, accept / update in place
//   phase C  H threads   : stage the proposals q = c - z (c - x) of half-step t+1
template <int L, int WPT, bool GS>          // GS: ensemble state in global memory instead of shared
__global__ void __launch_bounds__(512) ensemble_kernel(const EnsembleArgs args) {
    constexpr int LP = L | 1;
    constexpr int ROW = L + 1;

    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ __align__(8) uint64_t mbar;

    const int tid = threadIdx.x;
    const int T = blockDim.x;
    const int G = args.groups;
    const int Tw = T / G;
    const int tw = tid % Tw;
    const int g = tid / Tw;                       // uniform per warp (Tw is a multiple of 32)
    const int bin = blockIdx.x;
    const int W = args.walkers;
    const int H = W >> 1;                         // walkers updated per half-step
    const int n_pad = args.n_pad;

    const EnsembleSmem lay(W, L, n_pad, G, !GS);
    double2 *sC = reinterpret_cast<double2 *>(smem_raw + lay.consts);
    double *gPos = args.pos + (size_t)bin * W * L;
    double *gLnp = args.lnp + (size_t)bin * W;
    unsigned int *gAcc = args.nacc + (size_t)bin * W;
    double *sPos = GS ? gPos : reinterpret_cast<double *>(smem_raw + lay.pos);
    double *sLnp = GS ? gLnp : reinterpret_cast<double *>(smem_raw + lay.lnp);
    double *sQ = reinterpret_cast<double *>(smem_raw + lay.q);
    double *sPart = reinterpret_cast<double *>(smem_raw + lay.part);
    double *sZ = reinterpret_cast<double *>(smem_raw + lay.z);
    double *sUacc = reinterpret_cast<double *>(smem_raw + lay.uacc);
    float *sFast = reinterpret_cast<float *>(smem_raw + lay.fast);
    int *sPartner = reinterpret_cast<int *>(smem_raw + lay.partner);
    int *sOutside = reinterpret_cast<int *>(smem_raw + lay.outside);
    unsigned int *sAcc = GS ? gAcc : reinterpret_cast<unsigned int *>(smem_raw + lay.acc);

    if (tid == 0) {
        mbar_init(&mbar, 1);
        fence_mbar_init();
    }
    __syncthreads();
    if (tid == 0) {                                // constants, positions, log-probabilities: bulk copies
        const uint32_t cb = (uint32_t)n_pad * ROW * 8u, pb = GS ? 0u : (uint32_t)W * L * 8u, lb = GS ? 0u : (uint32_t)W * 8u;
        mbar_arrive_expect_tx(&mbar, cb + pb + lb);
        bulk_g2s(sC, reinterpret_cast<const double2 *>(args.consts) + (size_t)bin * (n_pad / 2) * ROW, cb, &mbar);
        if (!GS) {
            bulk_g2s(sPos, gPos, pb, &mbar);
            bulk_g2s(sLnp, gLnp, lb, &mbar);
        }
    }
    if (!GS)
        for (int w = tid; w < W; w += T) sAcc[w] = gAcc[w];
    const int n = args.n_pts[bin];
    const uint32_t bin_id = (uint32_t)args.bin_ids[bin];
    const uint32_t k0 = args.seed_lo, k1 = args.seed_hi;
    const double a_scale = args.a;
    const double dim_m1 = (double)(L - 1);
    const float dim_m1f = (float)(L - 1);

    // this thread's slice of the angle pairs (the last group takes an unpaired last angle)
    const int n_pairs = (n + 1) >> 1;
    const int pair0 = (int)(((long long)g * n_pairs) / G);
    const int pair1 = (int)(((long long)(g + 1) * n_pairs) / G);
    const int my_valid = max(0, min(2 * pair1, n) - 2 * pair0);

    // The Philox draw of half-step t (counter-based: independent of positions, so it is made
    // one half-step ahead).  emcee draws per half: u -> z, the partner index, then u'.
    auto draw = [&](uint32_t t) {
        const int c_base = (t & 1u) ? 0 : H;
        const int s_base = (t & 1u) ? H : 0;
        double *uacc = sUacc + (size_t)(t & 1u) * H;
        double *zbuf = sZ + (size_t)(t & 1u) * H;
        float *fast = sFast + (size_t)(t & 1u) * H;
        for (int w = tid; w < H; w += T) {
            const Philox4 r = philox4x32_10_keys(Philox4{(uint32_t)(s_base + w), bin_id, t, 0u}, args.round_keys);
            const double u = uniform53(r.x, r.y);
            // numpy arithmetic is one rounding per operation: no FMA contraction here
            const double zr = __dadd_rn(__dmul_rn(a_scale - 1.0, u), 1.0);
            const double zz = __dmul_rn(zr, zr);
            const double z = (a_scale == 2.0) ? zz * 0.5 : zz / a_scale;        // ((a-1) u + 1)^2 / a
            const double ua = uniform43(r.w, r.x, r.y);
            zbuf[w] = z;
            uacc[w] = ua;
            // single-precision estimate of (L-1) ln z - ln u'; decides all but near-ties
            fast[w] = dim_m1f * __logf((float)z) - __logf((float)ua);
            sPartner[w] = c_base + (int)(((uint64_t)r.z * (uint64_t)H) >> 32);
        }
    };
    // Proposals of half-step t: q = c - z (c - x), box prior flag.
    auto propose = [&](uint32_t t) {
        const int s_base = (t & 1u) ? H : 0;
        const double *zbuf = sZ + (size_t)(t & 1u) * H;
        for (int i = tid; i < H; i += T) {
            const double z = zbuf[i];
            const double *xs = sPos + (size_t)(s_base + i) * L;
            const double *xc = sPos + (size_t)sPartner[i] * L;
            double *q = sQ + (size_t)i * LP;
            bool outside = false;
#pragma unroll
            for (int l = 0; l < L; ++l) {
                const double c = xc[l];
                const double v = __dsub_rn(c, __dmul_rn(z, __dsub_rn(c, xs[l])));
                q[l] = v;
                outside = outside || (v < args.lo[l]) || (v > args.hi[l]);    // mda.py:1575
            }
            sOutside[i] = outside ? 1 : 0;
        }
    };

    // chain bookkeeping without 64-bit divisions: steps until the next kept sample, next slot
    const bool keeping = args.chain != nullptr;
    int until_keep = args.thin - (int)((uint64_t)args.step0 % (uint64_t)args.thin);
    long long slot = args.keep0;
    auto store_step = [&]() {                      // thread 0, after a barrier that followed phase B
        if (!keeping || slot >= args.keep_cap) return;
        bulk_s2g(args.chain + (((size_t)slot * args.bins + (size_t)bin) * W) * L, sPos, (uint32_t)W * L * 8u);
        bulk_s2g(args.lnp_chain + ((size_t)slot * args.bins + (size_t)bin) * W, sLnp, (uint32_t)W * 8u);
        bulk_commit();
    };
    auto copy_step = [&]() {                       // global-state variant: all threads, HBM -> HBM
        if (!keeping || slot >= args.keep_cap) return;
        double *gc = args.chain + (((size_t)slot * args.bins + (size_t)bin) * W) * L;
        for (int i = tid; i < W * L; i += T) gc[i] = sPos[i];
        double *gl = args.lnp_chain + ((size_t)slot * args.bins + (size_t)bin) * W;
        for (int w = tid; w < W; w += T) gl[w] = sLnp[w];
    };

    const bool prof = args.prof != nullptr && tid == 0;
    long long t_like = 0, t_acc = 0, t_prop = 0, t_chi = 0, t_draw = 0, t_mark = 0;
#define MDA_SECTION(accum)                        \
    if (prof) {                                   \
        const long long now__ = clock64();        \
        accum += now__ - t_mark;                  \
        t_mark = now__;                           \
    }

    const uint32_t t_begin = 2u * (uint32_t)args.step0;
    const uint32_t t_end = t_begin + 2u * (uint32_t)args.n_steps;
    mbar_wait(&mbar, 0);
    __syncthreads();
    draw(t_begin);
    __syncthreads();
    propose(t_begin);
    __syncthreads();
    bool saw_nan = false;
    bool store_pending = false;                    // a finished step waits to be bulk-stored
    if (prof) t_mark = clock64();

    for (uint32_t t = t_begin; t < t_end; ++t) {
        const int s_base = (t & 1u) ? H : 0;
        // ---- phase A: likelihood partials (+ bulk store of the finished step, + the draw for t+1)
        if (GS) { if (store_pending) copy_step(); }
        else if (tid == 0 && store_pending) store_step();
        for (int wb = 0; wb < H; wb += Tw * WPT) {
            double a[WPT][L], chi[WPT];
#pragma unroll
            for (int k = 0; k < WPT; ++k) {
                const double *q = sQ + (size_t)min(wb + k * Tw + tw, H - 1) * LP;
                chi[k] = 0.0;
#pragma unroll
                for (int l = 0; l < L; ++l) a[k][l] = q[l];
            }
            accumulate_chi<L, WPT, MDA_ENS_UNROLL>(sC + (size_t)pair0 * ROW, my_valid, a, chi);
#pragma unroll
            for (int k = 0; k < WPT; ++k) {
                const int i = wb + k * Tw + tw;
                if (i < H) sPart[(size_t)g * H + i] = chi[k];
            }
        }
        MDA_SECTION(t_chi)
        if (t + 1 < t_end) draw(t + 1);
        MDA_SECTION(t_draw)
        if (!GS && tid == 0 && store_pending) bulk_wait_read_all();   // the store has read sPos/sLnp: they may change again
        if (store_pending) { store_pending = false; ++slot; }
        __syncthreads();
        MDA_SECTION(t_like)
        // ---- phase B: accept / reject, update the active half in place
        {
            const double *uacc = sUacc + (size_t)(t & 1u) * H;
            const float *fast = sFast + (size_t)(t & 1u) * H;
            for (int i = tid; i < H; i += T) {
                double chi = sPart[i];
                for (int gg = 1; gg < G; ++gg) chi += sPart[(size_t)gg * H + i];
                const double newlnp = sOutside[i] ? -CUDART_INF : chi * -0.5;       // C/ChiSquare.c:212
                saw_nan = saw_nan || (newlnp != newlnp);
                const int gw = s_base + i;
                const double delta = __dsub_rn(newlnp, sLnp[gw]);
                const double est = delta + (double)fast[i];
                bool accept;
                if (fabs(est) > 1e-3) {
                    accept = est > 0.0;            // far from a tie: the float logs cannot change the sign
                } else {                           // emcee's test in full precision
                    const double zlog = __dmul_rn(dim_m1, log(sZ[(size_t)(t & 1u) * H + i]));
                    accept = __dsub_rn(__dadd_rn(zlog, newlnp), sLnp[gw]) > log(uacc[i]);
                }
                if (accept) {
                    const double *q = sQ + (size_t)i * LP;
                    double *xs = sPos + (size_t)gw * L;
#pragma unroll
                    for (int l = 0; l < L; ++l) xs[l] = q[l];
                    sLnp[gw] = newlnp;
                    sAcc[gw] += 1u;
                }
            }
            if (!GS && keeping && (t & 1u)) fence_proxy_async_smem();     // order these writes before a bulk store
        }
        __syncthreads();
        MDA_SECTION(t_acc)
        if (t & 1u) {                              // a full step is complete
            if (--until_keep == 0) {
                until_keep = args.thin;
                store_pending = keeping;
            }
        }
        // ---- phase C: stage the proposals of the next half-step against the updated half
        if (t + 1 < t_end) {
            propose(t + 1);
            __syncthreads();
        }
        MDA_SECTION(t_prop)
    }
    if (GS) {
        if (store_pending) copy_step();
    } else if (tid == 0) {
        if (store_pending) store_step();
        bulk_wait_all();
    }
    if (prof) {
        long long *o = args.prof + (size_t)bin * 128;
        o[0] = t_prop; o[1] = t_like; o[2] = t_acc; o[3] = t_chi; o[4] = t_draw;
    }
#undef MDA_SECTION

    if (!GS) {
        for (int i = tid; i < W * L; i += T) gPos[i] = sPos[i];
        for (int w = tid; w < W; w += T) {
            gLnp[w] = sLnp[w];
            gAcc[w] = sAcc[w];
        }
    }
    if (saw_nan) atomicOr(args.status, kStatusNaN);
}

using EnsKernel = void (*)(const EnsembleArgs);

template <int L>
EnsKernel pick_ens_wpt(int wpt, bool global_state) {
    if (global_state) return wpt == 2 ? ensemble_kernel<L, 2, true> : (wpt == 1 ? ensemble_kernel<L, 1, true> : nullptr);
    switch (wpt) {
        case 1: return ensemble_kernel<L, 1, false>;
        case 2: return ensemble_kernel<L, 2, false>;
        default: return nullptr;
    }
}

EnsKernel pick_ens_kernel(int n_l, int wpt, bool global_state) {
    switch (n_l) {
        case 1: return pick_ens_wpt<1>(wpt, global_state);
        case 2: return pick_ens_wpt<2>(wpt, global_state);
        case 3: return pick_ens_wpt<3>(wpt, global_state);
        case 4: return pick_ens_wpt<4>(wpt, global_state);
        case 5: return pick_ens_wpt<5>(wpt, global_state);
        case 6: return pick_ens_wpt<6>(wpt, global_state);
        case 7: return pick_ens_wpt<7>(wpt, global_state);
        case 8: return pick_ens_wpt<8>(wpt, global_state);
        case 9: return pick_ens_wpt<9>(wpt, global_state);
        case 10: return pick_ens_wpt<10>(wpt, global_state);
        case 11: return pick_ens_wpt<11>(wpt, global_state);
        case 12: return pick_ens_wpt<12>(wpt, global_state);
        case 13: return pick_ens_wpt<13>(wpt, global_state);
        case 14: return pick_ens_wpt<14>(wpt, global_state);
        case 15: return pick_ens_wpt<15>(wpt, global_state);
        case 16: return pick_ens_wpt<16>(wpt, global_state);
        default: return nullptr;
    }
}

}  // namespace

EnsembleShape choose_ensemble_shape(int bins, int walkers, int n_l, int n_pad, int sm_count, size_t smem_optin,
                                    int force_groups, int force_wpt) {
    EnsembleShape s{};
    const int half = walkers / 2;
    s.wpt = half >= 64 ? 2 : 1;
    if (force_wpt == 1 || force_wpt == 2) s.wpt = force_wpt;
    int tw = (half + s.wpt - 1) / s.wpt;
    tw = (tw + 31) / 32 * 32;
    if (tw > 256) tw = 256;
    // angle groups: enough warps to hide latency; two CTAs must still fit an SM (registers)
    // when there are more bins than SMs
    const int max_threads = bins > sm_count ? 256 : 512;
    int groups = force_groups > 0 ? force_groups : max_threads / tw;
    const int n_pairs = (n_pad + 1) / 2;
    if (groups > n_pairs / 2) groups = n_pairs / 2;
    if (groups > 8) groups = 8;
    if (groups < 1) groups = 1;
    while (groups > 1 && tw * groups > 512) --groups;
    s.groups = groups;
    s.threads = tw * groups;
    s.smem_bytes = EnsembleSmem(walkers, n_l, n_pad, groups, true).total;
    s.global_state = false;
    s.mma = false;
    if (s.smem_bytes > smem_optin) {               // the ensemble state moves to HBM/L2; shrink the partials if need be
        s.global_state = true;
        while (s.groups > 1 && EnsembleSmem(walkers, n_l, n_pad, s.groups, false).total > smem_optin) --s.groups;
        s.threads = tw * s.groups;
        s.smem_bytes = EnsembleSmem(walkers, n_l, n_pad, s.groups, false).total;
    }
    s.fits = n_l >= 1 && n_l <= kMaxRegL && s.smem_bytes <= smem_optin &&
             (size_t)n_pad * (size_t)(n_l + 1) * 8 < (1u << 20);
    return s;
}

cudaError_t configure_ensemble_kernels(size_t smem_optin) {
    for (int l = 1; l <= kMaxRegL; ++l)
        for (int wpt = 1; wpt <= 2; ++wpt)
            for (int gs = 0; gs < 2; ++gs) {
                cudaError_t e = cudaFuncSetAttribute(reinterpret_cast<const void *>(pick_ens_kernel(l, wpt, gs != 0)),
                                                     cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_optin);
                if (e != cudaSuccess) return e;
            }
    return cudaSuccess;
}

cudaError_t launch_ensemble(const EnsembleArgs &args, const EnsembleShape &shape, cudaStream_t stream) {
    if (args.bins <= 0 || args.n_steps <= 0) return cudaSuccess;
    EnsKernel k = pick_ens_kernel(args.n_l, shape.wpt, shape.global_state);
    if (!k || !shape.fits) return cudaErrorInvalidValue;
    EnsembleArgs a = args;
    a.groups = shape.groups;
    k<<<a.bins, shape.threads, shape.smem_bytes, stream>>>(a);
    return cudaGetLastError();
}

}  // namespace mda
