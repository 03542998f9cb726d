// ensemble_ws_kernel.cu -- device-resident ensemble stepping with the chi-square contraction on
// the fp64 tensor path (DMMA.8x8x4) and WARP-SPECIALISED bookkeeping, SURVEY.md 8f-1.
//
// The schedule for up to ~2 bins per SM: 16 warps work on ONE bin, so that the draw and the accept tests of
// the walker warps run under the proposals and the operand prefetch of the math warps.  With two or more bins
// per SM the homogeneous 8-warp kernel of ensemble_mma_kernel.cu (two CTAs per SM) is faster; the C ABI picks.
//
// Same job and same random streams as ensemble_kernel.cu (the reference's host loop
// emcee.EnsembleSampler.run_mcmc -> ln_post_prob per walker, mda.py:1128-1131, 1540-1578, as ONE
// persistent launch).  A CTA owns an Ex bin (for a chunk of steps) and has two kinds of warps:
//
//   * math warps: one item of 8 WT walkers per half-step.  Lane (g, tig) = (lane / 4, lane % 4) holds, for
//     walker g of each tile of 8, the proposal components l = R + 4 ks + tig (ks < Q = L / 4) -- the A fragment
//     of mma.m8n8k4 -- and a share of the R = L % 4 leading ones.  It gathers the partners' components c and
//     forms q = c - z (c - x) itself (numpy rounding, no FMA contraction) from operands x, z and partner
//     addresses it fetched while the walker warps decided the half-step before, stages q for the walker
//     warps' row update, tests the box prior (mda.py:1574-1576) on those registers (outside: chi^2 = +inf,
//     i.e. log-posterior -inf); residuals of an 8-walker x 8-angle tile are c = d + sum_{l<R} q_l (-D_l) by
//     DFMA, then Q DMMAs against B fragments of the NEGATED distributions staged fragment-major in shared
//     memory; chi^2 is accumulated from the C fragment and reduced over the quad.  Every term of the
//     reference's sum (C/ChiSquare.c:188-210) is evaluated exactly once, only the order of the fp64 additions
//     differs (|error| ~ 1e-16 relative, tested against the oracle).  Why DMMA: a DFMA with three distinct
//     register operands is register-bandwidth bound at ~72 % of the fp64 pipe on sm_100a; DMMA.8x8x4 runs on
//     the same pipe at its full rate with 1/8 of the issue slots (tools/ubench/dmma_ubench.cu).
//   * walker warps (lane = walker of warp = group of 32): the Philox draw of the NEXT half-step while the
//     math warps work on this one, then accept / reject, in-place row update from the staged proposals,
//     acceptance counters and the bulk store of kept steps.
//   * hand-offs are NAMED BARRIERS (bar.sync / bar.arrive), never polled: barrier 1 completes when the draw of
//     the next half-step is published (walker warps arrive, math warps wait), barrier 2 + g when the chi^2 of
//     group g is there (its math warps arrive, walker warp g waits), barrier 10 joins the walker warps around
//     a pending chain store, and ONE CTA barrier ends a half-step (the stretch move needs the other half
//     complete).  A bar.arrive costs ~150 cycles (it waits for the thread's shared-memory stores), so there
//     is one per warp and half-step.  An mbarrier wait compiles to a try_wait / NANOSLEEP.SYNCS loop that
//     wakes up every ~100 cycles: in the mbarrier version of this kernel 25 % of all executed instructions
//     were such probes (profiles/r01_ncu_full_c4_ensemble_mma.csv), taken from the math warps' issue slots.
//   * CTAs walk a list of (step chunk, bin) units in chunk-major order, unit u on CTA u mod grid; a
//     unit waits for the previous chunk of its bin through a release/acquire word per bin and hands
//     the ensemble state over through HBM/L2.  With more bins than SMs (C4: 200 on 148) every SM
//     carries the same share of the work instead of 52 SMs carrying two bins for the whole run.
//
// Shapes: every walker warp owns one group of 32 walkers of a half and every math warp one item, so
// ensembles of up to 512 walkers (all BASELINE.json shapes); larger ones run on ensemble_kernel.cu.
#include "mma_chi.cuh"
#include "philox.cuh"

namespace mda {

namespace {

constexpr int kStatusNaN = 1;
constexpr int kStatusStalled = 4;        // a unit waited 4 s (wall clock) for its predecessor: the launch gave up (never seen; a guard against a hung GPU)
constexpr int kWalkerWarps = 8;        // at most: one group of 32 walkers each
constexpr int kMathWarps = 8;          // at most: two per scheduler keep the fp64 pipe fed
constexpr int GW = 32;                 // walkers of a group (a walker warp's lanes)
constexpr int kDrawnBarrier = 1;       // named barrier: the draw of the next half-step is published (walker warps arrive, math warps wait)
constexpr int kDoneBarrier = 2;        // named barriers 2 .. 9: chi^2 of a group is there (its math warps arrive, its walker warp waits)
constexpr int kJoinBarrier = 10;       // named barrier of the walker warps
constexpr size_t kStaticSmem = 64;     // the mbarrier of the constants (static __shared__), rounded up

struct EnsembleWsSmem {
    size_t consts, pos, q, lnp, chi, z, bounds, fast, partner, acc, total;
    __host__ __device__ EnsembleWsSmem(int walkers, int n_l, int n_tiles) {
        const size_t h = (size_t)walkers / 2;
        size_t off = 0;
        consts = off;  off += (size_t)n_tiles * 8 * (size_t)(n_l + 1) * 8;     // [tiles][8 (1+L)] fragment-major
        pos = off;     off += (size_t)walkers * (size_t)n_l * 8;               // [W][L] dense, updated in place; bulk-stored as is
        lnp = off;     off += (size_t)walkers * 8;                             // [W] (16-byte aligned: W is even)
        q = off;       off += h * (size_t)n_l * 8;                             // [H][L] proposals of the half-step
        chi = off;     off += h * 8;                                           // [H] chi^2 of the proposals (+inf: outside the box)
        z = off;       off += h * 8;                                           // [H] stretch factor of the next proposal
        bounds = off;  off += 2 * (size_t)n_l * 8;                             // [2][L] box prior: lo, hi
        fast = off;    off += 2 * h * 4;                                       // [2][H] float (L-1) ln z - ln u'
        partner = off; off += h * 4;                                           // [H]
        acc = off;     off += (size_t)walkers * 4;                             // [W] accepted proposals
        total = (off + 15) & ~(size_t)15;
    }
};

__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void bulk_s2g(void *dst, const void *src, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(smem_u32(src)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read_all() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ unsigned long long ld_acquire_u64(const unsigned long long *p) {
    unsigned long long v;
    asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_u64(unsigned long long *p, unsigned long long v) {
    asm volatile("st.release.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

__host__ __device__ __forceinline__ double uniform43(uint32_t w, uint32_t x, uint32_t y) {
    return ((double)w * 2048.0 + (double)(((x & 31u) << 6) | (y & 63u))) * (1.0 / 8796093022208.0);
}

// Bank conflicts of one 64-bit access of a half-warp to the own rows: 4 rows (consecutive g), 4 consecutive
// components each, the rows `row_step` rows of L doubles apart.  Returns the worst number of rows sharing a bank.
__host__ __device__ constexpr int own_row_conflicts(int n_l, int row_step) {
    int worst = 0;
    for (int bank = 0; bank < 32; ++bank) {
        int hits = 0;
        for (int j = 0; j < 4; ++j) {
            const int start = (2 * n_l * row_step * j) % 32;
            if ((bank - start + 32) % 32 < 8) ++hits;
        }
        worst = hits > worst ? hits : worst;
    }
    return worst;
}

// Named barriers: a waiting warp is parked by the hardware and issues nothing.
__device__ __forceinline__ void named_sync(int id, int count) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory"); }
__device__ __forceinline__ void named_arrive(int id, int count) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(count) : "memory"); }

// blockDim = 32 (n_walker_warps + n_math_warps); warps [0, n_walker_warps) are walker warps, one per
// group of 32 walkers of a half; math warp m computes item m (8 WT walkers) of every half-step.
template <int L, int WT, bool PROF>               // WT: tiles of 8 walkers per math item (4: a whole group)
__global__ void __launch_bounds__(32 * (kWalkerWarps + kMathWarps), 1) ensemble_ws_kernel(const EnsembleArgs args) {
    constexpr int Q = L / 4, R = L % 4, RR = R > 0 ? R : 1;
    constexpr int IW = 8 * WT;                    // walkers of a math item
    constexpr int IPG = GW / IW;                  // math items per group
    constexpr int TS = 8 * (L + 1);               // doubles per angle tile of the constants
    constexpr int NLE = (WT * R + 3) / 4;         // leading elements (l < R) a lane proposes; the quad shares them
    constexpr int NLS = NLE > 0 ? NLE : 1;

    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ __align__(8) uint64_t mbar;                 // constants of a unit have landed

    const int tid = threadIdx.x;
    const int T = blockDim.x;
    const int lane = tid & 31, warp = tid >> 5;
    const int W = args.walkers;
    const int H = W >> 1;
    const int NT = args.n_tiles;
    const int n_items = (H + IW - 1) / IW;
    const int n_ww = args.groups;                 // walker warps = groups of a half (the launch sets it)
    const int NWT = 32 * n_ww;                    // walker threads
    const bool walker_warp = warp < n_ww;

    const EnsembleWsSmem lay(W, L, NT);
    const double *sC = reinterpret_cast<const double *>(smem_raw + lay.consts);
    double *sPos = reinterpret_cast<double *>(smem_raw + lay.pos);
    double *sQ = reinterpret_cast<double *>(smem_raw + lay.q);
    double *sLnp = reinterpret_cast<double *>(smem_raw + lay.lnp);
    double *sChi = reinterpret_cast<double *>(smem_raw + lay.chi);
    double *sZ = reinterpret_cast<double *>(smem_raw + lay.z);
    double *sLo = reinterpret_cast<double *>(smem_raw + lay.bounds);
    double *sHi = sLo + L;
    float *sFast = reinterpret_cast<float *>(smem_raw + lay.fast);
    int *sPartner = reinterpret_cast<int *>(smem_raw + lay.partner);
    unsigned int *sAcc = reinterpret_cast<unsigned int *>(smem_raw + lay.acc);

    __shared__ int s_abort;
    if (tid == 0) {
        s_abort = 0;
        mbar_init(&mbar, 1);
        fence_mbar_init();
    }
    if (tid < L) { sLo[tid] = args.lo[tid]; sHi[tid] = args.hi[tid]; }
    __syncthreads();

    const uint32_t seed0 = args.seed_lo, seed1 = args.seed_hi;
    const double a_scale = args.a;
    const double dim_m1 = (double)(L - 1);
    const float dim_m1f = (float)(L - 1);

    auto items_of = [&](int g) { return min(IPG, n_items - g * IPG); };

    bool saw_nan = false;
    uint32_t unit_phase = 0;                               // parity of mbar
    bool store_pending = false;                            // a bulk store of a kept step may still be reading the state
    const bool keeping = args.chain != nullptr;

    const long long n_units = (long long)args.n_chunks * args.bins;
    for (long long unit = blockIdx.x; unit < n_units; unit += gridDim.x) {
        const int chunk = (int)(unit / args.bins);
        const int bin = (int)(unit - (long long)chunk * args.bins);
        const long long step_first = args.step0 + (long long)chunk * args.chunk_steps;
        const int steps_here = min(args.chunk_steps, args.n_steps - chunk * args.chunk_steps);
        double *gPos = args.pos + (size_t)bin * W * L;
        double *gLnp = args.lnp + (size_t)bin * W;
        unsigned int *gAcc = args.nacc + (size_t)bin * W;

        // ---- take the bin over: constants by bulk copy, state written by the previous chunk's CTA
        if (tid == 0) {
            const uint32_t cb = (uint32_t)NT * TS * 8u;
            mbar_arrive_expect_tx(&mbar, cb);
            bulk_g2s(smem_raw + lay.consts, args.consts_mma + (size_t)bin * NT * TS, cb, &mbar);
            if (chunk > 0) {
                const unsigned long long want = args.launch_serial + (unsigned long long)chunk;
                const unsigned long long t_start = globaltimer_ns();
                while (ld_acquire_u64(args.progress + bin) != want) {
                    __nanosleep(64);
                    if (globaltimer_ns() - t_start > kHandOverTimeoutNs) { s_abort = 1; break; }     // 4 s of wall clock: the predecessor is not coming
                }
            }
        }
        __syncthreads();
        if (s_abort) {                                     // every CTA that depends on this bin gives up the same way
            mbar_wait(&mbar, unit_phase);                  // (the bulk copy of the constants must land before the CTA may exit)
            if (tid == 0) atomicOr(args.status, kStatusStalled);
            break;
        }
        for (int i = tid; i < W * L; i += T) sPos[i] = __ldcg(gPos + i);
        for (int w = tid; w < W; w += T) {
            sLnp[w] = __ldcg(gLnp + w);
            sAcc[w] = __ldcg(gAcc + w);
        }
        fence_proxy_async_smem();                          // these rows may be bulk-stored before being rewritten

        const uint32_t bin_id = (uint32_t)args.bin_ids[bin];
        int until_keep = args.thin - (int)((uint64_t)step_first % (uint64_t)args.thin);
        long long slot = args.keep0 + (step_first / args.thin - args.step0 / args.thin);
        const uint32_t t_begin = 2u * (uint32_t)step_first;
        const uint32_t t_end = t_begin + 2u * (uint32_t)steps_here;

        mbar_wait(&mbar, unit_phase);
        unit_phase ^= 1u;

        // debug timeline (mdaEnsembleProfile): clock stamps of one half-step, per warp
        const uint32_t t_prof = (PROF && args.prof != nullptr && chunk == 0) ? t_begin + min(81u, 2u * (uint32_t)steps_here - 1u) : 0xffffffffu;
#define MDA_STAMP(k) if (PROF && stamp) stamp[k] = clock64();
#define MDA_STAMP_AFTER(k, value) if (PROF && stamp && (value) != 123.456) stamp[k] = clock64();   // once `value` has arrived

        if (walker_warp) {
            // The Philox draw of walker i of the half updated in half-step t (counter based: it does not
            // depend on positions).  emcee draws per half: u -> z, the partner index, then u'.
            auto draw = [&](uint32_t t, int i) {
                const int s_base = (t & 1u) ? H : 0, c_base = (t & 1u) ? 0 : H;
                const Philox4 r = philox4x32_10_keys(Philox4{(uint32_t)(s_base + i), bin_id, t, 0u}, args.round_keys);
                const double u = uniform53(r.x, r.y);
                // numpy arithmetic is one rounding per operation: no FMA contraction here
                const double zr = __dadd_rn(__dmul_rn(a_scale - 1.0, u), 1.0);
                const double z2 = __dmul_rn(zr, zr);
                const double z = (a_scale == 2.0) ? z2 * 0.5 : z2 / a_scale;              // ((a-1) u + 1)^2 / a
                const double ua = uniform43(r.w, r.x, r.y);
                sZ[i] = z;
                // single-precision estimate of (L-1) ln z - ln u'; decides all but near-ties
                sFast[(t & 1u) * H + i] = dim_m1f * __logf((float)z) - __logf((float)ua);
                sPartner[i] = c_base + (int)__umulhi(r.z, (uint32_t)H);
            };
            const int my_i = min(warp * GW + lane, H - 1);     // my walker of my group
            const int live_w = min(GW, H - warp * GW);
            draw(t_begin, my_i);
            __syncthreads();                                   // state and first draw are there
            __syncthreads();                                   // ... and the math warps have fetched that draw: the next may overwrite it

            for (uint32_t t = t_begin; t < t_end; ++t) {
                const uint32_t half = t & 1u;
                const int s_base = half ? H : 0;              // the half being updated; partners come from the other one
                long long *stamp = (PROF && t == t_prof && lane == 0) ? args.prof + ((size_t)bin * 16 + warp) * 8 : nullptr;
                MDA_STAMP(0)
                // ---- the draw of the next half-step, while the math warps propose and evaluate this one
                if (t + 1 < t_end) {
                    draw(t + 1, my_i);
                    named_arrive(kDrawnBarrier, T);
                }
                MDA_STAMP(2)
                if (store_pending) {                       // the chain store must have read the rows before they change
                    if (tid == 0) bulk_wait_read_all();
                    named_sync(kJoinBarrier, NWT);
                    store_pending = false;
                }
                // ---- accept / reject my walker
                const int gw = s_base + my_i;
                const double oldlnp = sLnp[gw];
                const float fast = sFast[half * H + my_i];
                named_sync(kDoneBarrier + warp, 32 * (1 + items_of(warp)));
                // outside the box the math warps report chi^2 = +inf: log-posterior -inf (mda.py:1576), never accepted
                const double newlnp = sChi[my_i] * -0.5;       // C/ChiSquare.c:212
                MDA_STAMP_AFTER(3, newlnp)
                saw_nan = saw_nan || (newlnp != newlnp);
                const double est = __dsub_rn(newlnp, oldlnp) + (double)fast;
                bool accept;
                if (fabs(est) > 1e-4) {
                    accept = est > 0.0;                // far from a tie: the float logs (error < 2e-5) cannot change the sign
                } else {                               // emcee's test in full precision (rare: redo the draw)
                    const Philox4 r = philox4x32_10_keys(Philox4{(uint32_t)gw, bin_id, t, 0u}, args.round_keys);
                    const double zr = __dadd_rn(__dmul_rn(a_scale - 1.0, uniform53(r.x, r.y)), 1.0);
                    const double z2 = __dmul_rn(zr, zr);
                    const double z = (a_scale == 2.0) ? z2 * 0.5 : z2 / a_scale;
                    const double zlog = __dmul_rn(dim_m1, log(z));
                    accept = __dsub_rn(__dadd_rn(zlog, newlnp), oldlnp) > log(uniform43(r.w, r.x, r.y));
                }
                accept = accept && lane < live_w;
                if (accept) {                          // my row: proposal -> state (rows are 8 L bytes apart: no bank conflicts)
                    double *nrow = sPos + gw * L;
                    const double *qrow = sQ + my_i * L;
                    double qv[L];
#pragma unroll
                    for (int l = 0; l < L; ++l) qv[l] = qrow[l];
#pragma unroll
                    for (int l = 0; l < L; ++l) nrow[l] = qv[l];
                    sLnp[gw] = newlnp;
                    sAcc[gw] += 1u;
                }
                MDA_STAMP(4)
                const bool kept_step = half && until_keep == 1 && keeping && slot < args.keep_cap;
                if (kept_step) fence_proxy_async_smem();   // my rows before the bulk store reads them
                __syncthreads();                           // the half is complete
                if (half && --until_keep == 0) {           // a full step is complete and kept: one bulk store of the whole state
                    until_keep = args.thin;
                    if (kept_step) {
                        if (tid == 0) {
                            bulk_s2g(args.chain + (((size_t)slot * args.bins + (size_t)bin) * W) * L, sPos, (uint32_t)W * L * 8u);
                            bulk_s2g(args.lnp_chain + ((size_t)slot * args.bins + (size_t)bin) * W, sLnp, (uint32_t)W * 8u);
                            bulk_commit();
                        }
                        store_pending = true;
                    }
                    ++slot;
                }
            }
            if (store_pending) {                           // the last step of the unit was kept: the hand-back waits for its store
                if (tid == 0) bulk_wait_read_all();
                store_pending = false;
            }
        } else {
            // ---- my item (WT tiles of 8 walkers) of every half-step: propose, box prior, chi^2 against every angle tile
            const int g = lane >> 2, tig = lane & 3;
            const int item = warp - n_ww;
            const int grp = item / IPG;
            const int n_done = 32 * (1 + items_of(grp));
            // Which walker of the item is row g of tile k: 8 k + g, or WT g + k where that spreads a half-warp's accesses
            // to its own rows (operand prefetch, staged proposals) over more banks (L = 9: rows 4 apart are 8 banks apart).
            constexpr bool KMAJOR = own_row_conflicts(L, WT) < own_row_conflicts(L, 1);
            auto walker_of = [&](int k) { return item * IW + (KMAJOR ? WT * g + k : 8 * k + g); };
            int wi[WT];                                        // my walker of tile k (index within the half)
#pragma unroll
            for (int k = 0; k < WT; ++k) wi[k] = min(walker_of(k), H - 1);
            // leading elements (l < R) of the quad's WT walkers are shared out over its 4 lanes: slot n of lane tig is
            // element j = tig + 4 n of the list (tile j / R, component j % R)
            int lw[NLS], ll[NLS];
#pragma unroll
            for (int n = 0; n < NLE; ++n) {
                const int j = min(tig + 4 * n, WT * RR - 1);
                lw[n] = min(walker_of(j / RR), H - 1);
                ll[n] = j % RR;
            }
            // what the proposals of half-step t need and half-step t - 1 does not change: stretch factors, own
            // components, index of the partner's components
            double pz[WT], pxa[WT][Q], pzl[NLS], pxl[NLS];
            int pca[WT][Q], pcl[NLS];
            auto prefetch = [&](uint32_t t) {
                const double *xhalf = sPos + ((t & 1u) ? H : 0) * L;
#pragma unroll
                for (int k = 0; k < WT; ++k) {
                    pz[k] = sZ[wi[k]];
                    const int part = sPartner[wi[k]] * L;
#pragma unroll
                    for (int ks = 0; ks < Q; ++ks) {
                        pca[k][ks] = part + R + 4 * ks + tig;
                        pxa[k][ks] = xhalf[wi[k] * L + R + 4 * ks + tig];
                    }
                }
#pragma unroll
                for (int n = 0; n < NLE; ++n) {
                    pzl[n] = sZ[lw[n]];
                    pcl[n] = sPartner[lw[n]] * L + ll[n];
                    pxl[n] = xhalf[lw[n] * L + ll[n]];
                }
            };
            __syncthreads();                                   // state and first draw are there
            prefetch(t_begin);
            __syncthreads();                                   // only now may the walker warps publish the next draw

            for (uint32_t t = t_begin; t < t_end; ++t) {
                long long *stamp = (PROF && t == t_prof && lane == 0) ? args.prof + ((size_t)bin * 16 + warp) * 8 : nullptr;
                MDA_STAMP(0)
                // ---- proposals q = c - z (c - x) of my elements: gather the partners' components, 3 fp64 operations each
                double A[WT][Q], lead[WT][RR], myl[NLS];
                {
                    double ca[WT][Q], cl[NLS];
#pragma unroll
                    for (int k = 0; k < WT; ++k)
#pragma unroll
                        for (int ks = 0; ks < Q; ++ks) ca[k][ks] = sPos[pca[k][ks]];
#pragma unroll
                    for (int n = 0; n < NLE; ++n) cl[n] = sPos[pcl[n]];
#pragma unroll
                    for (int k = 0; k < WT; ++k)
#pragma unroll
                        for (int ks = 0; ks < Q; ++ks)
                            A[k][ks] = __dsub_rn(ca[k][ks], __dmul_rn(pz[k], __dsub_rn(ca[k][ks], pxa[k][ks])));
#pragma unroll
                    for (int n = 0; n < NLE; ++n) myl[n] = __dsub_rn(cl[n], __dmul_rn(pzl[n], __dsub_rn(cl[n], pxl[n])));
                }
#pragma unroll
                for (int k = 0; k < WT; ++k)
#pragma unroll
                    for (int l = 0; l < R; ++l) {
                        constexpr int dummy = 0;
                        (void)dummy;
                        const int j = k * R + l;               // compile-time: lane j % 4 of the quad holds it in slot j / 4
                        lead[k][l] = __shfl_sync(0xffffffffu, myl[j / 4], (lane & ~3) | (j & 3));
                    }
                MDA_STAMP_AFTER(1, A[WT - 1][Q - 1])
                double chi0[WT], chi1[WT];
#pragma unroll
                for (int k = 0; k < WT; ++k) { chi0[k] = 0.0; chi1[k] = 0.0; }
                // Staging the proposals for the walker warps and the box prior are independent of the contraction: placed
                // in the middle of it (8 angle tiles: two unrolled halves) their shared-memory traffic and compares
                // issue under the DMMA stream instead of ahead of it.
                bool outside;
                auto stage_and_test = [&]() {
                // the rows the walker warps copy on acceptance
#pragma unroll
                for (int k = 0; k < WT; ++k)
#pragma unroll
                    for (int ks = 0; ks < Q; ++ks) sQ[wi[k] * L + R + 4 * ks + tig] = A[k][ks];
#pragma unroll
                for (int n = 0; n < NLE; ++n)
                    if (tig + 4 * n < WT * R) sQ[lw[n] * L + ll[n]] = myl[n];
                // box prior (mda.py:1575) on these registers: any component below lo or above hi (a NaN compares
                // false, as in the reference).  Lane tig will report the walker g of tile tig % WT.
                {
                    uint32_t out_of[WT];
#pragma unroll
                    for (int k = 0; k < WT; ++k) {
                        bool o = false;
#pragma unroll
                        for (int ks = 0; ks < Q; ++ks) o = o | (A[k][ks] < sLo[R + 4 * ks + tig]) | (A[k][ks] > sHi[R + 4 * ks + tig]);
#pragma unroll
                        for (int l = 0; l < R; ++l) o = o | (lead[k][l] < sLo[l]) | (lead[k][l] > sHi[l]);
                        out_of[k] = __ballot_sync(0xffffffffu, o);   // bits 4 g .. 4 g + 3: walker g of tile k
                    }
                    uint32_t mine = out_of[0];
#pragma unroll
                    for (int k = 1; k < WT; ++k) mine = (tig & (WT - 1)) == k ? out_of[k] : mine;
                    outside = ((mine >> (4 * g)) & 0xfu) != 0u;
                }
                };
                if (NT == 8) {                             // one straight-line block: the scheduler interleaves the three parts
                    chi_tiles<L, WT, 4>(sC, 4, lane, A, lead, chi0, chi1);
                    stage_and_test();
                    chi_tiles<L, WT, 4>(sC + 4 * TS, 4, lane, A, lead, chi0, chi1);
                } else {
                    stage_and_test();
                    chi_tiles<L, WT, 0>(sC, NT, lane, A, lead, chi0, chi1);
                }
                // quad sums by a transposing butterfly: lane tig ends up with the sum of tile tig % WT
                double sum;
                if (WT == 4) {
                    double sv[4];
#pragma unroll
                    for (int k = 0; k < 4; ++k) sv[k] = chi0[k & (WT - 1)] + chi1[k & (WT - 1)];
                    const bool b0 = (tig & 1) != 0, b1 = (tig & 2) != 0;
                    const double a01 = (b0 ? sv[1] : sv[0]) + __shfl_xor_sync(0xffffffffu, b0 ? sv[0] : sv[1], 1);   // tile tig & 1
                    const double a23 = (b0 ? sv[3] : sv[2]) + __shfl_xor_sync(0xffffffffu, b0 ? sv[2] : sv[3], 1);   // tile 2 + (tig & 1)
                    sum = (b1 ? a23 : a01) + __shfl_xor_sync(0xffffffffu, b1 ? a01 : a23, 2);
                } else if (WT == 2) {
                    const double s0 = chi0[0] + chi1[0], s1 = chi0[WT - 1] + chi1[WT - 1];
                    const bool b0 = (tig & 1) != 0;
                    sum = (b0 ? s1 : s0) + __shfl_xor_sync(0xffffffffu, b0 ? s0 : s1, 1);
                    sum += __shfl_xor_sync(0xffffffffu, sum, 2);
                } else {
                    sum = chi0[0] + chi1[0];
                    sum += __shfl_xor_sync(0xffffffffu, sum, 1);
                    sum += __shfl_xor_sync(0xffffffffu, sum, 2);
                }
                if (outside) sum = CUDART_INF;
                const int wk = walker_of(tig);
                if (tig < WT && wk < H) sChi[wk] = sum;
                named_arrive(kDoneBarrier + grp, n_done);
                MDA_STAMP(3)
                // ---- while the walker warps decide: what the next half-step's proposals can fetch already
                if (t + 1 < t_end) {
                    named_sync(kDrawnBarrier, T);
                    prefetch(t + 1);
                }
                __syncthreads();                           // the half is complete
            }
        }
#undef MDA_STAMP
#undef MDA_STAMP_AFTER

        // ---- hand the bin back: state to HBM/L2, then the progress word
        __syncthreads();
        for (int i = tid; i < W * L; i += T) gPos[i] = sPos[i];
        for (int w = tid; w < W; w += T) {
            gLnp[w] = sLnp[w];
            gAcc[w] = sAcc[w];
        }
        __syncthreads();
        if (tid == 0 && args.n_chunks > 1) {
            __threadfence();
            st_release_u64(args.progress + bin, args.launch_serial + (unsigned long long)(chunk + 1));
        }
    }
    if (tid == 0) bulk_wait_all();
    if (saw_nan) atomicOr(args.status, kStatusNaN);
}

using EnsKernel = void (*)(const EnsembleArgs);

template <int L>
EnsKernel pick_ws_wt(int wt) {
    switch (wt) {
        case 1: return ensemble_ws_kernel<L, 1, false>;
        case 2: return ensemble_ws_kernel<L, 2, false>;
        case 4: return ensemble_ws_kernel<L, 4, false>;
        default: return nullptr;
    }
}

EnsKernel pick_ws_kernel(int n_l, int wt) {
    switch (n_l) {
        case 4: return pick_ws_wt<4>(wt);
        case 5: return pick_ws_wt<5>(wt);
        case 6: return pick_ws_wt<6>(wt);
        case 7: return pick_ws_wt<7>(wt);
        case 8: return pick_ws_wt<8>(wt);
        case 9: return pick_ws_wt<9>(wt);
        case 10: return pick_ws_wt<10>(wt);
        case 11: return pick_ws_wt<11>(wt);
        case 12: return pick_ws_wt<12>(wt);
        case 13: return pick_ws_wt<13>(wt);
        case 14: return pick_ws_wt<14>(wt);
        case 15: return pick_ws_wt<15>(wt);
        case 16: return pick_ws_wt<16>(wt);
        default: return nullptr;
    }
}

// The build with the timeline stamps (mdaEnsembleProfile): the C4 shape only.
constexpr int kProfL = 9, kProfWT = 4;
EnsKernel prof_kernel() { return ensemble_ws_kernel<kProfL, kProfWT, true>; }

}  // namespace

bool choose_ensemble_ws_shape(int bins, int walkers, int n_l, int n_tiles, int sm_count, size_t smem_optin, int force_wt,
                               EnsembleShape *out) {
    (void)bins;
    (void)sm_count;
    if (n_l < kMinMmaL || n_l > kMaxRegL) return false;
    EnsembleShape s{};
    const int half = walkers / 2;
    const int n_groups = (half + GW - 1) / GW;
    if (n_groups > kWalkerWarps) return false;                               // one group per walker warp
    s.smem_bytes = EnsembleWsSmem(walkers, n_l, n_tiles).total;
    // tiles per math item: the fewest that give every item its own math warp
    int wt = 8 * kMathWarps >= half ? 1 : (16 * kMathWarps >= half ? 2 : 4);
    if (force_wt == 1 || force_wt == 2 || force_wt == 4) wt = force_wt;
    const int n_items = (half + 8 * wt - 1) / (8 * wt);
    if (n_items > kMathWarps) return false;                                  // one item per math warp
    s.wpt = wt;
    s.groups = n_groups;                                                     // walker warps
    s.stage = 1;
    s.threads = 32 * (n_groups + n_items);
    s.global_state = false;
    s.mma = true;
    s.ws = true;
    s.fits = s.smem_bytes + kStaticSmem <= smem_optin && (size_t)n_tiles * 8 * (size_t)(n_l + 1) * 8 < (1u << 20);
    if (!s.fits) return false;
    *out = s;
    return true;
}

cudaError_t configure_ensemble_ws_kernels(size_t smem_optin) {
    auto raise = [&](EnsKernel fn) -> cudaError_t {
        const void *k = reinterpret_cast<const void *>(fn);
        cudaFuncAttributes fa;
        cudaError_t e = cudaFuncGetAttributes(&fa, k);
        if (e != cudaSuccess) return e;
        return cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(smem_optin - fa.sharedSizeBytes));
    };
    for (int l = kMinMmaL; l <= kMaxRegL; ++l)
        for (int wt = 1; wt <= 4; wt *= 2) {
            cudaError_t e = raise(pick_ws_kernel(l, wt));
            if (e != cudaSuccess) return e;
        }
    return raise(prof_kernel());
}

// Co-resident CTAs of this kernel on the whole device (the unit list needs every CTA of the grid
// resident: a unit may wait for one that runs on another CTA).
int ensemble_ws_capacity(int n_l, const EnsembleShape &shape, int sm_count) {
    EnsKernel k = pick_ws_kernel(n_l, shape.wpt);
    int per_sm = 0;
    if (!k || cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, reinterpret_cast<const void *>(k), shape.threads,
                                                            shape.smem_bytes) != cudaSuccess)
        return 0;
    return per_sm * sm_count;
}

cudaError_t launch_ensemble_ws(const EnsembleArgs &args, const EnsembleShape &shape, int grid, cudaStream_t stream) {
    if (args.bins <= 0 || args.n_steps <= 0) return cudaSuccess;
    EnsKernel k = pick_ws_kernel(args.n_l, shape.wpt);
    if (args.prof && args.n_l == kProfL && shape.wpt == kProfWT) k = prof_kernel();
    if (!k || !shape.fits || !shape.mma || !args.consts_mma || grid < 1 || args.chunk_steps < 1 || args.n_chunks < 1 ||
        (args.n_chunks > 1 && !args.progress))
        return cudaErrorInvalidValue;
    EnsembleArgs a = args;
    a.groups = shape.groups;                               // walker warps (the rest of the block are math warps)
    return launch_maybe_cooperative(k, a, a.n_chunks > 1, grid, shape.threads, shape.smem_bytes, stream);
}

}  // namespace mda
