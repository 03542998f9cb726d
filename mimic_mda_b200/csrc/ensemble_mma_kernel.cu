// ensemble_mma_kernel.cu -- device-resident ensemble stepping with the chi-square contraction on
// the fp64 tensor path (DMMA.8x8x4), SURVEY.md 8f-1.
//
// Same job and same random streams as ensemble_kernel.cu (the reference's host loop
// emcee.EnsembleSampler.run_mcmc -> ln_post_prob per walker, mda.py:1128-1131, 1540-1578, as ONE
// persistent launch).  A CTA owns an Ex bin (for a chunk of steps): constants, positions, log-
// probabilities and acceptance counters live in shared memory.  Every warp does the whole half-step
// for its own item of 8 WT walkers, and everything between "the other half is complete" and "my
// walkers are updated" stays in its registers:
//
//   * lane (g, tig) = (lane / 4, lane % 4) holds, for walker g of each of the WT tiles of 8, the
//     components l = R + 4 ks + tig (ks < Q = L / 4) -- the A fragment of mma.m8n8k4 -- and a share of the
//     R = L % 4 leading ones.  It gathers the partners' components c and forms the proposal
//     q = c - z (c - x) itself (numpy rounding: __dmul_rn / __dsub_rn, no FMA contraction), from operands
//     x, z and partner addresses it fetched before the barrier;
//   * box prior (mda.py:1574-1576) on those registers; residuals of an 8-walker x 8-angle tile are
//     c = d + sum_{l<R} q_l (-D_l) by DFMA, then Q DMMAs against B fragments of the NEGATED distributions
//     staged fragment-major in shared memory; chi^2 is accumulated from the C fragment and reduced over
//     the quad so that lane (g, tig = k) ends up with chi^2 of walker g of tile k: its decision lane.
//     Every term of the reference's sum (C/ChiSquare.c:188-210) is evaluated exactly once, only the
//     order of the fp64 additions differs (|error| ~ 1e-16 relative, tested against the oracle);
//   * the decision lane made the Philox draw of its walker (z, partner, u'), so the accept test needs
//     nothing from other warps; the accept bits are balloted and every lane writes the components it
//     holds of the accepted walkers straight into the state.  No proposal, chi^2 or draw ever goes
//     through shared memory, and the only synchronisation is ONE CTA barrier per half-step (the
//     stretch move needs the other half complete).
//   * a CTA is 8 warps with at most 128 registers: two CTAs -- two bins -- share an SM, and the gather,
//     decision and barrier latencies of one run under the DMMA stream of the other.
//   * CTAs walk a list of (step chunk, bin) units in chunk-major order, unit u on CTA u mod grid; a
//     unit waits for the previous chunk of its bin through a release/acquire word per bin and hands
//     the ensemble state over through HBM/L2, so that every SM carries the same share of the run.
//
// Why DMMA: a DFMA with three distinct register operands is register-bandwidth bound at ~72 % of the
// fp64 pipe on sm_100a; DMMA.8x8x4 runs on the same pipe at its full rate with 1/8 of the issue
// slots and lets integer work of other warps issue underneath (tools/ubench/dmma_ubench.cu,
// port_ubench.cu).  Earlier versions of this kernel (warp-specialised, mbarrier or named-barrier
// hand-offs through shared memory) are in the history; profiles/experiments/r01_stepper_variants.md.
//
// Shapes: one item per warp, at most 8 warps: ensembles of up to 512 walkers (all BASELINE.json
// shapes); larger ones run on ensemble_kernel.cu.
#include "mma_chi.cuh"
#include "philox.cuh"

namespace mda {

namespace {

constexpr int kStatusNaN = 1;
constexpr int kStatusStalled = 4;        // a unit waited 4 s (wall clock) for its predecessor: the launch gave up (never seen; a guard against a hung GPU)
constexpr int kMaxWarps = 8;           // items (warps) per CTA
constexpr int kDrawnBarrier = 1;       // named barrier: the draw of the next half-step is published (draw warps arrive, the others wait)
constexpr size_t kStaticSmem = 64;     // the mbarriers (static __shared__), rounded up

struct EnsembleMmaSmem {
    size_t consts, pos, lnp, bounds, z, acc, partner, fast, total;
    __host__ __device__ EnsembleMmaSmem(int walkers, int n_l, int n_tiles) {
        const size_t h = (size_t)walkers / 2;
        size_t off = 0;
        consts = off;  off += (size_t)n_tiles * 8 * (size_t)(n_l + 1) * 8;     // [tiles][8 (1+L)] fragment-major
        pos = off;     off += (size_t)walkers * (size_t)n_l * 8;               // [W][L] dense, updated in place; bulk-stored as is
        lnp = off;     off += (size_t)walkers * 8;                             // [W] (16-byte aligned: W is even)
        bounds = off;  off += 2 * (size_t)n_l * 8;                             // [2][L] box prior: lo, hi
        z = off;       off += h * 8;                                           // [H] draw warps: stretch factor of the next half-step
        acc = off;     off += (size_t)walkers * 4;                             // [W] accepted proposals
        partner = off; off += h * 4;                                           // [H] draw warps: L x partner row
        fast = off;    off += h * 4;                                           // [H] draw warps: float (L-1) ln z - ln u'
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

__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// blockDim = 32 x items of a half (8 WT walkers each); warp = item.  DW: as many DRAW WARPS again (warps items .. 2 items - 1),
// which make the Philox draws of the next half-step (lane = walker of item warp - items) and publish them in shared
// memory while the others propose and evaluate -- the schedule for one bin per SM, where nothing else would hide them.
template <int L, int WT, bool DW, bool PROF>      // WT: tiles of 8 walkers per item
__global__ void __launch_bounds__(32 * kMaxWarps * (DW ? 2 : 1), DW ? 1 : 2) ensemble_mma_kernel(const EnsembleArgs args) {
    constexpr int Q = L / 4, R = L % 4, RR = R > 0 ? R : 1;
    constexpr int IW = 8 * WT;                    // walkers of an item
    constexpr int TS = 8 * (L + 1);               // doubles per angle tile of the constants
    constexpr int NLE = (WT * R + 3) / 4;         // leading elements (l < R) a lane proposes; the quad shares them
    constexpr int NLS = NLE > 0 ? NLE : 1;
    // Which walker of the item is row g of tile k: 8 k + g, or WT g + k where that spreads a half-warp's accesses to
    // its own rows (operand fetch, update of accepted rows) over more banks (L = 9: rows 4 apart are 8 banks apart).
    constexpr bool KMAJOR = own_row_conflicts(L, WT) < own_row_conflicts(L, 1);

    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ __align__(8) uint64_t mbar;                 // constants of a unit have landed
    __shared__ __align__(8) uint64_t store_bar;            // the bulk store of a kept step has read the state

    const int tid = threadIdx.x;
    const int T = blockDim.x;
    const int lane = tid & 31, item = tid >> 5;
    const int n_items = DW ? (T >> 6) : (T >> 5);
    const bool draw_warp = DW && item >= n_items;
    const int g = lane >> 2, tig = lane & 3, quad = lane & ~3;
    const int W = args.walkers;
    const int H = W >> 1;
    const int NT = args.n_tiles;

    const EnsembleMmaSmem lay(W, L, NT);
    const double *sC = reinterpret_cast<const double *>(smem_raw + lay.consts);
    double *sPos = reinterpret_cast<double *>(smem_raw + lay.pos);
    double *sLnp = reinterpret_cast<double *>(smem_raw + lay.lnp);
    double *sLo = reinterpret_cast<double *>(smem_raw + lay.bounds);
    double *sHi = sLo + L;
    unsigned int *sAcc = reinterpret_cast<unsigned int *>(smem_raw + lay.acc);
    double *sZ = reinterpret_cast<double *>(smem_raw + lay.z);
    int *sPartner = reinterpret_cast<int *>(smem_raw + lay.partner);
    float *sFast = reinterpret_cast<float *>(smem_raw + lay.fast);

    __shared__ int s_abort;
    if (tid == 0) {
        s_abort = 0;
        mbar_init(&mbar, 1);
        mbar_init(&store_bar, 1);
        fence_mbar_init();
    }
    if (tid < L) { sLo[tid] = args.lo[tid]; sHi[tid] = args.hi[tid]; }
    __syncthreads();

    const uint32_t seed0 = args.seed_lo, seed1 = args.seed_hi;
    const double a_scale = args.a;
    const double dim_m1 = (double)(L - 1);
    const float dim_m1f = (float)(L - 1);

    // my walker of tile k (index within the half); rows beyond the half repeat the last walker and are never accepted
    auto walker_of = [&](int k) { return item * IW + (KMAJOR ? WT * g + k : 8 * k + g); };
    int wi[WT];
#pragma unroll
    for (int k = 0; k < WT; ++k) wi[k] = min(walker_of(k), H - 1);
    // the walker I decide for: walker g of tile tig
    const int dw = min(walker_of(tig & (WT - 1)), H - 1);
    const bool decides = tig < WT && walker_of(tig) < H;
    // leading elements (l < R) of the quad's WT walkers are shared out over its 4 lanes: slot n of lane tig is
    // element j = tig + 4 n of the list (tile j / R, component j % R)
    int lk[NLS], ll[NLS], lw[NLS];
#pragma unroll
    for (int n = 0; n < NLE; ++n) {
        const int j = min(tig + 4 * n, WT * RR - 1);
        lk[n] = j / RR;
        ll[n] = j % RR;
        lw[n] = min(walker_of(lk[n]), H - 1);
    }

    bool saw_nan = false;
    uint32_t unit_phase = 0;                               // parity of mbar
    uint32_t stores_issued = 0, stores_seen = 0;           // bulk stores of kept steps (store_bar: one phase each)
    bool owe_arrival = false;                              // the issuer: a bulk store is queued and store_bar not yet told
    // The thread that queues the chain stores and later waits until they have read the state: with draw warps one of
    // theirs (waiting costs it nothing), else thread 0 (the SM's other CTA covers the wait).
    const bool issuer = DW ? tid == 32 * n_items : tid == 0;
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
        // after the barrier that ends a half-step: a full step that is kept goes to the chain as one bulk store of the state
        auto end_of_half_step = [&](uint32_t half) {
            if (half && --until_keep == 0) {
                until_keep = args.thin;
                if (keeping && slot < args.keep_cap) {
                    if (issuer) {
                        bulk_s2g(args.chain + (((size_t)slot * args.bins + (size_t)bin) * W) * L, sPos, (uint32_t)W * L * 8u);
                        bulk_s2g(args.lnp_chain + ((size_t)slot * args.bins + (size_t)bin) * W, sLnp, (uint32_t)W * 8u);
                        bulk_commit();
                        owe_arrival = true;
                    }
                    ++stores_issued;
                }
                ++slot;
            }
        };
        auto settle_store = [&]() {                        // the queued chain store has read the state by now: tell the writers
            if (issuer && owe_arrival) {
                bulk_wait_read_all();
                mbar_arrive(&store_bar);
                owe_arrival = false;
            }
        };

        mbar_wait(&mbar, unit_phase);
        unit_phase ^= 1u;

        if (draw_warp) {
            // ---- draw warps: the Philox draw (counter based: it does not depend on positions) of every walker of the
            // half updated in half-step t + 1, published while the other warps work on half-step t
            const int wl = (item - n_items) * IW + lane;   // my walker (index within the half)
            const bool live = lane < IW && wl < H;
            auto publish = [&](uint32_t t) {
                const int s_base = (t & 1u) ? H : 0, c_base = (t & 1u) ? 0 : H;
                const Philox4 r = philox4x32_10_keys(Philox4{(uint32_t)(s_base + min(wl, H - 1)), bin_id, t, 0u}, args.round_keys);
                const double u = uniform53(r.x, r.y);
                const double zr = __dadd_rn(__dmul_rn(a_scale - 1.0, u), 1.0);          // numpy: one rounding per operation
                const double z2 = __dmul_rn(zr, zr);
                const double zd = (a_scale == 2.0) ? z2 * 0.5 : z2 / a_scale;           // ((a-1) u + 1)^2 / a
                const double ua = uniform43(r.w, r.x, r.y);
                if (live) {
                    sZ[wl] = zd;
                    sFast[wl] = dim_m1f * __logf((float)zd) - __logf((float)ua);
                    sPartner[wl] = (c_base + (int)__umulhi(r.z, (uint32_t)H)) * L;
                }
            };
            publish(t_begin);
            __syncthreads();                               // the state and the first draw are there
            __syncthreads();                               // ... and the item warps have fetched that draw: the next may overwrite it
            for (uint32_t t = t_begin; t < t_end; ++t) {
                if (t + 1 < t_end) {
                    publish(t + 1);
                    named_arrive(kDrawnBarrier, T);
                }
                settle_store();
                __syncthreads();                           // the half is complete
                end_of_half_step(t & 1u);
            }
        } else {
        // ---- what a half-step's proposals need and the half-step before does not change
        // The Philox draw of the walker I decide for, of the half updated in half-step t, is counter based: it does
        // not depend on positions.  emcee draws per half: u -> z, the partner index, then u'.
        double pz[WT], pxa[WT][Q], pzl[NLS], pxl[NLS];     // stretch factors, own components
        int pca[WT][Q], pcl[NLS];                          // index of the partner's components
        float fast = 0.f;                                  // decision lane: (L-1) ln z - ln u' in single precision
        // The Philox draw of the walker I decide for, of the half updated in half-step t (counter based: it
        // does not depend on positions).  emcee draws per half: u -> z, the partner index, then u'.
        double zd_next = 0.0;                              // my walker's draw of the coming half-step: stretch factor,
        int part_next = 0;                                 // L x its partner's row,
        float fast_next = 0.f;                             // (L-1) ln z - ln u'
        auto draw = [&](uint32_t t) {
            const int s_base = (t & 1u) ? H : 0, c_base = (t & 1u) ? 0 : H;
            const Philox4 r = philox4x32_10_keys(Philox4{(uint32_t)(s_base + dw), bin_id, t, 0u}, args.round_keys);
            const double u = uniform53(r.x, r.y);
            // numpy arithmetic is one rounding per operation: no FMA contraction here
            const double zr = __dadd_rn(__dmul_rn(a_scale - 1.0, u), 1.0);
            const double z2 = __dmul_rn(zr, zr);
            zd_next = (a_scale == 2.0) ? z2 * 0.5 : z2 / a_scale;                      // ((a-1) u + 1)^2 / a
            const double ua = uniform43(r.w, r.x, r.y);
            // single-precision estimate of (L-1) ln z - ln u'; decides all but near-ties
            fast_next = dim_m1f * __logf((float)zd_next) - __logf((float)ua);
            part_next = (c_base + (int)__umulhi(r.z, (uint32_t)H)) * L;
        };
        // hand the draws round the quad (lane (g, k) decided for walker g of tile k) and fetch the own components
        auto fetch = [&](uint32_t t) {
            const double *xhalf = sPos + ((t & 1u) ? H : 0) * L;
#pragma unroll
            for (int k = 0; k < WT; ++k) {
                pz[k] = DW ? sZ[wi[k]] : __shfl_sync(0xffffffffu, zd_next, quad | k);
                const int part = DW ? sPartner[wi[k]] : __shfl_sync(0xffffffffu, part_next, quad | k);
#pragma unroll
                for (int ks = 0; ks < Q; ++ks) {
                    pca[k][ks] = part + R + 4 * ks + tig;
                    pxa[k][ks] = xhalf[wi[k] * L + R + 4 * ks + tig];
                }
            }
#pragma unroll
            for (int n = 0; n < NLE; ++n) {
                pzl[n] = DW ? sZ[lw[n]] : __shfl_sync(0xffffffffu, zd_next, quad | lk[n]);
                pcl[n] = (DW ? sPartner[lw[n]] : __shfl_sync(0xffffffffu, part_next, quad | lk[n])) + ll[n];
                pxl[n] = xhalf[lw[n] * L + ll[n]];
            }
            fast = DW ? sFast[dw] : fast_next;
        };
        // Two warps share a scheduler (items i and i + 4): the first draws (integer work) before its chi^2 loop, the
        // second after, so that one's Philox rounds issue under the other's DMMA stream.
        const bool draws_early = item < 4;
        __syncthreads();                                   // the state (and with draw warps the first draw) is there
        if (!DW) draw(t_begin);
        fetch(t_begin);
        if (DW) __syncthreads();                           // only now may the draw warps publish the next draw

        // debug timeline (mdaEnsembleProfile): clock stamps of one half-step, per warp
        const uint32_t t_prof = (PROF && args.prof != nullptr && chunk == 0) ? t_begin + min(81u, 2u * (uint32_t)steps_here - 1u) : 0xffffffffu;
#define MDA_STAMP(k) if (PROF && stamp) stamp[k] = clock64();
#define MDA_STAMP_AFTER(k, value) if (PROF && stamp && (value) != 123.456) stamp[k] = clock64();   // once `value` has arrived

        for (uint32_t t = t_begin; t < t_end; ++t) {
            const uint32_t half = t & 1u;
            const int s_base = half ? H : 0;              // the half being updated; partners come from the other one
            long long *stamp = (PROF && t == t_prof && lane == 0) ? args.prof + ((size_t)bin * 16 + item) * 8 : nullptr;
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
                    const int j = k * R + l;               // compile-time: lane j % 4 of the quad holds it in slot j / 4
                    lead[k][l] = __shfl_sync(0xffffffffu, myl[j / 4], quad | (j & 3));
                }
            MDA_STAMP_AFTER(1, A[WT - 1][Q - 1])
            const int gw = s_base + dw;
            const double oldlnp = sLnp[gw];                // only my own decision below changes it
            if (!DW && draws_early && t + 1 < t_end) draw(t + 1);
            if (!DW) settle_store();
            // box prior (mda.py:1575) on these registers: any component below lo or above hi (a NaN compares
            // false, as in the reference).  Lane (g, k) will decide for walker g of tile k.
            bool outside;
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
            double chi0[WT], chi1[WT];
#pragma unroll
            for (int k = 0; k < WT; ++k) { chi0[k] = 0.0; chi1[k] = 0.0; }
            if (NT == 8) chi_tiles<L, WT, 8>(sC, 8, lane, A, lead, chi0, chi1);
            else chi_tiles<L, WT, 0>(sC, NT, lane, A, lead, chi0, chi1);
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
            MDA_STAMP_AFTER(2, sum)
            const float fast_now = fast;
            // ---- accept / reject the walker I decide for
            // outside the box the log-posterior is -inf (mda.py:1576): never accepted
            const double newlnp = outside ? -CUDART_INF : sum * -0.5;    // C/ChiSquare.c:212
            const double est = __dsub_rn(newlnp, oldlnp) + (double)fast_now;
            MDA_STAMP_AFTER(6, est)
            bool accept;
            if (fabs(est) > 1e-4 || !decides) {
                accept = est > 0.0;                    // far from a tie: the float logs (error < 2e-5) cannot change the sign
            } else {                                   // emcee's test in full precision (rare: redo the draw)
                const Philox4 r = philox4x32_10_keys(Philox4{(uint32_t)gw, bin_id, t, 0u}, args.round_keys);
                const double zr = __dadd_rn(__dmul_rn(a_scale - 1.0, uniform53(r.x, r.y)), 1.0);
                const double z2 = __dmul_rn(zr, zr);
                const double z = (a_scale == 2.0) ? z2 * 0.5 : z2 / a_scale;
                const double zlog = __dmul_rn(dim_m1, log(z));
                accept = __dsub_rn(__dadd_rn(zlog, newlnp), oldlnp) > log(uniform43(r.w, r.x, r.y));
            }
            accept = accept && decides;
            saw_nan = saw_nan || (decides && newlnp != newlnp);
            const uint32_t moved = __ballot_sync(0xffffffffu, accept);       // bit 4 g + k: walker g of tile k moves
            MDA_STAMP_AFTER(7, (double)moved)
            if (stores_seen != stores_issued) {            // the chain store must have read the rows before they change
                mbar_wait(&store_bar, (stores_issued - 1u) & 1u);
                stores_seen = stores_issued;
            }
            if (accept) {
                sLnp[gw] = newlnp;
                sAcc[gw] += 1u;
            }
            // every lane writes the components it holds of the walkers that move
            {
                double *nhalf = sPos + s_base * L;
#pragma unroll
                for (int k = 0; k < WT; ++k)
                    if ((moved >> (4 * g + k)) & 1u) {
#pragma unroll
                        for (int ks = 0; ks < Q; ++ks) nhalf[wi[k] * L + R + 4 * ks + tig] = A[k][ks];
                    }
#pragma unroll
                for (int n = 0; n < NLE; ++n)
                    if (tig + 4 * n < WT * R && ((moved >> (4 * g + lk[n])) & 1u)) nhalf[lw[n] * L + ll[n]] = myl[n];
            }
            MDA_STAMP(3)
            // ---- the draw of the next half-step and what its proposals can fetch already (their own rows do not change now).
            // After the decision, not before it: a warp issues in order, and the decision's few fp64 operations would
            // queue behind these loads (measured: 590 instead of ~100 cycles from chi^2 to the accept bits).
            if (t + 1 < t_end) {
                if (DW) named_sync(kDrawnBarrier, T);
                else if (!draws_early) draw(t + 1);
                fetch(t + 1);
            }
            MDA_STAMP_AFTER(4, pz[0])
            const bool kept_step = half && until_keep == 1 && keeping && slot < args.keep_cap;
            if (kept_step) fence_proxy_async_smem();       // my rows before the bulk store reads them
            __syncthreads();                               // the half is complete
            end_of_half_step(half);
        }
        }   // not a draw warp
#undef MDA_STAMP
#undef MDA_STAMP_AFTER

        // ---- hand the bin back: state to HBM/L2, then the progress word
        settle_store();                                    // the last step of the unit was kept: its store still reads the state
        if (stores_seen != stores_issued) {
            mbar_wait(&store_bar, (stores_issued - 1u) & 1u);
            stores_seen = stores_issued;
        }
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
    if (issuer) bulk_wait_all();
    if (saw_nan) atomicOr(args.status, kStatusNaN);
}

using EnsKernel = void (*)(const EnsembleArgs);

template <int L>
EnsKernel pick_mma_wt(int wt, bool dw) {
    switch (wt) {
        case 1: return dw ? ensemble_mma_kernel<L, 1, true, false> : ensemble_mma_kernel<L, 1, false, false>;
        case 2: return dw ? ensemble_mma_kernel<L, 2, true, false> : ensemble_mma_kernel<L, 2, false, false>;
        case 4: return dw ? ensemble_mma_kernel<L, 4, true, false> : ensemble_mma_kernel<L, 4, false, false>;
        default: return nullptr;
    }
}

EnsKernel pick_mma_kernel(int n_l, int wt, bool dw) {
    switch (n_l) {
        case 4: return pick_mma_wt<4>(wt, dw);
        case 5: return pick_mma_wt<5>(wt, dw);
        case 6: return pick_mma_wt<6>(wt, dw);
        case 7: return pick_mma_wt<7>(wt, dw);
        case 8: return pick_mma_wt<8>(wt, dw);
        case 9: return pick_mma_wt<9>(wt, dw);
        case 10: return pick_mma_wt<10>(wt, dw);
        case 11: return pick_mma_wt<11>(wt, dw);
        case 12: return pick_mma_wt<12>(wt, dw);
        case 13: return pick_mma_wt<13>(wt, dw);
        case 14: return pick_mma_wt<14>(wt, dw);
        case 15: return pick_mma_wt<15>(wt, dw);
        case 16: return pick_mma_wt<16>(wt, dw);
        default: return nullptr;
    }
}

// The build with the timeline stamps (mdaEnsembleProfile): the C4 shape only.
constexpr int kProfL = 9, kProfWT = 4;
EnsKernel prof_kernel(bool dw) { return dw ? ensemble_mma_kernel<kProfL, kProfWT, true, true> : ensemble_mma_kernel<kProfL, kProfWT, false, true>; }

}  // namespace

bool choose_ensemble_mma_shape(int bins, int walkers, int n_l, int n_tiles, int sm_count, size_t smem_optin, int force_wt,
                               int force_draw_warps, EnsembleShape *out) {
    if (n_l < kMinMmaL || n_l > kMaxRegL) return false;
    EnsembleShape s{};
    const int half = walkers / 2;
    s.smem_bytes = EnsembleMmaSmem(walkers, n_l, n_tiles).total;
    // tiles per item: the fewest that give every item its own warp
    int wt = 8 * kMaxWarps >= half ? 1 : (16 * kMaxWarps >= half ? 2 : 4);
    if (force_wt == 1 || force_wt == 2 || force_wt == 4) wt = force_wt;
    const int n_items = (half + 8 * wt - 1) / (8 * wt);
    if (n_items > kMaxWarps) return false;                                   // one item per warp
    // draw warps where an SM holds one bin at a time; from two bins per SM on, two 8-warp CTAs share an SM
    bool dw = bins < 2 * sm_count;
    if (force_draw_warps == 1) dw = true;
    if (force_draw_warps == 2) dw = false;
    s.wpt = wt;
    s.groups = n_items;
    s.stage = dw ? 1 : 0;
    s.threads = 32 * n_items * (dw ? 2 : 1);
    s.global_state = false;
    s.mma = true;
    s.fits = s.smem_bytes + kStaticSmem <= smem_optin && (size_t)n_tiles * 8 * (size_t)(n_l + 1) * 8 < (1u << 20);
    if (!s.fits) return false;
    *out = s;
    return true;
}

cudaError_t configure_ensemble_mma_kernels(size_t smem_optin) {
    auto raise = [&](EnsKernel fn) -> cudaError_t {
        const void *k = reinterpret_cast<const void *>(fn);
        cudaFuncAttributes fa;
        cudaError_t e = cudaFuncGetAttributes(&fa, k);
        if (e != cudaSuccess) return e;
        return cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(smem_optin - fa.sharedSizeBytes));
    };
    for (int l = kMinMmaL; l <= kMaxRegL; ++l)
        for (int dw = 0; dw <= 1; ++dw)
            for (int wt = 1; wt <= 4; wt *= 2) {
                cudaError_t e = raise(pick_mma_kernel(l, wt, dw != 0));
                if (e != cudaSuccess) return e;
            }
    cudaError_t e = raise(prof_kernel(false));
    return e != cudaSuccess ? e : raise(prof_kernel(true));
}

// Co-resident CTAs of this kernel on the whole device (the unit list needs every CTA of the grid
// resident: a unit may wait for one that runs on another CTA).
int ensemble_mma_capacity(int n_l, const EnsembleShape &shape, int sm_count) {
    EnsKernel k = pick_mma_kernel(n_l, shape.wpt, shape.stage == 1);
    int per_sm = 0;
    if (!k || cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, reinterpret_cast<const void *>(k), shape.threads,
                                                            shape.smem_bytes) != cudaSuccess)
        return 0;
    return per_sm * sm_count;
}

cudaError_t launch_ensemble_mma(const EnsembleArgs &args, const EnsembleShape &shape, int grid, cudaStream_t stream) {
    if (args.bins <= 0 || args.n_steps <= 0) return cudaSuccess;
    EnsKernel k = pick_mma_kernel(args.n_l, shape.wpt, shape.stage == 1);
    if (args.prof && args.n_l == kProfL && shape.wpt == kProfWT) k = prof_kernel(shape.stage == 1);
    if (!k || !shape.fits || !shape.mma || !args.consts_mma || grid < 1 || args.chunk_steps < 1 || args.n_chunks < 1 ||
        (args.n_chunks > 1 && !args.progress))
        return cudaErrorInvalidValue;
    EnsembleArgs a = args;
    return launch_maybe_cooperative(k, a, a.n_chunks > 1, grid, shape.threads, shape.smem_bytes, stream);
}

}  // namespace mda
