// ensemble_mma_kernel.cu -- device-resident ensemble stepping with the chi-square contraction on
// the fp64 tensor path (DMMA.8x8x4) and warp-specialised bookkeeping, SURVEY.md 8f-1.
//
// Same job and same random streams as ensemble_kernel.cu (the reference's host loop
// emcee.EnsembleSampler.run_mcmc -> ln_post_prob per walker, mda.py:1128-1131, 1540-1578, as ONE
// persistent launch).  A CTA has two kinds of warps:
//
//   * walker warps (lane = walker of a group of 32): Philox draw, proposal q = c - z (c - x), box
//     prior, accept / reject, in-place update and the bulk store of kept steps to the HBM chain.
//     Everything a walker needs between its proposal and its accept test stays in its lane.
//   * math warps: chi^2 of items of 16 walkers.  Lane (g, tig) = (lane / 4, lane % 4) loads, for
//     walker g of a tile of 8, the proposal elements l = R + 4 ks + tig (ks < Q = L / 4) -- the A
//     fragment of mma.m8n8k4 -- and the R = L % 4 leading ones; residuals of an 8-walker x 8-angle
//     tile are c = d + sum_{l<R} q_l (-D_l) by DFMA, then Q DMMAs against B fragments of the NEGATED
//     distributions staged fragment-major in shared memory; chi^2 is accumulated from the C fragment
//     and reduced over the quad.  Every term of the reference's sum (C/ChiSquare.c:188-210) is
//     evaluated exactly once, only the order of the fp64 additions differs (|error| ~ 1e-16
//     relative, tested against the oracle).  Why DMMA: a DFMA with three distinct register operands
//     is register-bandwidth bound at ~72 % of the fp64 pipe on sm_100a; DMMA.8x8x4 runs on the same
//     pipe at its full rate with 1/8 of the issue slots (tools/ubench/dmma_ubench.cu).
//   * the two kinds meet through shared memory and one mbarrier pair per group of 32 walkers
//     (proposals ready -> chi^2 ready), so the fp64 pipe works on the first groups while the others
//     are still being proposed, and the draw of the next half-step is made while the pipe is busy.
//     ONE CTA-wide barrier per half-step (the stretch move needs the other half complete).
//   * CTAs walk a list of (step chunk, bin) units in chunk-major order, unit u on CTA u mod grid; a
//     unit waits for the previous chunk of its bin through a release/acquire word per bin and hands
//     the ensemble state over through HBM/L2.  With more bins than SMs (C4: 200 on 148) every SM
//     carries the same share of the work instead of 52 SMs carrying two bins for the whole run.
#include "likelihood.cuh"
#include "philox.cuh"

namespace mda {

namespace {

constexpr int kStatusNaN = 1;
constexpr int kWalkerWarps = 8;        // at most: one group of 32 walkers each at C4
constexpr int kMathWarps = 8;          // at most: two per scheduler keep the fp64 pipe fed
constexpr int kMaxGroups = 64;         // groups of 32 walkers per half (mbarrier pairs)
constexpr int GW = 32;                 // walkers of a group (a walker warp's lanes)
constexpr size_t kStaticSmem = 8 * (2 + 2 * kMaxGroups + kWalkerWarps) + 16;   // the mbarriers (static __shared__)

struct EnsembleMmaSmem {
    size_t consts, pos, q, lnp, chi, z, fast, partner, acc, total;
    __host__ __device__ EnsembleMmaSmem(int walkers, int n_l, int n_tiles) {
        const size_t h = (size_t)walkers / 2;
        size_t off = 0;
        consts = off;  off += (size_t)n_tiles * 8 * (size_t)(n_l + 1) * 8;     // [tiles][8 (1+L)] fragment-major
        pos = off;     off += (size_t)walkers * (size_t)n_l * 8;               // [W][L] dense, updated in place; bulk-stored as is
        lnp = off;     off += (size_t)walkers * 8;                             // [W] (16-byte aligned: W is even)
        q = off;       off += h * (size_t)n_l * 8;                             // [H][L] proposals of the half-step
        chi = off;     off += h * 8;                                           // [H] chi^2 of the proposals
        z = off;       off += h * 8;                                           // [H] stretch factor of the next proposal
        fast = off;    off += 2 * h * 4;                                       // [2][H] float (L-1) ln z - ln u'
        partner = off; off += h * 4;                                           // [H]
        acc = off;     off += (size_t)walkers * 4;                             // [W] accepted proposals
        total = (off + 15) & ~(size_t)15;
    }
};

__device__ __forceinline__ void dmma884(double &c0, double &c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void bulk_s2g(void *dst, const void *src, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(smem_u32(src)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read_all() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// Wait on an mbarrier phase that is expected to take hundreds of cycles.  A DMMA holds the issue port of
// its scheduler for its whole duration (tools/ubench/dmma_ubench.cu), so every probe of a waiting warp is
// taken from the math warps: try_wait with a long suspend hint parks the warp in hardware.
__device__ __forceinline__ void mbar_wait_relaxed(uint64_t *bar, uint32_t parity) {
    const uint32_t a = smem_u32(bar);
    uint32_t ok;
    do {
        asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3; selp.u32 %0, 1, 0, p; }"
                     : "=r"(ok) : "r"(a), "r"(parity), "r"(100000u) : "memory");
    } while (!ok);
}
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

// chi^2 of WT tiles of 8 walkers (A fragments a[][], leading elements lead[][]) against NTC angle tiles
// (NTC = 0: nt of them, a run-time count).  Constants are fetched one tile ahead of their use.
template <int L, int WT, int NTC>
__device__ __forceinline__ void chi_tiles(const double *sC, int nt, int lane, const double (&a)[WT][L / 4],
                                          const double (&lead)[WT][(L % 4) > 0 ? (L % 4) : 1], double (&chi0)[WT], double (&chi1)[WT]) {
    constexpr int Q = L / 4, R = L % 4, RR = R > 0 ? R : 1, TS = 8 * (L + 1);
    const int tig = lane & 3;
    const int n = NTC > 0 ? NTC : nt;
    const double *tp = sC;
    double2 dv = *reinterpret_cast<const double2 *>(tp + 2 * tig);
    double2 Di[RR];
#pragma unroll
    for (int l = 0; l < R; ++l) Di[l] = *reinterpret_cast<const double2 *>(tp + 8 * (1 + l) + 2 * tig);
    double Bf[Q];
#pragma unroll
    for (int ks = 0; ks < Q; ++ks) Bf[ks] = tp[8 * (1 + R) + 32 * ks + lane];
#pragma unroll(NTC > 0 ? NTC : 2)
    for (int at = 0; at < n; ++at) {
        tp = sC + min(at + 1, n - 1) * TS;
        const double2 dv_n = *reinterpret_cast<const double2 *>(tp + 2 * tig);
        double2 Di_n[RR];
#pragma unroll
        for (int l = 0; l < R; ++l) Di_n[l] = *reinterpret_cast<const double2 *>(tp + 8 * (1 + l) + 2 * tig);
        double Bf_n[Q];
#pragma unroll
        for (int ks = 0; ks < Q; ++ks) Bf_n[ks] = tp[8 * (1 + R) + 32 * ks + lane];
        double c0[WT], c1[WT];
#pragma unroll
        for (int k = 0; k < WT; ++k) {
            c0[k] = dv.x;
            c1[k] = dv.y;
#pragma unroll
            for (int l = 0; l < R; ++l) {
                c0[k] = fma(lead[k][l], Di[l].x, c0[k]);
                c1[k] = fma(lead[k][l], Di[l].y, c1[k]);
            }
        }
#pragma unroll
        for (int ks = 0; ks < Q; ++ks)
#pragma unroll
            for (int k = 0; k < WT; ++k) dmma884(c0[k], c1[k], a[k][ks], Bf[ks]);
#pragma unroll
        for (int k = 0; k < WT; ++k) {
            chi0[k] = fma(c0[k], c0[k], chi0[k]);
            chi1[k] = fma(c1[k], c1[k], chi1[k]);
        }
        dv = dv_n;
#pragma unroll
        for (int l = 0; l < R; ++l) Di[l] = Di_n[l];
#pragma unroll
        for (int ks = 0; ks < Q; ++ks) Bf[ks] = Bf_n[ks];
    }
}

// blockDim = 32 (n_walker_warps + n_math_warps); warps [0, n_walker_warps) are walker warps.
// A walker warp owns groups grp = warp, warp + n_walker_warps, ... of BOTH halves for a whole chunk of
// steps: only it writes the rows of these walkers (in place, on acceptance); every other warp reads
// them as partners, and only in the half-step in which they are not being updated.
// CPS = 1: up to 8 walker + 8 math warps, one CTA per SM.  CPS = 2: up to 8 + 4 warps and at most 80
// registers, so that two CTAs (two bins) share an SM and the serial head and tail of one bin's
// half-step (proposals; accept tests) run under the chi^2 loop of the other.
template <int L, int WT, int CPS>               // WT: tiles of 8 walkers per math item (4: a whole group)
__global__ void __launch_bounds__(32 * (kWalkerWarps + kMathWarps / CPS), CPS) ensemble_mma_kernel(const EnsembleArgs args) {
    constexpr int Q = L / 4, R = L % 4, RR = R > 0 ? R : 1;
    constexpr int IW = 8 * WT;                    // walkers of a math item
    constexpr int TS = 8 * (L + 1);               // doubles per angle tile of the constants

    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ __align__(8) uint64_t mbar;                 // constants of a unit have landed
    __shared__ __align__(8) uint64_t full_bar[kMaxGroups]; // proposals of a group are staged
    __shared__ __align__(8) uint64_t done_bar[kMaxGroups]; // chi^2 of a group is there
    __shared__ __align__(8) uint64_t store_bar;            // the bulk store of a kept step has read the state
    __shared__ __align__(8) uint64_t turn_bar[kWalkerWarps]; // walker warp w may queue its gathers

    const int tid = threadIdx.x;
    const int T = blockDim.x;
    const int lane = tid & 31, warp = tid >> 5;
    const int W = args.walkers;
    const int H = W >> 1;
    const int NT = args.n_tiles;
    const int n_groups = (H + GW - 1) / GW;
    const int n_items = (H + IW - 1) / IW;
    const int n_ww = args.groups & 0xff;          // walker warps (the launch sets it)
    const int stage_w = max(1, args.groups >> 8); // walker warps that gather at the same time
    const int n_mw = (T >> 5) - n_ww;
    const bool walker_warp = warp < n_ww;

    const EnsembleMmaSmem lay(W, L, NT);
    const double *sC = reinterpret_cast<const double *>(smem_raw + lay.consts);
    double *sPos = reinterpret_cast<double *>(smem_raw + lay.pos);
    double *sQ = reinterpret_cast<double *>(smem_raw + lay.q);
    double *sLnp = reinterpret_cast<double *>(smem_raw + lay.lnp);
    double *sChi = reinterpret_cast<double *>(smem_raw + lay.chi);
    double *sZ = reinterpret_cast<double *>(smem_raw + lay.z);
    float *sFast = reinterpret_cast<float *>(smem_raw + lay.fast);
    int *sPartner = reinterpret_cast<int *>(smem_raw + lay.partner);
    unsigned int *sAcc = reinterpret_cast<unsigned int *>(smem_raw + lay.acc);

    if (tid == 0) {
        mbar_init(&mbar, 1);
        mbar_init(&store_bar, 1);
        for (int w = 0; w < kWalkerWarps; ++w) mbar_init(&turn_bar[w], 1);
        for (int grp = 0; grp < n_groups; ++grp) {
            mbar_init(&full_bar[grp], 1);
            mbar_init(&done_bar[grp], (uint32_t)min(GW / IW, n_items - grp * (GW / IW)));   // math items of the group
        }
        fence_mbar_init();
    }
    __syncthreads();

    const uint32_t seed0 = args.seed_lo, seed1 = args.seed_hi;
    const double a_scale = args.a;
    const double dim_m1 = (double)(L - 1);
    const float dim_m1f = (float)(L - 1);

    bool saw_nan = false;
    uint32_t unit_phase = 0;                               // parity of mbar
    uint32_t hs_phase = 0;                                 // parity of full_bar / done_bar: one phase per half-step
    uint32_t stores_issued = 0, stores_seen = 0;           // bulk stores of kept steps (store_bar: one phase each)
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
                while (ld_acquire_u64(args.progress + bin) != want) __nanosleep(64);
            }
        }
        __syncthreads();
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

        // The Philox draw of walker i of the half updated in half-step t (counter based: it does not
        // depend on positions), in two parts so that half of the walker warps can put the integer
        // part before their proposals and the other half after (see below).
        auto draw_bits = [&](uint32_t t, int i) {
            const int s_base = (t & 1u) ? H : 0;
            return philox4x32_10(Philox4{(uint32_t)(s_base + i), bin_id, t, 0u}, seed0, seed1);
        };
        // emcee draws per half: u -> z, the partner index, then u'
        auto draw_finish = [&](const Philox4 &r, uint32_t t, int i) {
            const int c_base = (t & 1u) ? 0 : H;
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

        mbar_wait(&mbar, unit_phase);
        unit_phase ^= 1u;
        __syncthreads();
        if (walker_warp)
            for (int grp = warp; grp < n_groups; grp += n_ww) {
                const int i = min(grp * GW + lane, H - 1);
                draw_finish(draw_bits(t_begin, i), t_begin, i);
            }

        // debug timeline (mdaEnsembleProfile): clock stamps of one half-step, per warp
        const uint32_t t_prof = (args.prof != nullptr && chunk == 0) ? t_begin + min(81u, 2u * (uint32_t)steps_here - 1u) : 0xffffffffu;
        for (uint32_t t = t_begin; t < t_end; ++t, hs_phase ^= 1u) {
            const uint32_t half = t & 1u;
            long long *stamp = (t == t_prof && lane == 0 && warp < 16) ? args.prof + ((size_t)bin * 16 + warp) * 8 : nullptr;
#define MDA_STAMP(k) if (stamp) stamp[k] = clock64();
            MDA_STAMP(0)
            const int s_base = half ? H : 0;          // the half being updated; partners come from the other one

            if (walker_warp) {
                const bool draw_next = t + 1 < t_end;
                // ---- proposals of my groups: q = c - z (c - x); box prior
                unsigned long long out_bits = 0ull;        // bit k: my walker of my k-th group is outside the box
                int k = 0;
                for (int grp = warp; grp < n_groups; grp += n_ww, ++k) {
                    const int i = min(grp * GW + lane, H - 1);
                    const double z = sZ[i];
                    const double *xs = sPos + (s_base + i) * L;
                    const double *xc = sPos + sPartner[i] * L;
                    double *q = sQ + i * L;
                    // The gathers of the partner rows are shared-memory-bandwidth bound (random rows: ~3.5-way
                    // bank conflicts).  The groups take turns at it, in order, so that the first group is staged
                    // after one group's worth of traffic and the math warps start while the others still load.
                    if (stamp && z != 123.456) stamp[7] = clock64();                    // debug: my draw has been read
                    if (k == 0 && warp >= stage_w) mbar_wait_relaxed(&turn_bar[warp], hs_phase);
                    double c[L], v[L];
#pragma unroll
                    for (int l = 0; l < L; ++l) { c[l] = xc[l]; v[l] = xs[l]; }        // all loads before any store
                    if (k == 0 && warp + stage_w < n_ww) {
                        __syncwarp();
                        if (lane == 0) mbar_arrive(&turn_bar[warp + stage_w]);          // my loads are queued: next, please
                    }
                    // box prior (mda.py:1575): v < lo or v > hi for any element  <=>  the smallest of the margins
                    // v - lo, hi - v is negative (x - y < 0 <=> x < y exactly; fmin drops NaN margins, and a
                    // NaN compares false in the reference too).  A min tree instead of 2 L chained compares.
                    if (stamp && c[L - 1] != 123.456 && v[L - 1] != 123.456) stamp[5] = clock64();   // debug: rows are here
                    double m[L];
#pragma unroll
                    for (int l = 0; l < L; ++l) {
                        v[l] = __dsub_rn(c[l], __dmul_rn(z, __dsub_rn(c[l], v[l])));
                        m[l] = fmin(__dsub_rn(v[l], args.lo[l]), __dsub_rn(args.hi[l], v[l]));
                    }
#pragma unroll
                    for (int l = 0; l < L; ++l) q[l] = v[l];
#pragma unroll
                    for (int w = 1; w < L; w *= 2)
#pragma unroll
                        for (int l = 0; l + w < L; l += 2 * w) m[l] = fmin(m[l], m[l + w]);
                    const bool out = m[0] < 0.0;
                    if (stamp && (out || m[0] != 123.456)) stamp[6] = clock64();                    // debug: box prior known
                    out_bits |= (unsigned long long)out << k;
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&full_bar[grp]);
                }
                MDA_STAMP(1)
                // ---- the draw of the next half-step, while the math warps work
                if (draw_next)
                    for (int grp = warp; grp < n_groups; grp += n_ww) {
                        const int i = min(grp * GW + lane, H - 1);
                        draw_finish(draw_bits(t + 1, i), t + 1, i);
                    }
                MDA_STAMP(2)
                if (stores_seen != stores_issued) {        // the chain store must have read my rows
                    mbar_wait(&store_bar, (stores_issued - 1u) & 1u);
                    stores_seen = stores_issued;
                }
                // ---- accept / reject
                k = 0;
                for (int grp = warp; grp < n_groups; grp += n_ww, ++k) {
                    const int i = min(grp * GW + lane, H - 1);
                    const int gw = s_base + i;
                    const double oldlnp = sLnp[gw];
                    const float fast = sFast[half * H + i];
                    double qv[L];
#pragma unroll
                    for (int l = 0; l < L; ++l) qv[l] = sQ[i * L + l];                  // my own proposal, staged above
                    mbar_wait_relaxed(&done_bar[grp], hs_phase);
                    MDA_STAMP(3)
                    // outside the box the log-posterior is -inf (mda.py:1576): never accepted
                    const double newlnp = ((out_bits >> k) & 1ull) ? -CUDART_INF : sChi[i] * -0.5;    // C/ChiSquare.c:212
                    saw_nan = saw_nan || (newlnp != newlnp);
                    const double est = __dsub_rn(newlnp, oldlnp) + (double)fast;
                    bool accept;
                    if (fabs(est) > 1e-4) {
                        accept = est > 0.0;            // far from a tie: the float logs (error < 2e-5) cannot change the sign
                    } else {                           // emcee's test in full precision (rare: redo the draw)
                        const Philox4 r = philox4x32_10(Philox4{(uint32_t)gw, bin_id, t, 0u}, seed0, seed1);
                        const double zr = __dadd_rn(__dmul_rn(a_scale - 1.0, uniform53(r.x, r.y)), 1.0);
                        const double z2 = __dmul_rn(zr, zr);
                        const double z = (a_scale == 2.0) ? z2 * 0.5 : z2 / a_scale;
                        const double zlog = __dmul_rn(dim_m1, log(z));
                        accept = __dsub_rn(__dadd_rn(zlog, newlnp), oldlnp) > log(uniform43(r.w, r.x, r.y));
                    }
                    if (accept && grp * GW + lane < H) {
                        double *nx = sPos + gw * L;
#pragma unroll
                        for (int l = 0; l < L; ++l) nx[l] = qv[l];
                        sLnp[gw] = newlnp;
                        sAcc[gw] += 1u;
                    }
                }
                if (half && until_keep == 1 && keeping) fence_proxy_async_smem();   // a kept step: rows before the bulk store
                MDA_STAMP(4)
            } else {
                // ---- chi^2 of items of WT tiles of 8 walkers against every angle tile
                const int g = lane >> 2, tig = lane & 3;
                for (int item = warp - n_ww; item < n_items; item += n_mw) {
                    const int grp = item / (GW / IW);
                    const double *row[WT];
#pragma unroll
                    for (int k = 0; k < WT; ++k) row[k] = sQ + min(item * IW + 8 * k + g, H - 1) * L;
                    mbar_wait_relaxed(&full_bar[grp], hs_phase);
                    MDA_STAMP(1)
                    double A[WT][Q], lead[WT][RR];
#pragma unroll
                    for (int k = 0; k < WT; ++k) {
#pragma unroll
                        for (int l = 0; l < R; ++l) lead[k][l] = row[k][l];
#pragma unroll
                        for (int ks = 0; ks < Q; ++ks) A[k][ks] = row[k][R + 4 * ks + tig];
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
                    const int wk = item * IW + 8 * tig + g;
                    if (tig < WT && wk < H) sChi[wk] = sum;
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&done_bar[grp]);
                    MDA_STAMP(3)
                }
            }
            if (!walker_warp) { MDA_STAMP(5) }
            __syncthreads();
            if (!walker_warp) { MDA_STAMP(6) }
#undef MDA_STAMP
            if (half && --until_keep == 0) {           // a full step is complete and kept: one bulk store of the whole state
                until_keep = args.thin;
                if (keeping && slot < args.keep_cap) {
                    if (warp == n_ww && lane == 0) {       // a math warp: it would only wait for proposals now
                        bulk_s2g(args.chain + (((size_t)slot * args.bins + (size_t)bin) * W) * L, sPos, (uint32_t)W * L * 8u);
                        bulk_s2g(args.lnp_chain + ((size_t)slot * args.bins + (size_t)bin) * W, sLnp, (uint32_t)W * 8u);
                        bulk_commit();
                        bulk_wait_read_all();
                        mbar_arrive(&store_bar);
                    }
                    ++stores_issued;
                }
                ++slot;
            }
        }

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
    if (lane == 0) bulk_wait_all();
    if (saw_nan) atomicOr(args.status, kStatusNaN);
}

using EnsKernel = void (*)(const EnsembleArgs);

template <int L>
EnsKernel pick_mma_wt(int wt, int cps) {
    if (cps == 2) return wt == 2 ? ensemble_mma_kernel<L, 2, 2> : nullptr;
    switch (wt) {
        case 1: return ensemble_mma_kernel<L, 1, 1>;
        case 2: return ensemble_mma_kernel<L, 2, 1>;
        case 4: return ensemble_mma_kernel<L, 4, 1>;
        default: return nullptr;
    }
}

EnsKernel pick_mma_kernel(int n_l, int wt, int cps) {
    switch (n_l) {
        case 4: return pick_mma_wt<4>(wt, cps);
        case 5: return pick_mma_wt<5>(wt, cps);
        case 6: return pick_mma_wt<6>(wt, cps);
        case 7: return pick_mma_wt<7>(wt, cps);
        case 8: return pick_mma_wt<8>(wt, cps);
        case 9: return pick_mma_wt<9>(wt, cps);
        case 10: return pick_mma_wt<10>(wt, cps);
        case 11: return pick_mma_wt<11>(wt, cps);
        case 12: return pick_mma_wt<12>(wt, cps);
        case 13: return pick_mma_wt<13>(wt, cps);
        case 14: return pick_mma_wt<14>(wt, cps);
        case 15: return pick_mma_wt<15>(wt, cps);
        case 16: return pick_mma_wt<16>(wt, cps);
        default: return nullptr;
    }
}

}  // namespace

bool choose_ensemble_mma_shape(int bins, int walkers, int n_l, int n_tiles, int sm_count, size_t smem_optin, int force_wt,
                               int force_cps, EnsembleShape *out) {
    if (n_l < kMinMmaL || n_l > kMaxRegL) return false;
    EnsembleShape s{};
    const int half = walkers / 2;
    const int n_groups = (half + GW - 1) / GW;
    if (n_groups > kMaxGroups) return false;
    s.smem_bytes = EnsembleMmaSmem(walkers, n_l, n_tiles).total;
    // two bins per SM when every SM gets two and two ensembles fit its shared memory (measured, C4 shape:
    // +8 % at 296 bins; with fewer bins the SMs that carry two slower CTAs finish last)
    int cps = (bins >= 2 * sm_count && 2 * (s.smem_bytes + kStaticSmem + 1024) <= smem_optin + 1024) ? 2 : 1;
    if (force_cps == 1 || force_cps == 2) cps = force_cps;
    // tiles per math item: a whole group where that still gives every scheduler two math warps
    int wt = n_groups >= kMathWarps ? 4 : (2 * n_groups >= kMathWarps ? 2 : 1);
    if (n_l > 12 && wt == 4) wt = 2;                                         // registers: A fragments of 4 tiles x 4 k-steps
    if (force_wt == 1 || force_wt == 2 || force_wt == 4) wt = force_wt;
    if (cps == 2) wt = 2;
    const int n_items = (half + 8 * wt - 1) / (8 * wt);
    s.wpt = wt;
    s.stage = cps;
    s.groups = n_groups < kWalkerWarps ? n_groups : kWalkerWarps;            // walker warps
    const int max_math = kMathWarps / cps;
    const int math = n_items < max_math ? n_items : max_math;
    s.threads = 32 * (s.groups + math);
    s.global_state = false;
    s.mma = true;
    s.fits = s.smem_bytes + kStaticSmem <= smem_optin && (size_t)n_tiles * 8 * (size_t)(n_l + 1) * 8 < (1u << 20);
    if (!s.fits) return false;
    *out = s;
    return true;
}

cudaError_t configure_ensemble_mma_kernels(size_t smem_optin) {
    for (int l = kMinMmaL; l <= kMaxRegL; ++l)
        for (int cps = 1; cps <= 2; ++cps)
            for (int wt = 1; wt <= 4; wt *= 2) {
                EnsKernel fn = pick_mma_kernel(l, wt, cps);
                if (!fn) continue;
                const void *k = reinterpret_cast<const void *>(fn);
                cudaFuncAttributes fa;
                cudaError_t e = cudaFuncGetAttributes(&fa, k);
                if (e != cudaSuccess) return e;
                e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(smem_optin - fa.sharedSizeBytes));
                if (e != cudaSuccess) return e;
            }
    return cudaSuccess;
}

// Co-resident CTAs of this kernel on the whole device (the unit list needs every CTA of the grid
// resident: a unit may wait for one that runs on another CTA).
int ensemble_mma_capacity(int n_l, const EnsembleShape &shape, int sm_count) {
    EnsKernel k = pick_mma_kernel(n_l, shape.wpt, shape.stage);
    int per_sm = 0;
    if (!k || cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, reinterpret_cast<const void *>(k), shape.threads,
                                                            shape.smem_bytes) != cudaSuccess)
        return 0;
    return per_sm * sm_count;
}

cudaError_t launch_ensemble_mma(const EnsembleArgs &args, const EnsembleShape &shape, int grid, cudaStream_t stream) {
    if (args.bins <= 0 || args.n_steps <= 0) return cudaSuccess;
    EnsKernel k = pick_mma_kernel(args.n_l, shape.wpt, shape.stage);
    if (!k || !shape.fits || !shape.mma || !args.consts_mma || grid < 1 || args.chunk_steps < 1 || args.n_chunks < 1 ||
        (args.n_chunks > 1 && !args.progress))
        return cudaErrorInvalidValue;
    EnsembleArgs a = args;
    a.groups = shape.groups | (2 << 8);                    // walker warps (the rest of the block are math warps); two gather at a time
    k<<<grid, shape.threads, shape.smem_bytes, stream>>>(a);
    return cudaGetLastError();
}

}  // namespace mda
