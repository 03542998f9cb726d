// ensemble_mma_kernel.cu -- device-resident ensemble stepping with the chi-square contraction on
// the fp64 tensor path (DMMA.8x8x4), SURVEY.md 8f-1.
//
// Same job and same random streams as ensemble_kernel.cu (the reference's host loop
// emcee.EnsembleSampler.run_mcmc -> ln_post_prob per walker, mda.py:1128-1131, 1540-1578, as ONE
// persistent launch).  How a half-step is laid out here:
//
//   * A warp owns work items of WT tiles of 8 walkers.  Two register layouts are used:
//       fragment layout      lane (g, tig) = (lane / 4, lane % 4) holds, for walker g of tile k, the
//                            proposal elements l = R + 4 ks + tig (ks < Q = L / 4) -- the A fragment of
//                            mma.m8n8k4 -- and, replicated over the quad, the R = L % 4 leading ones;
//       walker layout        lane j < 8 WT stands for walker j of the item (Philox draw, accept test).
//   * residuals of an 8-walker x 8-angle tile: c = d + sum_{l<R} q_l (-D_l) by DFMA, then Q DMMAs
//     against B fragments of the NEGATED distributions staged fragment-major in shared memory;
//     chi^2 is accumulated from the C fragment and reduced over the quad by a transposing butterfly.
//     Every term of the reference's sum (C/ChiSquare.c:188-210) is evaluated exactly once, only the
//     order of the fp64 additions differs (|error| ~ 1e-16 relative, tested against the oracle).
//     Why: a DFMA with three distinct register operands is register-bandwidth bound at ~72 % of
//     the fp64 pipe on sm_100a; DMMA.8x8x4 runs on the same pipe at its full rate with 1/8 of the
//     issue slots (tools/ubench/dmma_ubench.cu).
//   * the Philox rounds of the NEXT item's draw are issued between the DMMAs of the current item
//     (integer work in the issue slots the fp64 pipe leaves free); the warp that evaluated an item
//     also proposes and accepts it: no staging of proposals or partial sums, ONE barrier per half-step.
//   * a warp owns the same walkers for a whole chunk of steps: it alone writes their rows (in place,
//     on acceptance) and bulk-stores them (cp.async.bulk shared -> global) to the HBM chain.
//   * CTAs walk a list of (step chunk, bin) units in chunk-major order, unit u on CTA u mod grid; a
//     unit waits for the previous chunk of its bin through a release/acquire word per bin and hands
//     the ensemble state over through HBM/L2.  With more bins than SMs (C4: 200 on 148) every SM
//     carries the same share of the work instead of 52 SMs carrying two bins for the whole run.
#include "likelihood.cuh"
#include "philox.cuh"

namespace mda {

namespace {

constexpr int kStatusNaN = 1;
constexpr int kMaxWarps = 8;

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

// Philox4x32-10 (philox.cuh) as a state that is advanced a round at a time.
struct PhiloxState { uint32_t x, y, z, w, k0, k1; };
__device__ __forceinline__ void philox_round(PhiloxState &s) {
    constexpr uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
    const uint32_t hi0 = __umulhi(M0, s.x), lo0 = M0 * s.x;
    const uint32_t hi1 = __umulhi(M1, s.z), lo1 = M1 * s.z;
    s.x = hi1 ^ s.y ^ s.k0;
    s.y = lo1;
    s.z = hi0 ^ s.w ^ s.k1;
    s.w = lo0;
    s.k0 += W0;
    s.k1 += W1;
}

// what the stretch move of one walker needs (walker layout)
struct Draw { double z; float fast; int partner; };

// Work of a warp: items item = warp, warp + n_warps, ... of BOTH halves, for a whole chunk of steps.
// Only their owner writes the rows of these walkers (in place, on acceptance); the other warps read
// them as partners, and only in the half-step in which they are not being updated.
template <int L, int WT>
__global__ void __launch_bounds__(32 * kMaxWarps, 2) ensemble_mma_kernel(const EnsembleArgs args) {
    constexpr int Q = L / 4, R = L % 4, RR = R > 0 ? R : 1;
    constexpr int TS = 8 * (L + 1);               // doubles per angle tile of the constants
    constexpr int IW = 8 * WT;                    // walkers of a work item

    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ __align__(8) uint64_t mbar;

    const int tid = threadIdx.x;
    const int T = blockDim.x;
    const int lane = tid & 31, warp = tid >> 5, n_warps = T >> 5;
    const int g = lane >> 2, tig = lane & 3;
    const int W = args.walkers;
    const int H = W >> 1;
    const int NT = args.n_tiles;
    const int n_items = (H + IW - 1) / IW;

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

    const uint32_t seed0 = args.seed_lo, seed1 = args.seed_hi;
    const double a_scale = args.a;
    const double dim_m1 = (double)(L - 1);
    const float dim_m1f = (float)(L - 1);

    // box prior of the elements this lane holds (mda.py:1124, 1574-1576)
    double lo_r[RR], hi_r[RR], lo_q[Q], hi_q[Q];
#pragma unroll
    for (int l = 0; l < R; ++l) { lo_r[l] = args.lo[l]; hi_r[l] = args.hi[l]; }
#pragma unroll
    for (int ks = 0; ks < Q; ++ks) { lo_q[ks] = args.lo[R + 4 * ks + tig]; hi_q[ks] = args.hi[R + 4 * ks + tig]; }
    const int elem0 = R + tig;                            // first DMMA element of this lane within a row
    // walker layout: lane j < IW is walker j of the item; its chi^2 arrives from lane 4 (j % 8) + j / 8
    const int wl_src = 4 * (lane & 7) + ((lane >> 3) & (WT - 1));
    const bool bulk_ok = (H & 1) == 0;                    // 16-byte alignment of both halves' rows

    bool saw_nan = false;
    uint32_t mbar_phase = 0;
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

        auto philox_start = [&](uint32_t t, int item) {
            const int s_base = (t & 1u) ? H : 0;
            const int i = min(item * IW + (lane & (IW - 1)), H - 1);
            return PhiloxState{(uint32_t)(s_base + i), bin_id, t, 0u, seed0, seed1};
        };
        // emcee draws per half: u -> z, the partner index, then u'
        auto philox_finish = [&](const PhiloxState &r, uint32_t t) {
            const int c_base = (t & 1u) ? 0 : H;
            const double u = uniform53(r.x, r.y);
            // numpy arithmetic is one rounding per operation: no FMA contraction here
            const double zr = __dadd_rn(__dmul_rn(a_scale - 1.0, u), 1.0);
            const double z2 = __dmul_rn(zr, zr);
            Draw d;
            d.z = (a_scale == 2.0) ? z2 * 0.5 : z2 / a_scale;                  // ((a-1) u + 1)^2 / a
            const double ua = uniform43(r.w, r.x, r.y);
            // single-precision estimate of (L-1) ln z - ln u'; decides all but near-ties
            d.fast = dim_m1f * __logf((float)d.z) - __logf((float)ua);
            d.partner = c_base + (int)__umulhi(r.z, (uint32_t)H);
            return d;
        };

        mbar_wait(&mbar, mbar_phase);
        mbar_phase ^= 1u;
        __syncthreads();

        Draw dr;
        {
            PhiloxState ps = philox_start(t_begin, min(warp, n_items - 1));
#pragma unroll
            for (int r = 0; r < 10; ++r) philox_round(ps);
            dr = philox_finish(ps, t_begin);
        }
        bool store_in_flight = false;

        for (uint32_t t = t_begin; t < t_end; ++t) {
            const uint32_t half = t & 1u;
            const int s_base = half ? H : 0;          // the half being updated; partners come from the other one

            for (int item = warp; item < n_items; item += n_warps) {
                const bool last_pass = item + n_warps >= n_items;
                const uint32_t t_nxt = last_pass ? t + 1 : t;
                const int item_nxt = last_pass ? warp : item + n_warps;
                const int w_first = item * IW;            // within the half

                // ---- propose: q = c - z (c - x), in A-fragment order
                double A[WT][Q], lead[WT][RR];
                unsigned int out_mask[WT];
#pragma unroll
                for (int k = 0; k < WT; ++k) {
                    const double z = __shfl_sync(0xffffffffu, dr.z, 8 * k + g);
                    const int partner = __shfl_sync(0xffffffffu, dr.partner, 8 * k + g);
                    const double *xs = sPos + (s_base + min(w_first + 8 * k + g, H - 1)) * L;
                    const double *xc = sPos + partner * L;
                    bool out = false;
#pragma unroll
                    for (int l = 0; l < R; ++l) {
                        const double c = xc[l];
                        const double v = __dsub_rn(c, __dmul_rn(z, __dsub_rn(c, xs[l])));
                        lead[k][l] = v;
                        out = out || (v < lo_r[l]) || (v > hi_r[l]);                    // mda.py:1575
                    }
#pragma unroll
                    for (int ks = 0; ks < Q; ++ks) {
                        const double c = xc[elem0 + 4 * ks];
                        const double v = __dsub_rn(c, __dmul_rn(z, __dsub_rn(c, xs[elem0 + 4 * ks])));
                        A[k][ks] = v;
                        out = out || (v < lo_q[ks]) || (v > hi_q[ks]);
                    }
                    out_mask[k] = __ballot_sync(0xffffffffu, out);
                }
                const int my_w = s_base + min(w_first + (lane & (IW - 1)), H - 1);       // walker layout
                const double oldlnp = sLnp[my_w];

                // ---- chi^2 of WT tiles of 8 walkers against every angle tile, with the Philox rounds of
                //      the next item's draw in between
                PhiloxState ps = philox_start(t_nxt, item_nxt);
                double chi0[WT], chi1[WT];
#pragma unroll
                for (int k = 0; k < WT; ++k) { chi0[k] = 0.0; chi1[k] = 0.0; }
                auto chi_tile = [&](int at) {
                    const double *tp = sC + at * TS;
                    const double2 dv = *reinterpret_cast<const double2 *>(tp + 2 * tig);
                    double2 Di[RR];
#pragma unroll
                    for (int l = 0; l < R; ++l) Di[l] = *reinterpret_cast<const double2 *>(tp + 8 * (1 + l) + 2 * tig);
                    double Bf[Q];
#pragma unroll
                    for (int ks = 0; ks < Q; ++ks) Bf[ks] = tp[8 * (1 + R) + 32 * ks + lane];
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
                        for (int k = 0; k < WT; ++k) dmma884(c0[k], c1[k], A[k][ks], Bf[ks]);
#pragma unroll
                    for (int k = 0; k < WT; ++k) {
                        chi0[k] = fma(c0[k], c0[k], chi0[k]);
                        chi1[k] = fma(c1[k], c1[k], chi1[k]);
                    }
                };
                int at = 0;
                const int n_fused = min(NT, 5);
                for (; at < n_fused; ++at) {
                    chi_tile(at);
                    philox_round(ps);
                    philox_round(ps);
                }
                for (int r = 2 * n_fused; r < 10; ++r) philox_round(ps);
                Draw dn;
                if (at < NT) {                            // the conversions of the draw under one more tile
                    chi_tile(at);
                    ++at;
                    dn = philox_finish(ps, t_nxt);
                } else {
                    dn = philox_finish(ps, t_nxt);
                }
#pragma unroll 2
                for (; at < NT; ++at) chi_tile(at);

                // ---- quad sums, transposed into the walker layout
                double offer;
                if (WT == 4) {
                    // lane tig ends up with the quad's sum of tile tig
                    double s[4];
#pragma unroll
                    for (int k = 0; k < 4; ++k) s[k] = chi0[k & (WT - 1)] + chi1[k & (WT - 1)];
                    const bool b0 = (tig & 1) != 0, b1 = (tig & 2) != 0;
                    const double keepA = b0 ? s[1] : s[0], sendA = b0 ? s[0] : s[1];
                    const double keepB = b0 ? s[3] : s[2], sendB = b0 ? s[2] : s[3];
                    const double a01 = keepA + __shfl_xor_sync(0xffffffffu, sendA, 1);     // tile tig & 1
                    const double a23 = keepB + __shfl_xor_sync(0xffffffffu, sendB, 1);     // tile 2 + (tig & 1)
                    const double keep = b1 ? a23 : a01, send = b1 ? a01 : a23;
                    offer = keep + __shfl_xor_sync(0xffffffffu, send, 2);
                } else if (WT == 2) {
                    const double s0 = chi0[0] + chi1[0], s1 = chi0[WT - 1] + chi1[WT - 1];
                    const bool b0 = (tig & 1) != 0;
                    const double keep = b0 ? s1 : s0, send = b0 ? s0 : s1;
                    offer = keep + __shfl_xor_sync(0xffffffffu, send, 1);                  // tile tig & 1
                    offer += __shfl_xor_sync(0xffffffffu, offer, 2);
                } else {
                    offer = chi0[0] + chi1[0];
                    offer += __shfl_xor_sync(0xffffffffu, offer, 1);
                    offer += __shfl_xor_sync(0xffffffffu, offer, 2);
                }
                const double chi = __shfl_sync(0xffffffffu, offer, wl_src);

                // ---- accept / reject (walker layout)
                unsigned int om = out_mask[0];
#pragma unroll
                for (int k = 1; k < WT; ++k) om = ((lane >> 3) == k) ? out_mask[k] : om;
                const bool outside = ((om >> (4 * (lane & 7))) & 0xFu) != 0u;
                // outside the box the log-posterior is -inf (mda.py:1576): never accepted
                const double newlnp = outside ? -CUDART_INF : chi * -0.5;                  // C/ChiSquare.c:212
                saw_nan = saw_nan || (newlnp != newlnp);
                const double est = __dsub_rn(newlnp, oldlnp) + (double)dr.fast;
                bool accept;
                if (fabs(est) > 1e-3) {
                    accept = est > 0.0;            // far from a tie: the float logs cannot change the sign
                } else {                           // emcee's test in full precision (rare: redo the draw for u')
                    const Philox4 r = philox4x32_10(Philox4{(uint32_t)my_w, bin_id, t, 0u}, seed0, seed1);
                    const double zlog = __dmul_rn(dim_m1, log(dr.z));
                    accept = __dsub_rn(__dadd_rn(zlog, newlnp), oldlnp) > log(uniform43(r.w, r.x, r.y));
                }
                accept = accept && (lane < IW) && (w_first + lane < H);
                const unsigned int acc_mask = __ballot_sync(0xffffffffu, accept);
                if (store_in_flight) {                     // the chain store must have read these rows
                    if (lane == 0) bulk_wait_read_all();
                    __syncwarp();
                    store_in_flight = false;
                }
                if (accept) {
                    sLnp[my_w] = newlnp;
                    sAcc[my_w] += 1u;
                }
#pragma unroll
                for (int k = 0; k < WT; ++k) {
                    if ((acc_mask >> (8 * k + g)) & 1u) {
                        double *nx = sPos + (s_base + w_first + 8 * k + g) * L;
#pragma unroll
                        for (int l = 0; l < R; ++l)
                            if (tig == l) nx[l] = lead[k][l];
#pragma unroll
                        for (int ks = 0; ks < Q; ++ks) nx[elem0 + 4 * ks] = A[k][ks];
                    }
                }
                dr = dn;
            }
            if (half && until_keep == 1 && keeping) fence_proxy_async_smem();   // a kept step: order the rows before the bulk store
            __syncthreads();
            if (half && --until_keep == 0) {           // a full step is complete and kept: every warp copies out ITS rows
                until_keep = args.thin;
                if (keeping && slot < args.keep_cap) {
                    double *gc = args.chain + (((size_t)slot * args.bins + (size_t)bin) * W) * L;
                    double *gl = args.lnp_chain + ((size_t)slot * args.bins + (size_t)bin) * W;
                    for (int item = warp; item < n_items; item += n_warps) {
                        const int w0 = item * IW, nw = min(IW, H - w0);
#pragma unroll
                        for (int hb = 0; hb < 2; ++hb) {
                            const int row0 = hb * H + w0;
                            if (bulk_ok && nw == IW) {
                                if (lane == 0) {
                                    bulk_s2g(gc + (size_t)row0 * L, sPos + row0 * L, (uint32_t)(IW * L * 8));
                                    bulk_s2g(gl + row0, sLnp + row0, (uint32_t)(IW * 8));
                                }
                            } else {
                                for (int i = lane; i < nw * L; i += 32) gc[(size_t)row0 * L + i] = sPos[row0 * L + i];
                                if (lane < nw) gl[row0 + lane] = sLnp[row0 + lane];
                            }
                        }
                    }
                    if (lane == 0) bulk_commit();
                    store_in_flight = true;
                }
                ++slot;
            }
        }

        // ---- hand the bin back: state to HBM/L2, then the progress word
        if (lane == 0) bulk_wait_read_all();
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
    // a full warp of walkers per item where the ensemble gives every warp one; fewer tiles per item otherwise
    s.wpt = w_tiles >= 4 * kMaxWarps ? 4 : (w_tiles >= 2 * kMaxWarps ? 2 : 1);
    if (n_l > 12 && s.wpt == 4) s.wpt = 2;                     // registers: 4 tiles x 4 k-steps of A fragments
    if (force_wt == 1 || force_wt == 2 || force_wt == 4) s.wpt = force_wt;
    int warps = (w_tiles + s.wpt - 1) / s.wpt;
    if (warps > kMaxWarps) warps = kMaxWarps;
    s.threads = 32 * warps;
    s.groups = 1;
    s.smem_bytes = EnsembleMmaSmem(walkers, n_l, n_tiles).total;
    s.global_state = false;
    s.mma = true;
    s.fits = s.smem_bytes <= smem_optin && (size_t)n_tiles * 8 * (size_t)(n_l + 1) * 8 < (1u << 20);
    (void)bins;
    (void)sm_count;
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

// Co-resident CTAs of this kernel on the whole device (the unit list needs every CTA of the grid
// resident: a unit may wait for one that runs on another CTA).
int ensemble_mma_capacity(int n_l, const EnsembleShape &shape, int sm_count) {
    EnsKernel k = pick_mma_kernel(n_l, shape.wpt);
    int per_sm = 0;
    if (!k || cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, reinterpret_cast<const void *>(k), shape.threads,
                                                            shape.smem_bytes) != cudaSuccess)
        return 0;
    return per_sm * sm_count;
}

cudaError_t launch_ensemble_mma(const EnsembleArgs &args, const EnsembleShape &shape, int grid, cudaStream_t stream) {
    if (args.bins <= 0 || args.n_steps <= 0) return cudaSuccess;
    EnsKernel k = pick_mma_kernel(args.n_l, shape.wpt);
    if (!k || !shape.fits || !shape.mma || !args.consts_mma || grid < 1 || args.chunk_steps < 1 || args.n_chunks < 1 ||
        (args.n_chunks > 1 && !args.progress))
        return cudaErrorInvalidValue;
    k<<<grid, shape.threads, shape.smem_bytes, stream>>>(args);
    return cudaGetLastError();
}

}  // namespace mda
