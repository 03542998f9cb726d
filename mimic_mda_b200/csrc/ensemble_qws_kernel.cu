// ensemble_qws_kernel.cu -- the triangular-factor stepper, warp-specialised: walker warps and draw warps.
//
// Same job, same random streams, same arithmetic and therefore the same chains as ensemble_quad_kernel.cu (the
// reference's host loop emcee.EnsembleSampler.run_mcmc -> ln_post_prob per walker, mda.py:1128-1131, 1540-1578, with
// chi^2(a) = |T [a; -1]|^2 from the bin's triangular factor, quad_chi.cuh).  What changes is who does what, and when.
//
// A half-step of the plain kernel is ~400 instructions per walker thread of which only ~170 depend on the ensemble's
// state; the rest -- the Philox4x32-10 draw, the stretch factor z = ((a-1) u + 1)^2 / a, the partner index, the
// single-precision estimate of (L-1) ln z - ln u' -- depends on the counter alone.  With one bin per SM (the shapes
// BASELINE.json names: 100 bins, or 25 per GPU when C4 is spread over 8 GPUs) there are two walker warps per scheduler and
// every half-step is a chain of exposed latencies (ncu: a warp issues in 15 % of its cycles).  Here
//
//   * 8 WALKER warps (thread i = walker i of the half being updated) take what the draw warps left for this half-step,
//     gather the partner row, propose, test the box prior (mda.py:1574-1576), evaluate chi^2, decide, overwrite their row.
//     The accept test is two thresholds on chi^2 prepared while chi^2 is being summed;
//   * 4 DRAW warps stay two half-steps ahead with the draws (a ring of three slots in shared memory) -- work that is
//     independent of the walker warps' chain and fills their stall slots;
//   * one thread owns every bulk copy: the chain stores (issued ~100 cycles after the half-step's barrier, waited for before
//     the next one), the state of the next piece of work coming in and the state of the last one going back.  In the build
//     with two CTAs per SM it is a draw thread, in the build that has its SM to itself the first thread of a fourth
//     warpgroup that does nothing else;
//   * one __syncthreads per half-step, as before.
//
// A CTA that has its SM to itself (TREG build) moves registers from the draw and copy warps to the walker warps with
// setmaxnreg (208 for a walker thread, 56 for a draw thread, 40 for the copy warpgroup): T (55 doubles at L = 9) and the
// thread's own two rows stay in registers.  The other build keeps T in shared memory and fits two CTAs per SM.
//
// Schedule: see qws_next_unit.  With more bins than resident CTAs (C4: 200 bins on 148 SMs) every CTA carries the same
// number of bin-steps whatever the length of the launch; a CTA's pieces of work are double-buffered in shared memory, the
// next one's state (positions, log-probabilities, factor) arriving by bulk copy while the current one steps, and the draw
// pipeline runs across the boundary.  At a switch the walker warps arrive at the piece barrier and go on; the copy thread
// waits there and stores the finished piece.
//
// The walker loop is written for what the profiler showed, not for looks: values the compiler would re-derive in every
// half-step (owner test, shared addresses, barrier counts) are opaque asm results, chi^2 is summed by every lane and
// +inf selected afterwards, the timing switches exist in the profiling build only (profiles/experiments/r02_stepper_notes.md).
#include "philox.cuh"
#include "quad_chi.cuh"

#include <math_constants.h>

#include <algorithm>
#include <cstdlib>
#include <type_traits>
#include <utility>

namespace mda {

namespace {

constexpr int kStatusNaN = 1;
constexpr int kStatusStalled = 4;             // a piece of work waited 4 s for the other piece of its bin: the launch gave up
constexpr int kWalkerThreads = 256;           // 8 walker warps: one walker of the active half each
constexpr int kDrawThreads = 128;             // 4 draw warps (one warpgroup: setmaxnreg works on warpgroups)
constexpr int kThreads = kWalkerThreads + kDrawThreads;          // two CTAs per SM: a draw thread also copies
constexpr int kThreadsLone = kThreads + 128;                      // a CTA with its SM to itself: plus a copy warpgroup (see qws_service)

struct QwsSmem {
    size_t lnp, fac, buf_bytes, z, ua, pf, tri, pad, total;       // a piece's buffer: positions at 0, then lnp, then fac; buffer b at b * buf_bytes
    __host__ __device__ QwsSmem(int walkers, int n_l, bool tri_in_smem) {
        const size_t h = (size_t)walkers / 2;
        size_t off = 0;
        off += (size_t)walkers * (size_t)n_l * 8;                             // [W][L] dense, updated in place; bulk-stored as is
        lnp = off; off += (size_t)walkers * 8;                                // [W]
        fac = off; off += (size_t)quad_stride(n_l) * 8;                       // the bin's packed factor as it lies in the store
        buf_bytes = off;
        off *= 2;                                                             // two pieces of work: the current one and the next one coming in
        z = off; off += 3 * h * 8;                                            // [3][H] stretch factor of the half-step's draw (ring: t, t+1, t+2)
        ua = off; off += 3 * h * 8;                                           // [3][H] acceptance uniform
        pf = off; off += 3 * h * 8;                                           // [3][H] {partner index, float (L-1) ln z - ln u'}
        off = (off + 15) & ~(size_t)15;
        tri = off; off += (size_t)quad_padded_doubles(n_l) * 8;               // T, rows padded (the build with T in shared memory; the register build leaves it unused)
        off = (off + 15) & ~(size_t)15;
        // Register build, even L: a second copy of the positions with rows of L + 1 doubles, for the partner gather only.  Dense
        // rows of an even L are 8 L bytes apart -- at L = 8 only two residue classes of bank pairs, at L = 16 one: a random
        // gather collides 8- to 16-fold (2.54 us per step at L = 8 against 1.95 with the copy); an odd stride spreads rows evenly.
        pad = off; if (!tri_in_smem && (n_l & 1) == 0) off += (size_t)walkers * (size_t)(n_l + 1) * 8;
        total = (off + 15) & ~(size_t)15;
    }
};

__device__ __forceinline__ uint32_t q_smem_addr(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
// The same, opaque to the compiler: an address formed once per piece stays in its register (the plain conversion is
// re-materialised in every half-step -- a read of the CTA's shared-memory window, SR_CgaCtaId, in front of the gather).
__device__ __forceinline__ uint32_t q_smem_addr_once(const void *p) {
    uint32_t a;
    asm volatile("{\n.reg .u64 t;\ncvta.to.shared.u64 t, %1;\ncvt.u32.u64 %0, t;\n}" : "=r"(a) : "l"(p));
    return a;
}
__device__ __forceinline__ void q_fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void q_bulk_s2g(void *dst, const void *src, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(q_smem_addr(src)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void q_bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void q_bulk_wait_read_all() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void q_bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void q_mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(q_smem_addr(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void q_mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(q_smem_addr(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void q_bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(q_smem_addr(dst)), "l"(src),
                 "r"(bytes), "r"(q_smem_addr(bar))
                 : "memory");
}
__device__ __forceinline__ void q_mbar_wait(uint64_t *bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "QWS_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra QWS_DONE;\n"
        "bra QWS_WAIT;\n"
        "QWS_DONE:\n"
        "}\n" ::"r"(q_smem_addr(bar)),
        "r"(parity)
        : "memory");
}
// The same wait, given up when *abort is set (the copy thread's hand-over timed out: the state is not coming).
__device__ __forceinline__ bool q_mbar_wait_or_abort(uint64_t *bar, uint32_t parity, const int *abort) {
    for (;;) {
        uint32_t done;
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n"
            : "=r"(done)
            : "r"(q_smem_addr(bar)), "r"(parity)
            : "memory");
        if (done) return true;
        if (*reinterpret_cast<const volatile int *>(abort)) return false;
    }
}
__device__ __forceinline__ unsigned long long q_ld_acquire_u64(const unsigned long long *p) {
    unsigned long long v;
    asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long q_ld_relaxed_u64(const unsigned long long *p) {
    unsigned long long v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void q_fence_acquire_gpu() { asm volatile("fence.acq_rel.gpu;" ::: "memory"); }
__device__ __forceinline__ void q_st_release_u64(unsigned long long *p, unsigned long long v) {
    asm volatile("st.release.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long q_globaltimer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

// Shared-memory accesses by 32-bit address.  volatile + "memory": they keep their place among the barriers.
__device__ __forceinline__ double q_lds_f64(uint32_t addr) {
    double v;
    asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(addr) : "memory");
    return v;
}
__device__ __forceinline__ uint2 q_lds_u2(uint32_t addr) {
    uint2 v;
    asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(addr) : "memory");
    return v;
}
__device__ __forceinline__ void q_sts_f64(uint32_t addr, double v) { asm volatile("st.shared.f64 [%0], %1;" ::"r"(addr), "d"(v) : "memory"); }

__device__ __forceinline__ double q_uniform43(uint32_t w, uint32_t x, uint32_t y) {
    return ((double)w * 2048.0 + (double)(((x & 31u) << 6) | (y & 63u))) * (1.0 / 8796093022208.0);
}

// One piece of work of a CTA: steps [s0, s1) of the launch for bin `bin`.
struct QwsUnit {
    int bin, s0, s1;
    bool waits;        // the bin's earlier steps run on the next CTA: take the state over behind its progress word
    bool signals;      // the bin's later steps run on the previous CTA: publish the state when done
};

// The schedule.  Up to one bin per CTA every CTA simply owns a bin.  With more bins than CTAs (grid = SMs: C4's 200 bins on
// 148 SMs) the bins' step ranges are laid end to end on a ribbon of bins x n_steps bin-steps and the ribbon is cut into
// gridDim.x equal segments, one per CTA (McNaughton's wrap-around rule): every CTA carries the same number of bin-steps
// whatever the launch length, a bin is cut at most once, and its two pieces never overlap in time -- the piece at the START of
// CTA j+1's segment takes the bin's FIRST steps, the piece at the END of CTA j's segment the rest, which (a segment being at
// least n_steps long) begins after the other has ended.  The later piece still waits for the earlier one's progress word; the
// grid is launched cooperatively, so the CTA it waits for is resident by construction.
__device__ __forceinline__ bool qws_next_unit(const EnsembleArgs &args, long long &p, long long r1, QwsUnit &u) {
    if (p >= r1) return false;
    const long long K = args.n_steps;
    // (a 64-bit division is a subroutine of a few hundred cycles, at every piece switch; ribbons below 2^31 bin-steps take the 32-bit one)
    const long long b = (r1 >> 31) == 0 ? (long long)((uint32_t)p / (uint32_t)K) : p / K, o = p - b * K;
    const long long len = min(K - o, r1 - p);
    u.bin = (int)b;
    if (o == 0 && len == K) { u.s0 = 0; u.s1 = (int)K; u.waits = u.signals = false; }
    else if (o == 0) { u.s0 = (int)(K - len); u.s1 = (int)K; u.waits = true; u.signals = false; }     // the end of my segment: the bin's last steps
    else { u.s0 = 0; u.s1 = (int)len; u.waits = false; u.signals = true; }                              // the start of my segment: its first steps
    p += len;
    return true;
}

struct QwsCtx {
    unsigned char *smem_raw;
    uint64_t *s_bar;
    int *s_abort;
    int W, H;
    long long total, ribbon_end;
    bool keeping, split;
    QwsSmem lay{2, 1, false};
};

// Whether the walker warps of the build with a copy warpgroup gather in two groups (see the walker warps): more than four
// warps of owners, whole warps only.  Debug bit 2 switches it off.
__device__ __forceinline__ bool qws_skewed(int H, int debug) { return H > kWalkerThreads / 2 && (H & 31) == 0 && (debug & 4) == 0; }
constexpr long long kChainStoreHoldCycles = 100;  // see qws_service

// The barrier that only the walker warps and the thread that copies take part in (state of a piece complete / whole-state
// chain store read).  In the build with a copy warpgroup the draw warps do not track pieces of work at all: they count
// half-steps (qws_draw_loop), so these hand-offs are a named barrier of the other 384 threads.
template <bool COPYWG>
__device__ __forceinline__ void qws_piece_barrier() {
    if constexpr (COPYWG) asm volatile("bar.sync 2, 384;" ::: "memory");
    else __syncthreads();
}

__device__ __noinline__ double qws_divide(double x, double y) { return x / y; }

// One walker's draw of half-step tt into slot `slot` (= ring position * H + walker): counter (walker, bin, half-step, 0),
// round keys from the parameter block.  numpy arithmetic is one rounding per operation: no FMA contraction.
template <int L>
__device__ __forceinline__ void qws_draw(const EnsembleArgs &args, double *sZ, double *sUa, uint2 *sPF, uint32_t walker, uint32_t bin_id,
                                         uint32_t tt, uint32_t H, int slot, bool valid) {
    const Philox4 r = philox4x32_10_keys(Philox4{walker, bin_id, tt, 0u}, args.round_keys);
    const double zr = __dadd_rn(__dmul_rn(args.a - 1.0, uniform53(r.x, r.y)), 1.0);
    const double z2 = __dmul_rn(zr, zr);
    double z;                                                                    // ((a-1) u + 1)^2 / a
    if (args.a_inv != 0.0) z = z2 * args.a_inv; else z = qws_divide(z2, args.a);
    const double ua = q_uniform43(r.w, r.x, r.y);
    const float fast = (float)(L - 1) * __logf((float)z) - __logf((float)ua);   // decides all but near-ties
    if (valid) {
        sZ[slot] = z;
        sUa[slot] = ua;
        sPF[slot] = make_uint2(__umulhi(r.z, H), __float_as_uint(fast));
    }
}

// The draw warps of the build with a copy warpgroup: NDRAW threads, thread d draws for walkers d, d + NDRAW, ... of the half,
// two half-steps ahead of the walkers, over the CTA's whole sequence of half-steps (across its pieces of work).  Nothing
// but the cursor, the draw and one barrier per half-step: these warps are the pacemaker of a CTA that has its SM to itself
// (with the piece bookkeeping of qws_service in the same 56 registers a half-step's 256 draws took 1750 cycles).
template <int L, int NDRAW, bool PROF>
__device__ __forceinline__ void qws_draw_loop(const EnsembleArgs &args, unsigned char *smem_raw, const QwsSmem &lay, const int H,
                                              const long long seg_begin, const long long seg_end, const int d) {
    double *sZ = reinterpret_cast<double *>(smem_raw + lay.z);
    double *sUa = reinterpret_cast<double *>(smem_raw + lay.ua);
    uint2 *sPF = reinterpret_cast<uint2 *>(smem_raw + lay.pf);
    long long cursor = seg_begin;
    QwsUnit u;
    bool valid = qws_next_unit(args, cursor, seg_end, u);
    uint32_t bin_id = 0, t = 0, t_end = 0;
    auto open = [&]() {
        bin_id = (uint32_t)args.bin_ids[u.bin];
        t = 2u * (uint32_t)(args.step0 + u.s0);
        t_end = 2u * (uint32_t)(args.step0 + u.s1);
    };
    open();
    int ring_base = 0;                                   // ring position * H
    int made = 0;
    auto draw_one = [&]() {
        if (!valid) return;
        if (PROF && (args.groups & 2) && ++made > 3) return;     // (debug bit 1: the first three draws are reused for ever)
        const uint32_t w_base = (t & 1u) ? (uint32_t)H : 0u;
        if (NDRAW >= kWalkerThreads) {
            qws_draw<L>(args, sZ, sUa, sPF, w_base + (uint32_t)min(d, H - 1), bin_id, t, (uint32_t)H, ring_base + d, d < H);
        } else {                                         // two independent Philox chains interleave
            qws_draw<L>(args, sZ, sUa, sPF, w_base + (uint32_t)min(d, H - 1), bin_id, t, (uint32_t)H, ring_base + d, d < H);
            if (H > NDRAW)
                qws_draw<L>(args, sZ, sUa, sPF, w_base + (uint32_t)min(d + NDRAW, H - 1), bin_id, t, (uint32_t)H, ring_base + d + NDRAW, d + NDRAW < H);
        }
        ring_base = ring_base == 2 * H ? 0 : ring_base + H;
        if (++t == t_end) {
            valid = qws_next_unit(args, cursor, seg_end, u);
            if (valid) open();
        }
    };
    const int n_half = (int)(2 * (seg_end - seg_begin));
    long long p_d[2] = {0, 0};
    draw_one();
    draw_one();
    __syncthreads();                                     // (A) the first draws are in place
    for (int h = 0; h < n_half; ++h) {
        long long c0 = 0;
        if constexpr (PROF) c0 = clock64();
        draw_one();
        if constexpr (PROF) { const long long c1 = clock64(); p_d[0] += c1 - c0; c0 = c1; }
        __syncthreads();                                 // (B) the half is complete
        if constexpr (PROF) p_d[1] += clock64() - c0;
    }
    if constexpr (PROF) if (d == 0 && args.prof) {
        long long *o = args.prof + (size_t)blockIdx.x * 128 + 112;
        o[0] = p_d[0]; o[1] = 0; o[2] = p_d[1]; o[3] = 0; o[4] = n_half;
    }
}

// What the non-walker threads do.  DRAWS: this thread (d of 128) makes the draws, two half-steps ahead of the walkers.
// COPIES: thread d == 0 owns every bulk copy and the progress words.  One warpgroup does both (two CTAs per SM), or -- in the
// build whose CTA has its SM to itself -- the draw warps only draw and one thread of a fourth warpgroup only copies: issuing
// two bulk stores and computing their addresses cost the draw warp that did it ~500 cycles per half-step, and the draw warps
// are the last to arrive at the half-step's barrier as it is.
template <int L, bool DRAWS, bool COPIES, bool PROF>
__device__ __forceinline__ void qws_service(const EnsembleArgs &args, const QwsCtx &cx, QwsUnit unit, long long ribbon, const int d) {
    unsigned char *smem_raw = cx.smem_raw;
    uint64_t *s_bar = cx.s_bar;
    int &s_abort = *cx.s_abort;
    const int W = cx.W, H = cx.H;
    const long long total = cx.total, ribbon_end = cx.ribbon_end;
    const bool keeping = cx.keeping, split = cx.split;
    const bool hold = !DRAWS && qws_skewed(H, PROF ? args.groups : 0) && (!PROF || (args.groups & 8) == 0);     // (debug bit 3: no hold)
    const QwsSmem &lay = cx.lay;
    double *sZ = reinterpret_cast<double *>(smem_raw + lay.z);
    double *sUa = reinterpret_cast<double *>(smem_raw + lay.ua);
    uint2 *sPF = reinterpret_cast<uint2 *>(smem_raw + lay.pf);
    auto pos_of = [&](int b) { return reinterpret_cast<double *>(smem_raw + (size_t)b * lay.buf_bytes); };
    auto lnp_of = [&](int b) { return reinterpret_cast<double *>(smem_raw + (size_t)b * lay.buf_bytes + lay.lnp); };
    auto fac_of = [&](int b) { return reinterpret_cast<double *>(smem_raw + (size_t)b * lay.buf_bytes + lay.fac); };
    const bool copier = COPIES && d == 0;
    QwsUnit next;
    const uint32_t state_bytes = (uint32_t)W * (L + 1) * 8u + (uint32_t)quad_stride(L) * 8u;
    auto load_state = [&](int bin, int b) {              // (thread d == 0) positions, log-probabilities and factor of `bin` into buffer b
        q_mbar_expect_tx(&s_bar[b], state_bytes);
        q_bulk_g2s(pos_of(b), args.pos + (size_t)bin * W * L, (uint32_t)W * L * 8u, &s_bar[b]);
        q_bulk_g2s(lnp_of(b), args.lnp + (size_t)bin * W, (uint32_t)W * 8u, &s_bar[b]);
        q_bulk_g2s(fac_of(b), args.quad + (size_t)bin * quad_stride(L), (uint32_t)quad_stride(L) * 8u, &s_bar[b]);
    };
    auto make_draws = [&](uint32_t bin_id, uint32_t tt, int ring) {   // the draws of half-step tt into slot `ring` of the ring of three
        const int buf = ring * H;
        const uint32_t w_base = (tt & 1u) ? (uint32_t)H : 0u;
        // a thread's two walkers (d and d + 128) in one straight-line block: two independent Philox chains interleave
        qws_draw<L>(args, sZ, sUa, sPF, w_base + (uint32_t)min(d, H - 1), bin_id, tt, (uint32_t)H, buf + d, d < H);
        if (H > kDrawThreads)
            qws_draw<L>(args, sZ, sUa, sPF, w_base + (uint32_t)min(d + kDrawThreads, H - 1), bin_id, tt, (uint32_t)H, buf + d + kDrawThreads,
                        d + kDrawThreads < H);
    };
    // the draw cursor runs over the CTA's whole sequence of half-steps, across its pieces of work
    long long d_ribbon = (long long)blockIdx.x * total / gridDim.x;
    QwsUnit d_unit;
    bool d_valid = qws_next_unit(args, d_ribbon, ribbon_end, d_unit);
    uint32_t d_bin_id = 0, d_t = 0, d_t_end = 0;
    auto cursor_open = [&]() {
        d_bin_id = (uint32_t)args.bin_ids[d_unit.bin];
        d_t = 2u * (uint32_t)(args.step0 + d_unit.s0);
        d_t_end = 2u * (uint32_t)(args.step0 + d_unit.s1);
    };
    int d_ring = 0;
    auto draw_one = [&]() {                              // the next half-step of the sequence, if there is one
        if (!d_valid) return;
        make_draws(d_bin_id, d_t, d_ring);
        d_ring = d_ring == 2 ? 0 : d_ring + 1;
        if (++d_t == d_t_end) { d_valid = qws_next_unit(args, d_ribbon, ribbon_end, d_unit); if (d_valid) cursor_open(); }
    };
    cursor_open();
    int buf = 0;
    bool next_loaded = false;
    long long p_d[5] = {0, 0, 0, 0, 0};
    int pending_release = -1;                            // bin whose state store must complete before its progress word is set
    if (copier && !unit.waits) load_state(unit.bin, 0);
    if (copier && unit.waits) {                          // (a one-unit segment that is the tail of a bin)
        const unsigned long long want = args.launch_serial + 1ull, t_start = q_globaltimer_ns();
        while (q_ld_acquire_u64(args.progress + unit.bin) != want) {
            __nanosleep(100);
            if (q_globaltimer_ns() - t_start > 4000000000ull) { s_abort = 1; break; }
        }
        asm volatile("fence.proxy.async;" ::: "memory");
        if (!s_abort) load_state(unit.bin, 0);
    }
    if constexpr (DRAWS) { draw_one(); draw_one(); }
    __syncthreads();                                     // (A) the first draws are in place
    while (true) {
        const bool has_next = qws_next_unit(args, ribbon, ribbon_end, next);
        const int bin = unit.bin;
        if (s_abort) break;
        const long long step_first = args.step0 + unit.s0;
        const uint32_t t_begin = 2u * (uint32_t)step_first, t_end = 2u * (uint32_t)(args.step0 + unit.s1);
        // (step counters stay below 2^31 -- mdaEnsembleAdvance refuses more: 32-bit divisions, a tenth of the 64-bit ones' cost)
        const uint32_t thin_u = (uint32_t)args.thin, first_u = (uint32_t)step_first;
        int until_keep = args.thin - (int)(first_u % thin_u);
        long long slot = args.keep0 + (long long)(first_u / thin_u - (uint32_t)args.step0 / thin_u);
        double *sPos = pos_of(buf), *sLnp = lnp_of(buf);
        for (uint32_t t = t_begin; t < t_end; ++t) {
            const uint32_t half = t & 1u;
            const int s_base = half ? H : 0;
            const bool kept_step = keeping && until_keep == 1 && slot < args.keep_cap;
            long long c0 = 0;
            if constexpr (PROF) c0 = clock64();
            if constexpr (DRAWS) draw_one();
            if constexpr (PROF) { const long long c1 = clock64(); p_d[0] += c1 - c0; c0 = c1; }
            // the last store read the half that is written next, a whole half-step after it was issued: no wait in practice
            if (keeping && copier) q_bulk_wait_read_all();
            if constexpr (PROF) { const long long c1 = clock64(); p_d[1] += c1 - c0; c0 = c1; }
            __syncthreads();                             // (B) the half is complete
            if constexpr (PROF) { const long long c1 = clock64(); p_d[2] += c1 - c0; c0 = c1; }
            if (copier) {
                if (pending_release >= 0) {              // the previous piece's state is on its way out since the switch
                    q_bulk_wait_all();
                    asm volatile("fence.proxy.async;" ::: "memory");
                    __threadfence();
                    q_st_release_u64(args.progress + pending_release, args.launch_serial + 1ull);
                    pending_release = -1;
                }
                if (kept_step) {
                    // The store reads 8 H (L + 1) bytes of shared memory while the next half-step's partner gather saturates that
                    // pipe.  Issued right behind (B) its reads delay walker warps 0..3 and with them warps 4..7, which start
                    // their gather behind those (the skew, see the walker warps).  Held back ~100 cycles -- warps 0..3 have
                    // issued their loads by then -- a step takes 1.70 instead of 1.76 us (512 x 9; 2.31 instead of 2.38 at C4;
                    // 1 to 2.5 % at 384 x 9, 512 x 4, 512 x 7, 512 x 8); held 300 cycles or more the time comes back as a wait
                    // for the store's reads before the next (B), and with no skew (256 x 5) any hold costs 10 %.
                    if (hold) {
                        const long long until = clock64() + kChainStoreHoldCycles;
                        while (clock64() < until) __nanosleep(20);
                    }
                    double *dst = args.chain + (((size_t)slot * args.bins + (size_t)bin) * W) * L;
                    double *ldst = args.lnp_chain + ((size_t)slot * args.bins + (size_t)bin) * W;
                    if (split) {                         // the half just updated is final for this step
                        q_bulk_s2g(dst + (size_t)s_base * L, sPos + (size_t)s_base * L, (uint32_t)H * L * 8u);
                        q_bulk_s2g(ldst + s_base, sLnp + s_base, (uint32_t)H * 8u);
                        q_bulk_commit();
                    } else if (half) {
                        q_bulk_s2g(dst, sPos, (uint32_t)W * L * 8u);
                        q_bulk_s2g(ldst, sLnp, (uint32_t)W * 8u);
                        q_bulk_commit();
                    }
                }
                // the next piece's state comes in behind this one's steps (its buffer was last read by the store of the piece
                // before this one); a piece that continues another CTA's bin only once that CTA has published it
                if (has_next && !next_loaded && t > t_begin) {
                    // (a look every fourth half-step, relaxed, the acquire only once the word is there: an acquiring load
                    // invalidates the L1 and takes an L2 round trip -- with the chain store of the same half-step that was most
                    // of the copy thread's 1500 cycles, and the copy thread paced the CTA while it waited for a word)
                    bool ready = !next.waits;
                    if (!ready && (t & 3u) == 0u && q_ld_relaxed_u64(args.progress + next.bin) == args.launch_serial + 1ull) {
                        q_fence_acquire_gpu();
                        asm volatile("fence.proxy.async;" ::: "memory");
                        ready = true;
                    }
                    if (ready) {
                        q_bulk_wait_read_all();
                        load_state(next.bin, buf ^ 1);
                        next_loaded = true;
                    }
                }
            }
            if (kept_step && !split && half) {           // whole-state store: nobody may write until it has been read
                if (copier) q_bulk_wait_read_all();
                qws_piece_barrier<!DRAWS>();             // (C)
            }
            if (half) {
                if (--until_keep == 0) { until_keep = args.thin; if (keeping) ++slot; }
            }
            if constexpr (PROF) { const long long c1 = clock64(); p_d[3] += c1 - c0; p_d[4] += 1; }
        }
        // ---- switch: this piece's state goes back, the next one's must be on its way
        if (copier && has_next && !next_loaded) {
            const unsigned long long want = args.launch_serial + 1ull, t_start = q_globaltimer_ns();
            while (next.waits && q_ld_acquire_u64(args.progress + next.bin) != want) {
                __nanosleep(100);
                if (q_globaltimer_ns() - t_start > 4000000000ull) { s_abort = 1; break; }     // 4 s: it is not coming
            }
            asm volatile("fence.proxy.async;" ::: "memory");
            if (!s_abort) { q_bulk_wait_read_all(); load_state(next.bin, buf ^ 1); }
        }
        qws_piece_barrier<!DRAWS>();                     // (D) every walker thread has fenced its last writes; abort flag settled
        if (copier) {
            q_bulk_s2g(args.pos + (size_t)bin * W * L, sPos, (uint32_t)W * L * 8u);
            q_bulk_s2g(args.lnp + (size_t)bin * W, sLnp, (uint32_t)W * 8u);
            q_bulk_commit();
            if (unit.signals) pending_release = bin;     // the rest of this bin's steps is waiting on another CTA
        }
        if (!has_next) break;
        unit = next;
        buf ^= 1;
        next_loaded = false;
    }
    if constexpr (PROF) if (d == 0 && args.prof) for (int k = 0; k < 5; ++k) args.prof[(size_t)blockIdx.x * 128 + (DRAWS ? 112 : 120) + k] = p_d[k];
    if constexpr (PROF) if (copier && args.prof) args.prof[(size_t)blockIdx.x * 128 + 125] = (long long)q_globaltimer_ns();     // before the last copies are waited for
    if (copier) {
        // Shared memory must outlive the copies' reads; their writes are complete when the grid is (nobody inside this launch
        // reads them, except the CTA that waits for a progress word: that one needs the full wait before the release).
        if (pending_release >= 0) {
            q_bulk_wait_all();
            asm volatile("fence.proxy.async;" ::: "memory");
            __threadfence();
            q_st_release_u64(args.progress + pending_release, args.launch_serial + 1ull);
        } else {
            q_bulk_wait_read_all();
        }
    }
    if constexpr (PROF) if (copier && args.prof) args.prof[(size_t)blockIdx.x * 128 + 127] = (long long)q_globaltimer_ns();     // the CTA's last instruction but a few
}

// PROF: thread 0 stamps the phases of every piece of work into args.prof[cta][...] (debug build, selected when
// mdaEnsembleProfile switched the counters on).
template <int L, bool TREG, bool PROF = false>
__global__ void __launch_bounds__(TREG ? kThreadsLone : kThreads, TREG ? 1 : 2) ensemble_qws_kernel(const EnsembleArgs args) {
    constexpr bool COPYWG = TREG;                            // a fourth warpgroup whose first thread does nothing but the bulk copies
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ __align__(8) uint64_t s_bar[2];
    __shared__ int s_abort;
    const int tid = threadIdx.x;
    const bool is_walker = tid < kWalkerThreads;
    const int W = args.walkers, H = W >> 1;
    const QwsSmem lay(W, L, !TREG);
    double *sZ = reinterpret_cast<double *>(smem_raw + lay.z);
    double *sUa = reinterpret_cast<double *>(smem_raw + lay.ua);
    uint2 *sPF = reinterpret_cast<uint2 *>(smem_raw + lay.pf);
    double *sTri = reinterpret_cast<double *>(smem_raw + lay.tri);
    auto pos_of = [&](int b) { return reinterpret_cast<double *>(smem_raw + (size_t)b * lay.buf_bytes); };
    auto lnp_of = [&](int b) { return reinterpret_cast<double *>(smem_raw + (size_t)b * lay.buf_bytes + lay.lnp); };
    auto fac_of = [&](int b) { return reinterpret_cast<double *>(smem_raw + (size_t)b * lay.buf_bytes + lay.fac); };

    if constexpr (PROF) if (tid == 0 && args.prof) args.prof[(size_t)blockIdx.x * 128 + 126] = (long long)q_globaltimer_ns();   // kernel entry (wall clock)
    if (tid == 0) {
        q_mbar_init(&s_bar[0], 1);
        q_mbar_init(&s_bar[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        s_abort = 0;
    }
    const bool keeping = args.chain != nullptr;
    // a half can be stored on its own if its rows and log-probabilities are 16-byte multiples (bulk-copy granularity)
    const bool split = ((H * L) & 1) == 0 && (H & 1) == 0;
    // my segment of the ribbon of bins x n_steps bin-steps (a bin per CTA when the grid is the number of bins)
    const long long total = (long long)args.bins * args.n_steps;
    long long ribbon = (long long)blockIdx.x * total / gridDim.x;
    const long long ribbon_end = (long long)(blockIdx.x + 1) * total / gridDim.x;
    QwsUnit unit, next;
    bool has_unit = qws_next_unit(args, ribbon, ribbon_end, unit);
    __syncthreads();                                         // (I) the mbarriers are initialised
    if (!has_unit) return;

    if (!is_walker) {
        // =========================================================== draw warps and the copy thread
        QwsCtx cx;
        cx.smem_raw = smem_raw; cx.s_bar = s_bar; cx.s_abort = &s_abort; cx.W = W; cx.H = H; cx.total = total; cx.ribbon_end = ribbon_end;
        cx.keeping = keeping; cx.split = split; cx.lay = lay;
        if constexpr (COPYWG) {
            if (tid < kWalkerThreads + kDrawThreads) {
                asm volatile("setmaxnreg.dec.sync.aligned.u32 56;");
                qws_draw_loop<L, kDrawThreads, PROF>(args, smem_raw, lay, H, (long long)blockIdx.x * total / gridDim.x, ribbon_end, tid - kWalkerThreads);
            } else {
                asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
                qws_service<L, false, true, PROF>(args, cx, unit, ribbon, tid - kWalkerThreads - kDrawThreads);
            }
        } else {
            qws_service<L, true, true, PROF>(args, cx, unit, ribbon, tid - kWalkerThreads);
        }
        return;
    }

    // =============================================================== walker warps
    if constexpr (TREG) asm volatile("setmaxnreg.inc.sync.aligned.u32 208;");
    // (an opaque flag: the compiler would otherwise re-read %tid.x and compare in every half-step -- a special-register read of
    // ~40 cycles in front of the half-step's fence and barrier -- rather than keep one of 208 registers for it)
    unsigned int owner_flag;
    asm volatile("{\n.reg .pred p;\nsetp.lt.s32 p, %1, %2;\nselp.u32 %0, 1, 0, p;\n}" : "=r"(owner_flag) : "r"(tid), "r"(H));
    const bool owner = owner_flag != 0u;
    // (likewise H, the skew barrier's count: re-derived from the parameter block -- a constant-bank load and a shift of ~40
    // cycles in front of that barrier -- unless it is opaque)
    int h_pinned;
    asm volatile("mov.s32 %0, %1;" : "=r"(h_pinned) : "r"(H));
    const int me = min(tid, H - 1);
    const double dim_m1 = (double)(L - 1);
    bool saw_nan = false;
    int buf = 0, ring = 0, unit_index = 0;
    uint32_t load_parity[2] = {0u, 0u};
    double z = 0.0;
    uint2 pf = make_uint2(0u, 0u);
    __syncthreads();                                         // (A)
    if constexpr (PROF) if (tid == 0 && args.prof) args.prof[(size_t)blockIdx.x * 128 + 119] = (long long)q_globaltimer_ns();   // first draws made
    if (owner) { z = sZ[me]; pf = sPF[me]; }                 // the first half-step's draw; from then on fetched one half-step ahead
    while (true) {
        const bool has_next = qws_next_unit(args, ribbon, ribbon_end, next);
        const int bin = unit.bin;
        if (s_abort) {
            if (tid == 0) atomicOr(args.status, kStatusStalled);
            break;
        }
        long long *ustamp = (PROF && tid == 0 && args.prof && unit_index < 12) ? args.prof + (size_t)blockIdx.x * 128 + 16 + 8 * unit_index : nullptr;
        ++unit_index;
        if (ustamp) { ustamp[0] = clock64(); ustamp[6] = unit.s1 - unit.s0; }
        double *sPos = pos_of(buf), *sLnp = lnp_of(buf);
        // the piece's state has arrived (or never will: the walker warps do not wait at (D), where the copy thread settles the flag)
        if (!q_mbar_wait_or_abort(&s_bar[buf], buf ? load_parity[1] : load_parity[0], &s_abort)) {
            if (tid == 0) atomicOr(args.status, kStatusStalled);
            break;
        }
        if (buf) load_parity[1] ^= 1u; else load_parity[0] ^= 1u;
        if (ustamp) ustamp[1] = clock64();
        // The register build keeps the first kSplit rows of T in registers and the rest in shared memory (16-byte broadcast
        // loads).  With 32-bit shared addressing there is room for ALL of T and the thread's own two rows (x0, x1 below) in
        // the 208 registers of a walker thread, so kSplit = L + 1; measured at L = 9 (us per step, one bin per SM): 6 rows 1.88,
        // 7: 1.87, 8: 1.92, 9: 1.76, 10: 1.76.
        constexpr int kSplit = TREG ? L + 1 : 0;
        constexpr bool kPadGather = TREG && (L & 1) == 0;
        constexpr int kTriRegs = quad_padded_offset(L, kSplit) > 0 ? quad_padded_offset(L, kSplit) : 1;
        double tri_regs[kTriRegs];
        if constexpr (TREG) {
            // all of T into registers: the packed factor lies 16-byte aligned in the piece's buffer (quad_stride is even), so it
            // is read with 16-byte broadcast loads and sorted into the padded register layout at compile time
            const double2 *packed2 = reinterpret_cast<const double2 *>(fac_of(buf));
            double2 flat[quad_stride(L) / 2];
#pragma unroll
            for (int k = 0; k < quad_stride(L) / 2; ++k) flat[k] = packed2[k];
#pragma unroll
            for (int k = 0; k < kTriRegs; ++k) tri_regs[k] = 0.0;
#pragma unroll
            for (int i = 0; i <= L; ++i)
#pragma unroll
                for (int k = 0; k < L + 1 - i; ++k) {
                    const int e = quad_row_offset(L, i) + k;
                    tri_regs[quad_padded_offset(L, i) + k] = (e & 1) ? flat[e >> 1].y : flat[e >> 1].x;
                }
            if constexpr (kPadGather) {                      // the odd-stride copy of the positions the partner gather reads (see QwsSmem)
                if (owner) {
                    double *sPad = reinterpret_cast<double *>(smem_raw + lay.pad);
#pragma unroll
                    for (int l = 0; l < L; ++l) {
                        sPad[(size_t)me * (L + 1) + l] = sPos[(size_t)me * L + l];
                        sPad[(size_t)(H + me) * (L + 1) + l] = sPos[(size_t)(H + me) * L + l];
                    }
                }
                asm volatile("bar.sync 1, 256;" ::: "memory");     // walker warps only (named barrier 1)
            }
        } else {
            // T into shared memory, rows padded for 16-byte broadcast reads (walker warps only: named barrier 1, 256 threads; the
            // previous piece's last read of sTri lies before its barrier (D))
            const double *packed = fac_of(buf);
            tri_regs[0] = 0.0;
            for (int k = tid; k < quad_padded_doubles(L); k += kWalkerThreads) sTri[k] = 0.0;
            asm volatile("bar.sync 1, 256;" ::: "memory");
#pragma unroll
            for (int i = 0; i <= L; ++i)
                for (int k = tid; k < L + 1 - i; k += kWalkerThreads) sTri[quad_padded_offset(L, i) + k] = packed[quad_row_offset(L, i) + k];
            asm volatile("bar.sync 1, 256;" ::: "memory");
        }
        const long long step_first = args.step0 + unit.s0;
        const uint32_t t_begin = 2u * (uint32_t)step_first, t_end = 2u * (uint32_t)(args.step0 + unit.s1);
        // (step counters stay below 2^31 -- mdaEnsembleAdvance refuses more: 32-bit divisions, a tenth of the 64-bit ones' cost)
        const uint32_t thin_u = (uint32_t)args.thin, first_u = (uint32_t)step_first;
        int until_keep = args.thin - (int)(first_u % thin_u);
        long long slot = args.keep0 + (long long)(first_u / thin_u - (uint32_t)args.step0 / thin_u);
        unsigned int acc0 = 0u, acc1 = 0u;                   // proposals accepted in this piece, per walker of either half
        long long p_w[PROF ? 5 : 1] = {};
        if (ustamp) ustamp[2] = clock64();
        if constexpr (PROF) if (tid == 0 && args.prof && unit_index == 1) args.prof[(size_t)blockIdx.x * 128 + 118] = (long long)q_globaltimer_ns();   // first piece ready to step

        // 32-bit shared-memory addresses, formed once per piece: the partner row's is then one multiply-add from the draw
        const uint32_t pos_a = q_smem_addr_once(sPos), lnp_a = q_smem_addr_once(sLnp), pad_a = q_smem_addr_once(smem_raw + lay.pad);
        const uint32_t draw_a = q_smem_addr_once(sZ) + (uint32_t)me * 8u;       // my slot of ring position 0; Ua and PF lie 3 H and 6 H doubles on
        uint32_t ring_bytes, ring_span;                                    // (opaque: kept, not re-derived from the parameter block)
        asm volatile("shl.b32 %0, %2, 3;\n\tmul.lo.u32 %1, %2, 24;" : "=r"(ring_bytes), "=r"(ring_span) : "r"(h_pinned));
        uint32_t ring_off = (uint32_t)ring * ring_bytes;                   // ring position of this half-step's draw
        // My two walkers' rows and log-probabilities.  Register build: both stay in registers for the whole piece (reloading
        // the next half-step's row from shared memory at the end of every half-step cost 6 % of the step).  Other build: the
        // next half-step's row is fetched at the end of the half-step before (its half is not written meanwhile).
        constexpr bool XREG = TREG;
        double x0[L], x1[XREG ? L : 1], lnp0 = 0.0, lnp1 = 0.0;
        if (owner) {
#pragma unroll
            for (int l = 0; l < L; ++l) x0[l] = q_lds_f64(pos_a + (uint32_t)me * (uint32_t)(L * 8) + 8u * l);
            lnp0 = q_lds_f64(lnp_a + (uint32_t)me * 8u);
            if constexpr (XREG) {
#pragma unroll
                for (int l = 0; l < L; ++l) x1[XREG ? l : 0] = q_lds_f64(pos_a + (uint32_t)(H + me) * (uint32_t)(L * 8) + 8u * l);
                lnp1 = q_lds_f64(lnp_a + (uint32_t)(H + me) * 8u);
            }
        }
        // x, oldlnp: the row and log-probability of the walker updated in this half-step (register build: x0/lnp0 or x1/lnp1;
        // other build: always x0/lnp0, reloaded at the end of every half-step)
        auto half_step = [&](const int half, uint32_t t, double (&x)[L], double &oldlnp, unsigned int &acc) {   // half: a literal at both call sites
            const int s_base = half ? h_pinned : 0, c_base = half ? 0 : h_pinned;   // updated half; partners come from the other one
            const bool kept_step = keeping && until_keep == 1 && slot < args.keep_cap;
            long long w0 = 0;
            if constexpr (PROF) w0 = clock64();
            const uint32_t next_off = ring_off + ring_bytes == ring_span ? 0u : ring_off + ring_bytes;
            // The partner rows: 256 random 72-byte rows are ~530 shared-memory wavefronts (3.7 per half-warp access: bank
            // collisions of random rows), a burst that all eight walker warps would wait out together before they queue
            // for the fp64 pipe together.  Skewed instead: warps 4..7 issue their loads only after warps 0..3 have issued
            // theirs (the load queue of a scheduler is served in order), so each scheduler's first warp sums its chi^2 while
            // the second one's rows arrive, and has the fp64 pipe to itself.
            // (the lone-CTA build only, and only when whole warps own walkers -- H a multiple of 32 -- so that every lane of warps
            // 0..3 arrives and every lane of the owning warps 4.. waits: the barrier's count is the H owning threads.  Debug bit 2
            // switches it off.)
            const bool skew = TREG && qws_skewed(h_pinned, PROF ? args.groups : 0);
            if (owner) {
                if (skew && tid >= kWalkerThreads / 2) asm volatile("bar.sync 3, %0;" ::"r"(h_pinned) : "memory");
                const uint32_t partner = (uint32_t)(c_base + ((PROF && (args.groups & 1)) ? me : (int)pf.x));   // (debug bit 0: no gather conflicts)
                const uint32_t crow = kPadGather ? pad_a + partner * (uint32_t)((L + 1) * 8) : pos_a + partner * (uint32_t)(L * 8);
                double c[L];
#pragma unroll
                for (int l = 0; l < L; ++l) c[l] = q_lds_f64(crow + 8u * l);
                if (skew && tid < kWalkerThreads / 2) asm volatile("bar.arrive 3, %0;" ::"r"(h_pinned) : "memory");
                const uint32_t xrow = pos_a + (uint32_t)(s_base + me) * (uint32_t)(L * 8);
                const uint32_t nrow = pos_a + (uint32_t)(c_base + me) * (uint32_t)(L * 8);      // my walker of the other half: the next one
                // accept for sure if lnp(q) - lnp(x) + fast > 1e-4, reject for sure if < -1e-4 (fast: the float estimate of
                // (L-1) ln z - ln u', error < 2e-5); with lnp = -chi^2 / 2 both are thresholds on chi^2, ready before it is
                const double fast_d = (double)__uint_as_float(pf.y);
                const double chi_accept = -2.0 * ((oldlnp - fast_d) + 1e-4), chi_reject = -2.0 * ((oldlnp - fast_d) - 1e-4);
                double q[L];
                bool outside = false;
#pragma unroll
                for (int l = 0; l < L; ++l) {
                    q[l] = __dsub_rn(c[l], __dmul_rn(z, __dsub_rn(c[l], x[l])));     // numpy rounding: no FMA contraction
                    outside = outside | (q[l] < args.lo[l]) | (q[l] > args.hi[l]);   // mda.py:1575; a NaN compares false, as there
                }
                if constexpr (PROF) { const long long w1 = clock64(); p_w[0] += w1 - w0; w0 = w1; }
                // chi^2 for every lane, +inf put in afterwards for the proposals outside the box: no divergence region around the sum
                // (its reconvergence cost every warp ~50 cycles for lanes that are almost never there)
                double chi;
                if constexpr (TREG) chi = chi_from_factor_split<L, kSplit>(tri_regs, sTri, q);
                else chi = chi_from_factor<L>(sTri, q);
                asm volatile("{\n.reg .pred p;\nsetp.ne.u32 p, %1, 0;\n@p mov.f64 %0, 0d7FF0000000000000;\n}" : "+d"(chi) : "r"(outside ? 1u : 0u));
                if constexpr (PROF) { const long long w1 = clock64() + (long long)(chi == -1.0); p_w[1] += w1 - w0; w0 = w1; }
                bool accept = chi < chi_accept;
                if (!accept && !(chi > chi_reject)) {        // near a tie (or NaN: both comparisons false): emcee's test in full precision
                    saw_nan = saw_nan || (chi != chi);
                    const double newlnp = chi * -0.5;
                    const double zlog = __dmul_rn(dim_m1, log(z));
                    accept = __dsub_rn(__dadd_rn(zlog, newlnp), oldlnp) > log(q_lds_f64(draw_a + ring_span + ring_off));
                }
                if (accept) {                                // C/ChiSquare.c:212; outside the box: -inf (mda.py:1576), never accepted
#pragma unroll
                    for (int l = 0; l < L; ++l) q_sts_f64(xrow + 8u * l, q[l]);
                    q_sts_f64(lnp_a + (uint32_t)(s_base + me) * 8u, chi * -0.5);
                    acc += 1u;
                    if constexpr (kPadGather) {              // and the copy the partners read
                        const uint32_t prow = pad_a + (uint32_t)(s_base + me) * (uint32_t)((L + 1) * 8);
#pragma unroll
                        for (int l = 0; l < L; ++l) q_sts_f64(prow + 8u * l, q[l]);
                    }
                    if constexpr (XREG) {                    // the register copy follows
#pragma unroll
                        for (int l = 0; l < L; ++l) x[l] = q[l];
                        oldlnp = chi * -0.5;
                    }
                }
                // the next half-step's operands: its draw (in place since the last barrier; of this piece or of the next one) and
                // my walker of the other half -- not the next piece's, which is fetched when that piece starts.  (Fetched here, not
                // under the chi^2 sum: behind the other warp group's gather in the scheduler's load queue they would hold this warp
                // at their issue -- measured 4 % slower.)
                if (t + 1 < t_end || has_next) { z = q_lds_f64(draw_a + next_off); pf = q_lds_u2(draw_a + 2u * ring_span + next_off); }
                if constexpr (!XREG) {
                    if (t + 1 < t_end) {
#pragma unroll
                        for (int l = 0; l < L; ++l) x[l] = q_lds_f64(nrow + 8u * l);
                        oldlnp = q_lds_f64(lnp_a + (uint32_t)(c_base + me) * 8u);
                    }
                }
            }
            ring_off = next_off;
            ring = ring == 2 ? 0 : ring + 1;
            if (kept_step) q_fence_proxy_async();            // my rows before the bulk store reads them
            if constexpr (PROF) { const long long w1 = clock64(); p_w[2] += w1 - w0; w0 = w1; }
            __syncthreads();                                 // (B)
            if constexpr (PROF) { const long long w1 = clock64(); p_w[3] += w1 - w0; p_w[4] += 1; }
            if (kept_step && !split && half) qws_piece_barrier<COPYWG>();   // (C)
            if (half) {
                if (--until_keep == 0) { until_keep = args.thin; if (keeping) ++slot; }
            }
        };
        for (uint32_t t = t_begin; t < t_end; t += 2) {      // t_begin is even: a step is the update of half 0, then of half 1
            half_step(0, t, x0, lnp0, acc0);
            if constexpr (XREG) half_step(1, t + 1, x1, lnp1, acc1);
            else half_step(1, t + 1, x0, lnp0, acc1);
        }
        if (ustamp) ustamp[3] = clock64();
        q_fence_proxy_async();
        // (D) my last writes are fenced.  With a copy warpgroup the walker warps only say so and go on to the next piece (its
        // buffer is the other one; the copy thread, which waits here, stores this one): they met the copy thread ~700 cycles
        // late at every switch.
        if constexpr (COPYWG) asm volatile("bar.arrive 2, 384;" ::: "memory");
        else qws_piece_barrier<COPYWG>();
        if (owner) {                                         // the bin is mine alone: plain adds; issued after the barrier, so that no fence
            unsigned int *gAcc = args.nacc + (size_t)bin * W;   // of this piece waits for them (the next barrier orders them before a release)
            if (acc0) atomicAdd(gAcc + tid, acc0);
            if (acc1) atomicAdd(gAcc + H + tid, acc1);
        }
        if (ustamp) ustamp[4] = clock64();
        if constexpr (PROF) if (tid == 0 && args.prof) args.prof[(size_t)blockIdx.x * 128 + 117] = (long long)q_globaltimer_ns();   // (last) piece done
        if constexpr (PROF) if ((tid == 0 || tid == 255) && args.prof) for (int k = 0; k < 5; ++k) atomicAdd((unsigned long long *)args.prof + (size_t)blockIdx.x * 128 + (tid ? 8 : 0) + k, (unsigned long long)p_w[k]);
        if (!has_next) break;
        unit = next;
        buf ^= 1;
    }
    if (saw_nan) atomicOr(args.status, kStatusNaN);
}

using EnsKernel = void (*)(const EnsembleArgs);

template <int L>
EnsKernel pick_for_l(bool treg) {
    if constexpr (quad_padded_doubles(L) <= 60) return treg ? ensemble_qws_kernel<L, true> : ensemble_qws_kernel<L, false>;
    else return ensemble_qws_kernel<L, false>;
}

template <int... Ls>
EnsKernel pick_from(int n_l, bool treg, std::integer_sequence<int, Ls...>) {
    EnsKernel k = nullptr;
    ((n_l == Ls + 1 ? (k = pick_for_l<Ls + 1>(treg), 0) : 0), ...);
    return k;
}

EnsKernel pick_qws_kernel(int n_l, bool treg) { return pick_from(n_l, treg, std::make_integer_sequence<int, kMaxRegL>{}); }

}  // namespace

bool choose_ensemble_qws_shape(int bins, int walkers, int n_l, int sm_count, size_t smem_optin, EnsembleShape *out) {
    if (n_l < 1 || n_l > kMaxRegL || walkers < 2 || (walkers & 1) || walkers / 2 > kWalkerThreads || bins < 1) return false;
    EnsembleShape s{};
    s.threads = kThreads;                                   // (kThreadsLone with T in registers, set below)
    s.wpt = 1;
    s.groups = 1;
    s.qws = true;
    s.quad = true;
    // T in registers for a CTA that has its SM to itself (setmaxnreg gives the walker warps 208 registers) wherever T fits:
    // one CTA per SM carrying bins / SMs bins' worth of steps, one after the other, at 1.51 us per bin-step (512 x 9) beats two
    // CTAs per SM with T in shared memory (3.80 us per pair) and ensemble_quad_kernel's four (6.54 us per four) at any
    // number of bins
    (void)bins; (void)sm_count;
    s.treg = quad_padded_doubles(n_l) <= 60;
    const char *force_treg = getenv("MDA_ENS_TREG");
    if (force_treg && (force_treg[0] == '0' || force_treg[0] == '1')) s.treg = quad_padded_doubles(n_l) <= 60 && force_treg[0] == '1';
    if (s.treg) s.threads = kThreadsLone;
    s.smem_bytes = QwsSmem(walkers, n_l, !s.treg).total;
    if (s.smem_bytes > smem_optin) return false;
    s.fits = true;
    *out = s;
    return true;
}

cudaError_t configure_ensemble_qws_kernels(size_t smem_optin) {
    for (int l = 1; l <= kMaxRegL; ++l)
        for (int treg = 0; treg < 2; ++treg) {
            const void *k = reinterpret_cast<const void *>(pick_qws_kernel(l, treg != 0));
            cudaFuncAttributes fa;
            cudaError_t e = cudaFuncGetAttributes(&fa, k);
            if (e != cudaSuccess) return e;
            e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(smem_optin - fa.sharedSizeBytes));
            if (e != cudaSuccess) return e;
        }
    for (const void *k : {reinterpret_cast<const void *>(ensemble_qws_kernel<9, true, true>), reinterpret_cast<const void *>(ensemble_qws_kernel<9, false, true>)}) {
        cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(smem_optin - 64));
        if (e != cudaSuccess) return e;
    }
    return cudaSuccess;
}

// Co-resident CTAs on the whole device (a grid smaller than the number of bins is launched cooperatively).
int ensemble_qws_capacity(int n_l, const EnsembleShape &shape, int sm_count) {
    EnsKernel k = pick_qws_kernel(n_l, shape.treg);
    int per_sm = 0;
    if (!k || cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, reinterpret_cast<const void *>(k), shape.threads, shape.smem_bytes) != cudaSuccess)
        return 0;
    return per_sm * sm_count;
}

cudaError_t launch_ensemble_qws(const EnsembleArgs &args, const EnsembleShape &shape, int grid, cudaStream_t stream) {
    if (args.bins <= 0 || args.n_steps <= 0) return cudaSuccess;
    EnsKernel k = pick_qws_kernel(args.n_l, shape.treg);
    if (args.prof && args.n_l == 9) k = shape.treg ? ensemble_qws_kernel<9, true, true> : ensemble_qws_kernel<9, false, true>;
    if (!k || !shape.fits || !shape.qws || !args.quad || grid < 1 || grid > args.bins || (grid < args.bins && !args.progress)) return cudaErrorInvalidValue;
    EnsembleArgs copy = args;
    copy.groups = 0;
    // timing experiments (see the kernel; wrong chains): read by the PROFILING build only -- in the product build a test of the
    // word in every half-step cost a constant-bank load in front of the gather
    if (const char *dbg = getenv("MDA_QWS_DEBUG")) copy.groups = atoi(dbg);
    if (grid == args.bins) {                                 // a bin per CTA: no CTA waits for another
        k<<<grid, shape.threads, shape.smem_bytes, stream>>>(copy);
        return cudaGetLastError();
    }
    // bins cut across CTAs: the piece that waits needs the other one resident -- a cooperative launch fails instead of hanging
    void *params[] = {&copy};
    return cudaLaunchCooperativeKernel(reinterpret_cast<const void *>(k), dim3((unsigned)grid), dim3((unsigned)shape.threads), params, shape.smem_bytes, stream);
}

}  // namespace mda
