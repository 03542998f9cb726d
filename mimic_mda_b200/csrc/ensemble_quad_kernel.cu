// ensemble_quad_kernel.cu -- device-resident ensemble stepping with chi^2 from the bin's triangular factor.
//
// Same job, same random streams and same decisions as the other steppers (the reference's host loop
// emcee.EnsembleSampler.run_mcmc -> ln_post_prob per walker, mda.py:1128-1131, 1540-1578, as ONE launch), with the
// likelihood evaluated from sufficient statistics instead of angle by angle:
//
//   chi^2(a) = sum_j (d_j - sum_l a_l D_lj)^2 (C/ChiSquare.c:188-210) = |M [a; -1]|^2,  M = [D^T | d]  (N x (L+1))
//            = |T [a; -1]|^2 with M = Q T, T upper triangular ((L+1) x (L+1), chisq_abi.cu: factor_bin, Householder
//              reflections in long double, rounded to double once per bin)
//
// i.e. (L+1)(L+2)/2 multiply-adds per evaluation (55 at L = 9) instead of N (L+1) (600 at N = 60), independent of the
// number of angles.  Nothing is inverted (rank-deficient bins are fine) and every row contributes a square; the only
// cancellation is inside a row (T a against the transformed data), the same as inside a residual of the reference.
// Results agree with the reference arithmetic to ~1e-14 relative (tested against the oracle at 1e-12).
//
// With the contraction this cheap a half-step is no longer fp64 bound, and the layout is the plain one: a CTA owns
// an Ex bin for the whole launch, thread i owns walker i of the half being updated -- it draws (Philox4x32-10,
// counter = (walker, bin id, half-step, 0)), gathers its partner's row from the other half, forms the proposal
// q = c - z (c - x) (numpy rounding), tests the box prior (mda.py:1574-1576), evaluates chi^2(q), decides
// ((L-1) ln z + lnp(q) - lnp(x) > ln u': single-precision estimate, emcee's full-precision test for near-ties) and
// overwrites its own row.  Rows of the updated half are written only by their owners and rows of the other half are
// only read, so ONE barrier per half-step is all the synchronisation there is.  The half just updated goes to the
// HBM chain by a bulk-TMA store right after that barrier; it is next written a whole half-step later, so the store's
// read of shared memory never delays anyone.  Several CTAs -- bins -- share an SM (41 KB of shared memory at
// 512 walkers x 9 parameters) and hide each other's latencies; what bounds the run is the chain: W (L+1) 8 bytes
// per bin and kept step to HBM.
#include "philox.cuh"
#include "quad_chi.cuh"

#include <math_constants.h>

#include <algorithm>
#include <utility>
#include <cstdlib>

namespace mda {

namespace {

constexpr int kStatusNaN = 1;
constexpr int kStatusStalled = 4;        // a unit waited 4 s (wall clock) for its predecessor: the launch gave up (a guard against a hung GPU)
constexpr int kQuadMaxThreads = 256;

struct QuadSmem {
    size_t pos, lnp, tri, acc, total;
    __host__ __device__ QuadSmem(int walkers, int n_l) {
        size_t off = 0;
        pos = off; off += (size_t)walkers * (size_t)n_l * 8;                  // [W][L] dense, updated in place; bulk-stored as is
        lnp = off; off += (size_t)walkers * 8;                                // [W]
        tri = off; off += (size_t)quad_padded_doubles(n_l) * 8;               // T, rows T[i][i..L] padded to 16 bytes
        acc = off; off += (size_t)walkers * 4;                                // [W] accepted proposals
        total = (off + 15) & ~(size_t)15;
    }
};

__device__ __forceinline__ uint32_t smem_addr(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void bulk_s2g(void *dst, const void *src, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(smem_addr(src)), "r"(bytes) : "memory");
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

__device__ __forceinline__ double uniform43(uint32_t w, uint32_t x, uint32_t y) {
    return ((double)w * 2048.0 + (double)(((x & 31u) << 6) | (y & 63u))) * (1.0 / 8796093022208.0);
}

// OCC: CTAs per SM the build is made for.  4: 64 registers.  3 and 1: the Philox draw of the next half-step is made ahead of
// the evaluation of this one (shorter half-steps; 78 registers) -- for up to three bins per SM; 1: no register cap at all
// (158 registers) for a CTA that has its SM to itself (up to one bin per SM, and the walking schedule).
template <int L, int OCC>
__global__ void __launch_bounds__(kQuadMaxThreads, OCC) ensemble_quad_kernel(const EnsembleArgs args) {
    constexpr bool kTriInRegs = OCC == 1 && quad_padded_doubles(L) <= 60;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int tid = threadIdx.x, T = blockDim.x;
    const int W = args.walkers, H = W >> 1;
    const QuadSmem lay(W, L);
    double *sPos = reinterpret_cast<double *>(smem_raw + lay.pos);
    double *sLnp = reinterpret_cast<double *>(smem_raw + lay.lnp);
    double *sTri = reinterpret_cast<double *>(smem_raw + lay.tri);
    unsigned int *sAcc = reinterpret_cast<unsigned int *>(smem_raw + lay.acc);

    const uint32_t seed0 = args.seed_lo, seed1 = args.seed_hi;
    const double a_scale = args.a;
    const double dim_m1 = (double)(L - 1);
    const float dim_m1f = (float)(L - 1);
    const bool keeping = args.chain != nullptr;
    // a half can be stored on its own if its rows and log-probabilities are 16-byte multiples (bulk-copy granularity)
    const bool split = ((H * L) & 1) == 0 && (H & 1) == 0;
    bool saw_nan = false;

    // Units of work are (chunk of steps, bin), chunk-major, unit u on CTA u mod grid (one chunk, grid = bins: a CTA keeps its
    // bin for the whole launch).  With fewer CTAs than bins every SM carries the same share of the run; a unit takes the
    // bin's state over from HBM/L2 behind a release/acquire progress word per bin, as in the DMMA steppers.
    __shared__ int s_abort;
    if (tid == 0) s_abort = 0;
    __syncthreads();
    const long long n_units = (long long)args.n_chunks * args.bins;
    for (long long unit = blockIdx.x; unit < n_units; unit += gridDim.x) {
        const int chunk = (int)(unit / args.bins);
        const int bin = (int)(unit - (long long)chunk * args.bins);
        const long long step_first = args.step0 + (long long)chunk * args.chunk_steps;
        const int steps_here = min(args.chunk_steps, args.n_steps - chunk * args.chunk_steps);
        double *gPos = args.pos + (size_t)bin * W * L;
        double *gLnp = args.lnp + (size_t)bin * W;
        unsigned int *gAcc = args.nacc + (size_t)bin * W;
        if (tid == 0 && chunk > 0) {
            const unsigned long long want = args.launch_serial + (unsigned long long)chunk;
            const unsigned long long t_start = globaltimer_ns();
            while (ld_acquire_u64(args.progress + bin) != want) {
                __nanosleep(64);
                if (globaltimer_ns() - t_start > kHandOverTimeoutNs) { s_abort = 1; break; }     // 4 s of wall clock: the predecessor is not coming
            }
        }
        __syncthreads();
        if (s_abort) {                                         // every CTA that depends on this bin gives up the same way
            if (tid == 0) atomicOr(args.status, kStatusStalled);
            break;
        }
        for (int i = tid; i < W * L; i += T) sPos[i] = __ldcg(gPos + i);
        for (int w = tid; w < W; w += T) { sLnp[w] = __ldcg(gLnp + w); sAcc[w] = __ldcg(gAcc + w); }
        load_factor<L>(sTri, args.quad + (size_t)bin * quad_stride(L), tid, T);
        fence_proxy_async_smem();
        __syncthreads();

        // the lone-CTA build keeps T in registers for the unit (up to 60 doubles): no shared-memory reads in the contraction
        double tri_regs[kTriInRegs ? quad_padded_doubles(L) : 1];
        if constexpr (kTriInRegs) {
#pragma unroll
            for (int k = 0; k < quad_padded_doubles(L); ++k) tri_regs[k] = sTri[k];
        }
        const uint32_t bin_id = (uint32_t)args.bin_ids[bin];
        int until_keep = args.thin - (int)((uint64_t)step_first % (uint64_t)args.thin);
        long long slot = args.keep0 + (step_first / args.thin - args.step0 / args.thin);
        const uint32_t t_begin = 2u * (uint32_t)step_first, t_end = t_begin + 2u * (uint32_t)steps_here;

        auto draw = [&](uint32_t tt, int i) {
            return philox4x32_10_keys(Philox4{(uint32_t)(((tt & 1u) ? H : 0) + i), bin_id, tt, 0u}, args.round_keys);
        };
        const bool one_each = OCC != 4 && H <= T;                   // one walker per thread (every BASELINE.json shape)
        const int my_i = min(tid, H - 1);
        Philox4 r_cur = draw(t_begin, my_i);
        for (uint32_t t = t_begin; t < t_end; ++t) {
            const uint32_t half = t & 1u;
            const int s_base = half ? H : 0, c_base = half ? 0 : H;            // updated half; partners come from the other one
            auto update = [&](int i, const Philox4 &r) {
                const int gw = s_base + i;
                // numpy arithmetic is one rounding per operation: no FMA contraction in the proposal
                const double zr = __dadd_rn(__dmul_rn(a_scale - 1.0, uniform53(r.x, r.y)), 1.0);
                const double z2 = __dmul_rn(zr, zr);
                const double z = (a_scale == 2.0) ? z2 * 0.5 : z2 / a_scale;   // ((a-1) u + 1)^2 / a
                const double ua = uniform43(r.w, r.x, r.y);
                const double *crow = sPos + (size_t)(c_base + (int)__umulhi(r.z, (uint32_t)H)) * L;
                double *xrow = sPos + (size_t)gw * L;
                double q[L];
                bool outside = false;
#pragma unroll
                for (int l = 0; l < L; ++l) {
                    const double c = crow[l];
                    q[l] = __dsub_rn(c, __dmul_rn(z, __dsub_rn(c, xrow[l])));
                    outside = outside | (q[l] < args.lo[l]) | (q[l] > args.hi[l]);   // mda.py:1575; a NaN compares false, as there
                }
                double chi;
                if constexpr (kTriInRegs) chi = outside ? CUDART_INF : chi_from_factor_regs<L>(tri_regs, q);
                else chi = outside ? CUDART_INF : chi_from_factor<L>(sTri, q);
                const double newlnp = chi * -0.5;                              // C/ChiSquare.c:212; outside the box: -inf (mda.py:1576)
                saw_nan = saw_nan || (newlnp != newlnp);
                const double oldlnp = sLnp[gw];
                // single-precision estimate of (L-1) ln z - ln u'; decides all but near-ties
                const float fast = dim_m1f * __logf((float)z) - __logf((float)ua);
                const double est = __dsub_rn(newlnp, oldlnp) + (double)fast;
                bool accept;
                if (fabs(est) > 1e-4) {
                    accept = est > 0.0;                        // far from a tie: the float logs (error < 2e-5) cannot change the sign
                } else {                                       // emcee's test in full precision
                    const double zlog = __dmul_rn(dim_m1, log(z));
                    accept = __dsub_rn(__dadd_rn(zlog, newlnp), oldlnp) > log(ua);
                }
                if (accept) {
#pragma unroll
                    for (int l = 0; l < L; ++l) xrow[l] = q[l];
                    sLnp[gw] = newlnp;
                    sAcc[gw] += 1u;
                }
            };
            if (one_each) {
                // the draw of the next half-step does not depend on the state: issued here, it overlaps the evaluation below
                const Philox4 r_next = t + 1 < t_end ? draw(t + 1, my_i) : r_cur;
                if (tid < H) update(tid, r_cur);
                r_cur = r_next;
            } else {
                for (int i = tid; i < H; i += T) update(i, draw(t, i));
            }
            const bool kept_step = keeping && until_keep == 1 && slot < args.keep_cap;
            if (kept_step) fence_proxy_async_smem();           // my rows before the bulk store reads them
            // The last store (if any) read the half that is written next, a whole half-step after it was issued: no wait in practice.
            if (keeping && tid == 0) bulk_wait_read_all();
            __syncthreads();                                   // the half is complete
            if (kept_step && tid == 0) {
                double *dst = args.chain + (((size_t)slot * args.bins + (size_t)bin) * W) * L;
                double *ldst = args.lnp_chain + ((size_t)slot * args.bins + (size_t)bin) * W;
                if (split) {                                   // the half just updated is final for this step
                    bulk_s2g(dst + (size_t)s_base * L, sPos + (size_t)s_base * L, (uint32_t)H * L * 8u);
                    bulk_s2g(ldst + s_base, sLnp + s_base, (uint32_t)H * 8u);
                    bulk_commit();
                } else if (half) {
                    bulk_s2g(dst, sPos, (uint32_t)W * L * 8u);
                    bulk_s2g(ldst, sLnp, (uint32_t)W * 8u);
                    bulk_commit();
                }
            }
            if (kept_step && !split && half) {                 // whole-state store: nobody may write until it has been read
                if (tid == 0) bulk_wait_read_all();
                __syncthreads();
            }
            if (half) {
                if (--until_keep == 0) { until_keep = args.thin; if (keeping) ++slot; }
            }
        }
        // ---- hand the bin back
        if (tid == 0) bulk_wait_read_all();
        __syncthreads();
        for (int i = tid; i < W * L; i += T) gPos[i] = sPos[i];
        for (int w = tid; w < W; w += T) { gLnp[w] = sLnp[w]; gAcc[w] = sAcc[w]; }
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

// every numL 1 .. kMaxRegL x the three occupancy builds (1, 3 or 4 CTAs per SM)
template <int L>
EnsKernel pick_occ(int occ) {
    return occ == 1 ? ensemble_quad_kernel<L, 1> : (occ == 3 ? ensemble_quad_kernel<L, 3> : ensemble_quad_kernel<L, 4>);
}

template <int... Ls>
EnsKernel pick_from(int n_l, int occ, std::integer_sequence<int, Ls...>) {
    EnsKernel k = nullptr;
    ((n_l == Ls + 1 ? (k = pick_occ<Ls + 1>(occ), 0) : 0), ...);
    return k;
}

EnsKernel pick_quad_kernel(int n_l, int occ) { return pick_from(n_l, occ, std::make_integer_sequence<int, kMaxRegL>{}); }

}  // namespace

bool choose_ensemble_quad_shape(int walkers, int n_l, size_t smem_optin, EnsembleShape *out) {
    if (n_l < 1 || n_l > kMaxRegL || walkers < 2 || (walkers & 1)) return false;
    EnsembleShape s{};
    s.smem_bytes = QuadSmem(walkers, n_l).total;
    if (s.smem_bytes > smem_optin) return false;
    const int half = walkers / 2;
    s.threads = std::min(kQuadMaxThreads, ((half + 31) / 32) * 32);
    s.wpt = (half + s.threads - 1) / s.threads;
    s.groups = 1;
    s.fits = true;
    s.quad = true;
    *out = s;
    return true;
}

cudaError_t configure_ensemble_quad_kernels(size_t smem_optin) {
    for (int l = 1; l <= kMaxRegL; ++l)
        for (int occ : {1, 3, 4}) {
            const void *k = reinterpret_cast<const void *>(pick_quad_kernel(l, occ));
            cudaFuncAttributes fa;
            cudaError_t e = cudaFuncGetAttributes(&fa, k);
            if (e != cudaSuccess) return e;
            e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(smem_optin - fa.sharedSizeBytes));
            if (e != cudaSuccess) return e;
        }
    return cudaSuccess;
}

// The build for `grid` CTAs.
static int quad_occ_for(int grid, int sm_count) {
    const char *force = getenv("MDA_ENS_QUAD_OCC");                          // experiments: 1 | 3 | 4
    if (force && (force[0] == '1' || force[0] == '3' || force[0] == '4')) return force[0] - '0';
    return grid <= sm_count ? 1 : (grid <= 3 * sm_count ? 3 : 4);
}

// Co-resident CTAs on the whole device (a walked unit list needs every CTA of the grid resident).
int ensemble_quad_capacity(int n_l, const EnsembleShape &shape, int grid, int sm_count) {
    EnsKernel k = pick_quad_kernel(n_l, quad_occ_for(grid, sm_count));
    int per_sm = 0;
    if (!k || cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, reinterpret_cast<const void *>(k), shape.threads, shape.smem_bytes) != cudaSuccess)
        return 0;
    return per_sm * sm_count;
}

cudaError_t launch_ensemble_quad(const EnsembleArgs &args, const EnsembleShape &shape, int grid, int sm_count, cudaStream_t stream) {
    if (args.bins <= 0 || args.n_steps <= 0) return cudaSuccess;
    EnsKernel k = pick_quad_kernel(args.n_l, quad_occ_for(grid, sm_count));
    if (!k || !shape.fits || !shape.quad || !args.quad || grid < 1 || args.chunk_steps < 1 || args.n_chunks < 1 ||
        (args.n_chunks > 1 && !args.progress))
        return cudaErrorInvalidValue;
    return launch_maybe_cooperative(k, args, args.n_chunks > 1, grid, shape.threads, shape.smem_bytes, stream);
}

}  // namespace mda
