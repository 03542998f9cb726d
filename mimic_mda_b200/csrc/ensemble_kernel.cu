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
// Nothing crosses PCIe per step; kept samples go to an HBM chain [bin][kept][W][L].
#include "likelihood.cuh"
#include "philox.cuh"

namespace mda {

namespace {

constexpr int kStatusNaN = 1;          // a log-probability came out NaN (emcee raises on this)

template <int L, int WPT>
__global__ void __launch_bounds__(256) ensemble_kernel(const EnsembleArgs args) {
    constexpr int LP = L | 1;
    constexpr int ROW = L + 1;

    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ __align__(8) uint64_t mbar;

    const int tid = threadIdx.x;
    const int T = blockDim.x;
    const int bin = blockIdx.x;
    const int W = args.walkers;
    const int n_pad = args.n_pad;
    const int halfk = W >> 1;

    double2 *sC = reinterpret_cast<double2 *>(smem_raw);                       // [n_pad/2][ROW]
    double *sPos = reinterpret_cast<double *>(smem_raw) + (size_t)n_pad * ROW;   // [W][LP]
    double *sLnp = sPos + (size_t)W * LP;                                       // [W]
    unsigned int *sAcc = reinterpret_cast<unsigned int *>(sLnp + W);            // [W]

    if (tid == 0) {
        mbar_init(&mbar, 1);
        fence_mbar_init();
    }
    __syncthreads();
    if (tid == 0) {
        const uint32_t bytes = (uint32_t)n_pad * ROW * 8u;
        mbar_arrive_expect_tx(&mbar, bytes);
        bulk_g2s(sC, reinterpret_cast<const double2 *>(args.consts) + (size_t)bin * (n_pad / 2) * ROW, bytes, &mbar);
    }
    double *gPos = args.pos + (size_t)bin * W * L;
    double *gLnp = args.lnp + (size_t)bin * W;
    unsigned int *gAcc = args.nacc + (size_t)bin * W;
    for (int i = tid; i < W * L; i += T) {
        const int w = i / L;
        sPos[w * LP + (i - w * L)] = gPos[i];
    }
    for (int w = tid; w < W; w += T) {
        sLnp[w] = gLnp[w];
        sAcc[w] = gAcc[w];
    }
    const int n = args.n_pts[bin];
    const uint32_t bin_id = (uint32_t)args.bin_ids[bin];
    mbar_wait(&mbar, 0);
    __syncthreads();

    const double a_scale = args.a;
    const double dim_m1 = (double)(L - 1);
    bool saw_nan = false;

    for (int step = 0; step < args.n_steps; ++step) {
        const uint64_t t_step = (uint64_t)args.step0 + (uint64_t)step;
        for (int half = 0; half < 2; ++half) {
            const int s_base = half ? halfk : 0;            // active half
            const int c_base = half ? 0 : halfk;            // complementary half
            const int n_s = half ? W - halfk : halfk;
            const int n_c = W - n_s;
            const uint32_t t = (uint32_t)(2u * (uint32_t)t_step + (uint32_t)half);
            for (int w0 = 0; w0 < n_s; w0 += T * WPT) {
                double q[WPT][L];
                double zlog[WPT];
                bool outside[WPT], live[WPT];
#pragma unroll
                for (int k = 0; k < WPT; ++k) {
                    const int wl = w0 + k * T + tid;
                    live[k] = wl < n_s;
                    const int gw = s_base + min(wl, n_s - 1);
                    const Philox4 r = philox4x32_10(Philox4{(uint32_t)gw, bin_id, t, 0u}, args.seed_lo, args.seed_hi);
                    const double u = uniform53(r.x, r.y);
                    // emcee computes these with numpy, i.e. one rounding per operation: keep the
                    // compiler from contracting them into FMAs so positions match bit for bit
                    const double zr = __dadd_rn(__dmul_rn(a_scale - 1.0, u), 1.0);
                    const double z = __dmul_rn(zr, zr) / a_scale;      // ((a-1) u + 1)^2 / a
                    zlog[k] = __dmul_rn(dim_m1, log(z));
                    const int partner = c_base + (int)(((uint64_t)r.z * (uint64_t)n_c) >> 32);
                    const double *xs = sPos + (size_t)gw * LP;
                    const double *xc = sPos + (size_t)partner * LP;
                    outside[k] = false;
#pragma unroll
                    for (int l = 0; l < L; ++l) {
                        const double c = xc[l];
                        q[k][l] = __dsub_rn(c, __dmul_rn(z, __dsub_rn(c, xs[l])));   // q = c - z (c - x)
                        outside[k] = outside[k] || (q[k][l] < args.lo[l]) || (q[k][l] > args.hi[l]);
                    }
                }
                double chi[WPT];
#pragma unroll
                for (int k = 0; k < WPT; ++k) chi[k] = 0.0;
                accumulate_chi<L, WPT>(sC, n, q, chi);
#pragma unroll
                for (int k = 0; k < WPT; ++k) {
                    if (!live[k]) continue;
                    const int gw = s_base + w0 + k * T + tid;
                    const double newlnp = outside[k] ? -CUDART_INF : chi[k] * -0.5;
                    saw_nan = saw_nan || (newlnp != newlnp);
                    const Philox4 r2 = philox4x32_10(Philox4{(uint32_t)gw, bin_id, t, 1u}, args.seed_lo, args.seed_hi);
                    const double lnpdiff = __dsub_rn(__dadd_rn(zlog[k], newlnp), sLnp[gw]);
                    if (lnpdiff > log(uniform53(r2.x, r2.y))) {
                        double *xs = sPos + (size_t)gw * LP;
#pragma unroll
                        for (int l = 0; l < L; ++l) xs[l] = q[k][l];
                        sLnp[gw] = newlnp;
                        sAcc[gw] += 1u;
                    }
                }
            }
            __syncthreads();          // the updated half becomes the next half-step's partners
        }
        // keep this step?  (emcee: chain[:, i, :] = p after both halves, thinned)
        if (args.chain != nullptr && ((t_step + 1) % (uint64_t)args.thin) == 0) {
            const long long kept = args.keep0 + (long long)((t_step + 1) / (uint64_t)args.thin) -
                                   (long long)((uint64_t)args.step0 / (uint64_t)args.thin) - 1;
            if (kept < args.keep_cap) {
                double *gc = args.chain + (((size_t)bin * args.keep_cap + (size_t)kept) * W) * L;
                for (int i = tid; i < W * L; i += T) {
                    const int w = i / L;
                    gc[i] = sPos[w * LP + (i - w * L)];
                }
                double *gl = args.lnp_chain + ((size_t)bin * args.keep_cap + (size_t)kept) * W;
                for (int w = tid; w < W; w += T) gl[w] = sLnp[w];
            }
            __syncthreads();          // the copy reads every row; the next half-step rewrites rows in place
        }
    }

    for (int i = tid; i < W * L; i += T) {
        const int w = i / L;
        gPos[i] = sPos[w * LP + (i - w * L)];
    }
    for (int w = tid; w < W; w += T) {
        gLnp[w] = sLnp[w];
        gAcc[w] = sAcc[w];
    }
    if (saw_nan) atomicOr(args.status, kStatusNaN);
}

using EnsKernel = void (*)(const EnsembleArgs);

template <int L>
EnsKernel pick_ens_wpt(int wpt) {
    switch (wpt) {
        case 1: return ensemble_kernel<L, 1>;
        case 2: return ensemble_kernel<L, 2>;
        default: return nullptr;
    }
}

EnsKernel pick_ens_kernel(int n_l, int wpt) {
    switch (n_l) {
        case 1: return pick_ens_wpt<1>(wpt);
        case 2: return pick_ens_wpt<2>(wpt);
        case 3: return pick_ens_wpt<3>(wpt);
        case 4: return pick_ens_wpt<4>(wpt);
        case 5: return pick_ens_wpt<5>(wpt);
        case 6: return pick_ens_wpt<6>(wpt);
        case 7: return pick_ens_wpt<7>(wpt);
        case 8: return pick_ens_wpt<8>(wpt);
        case 9: return pick_ens_wpt<9>(wpt);
        case 10: return pick_ens_wpt<10>(wpt);
        case 11: return pick_ens_wpt<11>(wpt);
        case 12: return pick_ens_wpt<12>(wpt);
        case 13: return pick_ens_wpt<13>(wpt);
        case 14: return pick_ens_wpt<14>(wpt);
        case 15: return pick_ens_wpt<15>(wpt);
        case 16: return pick_ens_wpt<16>(wpt);
        default: return nullptr;
    }
}

}  // namespace

size_t ensemble_smem_bytes(int walkers, int n_l, int n_pad) {
    const size_t lp = (size_t)(n_l | 1);
    return (size_t)n_pad * (size_t)(n_l + 1) * 8 + (size_t)walkers * lp * 8 + (size_t)walkers * 8 +
           (size_t)walkers * 4 + 16;
}

EnsembleShape choose_ensemble_shape(int walkers, int n_l, int n_pad, size_t smem_optin) {
    EnsembleShape s{};
    const int half = (walkers + 1) / 2;
    s.wpt = half >= 128 ? 2 : 1;
    int threads = (half + s.wpt - 1) / s.wpt;
    threads = (threads + 31) / 32 * 32;
    if (threads > 256) threads = 256;
    if (threads < 32) threads = 32;
    s.threads = threads;
    s.smem_bytes = ensemble_smem_bytes(walkers, n_l, n_pad);
    s.fits = n_l >= 1 && n_l <= kMaxRegL && s.smem_bytes <= smem_optin &&
             (size_t)n_pad * (size_t)(n_l + 1) * 8 < (1u << 20);
    return s;
}

cudaError_t configure_ensemble_kernels(size_t smem_optin) {
    for (int l = 1; l <= kMaxRegL; ++l)
        for (int wpt = 1; wpt <= 2; ++wpt) {
            cudaError_t e = cudaFuncSetAttribute(reinterpret_cast<const void *>(pick_ens_kernel(l, wpt)),
                                                 cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_optin);
            if (e != cudaSuccess) return e;
        }
    return cudaSuccess;
}

cudaError_t launch_ensemble(const EnsembleArgs &args, const EnsembleShape &shape, cudaStream_t stream) {
    if (args.bins <= 0 || args.n_steps <= 0) return cudaSuccess;
    EnsKernel k = pick_ens_kernel(args.n_l, shape.wpt);
    if (!k || !shape.fits) return cudaErrorInvalidValue;
    k<<<args.bins, shape.threads, shape.smem_bytes, stream>>>(args);
    return cudaGetLastError();
}

}  // namespace mda
