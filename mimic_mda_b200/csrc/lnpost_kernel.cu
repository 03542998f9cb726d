// lnpost_kernel.cu -- the batched chi-square log-posterior kernel for sm_100a.
//
// Replaces, for every (Ex bin, walker) pair of one launch, the reference's
//   ln_post_prob            mda.py:1540-1578      (flat box prior -> -inf)
//   calculateLnLiklihood    C/ChiSquare.c:176-213 (r_j = d_j - sum_l a_l D_lj; -chi^2/2)
//   calculateChi            C/ChiSquare.c:114-172
//
// Work decomposition (fp64 DFMA-pipe bound):
//   * one CTA = one bin x one tile of walkers; a thread owns WPT walkers and keeps
//     their L parameters in registers (L is a template parameter), so the inner loop
//     is pure DFMA fed by broadcast LDS.128 (one 16-byte shared load feeds 2*WPT DFMAs);
//   * the bin's constants live in HBM angle-pair-major: consts[bin][n_pad/2][1+L][2]
//     (for an angle pair: data pair, then dist_0 pair ... dist_{L-1} pair).  That block
//     (or a chunk of it when it does not fit) is brought into shared memory by ONE
//     bulk-TMA copy (cp.async.bulk ... mbarrier::complete_tx) issued by one thread;
//   * the tile's walker parameters [tile][L] are gathered with coalesced 8-byte
//     cp.async into rows padded to an odd stride, so the per-thread row reads that
//     follow are bank-conflict free for every L;
//   * per walker the operation order is exactly the reference's: residual updates in
//     ascending l, squares accumulated in ascending j from 0, one multiply by -1/2.
//     No cross-lane reduction exists, so results are independent of the tiling.
//
// Tensor cores are deliberately not used: the contraction is K = L <= 16 deep in fp64.
#include "likelihood.cuh"

#include <algorithm>
#include <cstdlib>
#include <utility>

namespace mda {

namespace {

// Largest chunk of staged constants a single bulk copy may carry (tx-count limit is
// 2^20-1 bytes per mbarrier phase); chunks are always far below this.
constexpr uint32_t kMaxBulkBytes = 1u << 19;

// ---- register-tiled kernel: L in registers --------------------------------------
template <int L, int WPT>
__global__ void __launch_bounds__(256) lnpost_tile_kernel(const BatchArgs args) {
    constexpr int LP = L | 1;          // odd row stride (doubles) of the staged parameters
    constexpr int ROW = L + 1;         // double2 entries per angle pair: data, dist_0..dist_{L-1}

    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ __align__(8) uint64_t mbar;

    const int tid = threadIdx.x;
    const int T = blockDim.x;
    const int tile_w = T * WPT;
    const int bin = blockIdx.x / args.tiles_per_bin;
    const int tile = blockIdx.x - bin * args.tiles_per_bin;
    const int w0 = tile * tile_w;
    const int nw = min(tile_w, args.walkers - w0);
    const int chunk = args.chunk;                    // angle columns per staged chunk (even)
    const int n_pad = args.n_pad;

    double2 *sC = reinterpret_cast<double2 *>(smem_raw);                    // [chunk/2][ROW]
    double *sP = reinterpret_cast<double *>(smem_raw) + (size_t)chunk * ROW;  // [tile_w][LP]

    const double2 *gC = reinterpret_cast<const double2 *>(args.consts) + (size_t)bin * (n_pad / 2) * ROW;

    if (tid == 0) {
        mbar_init(&mbar, 1);
        fence_mbar_init();
    }
    __syncthreads();

    // The tile's walker parameters are one contiguous block [nw][L] of the caller's array: ONE bulk copy brings it in
    // (16-byte aligned and sized whenever the block starts on an even double and holds an even number of them -- always
    // for even L or even tiles; args.params_bulk says the array itself is 16-byte aligned).  Otherwise: coalesced 8-byte
    // cp.async into rows padded to an odd stride.  Both land on the same mbarrier phase as the first chunk of constants.
    const bool bulk_p = args.params_bulk && ((((size_t)bin * args.walkers + w0) * L) & 1) == 0 && ((nw * L) & 1) == 0;
    const int lp = bulk_p ? L : LP;                  // row stride of the staged parameters (doubles)
    const int first_cols = min(chunk, n_pad);
    const double *gP = args.params + ((size_t)bin * args.walkers + w0) * L;
    if (tid == 0 && (first_cols > 0 || bulk_p)) {
        const uint32_t c_bytes = (uint32_t)first_cols * ROW * 8u, p_bytes = bulk_p ? (uint32_t)nw * L * 8u : 0u;
        mbar_arrive_expect_tx(&mbar, c_bytes + p_bytes);
        if (c_bytes) bulk_g2s(sC, gC, c_bytes, &mbar);
        if (p_bytes) bulk_g2s(sP, gP, p_bytes, &mbar);
    }
    if (!bulk_p) {
        const int count = nw * L;
        for (int i = tid; i < count; i += T) {
            const int w = i / L;
            const int l = i - w * L;
            cp_async_8(&sP[w * LP + l], gP + i);
        }
    }
    const int n = args.n_pts[bin];                   // latency overlaps the copies above
    if (!bulk_p) {
        cp_async_commit_wait_all();
        __syncthreads();
    }
    uint32_t parity = 0;
    if (first_cols > 0 || bulk_p) {                  // parameters (bulk path) and the first chunk of constants are in
        mbar_wait(&mbar, 0);
        parity = 1u;
    }

    double a[WPT][L];
    bool outside[WPT];
#pragma unroll
    for (int k = 0; k < WPT; ++k) {
        const int w = min(tid + k * T, nw - 1);      // clamp: idle slots repeat the last walker
        outside[k] = false;
        if ((L & 1) == 0 && bulk_p) {                // dense rows of an even L are 16-byte aligned: 128-bit reads
            const double2 *row = reinterpret_cast<const double2 *>(sP + (size_t)w * L);
#pragma unroll
            for (int l2 = 0; l2 < L / 2; ++l2) {
                const double2 v = row[l2];
                a[k][2 * l2] = v.x;
                a[k][2 * l2 + 1] = v.y;
            }
        } else {
#pragma unroll
            for (int l = 0; l < L; ++l) a[k][l] = sP[w * lp + l];
        }
#pragma unroll
        for (int l = 0; l < L; ++l)                  // mda.py:1575 -- "params[i] < lo or params[i] > hi"; NaN compares false
            outside[k] = outside[k] || (a[k][l] < args.lo[l]) || (a[k][l] > args.hi[l]);
    }

    double chi[WPT];
#pragma unroll
    for (int k = 0; k < WPT; ++k) chi[k] = 0.0;

    for (int c0 = 0; c0 < n; c0 += chunk) {
        if (c0 > 0) {
            __syncthreads();                          // every thread is done with the old chunk
            if (tid == 0) {
                const uint32_t bytes = (uint32_t)min(chunk, n_pad - c0) * ROW * 8u;
                mbar_arrive_expect_tx(&mbar, bytes);
                bulk_g2s(sC, gC + (size_t)(c0 / 2) * ROW, bytes, &mbar);
            }
            mbar_wait(&mbar, parity);
            parity ^= 1u;
        }
        accumulate_chi<L, WPT>(sC, min(chunk, n - c0), a, chi);
    }

    double *out = args.out + (size_t)bin * args.walkers + w0;
#pragma unroll
    for (int k = 0; k < WPT; ++k) {
        const int w = tid + k * T;
        if (w < nw) out[w] = finish(chi[k], outside[k], args.mode);
    }
}

// ---- generic kernel: any L, parameters stay in shared memory ---------------------
__global__ void __launch_bounds__(256) lnpost_generic_kernel(const BatchArgs args, const double *__restrict__ lo,
                                                            const double *__restrict__ hi) {
    const int L = args.n_l;
    const int LP = L | 1;
    const int ROW = L + 1;

    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ __align__(8) uint64_t mbar;

    const int tid = threadIdx.x;
    const int T = blockDim.x;
    const int bin = blockIdx.x / args.tiles_per_bin;
    const int tile = blockIdx.x - bin * args.tiles_per_bin;
    const int w0 = tile * T;
    const int nw = min(T, args.walkers - w0);
    const int chunk = args.chunk;
    const int n_pad = args.n_pad;

    double2 *sC = reinterpret_cast<double2 *>(smem_raw);
    double *sP = reinterpret_cast<double *>(smem_raw) + (size_t)chunk * ROW;
    const double2 *gC = reinterpret_cast<const double2 *>(args.consts) + (size_t)bin * (n_pad / 2) * ROW;

    if (tid == 0) {
        mbar_init(&mbar, 1);
        fence_mbar_init();
    }
    __syncthreads();
    const int first_cols = min(chunk, n_pad);
    if (tid == 0 && first_cols > 0) {
        const uint32_t bytes = (uint32_t)first_cols * ROW * 8u;
        mbar_arrive_expect_tx(&mbar, bytes);
        bulk_g2s(sC, gC, bytes, &mbar);
    }
    {
        const double *gP = args.params + ((size_t)bin * args.walkers + w0) * L;
        const int count = nw * L;
        for (int i = tid; i < count; i += T) {
            const int w = i / L;
            const int l = i - w * L;
            cp_async_8(&sP[w * LP + l], gP + i);
        }
    }
    const int n = args.n_pts[bin];
    cp_async_commit_wait_all();
    __syncthreads();

    const double *my = sP + (size_t)min(tid, nw - 1) * LP;
    bool outside = false;
    for (int l = 0; l < L; ++l) outside = outside || (my[l] < lo[l]) || (my[l] > hi[l]);

    double chi = 0.0;
    uint32_t parity = 0;
    for (int c0 = 0; c0 < n; c0 += chunk) {
        if (c0 > 0) {
            __syncthreads();
            if (tid == 0) {
                const uint32_t bytes = (uint32_t)min(chunk, n_pad - c0) * ROW * 8u;
                mbar_arrive_expect_tx(&mbar, bytes);
                bulk_g2s(sC, gC + (size_t)(c0 / 2) * ROW, bytes, &mbar);
            }
        }
        mbar_wait(&mbar, parity);
        parity ^= 1u;
        const int valid = min(chunk, n - c0);
        const int pairs = valid >> 1;
        const double2 *p = sC;
        for (int j2 = 0; j2 < pairs; ++j2, p += ROW) {
            double r0 = p[0].x, r1 = p[0].y;
            for (int l = 0; l < L; ++l) {
                const double al = my[l];
                const double2 D = p[1 + l];
                r0 = fma(-al, D.x, r0);
                r1 = fma(-al, D.y, r1);
            }
            chi = fma(r0, r0, chi);
            chi = fma(r1, r1, chi);
        }
        if (valid & 1) {
            double r0 = p[0].x;
            for (int l = 0; l < L; ++l) r0 = fma(-my[l], p[1 + l].x, r0);
            chi = fma(r0, r0, chi);
        }
    }
    if (n <= 0 && first_cols > 0) mbar_wait(&mbar, 0);
    if (tid < nw) args.out[(size_t)bin * args.walkers + w0 + tid] = finish(chi, outside, args.mode);
}

// ---- dispatch ---------------------------------------------------------------------
using TileKernel = void (*)(const BatchArgs);

template <int L>
TileKernel pick_wpt(int wpt) {
    switch (wpt) {
        case 1: return lnpost_tile_kernel<L, 1>;
        case 2: return lnpost_tile_kernel<L, 2>;
        case 4: return lnpost_tile_kernel<L, 4>;
        default: return nullptr;
    }
}

template <int... Ls>
TileKernel pick_from(int n_l, int wpt, std::integer_sequence<int, Ls...>) {
    TileKernel k = nullptr;
    ((n_l == Ls + 1 ? (k = pick_wpt<Ls + 1>(wpt), 0) : 0), ...);
    return k;
}

TileKernel pick_kernel(int n_l, int wpt) { return pick_from(n_l, wpt, std::make_integer_sequence<int, kMaxRegL>{}); }

inline int round_down_even(long v) { return (int)(v & ~1L); }


}  // namespace

TileShape choose_tile(int bins, int walkers, int n_l, int n_pad, int sm_count, size_t smem_optin,
                      int force_threads, int force_wpt) {
    TileShape t{};
    int threads = force_threads, wpt = force_wpt;
    const bool generic = n_l > kMaxRegL;
    if (generic) wpt = 1;
    if (threads <= 0 || wpt <= 0) {
        // Heuristic: the largest register tile that still leaves every SM a few CTAs
        // to overlap staging with arithmetic; small launches fall back to small tiles
        // so the work spreads over all 148 SMs.
        // Measured on B200 (profiles/r01_sweep_*): every register tile with WPT >= 2 performs alike
        // once there are >= 2 CTAs per SM; WPT 4 wins ~4 % on launches of many waves (one shared
        // load feeds 8 DFMAs) as long as its 4 L parameter registers do not cut occupancy.
        static const int cand[][2] = {{64, 4}, {128, 2}, {64, 2}, {64, 1}, {32, 1}};
        const int n_cand = 5;
        const int walkers32 = (walkers + 31) / 32 * 32;
        int pick = n_cand - 1;
        for (int c = 0; c < n_cand; ++c) {
            const int tw = cand[c][0] * (generic ? 1 : cand[c][1]);
            if (tw > walkers32 && c < n_cand - 1) continue;
            const long tiles = (long)bins * ((walkers + tw - 1) / tw);
            const long need = (cand[c][1] == 4) ? 8L * sm_count : 2L * sm_count;
            if (cand[c][1] == 4 && (n_l > 10 || generic)) continue;
            if (tiles >= need || c == n_cand - 1) { pick = c; break; }
        }
        if (threads <= 0) threads = cand[pick][0];
        if (wpt <= 0) wpt = generic ? 1 : cand[pick][1];
    }
    const int tile_w = threads * wpt;
    const int lp = n_l | 1;
    const size_t params_bytes = (size_t)tile_w * lp * 8;
    const size_t row_bytes = (size_t)(n_l + 1) * 8;     // per angle column
    // keep the staged constants <= 96 KB so at least two CTAs fit on an SM
    size_t budget = 96u * 1024u;
    if (params_bytes + budget + 64 > smem_optin) budget = smem_optin > params_bytes + 64 ? smem_optin - params_bytes - 64 : 0;
    int chunk = round_down_even((long)(budget / row_bytes));
    if (chunk > n_pad) chunk = n_pad;
    if (chunk < 2) chunk = 2;
    t.threads = threads;
    t.wpt = wpt;
    t.chunk = chunk;
    t.grid = bins * ((walkers + tile_w - 1) / tile_w);
    t.smem_bytes = (size_t)chunk * row_bytes + params_bytes;
    return t;
}

cudaError_t configure_kernels(size_t smem_optin) {
    static const int wpts[3] = {1, 2, 4};
    for (int l = 1; l <= kMaxRegL; ++l)
        for (int i = 0; i < 3; ++i) {
            cudaError_t e = cudaFuncSetAttribute(reinterpret_cast<const void *>(pick_kernel(l, wpts[i])),
                                                 cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_optin);
            if (e != cudaSuccess) return e;

        }
    return cudaFuncSetAttribute(reinterpret_cast<const void *>(lnpost_generic_kernel),
                                cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_optin);
}

cudaError_t launch_batch(BatchArgs &args, const TileShape &tile, cudaStream_t stream, const double *lo_dev,
                         const double *hi_dev) {
    if (args.bins <= 0 || args.walkers <= 0) return cudaSuccess;
    const int tile_w = tile.threads * tile.wpt;
    args.tiles_per_bin = (args.walkers + tile_w - 1) / tile_w;
    args.chunk = tile.chunk;
    if ((size_t)tile.chunk * (args.n_l + 1) * 8 > kMaxBulkBytes) return cudaErrorInvalidValue;
    if (args.n_l <= kMaxRegL) {
        TileKernel k = pick_kernel(args.n_l, tile.wpt);
        if (!k) return cudaErrorInvalidValue;
        k<<<tile.grid, tile.threads, tile.smem_bytes, stream>>>(args);
    } else {
        if (!lo_dev || !hi_dev) return cudaErrorInvalidValue;
        lnpost_generic_kernel<<<tile.grid, tile.threads, tile.smem_bytes, stream>>>(args, lo_dev, hi_dev);
    }
    return cudaGetLastError();
}

}  // namespace mda
