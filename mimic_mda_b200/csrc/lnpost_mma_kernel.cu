// lnpost_mma_kernel.cu -- the batched chi-square log-posterior on the fp64 tensor path (DMMA.8x8x4).
//
// Same job as lnpost_kernel.cu -- for every (Ex bin, walker) pair of one launch the reference's
//   ln_post_prob            mda.py:1540-1578      (flat box prior -> -inf)
//   calculateLnLiklihood    C/ChiSquare.c:176-213 (r_j = d_j - sum_l a_l D_lj; -chi^2/2)
//   calculateChi            C/ChiSquare.c:114-172
// -- with the contraction of the device-resident sampler (mma_chi.cuh): a warp takes an item of 32 walkers of a
// bin, lane (g, tig) = (lane / 4, lane % 4) reads the components l = R + 4 ks + tig of walkers 4 g .. 4 g + 3 of the
// item straight from the parameter block in global memory (32-byte runs: whole sectors) as the A fragments of
// mma.m8n8k4, the bin's constants sit fragment-major in shared memory (one bulk copy per CTA), and after the quad
// butterfly lane i of the warp holds chi^2 of walker i of the item: one coalesced store.  Every term of the
// reference's sum is evaluated exactly once; only the order of the fp64 additions differs from C/ChiSquare.c
// (<= 1e-15 relative, tested), and it does not depend on the launch shape: a walker's value is the same
// whichever item, CTA or launch it is evaluated in (tested bit for bit).
//
// Why: the register-tiled DFMA loop of lnpost_kernel.cu plateaus at 65-70 % of the fp64 pipe (a DFMA with three
// distinct register operands is register-bandwidth bound) and needs 64-thread CTAs with staged parameters at the
// half-step shapes; this one has no staging of parameters at all.  numL 4 .. 16; lnpost_kernel.cu keeps the rest.
#include "mma_chi.cuh"

namespace mda {

namespace {

constexpr int kWarps = 4;              // per CTA
constexpr int WT = 4;                  // tiles of 8 walkers per item: an item is a warp's worth of walkers

template <int L>
__global__ void __launch_bounds__(32 * kWarps, 4) lnpost_mma_kernel(const BatchArgs args) {
    constexpr int Q = L / 4, R = L % 4, RR = R > 0 ? R : 1;
    constexpr int TS = 8 * (L + 1);               // doubles per angle tile of the constants

    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ __align__(8) uint64_t mbar;
    double *sC = reinterpret_cast<double *>(smem_raw);

    const int tid = threadIdx.x;
    const int lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, tig = lane & 3;
    const int W = args.walkers;
    const int NT = args.n_tiles;
    const int bin = blockIdx.x / args.tiles_per_bin;                 // tiles_per_bin: CTAs per bin
    const int part = blockIdx.x - bin * args.tiles_per_bin;
    const int n_items = (W + 31) >> 5;
    const int item_end = min(n_items, (part + 1) * args.chunk);     // chunk: items per CTA

    if (tid == 0) {
        mbar_init(&mbar, 1);
        fence_mbar_init();
        const uint32_t cb = (uint32_t)NT * TS * 8u;
        mbar_arrive_expect_tx(&mbar, cb);
        bulk_g2s(sC, args.consts_mma + (size_t)bin * NT * TS, cb, &mbar);
    }
    // box prior of the components this lane holds (mda.py:1124, 1574-1576)
    double a_lo[Q], a_hi[Q];
#pragma unroll
    for (int ks = 0; ks < Q; ++ks) {
        a_lo[ks] = args.lo[R + 4 * ks + tig];
        a_hi[ks] = args.hi[R + 4 * ks + tig];
    }
    __syncthreads();                                                 // the barrier word is initialised
    const double *pbin = args.params + (size_t)bin * W * L;
    double *obin = args.out + (size_t)bin * W;
    bool landed = false;

    // fragments of an item: walker 4 g + k of the item is row g of tile k, so that lane i ends up with walker i
    auto load_item = [&](int item, double (&A)[WT][Q], double (&lead)[WT][RR]) {
#pragma unroll
        for (int k = 0; k < WT; ++k) {
            const double *row = pbin + (size_t)min(item * 32 + 4 * g + k, W - 1) * L;
#pragma unroll
            for (int ks = 0; ks < Q; ++ks) A[k][ks] = __ldg(row + R + 4 * ks + tig);
#pragma unroll
            for (int l = 0; l < R; ++l) lead[k][l] = __ldg(row + l);
        }
    };
    double A_next[WT][Q], lead_next[WT][RR];
    const int item0 = part * args.chunk + warp;
    if (item0 < item_end) load_item(item0, A_next, lead_next);

    for (int item = item0; item < item_end; item += kWarps) {
        double A[WT][Q], lead[WT][RR];
#pragma unroll
        for (int k = 0; k < WT; ++k) {
#pragma unroll
            for (int ks = 0; ks < Q; ++ks) A[k][ks] = A_next[k][ks];
#pragma unroll
            for (int l = 0; l < RR; ++l) lead[k][l] = lead_next[k][l];
        }
        if (item + kWarps < item_end) load_item(item + kWarps, A_next, lead_next);   // in flight during this item's chi^2 loop
        bool outside = false;
        if (args.mode == kLnPost) {
            uint32_t out_of[WT];
#pragma unroll
            for (int k = 0; k < WT; ++k) {
                bool o = false;
#pragma unroll
                for (int ks = 0; ks < Q; ++ks) o = o | (A[k][ks] < a_lo[ks]) | (A[k][ks] > a_hi[ks]);   // a NaN compares false, as in the reference
#pragma unroll
                for (int l = 0; l < R; ++l) o = o | (lead[k][l] < args.lo[l]) | (lead[k][l] > args.hi[l]);
                out_of[k] = __ballot_sync(0xffffffffu, o);           // bits 4 g .. 4 g + 3: walker 4 g + k
            }
            uint32_t mine = out_of[0];
#pragma unroll
            for (int k = 1; k < WT; ++k) mine = tig == k ? out_of[k] : mine;
            outside = ((mine >> (4 * g)) & 0xfu) != 0u;
        }
        if (!landed) {
            mbar_wait(&mbar, 0);
            landed = true;
        }
        double chi0[WT], chi1[WT];
#pragma unroll
        for (int k = 0; k < WT; ++k) { chi0[k] = 0.0; chi1[k] = 0.0; }
        if (NT == 8) chi_tiles<L, WT, 8>(sC, 8, lane, A, lead, chi0, chi1);
        else chi_tiles<L, WT, 0>(sC, NT, lane, A, lead, chi0, chi1);
        // quad sums by a transposing butterfly: lane tig ends up with the sum of tile tig
        double sv[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) sv[k] = chi0[k] + chi1[k];
        const bool b0 = (tig & 1) != 0, b1 = (tig & 2) != 0;
        const double a01 = (b0 ? sv[1] : sv[0]) + __shfl_xor_sync(0xffffffffu, b0 ? sv[0] : sv[1], 1);   // tile tig & 1
        const double a23 = (b0 ? sv[3] : sv[2]) + __shfl_xor_sync(0xffffffffu, b0 ? sv[2] : sv[3], 1);   // tile 2 + (tig & 1)
        const double sum = (b1 ? a23 : a01) + __shfl_xor_sync(0xffffffffu, b1 ? a01 : a23, 2);
        const int w = item * 32 + lane;                              // = 4 g + tig of the item
        if (w < W) obin[w] = finish(sum, outside, args.mode);
    }
}

using MmaKernel = void (*)(const BatchArgs);

MmaKernel pick_lnpost_mma(int n_l) {
    switch (n_l) {
        case 4: return lnpost_mma_kernel<4>;
        case 5: return lnpost_mma_kernel<5>;
        case 6: return lnpost_mma_kernel<6>;
        case 7: return lnpost_mma_kernel<7>;
        case 8: return lnpost_mma_kernel<8>;
        case 9: return lnpost_mma_kernel<9>;
        case 10: return lnpost_mma_kernel<10>;
        case 11: return lnpost_mma_kernel<11>;
        case 12: return lnpost_mma_kernel<12>;
        case 13: return lnpost_mma_kernel<13>;
        case 14: return lnpost_mma_kernel<14>;
        case 15: return lnpost_mma_kernel<15>;
        case 16: return lnpost_mma_kernel<16>;
        default: return nullptr;
    }
}

}  // namespace

bool lnpost_mma_supported(int n_l, int n_tiles, size_t smem_optin) {
    const size_t bytes = (size_t)n_tiles * 8 * (size_t)(n_l + 1) * 8;
    return n_l >= kMinMmaL && n_l <= kMaxRegL && n_tiles >= 1 && bytes + 64 <= smem_optin && bytes < (1u << 20);
}

cudaError_t configure_lnpost_mma_kernels(size_t smem_optin) {
    for (int l = kMinMmaL; l <= kMaxRegL; ++l) {
        cudaError_t e = cudaFuncSetAttribute(reinterpret_cast<const void *>(pick_lnpost_mma(l)),
                                             cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(smem_optin - 64));
        if (e != cudaSuccess) return e;
    }
    return cudaSuccess;
}

// args.consts_mma / n_tiles set by the caller; tiles_per_bin (CTAs per bin) and chunk (items per CTA) are filled in here:
// one item per warp where that keeps the launch within ~8 CTAs per SM, more passes per warp beyond.
cudaError_t launch_batch_mma(BatchArgs &args, int sm_count, cudaStream_t stream, int *grid_out, int *threads_out) {
    if (args.bins <= 0 || args.walkers <= 0) return cudaSuccess;
    MmaKernel k = pick_lnpost_mma(args.n_l);
    if (!k || !args.consts_mma) return cudaErrorInvalidValue;
    const long n_items = (args.walkers + 31) / 32, total = n_items * args.bins;
    const long target_ctas = 8L * sm_count;
    long passes = (total + kWarps * target_ctas - 1) / (kWarps * target_ctas);
    if (passes < 1) passes = 1;
    args.chunk = (int)(kWarps * passes);
    args.tiles_per_bin = (int)((n_items + args.chunk - 1) / args.chunk);
    const int grid = args.bins * args.tiles_per_bin;
    const size_t smem = (size_t)args.n_tiles * 8 * (size_t)(args.n_l + 1) * 8;
    if (grid_out) *grid_out = grid;
    if (threads_out) *threads_out = 32 * kWarps;
    k<<<grid, 32 * kWarps, smem, stream>>>(args);
    return cudaGetLastError();
}

}  // namespace mda
