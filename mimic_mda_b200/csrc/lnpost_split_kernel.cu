// lnpost_split_kernel.cu -- the batched chi-square log-posterior for SMALL launches: G lanes per walker.
//
// Same job and same results, bit for bit, as lnpost_tile_kernel (lnpost_kernel.cu; reference: ln_post_prob
// mda.py:1540-1578, calculateLnLiklihood / calculateChi C/ChiSquare.c:114-213).  The tile kernel gives a walker to a
// thread, which is the right shape when a launch carries many walkers per SM; the half-step of a host-driven sampler
// on the shapes BASELINE.json names does not (C4: 200 bins x 256 walkers = 346 evaluations per SM), and a thread that
// walks 60 angles x 10 dependent multiply-adds alone is a 1.5 us latency chain inside an 8 us launch.  Here
//
//   * a walker belongs to a group of G consecutive lanes (G = 2, 4, ... 32, the smallest power of two with
//     kPairsPerLane angle pairs per lane); lane s of the group evaluates the residuals
//     r_j = d_j - sum_l a_l D_lj (ascending l, one fused multiply-add each: C/ChiSquare.c:188-202) of ITS angles,
//     a contiguous run of kPairsPerLane pairs, and keeps them in registers -- the part that is 9/10 of the work runs
//     G-fold parallel;
//   * chi^2 = sum_j r_j^2 is then accumulated in the reference's order, ascending j from 0 (C/ChiSquare.c:204-210), by
//     handing the running sum from lane to lane with a shuffle: lane 0 adds its squares, lane 1 continues from lane
//     0's value, ...  No partial sums are ever combined, so the value is the one a single thread would get --
//     results do not depend on which kernel, tiling or batch shape evaluated them (tested bit for bit);
//   * the bin's constants come in by bulk copies, one per lane segment, each segment padded by 16 bytes so that the G
//     lanes of a group read G different banks (segments of 8 pairs x (L+1) x 16 B are a multiple of 128 B apart otherwise);
//   * parameters are read straight from global memory (the G lanes of a group read the same 8 L bytes: one sector
//     fetch per walker), the last lane of a group applies the box prior and writes the result.
//
// Angles beyond a bin's own count are skipped element by element (never multiplied by padding: an infinite parameter
// times a zero pad would turn chi^2 = inf into NaN).
#include "likelihood.cuh"

#include <utility>

namespace mda {

namespace {

constexpr int kPairsPerLane = 8;          // angle pairs (16 angles) whose residuals a lane keeps in registers
constexpr int kSplitThreads = 256;

template <int L>
__global__ void __launch_bounds__(kSplitThreads) lnpost_split_kernel(const BatchArgs args, const int G) {
    constexpr int ROW = L + 1;                               // double2 entries per angle pair: data, dist_0 .. dist_{L-1}
    constexpr int SEG = kPairsPerLane * ROW + 1;             // double2 per lane segment, padded by one
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ __align__(8) uint64_t mbar;
    double2 *sC = reinterpret_cast<double2 *>(smem_raw);     // [G][SEG]

    const int tid = threadIdx.x;
    const int wpc = kSplitThreads / G;                       // walkers per CTA
    const int bin = blockIdx.x / args.tiles_per_bin;
    const int tile = blockIdx.x - bin * args.tiles_per_bin;
    const int sub = tid & (G - 1), group = tid / G;
    const int w = tile * wpc + group;
    const bool live = w < args.walkers;
    const int n_pairs_pad = args.n_pad / 2;

    if (tid == 0) {
        mbar_init(&mbar, 1);
        fence_mbar_init();
        const double2 *gC = reinterpret_cast<const double2 *>(args.consts) + (size_t)bin * n_pairs_pad * ROW;
        uint32_t total = 0;
        for (int s = 0; s * kPairsPerLane < n_pairs_pad; ++s) total += (uint32_t)min(kPairsPerLane, n_pairs_pad - s * kPairsPerLane) * ROW * 16u;
        mbar_arrive_expect_tx(&mbar, total);
        for (int s = 0; s * kPairsPerLane < n_pairs_pad; ++s)
            bulk_g2s(sC + (size_t)s * SEG, gC + (size_t)s * kPairsPerLane * ROW, (uint32_t)min(kPairsPerLane, n_pairs_pad - s * kPairsPerLane) * ROW * 16u, &mbar);
    }
    // my walker's parameters, while the constants are on their way
    const double *gP = args.params + ((size_t)bin * args.walkers + (live ? w : args.walkers - 1)) * L;
    double a[L];
#pragma unroll
    for (int l = 0; l < L; ++l) a[l] = __ldg(gP + l);
    const int n = args.n_pts[bin];
    __syncthreads();                                         // the mbarrier is initialised (and, for thread 0, armed)
    mbar_wait(&mbar, 0);

    // ---- residuals of my run of angles (parallel part)
    double r0[kPairsPerLane], r1[kPairsPerLane];
    const double2 *p = sC + (size_t)sub * SEG;
    const int j0 = 2 * sub * kPairsPerLane;                  // my first angle
#pragma unroll
    for (int k = 0; k < kPairsPerLane; ++k, p += ROW) {
        const int j = j0 + 2 * k;
        if (j < n) {
            const double2 dv = p[0];
            double x = dv.x, y = dv.y;
#pragma unroll
            for (int l = 0; l < L; ++l) {
                const double2 D = p[1 + l];
                x = fma(-a[l], D.x, x);                      // C/ChiSquare.c:190,200
                y = fma(-a[l], D.y, y);
            }
            r0[k] = x;
            r1[k] = j + 1 < n ? y : 0.0;                     // the pad of an odd bin contributes an exact zero
        } else {
            r0[k] = 0.0;
            r1[k] = 0.0;
        }
    }
    // ---- chi^2 in ascending j: the running sum travels through the group's lanes
    double chi = 0.0;
    for (int g = 0; g < G; ++g) {
        if (sub == g) {
#pragma unroll
            for (int k = 0; k < kPairsPerLane; ++k) {
                chi = fma(r0[k], r0[k], chi);                // C/ChiSquare.c:209; fma(0, 0, chi) == chi for skipped angles
                chi = fma(r1[k], r1[k], chi);
            }
        }
        const double handed = __shfl_up_sync(0xffffffffu, chi, 1, G);
        if (sub == g + 1) chi = handed;
    }
    if (live && sub == G - 1) {
        bool outside = false;
#pragma unroll
        for (int l = 0; l < L; ++l) outside = outside || (a[l] < args.lo[l]) || (a[l] > args.hi[l]);   // mda.py:1575; NaN compares false
        args.out[(size_t)bin * args.walkers + w] = finish(chi, outside, args.mode);
    }
}

using SplitKernel = void (*)(const BatchArgs, const int);

template <int... Ls>
SplitKernel pick_from(int n_l, std::integer_sequence<int, Ls...>) {
    SplitKernel k = nullptr;
    ((n_l == Ls + 1 ? (k = lnpost_split_kernel<Ls + 1>, 0) : 0), ...);
    return k;
}

SplitKernel pick_split_kernel(int n_l) { return pick_from(n_l, std::make_integer_sequence<int, kMaxRegL>{}); }

}  // namespace

int lnpost_split_lanes(int n_l, int n_pad) {
    if (n_l < 1 || n_l > kMaxRegL || n_pad < 2) return 0;
    const int segments = (n_pad / 2 + kPairsPerLane - 1) / kPairsPerLane;
    int g = 1;
    while (g < segments) g <<= 1;
    if (g < 2 || g > 32) return 0;                           // one lane would be the tile kernel with extra steps
    return (size_t)g * (size_t)(kPairsPerLane * (n_l + 1) + 1) * 16 <= 48u * 1024u ? g : 0;   // static shared-memory limit
}

cudaError_t launch_batch_split(BatchArgs &args, cudaStream_t stream, int *grid_out) {
    if (args.bins <= 0 || args.walkers <= 0) return cudaSuccess;
    const int g = lnpost_split_lanes(args.n_l, args.n_pad);
    SplitKernel k = pick_split_kernel(args.n_l);
    if (!g || !k) return cudaErrorInvalidValue;
    const int wpc = kSplitThreads / g;
    args.tiles_per_bin = (args.walkers + wpc - 1) / wpc;
    const int grid = args.bins * args.tiles_per_bin;
    const size_t smem = (size_t)g * (size_t)(kPairsPerLane * (args.n_l + 1) + 1) * 16;
    if (grid_out) *grid_out = grid;
    k<<<grid, kSplitThreads, smem, stream>>>(args, g);
    return cudaGetLastError();
}

}  // namespace mda
