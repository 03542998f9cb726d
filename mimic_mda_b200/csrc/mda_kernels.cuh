// mda_kernels.cuh -- internal launch interface between the C ABI (chisq_abi.cu)
// and the sm_100a kernels (lnpost_kernel.cu, scalar_kernel.cu, peaks.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace mda {

constexpr int kMaxRegL = 16;        // largest numL the register-tiled kernel is instantiated for
constexpr int kMinMmaL = 4;         // smallest numL with at least one DMMA k-step (ensemble_mma_kernel.cu)

enum EvalMode : int {
    kLnPost = 0,   // box prior (-inf outside) then chi^2 / -2   (mda.py:1574-1578)
    kChi = 1,      // chi^2, no prior                            (C/ChiSquare.c:114-172)
    kLnLike = 2    // chi^2 / -2, no prior                       (C/ChiSquare.c:176-213)
};

// One batched evaluation: every bin x walkers_per_bin walkers.
struct BatchArgs {
    const double *consts;   // [bins][n_pad/2][1+L][2]  per angle pair: data pair, dist_0 pair, ...
    const int *n_pts;       // [bins]               real angle count per bin (<= n_pad)
    const double *params;   // [bins][walkers][L]
    double *out;            // [bins][walkers]
    int bins;
    int walkers;            // per bin in this launch
    int n_l;
    int n_pad;              // even
    int tiles_per_bin;
    int chunk;              // angle columns staged in shared memory at a time (even)
    int mode;               // EvalMode
    double lo[kMaxRegL];    // box prior, used by the register-tiled kernel (numL <= kMaxRegL)
    double hi[kMaxRegL];
    const double *consts_mma;   // DMMA kernel: fragment-major constants [bins][n_tiles][8 (1+L)] (see mma_chi.cuh)
    int n_tiles;                // DMMA kernel: angle tiles of 8
    const double *quad;         // triangular-factor kernel: [bins][quad_stride(L)] packed T (see quad_chi.cuh)
    int params_bulk;            // params is 16-byte aligned: a tile's parameter block may come in by one bulk copy
};

struct TileShape {
    int threads;            // per CTA
    int wpt;                // walkers per thread (register tile)
    int grid;               // CTAs
    int chunk;              // angle chunk (even)
    size_t smem_bytes;      // dynamic shared memory
};

// Picks the tiling for a launch (force_threads / force_wpt = 0 -> heuristic).
TileShape choose_tile(int bins, int walkers, int n_l, int n_pad, int sm_count,
                      size_t smem_optin, int force_threads, int force_wpt);

// Enqueue one batched evaluation.  args.tiles_per_bin / args.chunk are filled in.
// lo_dev / hi_dev ([numL], device) are only read when numL > kMaxRegL.
cudaError_t launch_batch(BatchArgs &args, const TileShape &tile, cudaStream_t stream,
                         const double *lo_dev, const double *hi_dev);

// One-time per device: raise the dynamic shared memory limit of every instantiation.
cudaError_t configure_kernels(size_t smem_optin);

// The same evaluation on the fp64 tensor path (lnpost_mma_kernel.cu), numL kMinMmaL .. kMaxRegL: a warp per item of
// 32 walkers, parameters read straight from global memory.  args.consts_mma / n_tiles set by the caller.
bool lnpost_mma_supported(int n_l, int n_tiles, size_t smem_optin);
cudaError_t configure_lnpost_mma_kernels(size_t smem_optin);
cudaError_t launch_batch_mma(BatchArgs &args, int sm_count, cudaStream_t stream, int *grid_out, int *threads_out);

// The same evaluation with G lanes per walker for small launches (lnpost_split_kernel.cu): residuals of a lane's run of
// angles in parallel, chi^2 handed from lane to lane in the reference's order -- bit-identical to the tile kernel.
int lnpost_split_lanes(int n_l, int n_pad);                    // G (2 .. 32), or 0 when the shape is not served
cudaError_t launch_batch_split(BatchArgs &args, cudaStream_t stream, int *grid_out);

// The same evaluation from the bins' triangular factors (lnpost_quad_kernel.cu), numL <= kMaxRegL, args.quad set.
bool lnpost_quad_supported(int n_l);
cudaError_t launch_batch_quad(const BatchArgs &args, cudaStream_t stream);

// ---- device-resident ensemble stepping (ensemble_kernel.cu) -----------------------------
struct EnsembleArgs {
    const double *consts;      // store layout [bins][n_pad/2][1+L][2]
    const double *consts_mma;  // fragment-major copy [bins][n_tiles][8 (1+L)] (see mma_tile_doubles), or null
    const double *quad;        // [bins][quad_stride(L)] packed triangular factor T of [D^T | d]: chi^2(a) = |T [a; -1]|^2, or null
    const int *n_pts;          // [bins]
    const int *bin_ids;        // [bins] global bin index (Philox counter word; partition invariant)
    double *pos;               // [bins][W][L]   ensemble state, updated in place
    double *lnp;               // [bins][W]
    unsigned int *nacc;        // [bins][W]      accepted proposals per walker
    double *chain;             // [keep_cap][bins][W][L] or null  (step-major: a run streams out contiguously)
    double *lnp_chain;         // [keep_cap][bins][W]    or null
    int *status;               // bit flags (1 = NaN log-probability seen)
    long long *prof;           // optional [bins][8] cycle counters per section, or null
    int bins, walkers, n_l, n_pad;
    int n_tiles;               // angle tiles of 8 in consts_mma
    int n_steps, thin;
    int chunk_steps, n_chunks; // DMMA stepper: units of work are (chunk of steps, bin), chunk-major
    unsigned long long *progress;      // [bins] launch_serial + chunks done: the hand-over word between units of a bin
    unsigned long long launch_serial;  // distinct per launch, low 32 bits zero
    int groups;                // angle groups per walker tile (filled in by launch_ensemble)
    long long step0;           // index of the first step of this launch (RNG counter)
    long long keep0, keep_cap; // next free slot / capacity of the chain
    unsigned int seed_lo, seed_hi;
    double a;                  // stretch scale (emcee default 2)
    unsigned int round_keys[20];       // Philox round keys of the seed (philox_round_keys), for the kernels that take them from here
    double a_inv;              // 1 / a when a is a power of two (then z^2 / a == z^2 * a_inv exactly), else 0
    double lo[kMaxRegL];
    double hi[kMaxRegL];
};

struct EnsembleShape {
    int threads, wpt, groups;  // blockDim = walker threads x angle groups
    int stage;                 // ensemble_mma_kernel: 1 = with draw warps (one CTA of 2 x items warps per SM), 0 = two CTAs per SM
    size_t smem_bytes;
    bool fits;                 // the launch is possible (shared memory, numL <= kMaxRegL)
    bool global_state;         // positions / log-probabilities stay in HBM (ensemble too large for shared memory)
    bool mma;                  // a DMMA stepper (wpt = tiles of 8 walkers per item)
    bool ws;                   // the warp-specialised one, ensemble_ws_kernel (groups = walker warps); else ensemble_mma_kernel
    bool quad;                 // ensemble_quad_kernel: chi^2 from the bin's triangular factor, one thread per walker
    bool qws;                  // ensemble_qws_kernel: the same, warp-specialised (walker warps + draw warps); quad is set too
    bool treg;                 // ensemble_qws_kernel: the factor T in registers (else in shared memory)
};

__host__ __device__ constexpr int quad_doubles(int n_l) { return (n_l + 1) * (n_l + 2) / 2; }
// doubles between two bins' packed factors in the store: rounded up to a 16-byte multiple, so that a factor is a bulk-copy source
__host__ __device__ constexpr int quad_stride(int n_l) { return (quad_doubles(n_l) + 1) & ~1; }

EnsembleShape choose_ensemble_shape(int bins, int walkers, int n_l, int n_pad, int sm_count, size_t smem_optin,
                                    int force_groups, int force_wpt);
cudaError_t configure_ensemble_kernels(size_t smem_optin);
cudaError_t launch_ensemble(const EnsembleArgs &args, const EnsembleShape &shape, cudaStream_t stream);

// DMMA variant (ensemble_mma_kernel.cu).  Constants per bin and angle tile t (8 angles, zero padded):
//   rows 0..R  (R = L % 4): [8] data, then -D_l for l < R           -> 8 (1 + R) doubles
//   k-step ks < L / 4     : [32] B fragment, lane (g, tig) holds -D_{R + 4 ks + tig}[8 t + g]
// i.e. 8 (1 + L) doubles per tile; the distributions are stored negated so that the residual
// d - sum a_l D_l is a pure multiply-accumulate onto d.  Returns false when the shape cannot run (numL < kMinMmaL,
// numL > kMaxRegL, more than 512 walkers or the ensemble does not fit shared memory).
bool choose_ensemble_mma_shape(int bins, int walkers, int n_l, int n_tiles, int sm_count, size_t smem_optin, int force_wt,
                               int force_draw_warps, EnsembleShape *out);   // force_draw_warps: 0 auto, 1 with, 2 without
cudaError_t configure_ensemble_mma_kernels(size_t smem_optin);
int ensemble_mma_capacity(int n_l, const EnsembleShape &shape, int sm_count);   // co-resident CTAs on the device
cudaError_t launch_ensemble_mma(const EnsembleArgs &args, const EnsembleShape &shape, int grid, cudaStream_t stream);
// Warp-specialised DMMA variant (ensemble_ws_kernel.cu): same constants, same unit schedule, one CTA of 16 warps per SM.
bool choose_ensemble_ws_shape(int bins, int walkers, int n_l, int n_tiles, int sm_count, size_t smem_optin, int force_wt,
                              EnsembleShape *out);
cudaError_t configure_ensemble_ws_kernels(size_t smem_optin);
int ensemble_ws_capacity(int n_l, const EnsembleShape &shape, int sm_count);
cudaError_t launch_ensemble_ws(const EnsembleArgs &args, const EnsembleShape &shape, int grid, cudaStream_t stream);

// ensemble_quad_kernel.cu: one CTA per bin at a time, one thread per walker of the half being updated.
bool choose_ensemble_quad_shape(int walkers, int n_l, size_t smem_optin, EnsembleShape *out);
cudaError_t configure_ensemble_quad_kernels(size_t smem_optin);
int ensemble_quad_capacity(int n_l, const EnsembleShape &shape, int grid, int sm_count);
cudaError_t launch_ensemble_quad(const EnsembleArgs &args, const EnsembleShape &shape, int grid, int sm_count, cudaStream_t stream);

// ensemble_qws_kernel.cu: the same stepper warp-specialised -- 8 walker warps + 4 draw warps per bin (walkers / 2 <= 256).
bool choose_ensemble_qws_shape(int bins, int walkers, int n_l, int sm_count, size_t smem_optin, EnsembleShape *out);
cudaError_t configure_ensemble_qws_kernels(size_t smem_optin);
int ensemble_qws_capacity(int n_l, const EnsembleShape &shape, int sm_count);   // co-resident CTAs on the device
// grid == bins: a CTA per bin; grid < bins: the bins' step ranges cut evenly over `grid` CTAs (cooperative launch, args.progress set)
cudaError_t launch_ensemble_qws(const EnsembleArgs &args, const EnsembleShape &shape, int grid, cudaStream_t stream);

// A stepping launch whose CTAs wait for each other's progress words (units of one bin on different CTAs) must have every
// CTA resident: launched cooperatively, a grid that cannot be co-resident fails at launch instead of spinning.
template <typename Kernel>
inline cudaError_t launch_maybe_cooperative(Kernel k, const EnsembleArgs &args, bool ctas_wait_for_each_other, int grid, int threads,
                                            size_t smem_bytes, cudaStream_t stream) {
    if (!ctas_wait_for_each_other) {
        k<<<grid, threads, smem_bytes, stream>>>(args);
        return cudaGetLastError();
    }
    EnsembleArgs copy = args;
    void *params[] = {&copy};
    return cudaLaunchCooperativeKernel(reinterpret_cast<const void *>(k), dim3((unsigned)grid), dim3((unsigned)threads), params, smem_bytes, stream);
}

// Wall-clock bound of a wait for another CTA's progress word (nanoseconds of %globaltimer): generous enough for a
// profiler's replay or a sanitizer, short enough not to look like a hung GPU.
constexpr unsigned long long kHandOverTimeoutNs = 4000000000ull;
__device__ __forceinline__ unsigned long long globaltimer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

// ---- batched initial fits (fit_kernel.cu) -----------------------------------------------------
// One thread per (bin, start): box-constrained least-squares minimiser on the bin's packed triangular factor (quad:
// [bins][quad_stride(L)], see quad_chi.cuh), warm-started at the start.  starts is [starts_per_bin][L] when shared, else
// [bins][starts_per_bin][L]; out [bins][starts_per_bin][L]; iters (optional) [bins][starts_per_bin]: iterations, negative
// when the limit was hit, bit 16 set when a free set met a numerically null column.
cudaError_t launch_bvls(const double *quad, const double *lo, const double *hi, const double *starts,
                        int bins, int n_l, int starts_per_bin, bool shared_starts, double *out, int *iters,
                        cudaStream_t stream);

// ---- posterior reductions on the device chain (summary_kernel.cu) ---------------------------------
constexpr int kMaxRanks = 12;          // order statistics selected together per (bin, dim) column
// minmax[cols][2]: order-preserving keys of the smallest / largest sample (init ~0 / 0)
cudaError_t launch_chain_minmax(const double *chain, long long k0, long long k1, int bins, int walkers, int n_l,
                                unsigned long long *minmax, int sm_count, cudaStream_t stream);
// numpy.histogram(values, bins=n_bins) of every column over [min, max] (edges[cols][n_bins + 1] and
// hist[cols][n_bins] are scratch), then per column the first fullest bin and the number of
// samples up to and including it (mda.py:1466-1473).
cudaError_t launch_chain_hist(const double *chain, long long k0, long long k1, int bins, int walkers, int n_l,
                              const unsigned long long *minmax, double *edges, int n_bins, unsigned int *hist,
                              int *peak_bin, long long *peak_cum, int sm_count, cudaStream_t stream);
// radix selection, step 1: one histogram pass over the chain; on return prefix / remaining /
// count [cols][n_ranks] describe the bucket (12 leading undecided bits fixed) of every rank.
// hist is [cols][4096] scratch, top[cols] = undecided low key bits per column.
cudaError_t launch_select_first(const double *chain, long long k0, long long k1, int bins, int walkers, int n_l,
                                unsigned long long *prefix, long long *remaining, unsigned int *count, const int *top,
                                unsigned int *hist, int n_ranks, int sm_count, cudaStream_t stream);
// step 2: gather the keys of the buckets in play (lead / offset [cols][n_ranks] from the host,
// cursor scratch) into buf and finish every rank on its bucket; prefix then holds the exact keys.
cudaError_t launch_select_finish(const double *chain, long long k0, long long k1, int bins, int walkers, int n_l,
                                 unsigned long long *prefix, long long *remaining, unsigned int *count, const int *top,
                                 const unsigned char *lead, const long long *offset, unsigned int *cursor,
                                 unsigned long long *buf, int n_ranks, int sm_count, cudaStream_t stream);

// ---- autocorrelation of the ensemble-mean series (autocorr_kernel.cu) ------------------------------
// series[cols][stride]: mean over the walkers of every kept sample, col = bin * L + l.
cudaError_t launch_ensemble_mean(const double *chain, long long kept, int bins, int walkers, int n_l, double *series,
                                 long long stride, cudaStream_t stream);
// centred = series - mean over its first n samples; acf[cols][n_lags] = linear autocorrelation sums of those n.
cudaError_t launch_series_acf(const double *series, long long stride, int n, int cols, int n_lags, double *centred, double *acf,
                              cudaStream_t stream);

// Legacy scalar path (one bin, one parameter vector), see scalar_kernel.cu.
//   consts [1+L][n] dense (row 0 data, rows 1..L dists), params[L], result[0] = value,
//   resid (may be null) [n]
cudaError_t launch_scalar(const double *consts, int n, int n_l, const double *params_host, const double *params_dev, double *result,
                          double *resid, int mode, unsigned long long *done, unsigned long long ticket, cudaStream_t stream);

// Roofline denominators, see peaks.cu.
cudaError_t measure_fp64_peak(int repeats, int sm_count, double *tflops);
cudaError_t measure_hbm_copy(size_t bytes, int repeats, double *gbs);

}  // namespace mda
