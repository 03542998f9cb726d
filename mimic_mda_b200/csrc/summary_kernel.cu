// summary_kernel.cu -- posterior reductions on the device-resident chain.
//
// The reference pulls the whole chain into numpy and, per Ex bin and parameter, takes
// percentiles of the post-burn-in samples and a histogram peak (perform_sample_manips,
// calc_param_values, find_most_likely_values: mda.py:1253-1259, 1399-1491).  The chain of one
// run is gigabytes (mda_config0.py:192-193 warns about it) while the results are a few
// numbers per bin, so the reductions run where the chain lives:
//   minmax_kernel     per (bin, dim) column: smallest and largest sample (order-preserving
//                     integer keys, so the results are exact doubles)
//   hist_kernel       numpy.histogram-compatible equal-width counts over [min, max]
//   first_* / gather / finish   exact order statistics by most-significant-digit radix
//                     selection below the bits min and max share: one 12-bit histogram pass,
//                     then the few keys of the buckets in play are gathered and finished
//                     off-chain; up to kMaxRanks ranks per column together
// Everything that depends on numpy's floating-point conventions (virtual index, linear
// interpolation) is evaluated on the host from these exact order statistics.
#include "mda_kernels.cuh"

#include <cstdio>
#include <cstdlib>

namespace mda {

namespace {

constexpr int kDigitBits = 12;
constexpr int kDigits = 1 << kDigitBits;

__device__ __forceinline__ unsigned long long to_key(double x) {
    const unsigned long long u = (unsigned long long)__double_as_longlong(x);
    return (u >> 63) ? ~u : (u | 0x8000000000000000ull);        // monotone in x
}

// Thread layout shared by the kernels: blockDim = (L, ty); a warp covers consecutive
// (walker, dim) pairs, i.e. contiguous memory; a thread's dim is fixed.  grid = (bins, chunks):
// CTA (b, c) walks kept samples k = k0 + c, k0 + c + chunks, ...
struct ChainView {
    const double *chain;       // [kept][bins][W][L]
    long long k0, k1;          // kept-sample range [k0, k1)
    int bins, walkers, n_l;
};

// Visit every sample of column (bin, l) owned by this thread: eight independent loads are in
// flight per thread (64 B; ~2000 resident threads per SM keep > 100 KB in flight, HBM rate).
template <typename F>
__device__ __forceinline__ void for_each_sample(const ChainView &v, int bin, int l, F visit) {
    // The rows of the block (blockDim.y) split into G groups that take different kept samples, so that a
    // thread's share of a slab is about one batch of eight loads whatever the block and ensemble sizes.
    int G = ((int)blockDim.y * 8 + v.walkers / 2) / v.walkers;
    G = G < 1 ? 1 : (G > (int)blockDim.y ? (int)blockDim.y : G);
    const int stride = (int)blockDim.y / G;
    const int g = (int)threadIdx.y / stride, wy = (int)threadIdx.y % stride;
    if (g >= G) return;
    for (long long k = v.k0 + (long long)blockIdx.y * G + g; k < v.k1; k += (long long)gridDim.y * G) {
        const double *slab = v.chain + ((size_t)k * v.bins + bin) * v.walkers * v.n_l + l;
        int w = wy;
        for (; w + 7 * stride < v.walkers; w += 8 * stride) {
            double x[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) x[i] = __ldcs(slab + (size_t)(w + i * stride) * v.n_l);
#pragma unroll
            for (int i = 0; i < 8; ++i) visit(x[i]);
        }
        double x[8];
#pragma unroll
        for (int i = 0; i < 8; ++i)
            if (w + i * stride < v.walkers) x[i] = __ldcs(slab + (size_t)(w + i * stride) * v.n_l);
#pragma unroll
        for (int i = 0; i < 8; ++i)
            if (w + i * stride < v.walkers) visit(x[i]);
    }
}

__global__ void __launch_bounds__(256) minmax_kernel(ChainView v, unsigned long long *__restrict__ mm) {
    const int l = threadIdx.x, bin = blockIdx.x;
    unsigned long long lo = ~0ull, hi = 0ull;
    for_each_sample(v, bin, l, [&](double x) {
        const unsigned long long key = to_key(x);
        lo = key < lo ? key : lo;
        hi = key > hi ? key : hi;
    });
    __shared__ unsigned long long s_lo[32], s_hi[32];
    if (threadIdx.y == 0) { s_lo[l] = ~0ull; s_hi[l] = 0ull; }
    __syncthreads();
    atomicMin(&s_lo[l], lo);
    atomicMax(&s_hi[l], hi);
    __syncthreads();
    if (threadIdx.y == 0) {
        atomicMin(&mm[((size_t)bin * v.n_l + l) * 2], s_lo[l]);
        atomicMax(&mm[((size_t)bin * v.n_l + l) * 2 + 1], s_hi[l]);
    }
}

// numpy.histogram(values, bins=n_bins) for every column: the bin of x is
// floor((x - first) / (last - first) * n_bins), corrected against the linspace edges exactly
// as numpy does; edges[col][n_bins + 1] are computed on the host with numpy's formula.
// SMEM: the CTA counts into a private copy [L][n_bins] in shared memory and adds its non-zero
// counts to the global histogram once (global atomics: one per populated bin and CTA instead of one per sample).
template <bool SMEM>
__global__ void __launch_bounds__(256) hist_kernel(ChainView v, const double *__restrict__ edges, int n_bins,
                                                   unsigned int *__restrict__ hist) {
    extern __shared__ unsigned int s_hist[];
    const int l = threadIdx.x, bin = blockIdx.x;
    const int tid = threadIdx.y * blockDim.x + threadIdx.x, nthr = blockDim.x * blockDim.y;
    const size_t col = (size_t)bin * v.n_l + l;
    const double *e = edges + col * (size_t)(n_bins + 1);
    const double first = e[0], last = e[n_bins];
    const double norm = (double)n_bins / (last - first);
    // numpy.linspace's edges (edges_kernel) recomputed in registers: two fp64 operations instead of a
    // data-dependent load per comparison (32 distinct sectors per warp made the loads the bottleneck)
    const double step = __ddiv_rn(__dsub_rn(last, first), (double)n_bins);
    auto edge = [&](int i) { return i == n_bins ? last : __dadd_rn(__dmul_rn((double)i, step), first); };
    unsigned int *h = SMEM ? s_hist + (size_t)l * n_bins : hist + col * (size_t)n_bins;
    if (SMEM) {
        for (int i = tid; i < v.n_l * n_bins; i += nthr) s_hist[i] = 0u;
        __syncthreads();
    }
    if (last > first)
        for_each_sample(v, bin, l, [&](double x) {
            // any estimate within one bin of the truth will do: numpy corrects its own estimate
            // against the edges in exactly this way
            int idx = (int)((x - first) * norm);
            idx = idx < 0 ? 0 : (idx >= n_bins ? n_bins - 1 : idx);
            if (x < edge(idx)) idx -= 1;
            else if (idx != n_bins - 1 && x >= edge(idx + 1)) idx += 1;
            atomicAdd(&h[idx], 1u);
        });
    if (SMEM) {
        __syncthreads();
        unsigned int *g = hist + (size_t)bin * v.n_l * (size_t)n_bins;
        for (int i = tid; i < v.n_l * n_bins; i += nthr)
            if (s_hist[i]) atomicAdd(&g[i], s_hist[i]);
    }
}

// numpy.linspace(first, last, n_bins + 1) per column, with numpy's two roundings per edge
// (i * step, then + first) and the exact end point.
__global__ void edges_kernel(const unsigned long long *__restrict__ mm, int n_bins, double *__restrict__ edges) {
    const size_t col = blockIdx.x;
    const unsigned long long klo = mm[2 * col], khi = mm[2 * col + 1];
    const unsigned long long ulo = (klo >> 63) ? (klo & 0x7FFFFFFFFFFFFFFFull) : ~klo;
    const unsigned long long uhi = (khi >> 63) ? (khi & 0x7FFFFFFFFFFFFFFFull) : ~khi;
    const double first = __longlong_as_double((long long)ulo), last = __longlong_as_double((long long)uhi);
    const double step = __ddiv_rn(__dsub_rn(last, first), (double)n_bins);
    double *e = edges + col * (size_t)(n_bins + 1);
    for (int i = threadIdx.x; i <= n_bins; i += blockDim.x)
        e[i] = i == n_bins ? last : __dadd_rn(__dmul_rn((double)i, step), first);
}

// Per column: index of the first fullest bin (numpy.argmax) and the number of samples in
// bins 0..that one (the loop of mda.py:1471-1473).
__global__ void __launch_bounds__(256) peak_kernel(const unsigned int *__restrict__ hist, int n_bins,
                                                   int *__restrict__ peak_bin, long long *__restrict__ peak_cum) {
    const size_t col = blockIdx.x;
    const unsigned int *h = hist + col * (size_t)n_bins;
    __shared__ unsigned int s_max[256];
    __shared__ int s_arg[256];
    unsigned int best = 0;
    int arg = n_bins;
    for (int i = threadIdx.x; i < n_bins; i += blockDim.x)
        if (h[i] > best || (h[i] == best && i < arg)) { best = h[i]; arg = i; }
    s_max[threadIdx.x] = best;
    s_arg[threadIdx.x] = arg;
    __syncthreads();
    for (int w = 128; w > 0; w >>= 1) {
        if (threadIdx.x < w) {
            const unsigned int ob = s_max[threadIdx.x + w];
            const int oa = s_arg[threadIdx.x + w];
            if (ob > s_max[threadIdx.x] || (ob == s_max[threadIdx.x] && oa < s_arg[threadIdx.x])) {
                s_max[threadIdx.x] = ob;
                s_arg[threadIdx.x] = oa;
            }
        }
        __syncthreads();
    }
    const int ind = s_arg[0] < n_bins ? s_arg[0] : 0;
    __shared__ long long s_sum[256];
    long long sum = 0;
    for (int i = threadIdx.x; i <= ind; i += blockDim.x) sum += (long long)h[i];
    s_sum[threadIdx.x] = sum;
    __syncthreads();
    for (int w = 128; w > 0; w >>= 1) {
        if (threadIdx.x < w) s_sum[threadIdx.x] += s_sum[threadIdx.x + w];
        __syncthreads();
    }
    if (threadIdx.x == 0) { peak_bin[col] = ind; peak_cum[col] = s_sum[0]; }
}

// ---- exact order statistics ---------------------------------------------------------------
// Keys of a column differ only in their low top[col] bits (everything above is shared by the
// column's min and max).  Selection of up to kMaxRanks ranks per column:
//   1. first_hist_kernel   histogram of the leading 12 undecided bits, one per column
//   2. first_scan_kernel   per (column, rank): the bucket that holds the rank, its size and
//                          the rank inside it
//   3. gather_kernel       second read of the chain: the (few) keys of the buckets in play
//                          are appended to exactly-sized buffers (offsets from the counts)
//   4. finish_kernel       one CTA per (column, rank) finishes the radix selection on its
//                          bucket buffer, 8 bits per pass
// SMEM: private counts [L][4096] in shared memory (147 KB at L = 9: one CTA of ~1000 threads per SM),
// non-zero counts added to the global histogram once per CTA.
template <bool SMEM>
__global__ void __launch_bounds__(1024) first_hist_kernel(ChainView v, const int *__restrict__ top,
                                                          unsigned int *__restrict__ hist) {
    extern __shared__ unsigned int s_hist[];
    const int l = threadIdx.x, bin = blockIdx.x;
    const int tid = threadIdx.y * blockDim.x + threadIdx.x, nthr = blockDim.x * blockDim.y;
    const size_t col = (size_t)bin * v.n_l + l;
    const int hi_bit = top[col];
    const int lo_bit = hi_bit > kDigitBits ? hi_bit - kDigitBits : 0;
    const unsigned long long digit_mask = hi_bit > 0 ? (1ull << (hi_bit - lo_bit)) - 1ull : 0ull;
    unsigned int *h = SMEM ? s_hist + (size_t)l * kDigits : hist + col * kDigits;
    if (SMEM) {
        for (int i = tid; i < v.n_l * kDigits; i += nthr) s_hist[i] = 0u;
        __syncthreads();
    }
    if (hi_bit > 0)
        for_each_sample(v, bin, l, [&](double x) {
            atomicAdd(&h[(unsigned int)((to_key(x) >> lo_bit) & digit_mask)], 1u);
        });
    if (SMEM) {
        __syncthreads();
        unsigned int *g = hist + (size_t)bin * v.n_l * kDigits;
        for (int i = tid; i < v.n_l * kDigits; i += nthr)
            if (s_hist[i]) atomicAdd(&g[i], s_hist[i]);
    }
}

struct SelectState {
    unsigned long long *prefix;     // [cols][R]  key with the decided bits, undecided bits zero
    long long *remaining;           // [cols][R]  rank inside the decided prefix
    unsigned int *count;            // [cols][R]  keys sharing the decided prefix
    const int *top;                 // [cols]
    int n_ranks;
};

__global__ void __launch_bounds__(256) first_scan_kernel(SelectState st, const unsigned int *__restrict__ hist) {
    const size_t col = blockIdx.x;
    const int r = blockIdx.y;
    const size_t i = col * st.n_ranks + r;
    const int hi_bit = st.top[col];
    if (hi_bit <= 0) {                                   // constant column: the common prefix is the key
        if (threadIdx.x == 0) st.count[i] = 0u;
        return;
    }
    const int lo_bit = hi_bit > kDigitBits ? hi_bit - kDigitBits : 0;
    const unsigned int *h = hist + col * kDigits;
    __shared__ unsigned int part[256];
    const int per = kDigits / 256;
    unsigned int sum = 0;
    for (int k = 0; k < per; ++k) sum += h[threadIdx.x * per + k];
    part[threadIdx.x] = sum;
    __syncthreads();
    if (threadIdx.x == 0) {
        const long long want = st.remaining[i];
        long long before = 0;
        int t = 0;
        while (t < 255 && before + (long long)part[t] <= want) { before += part[t]; ++t; }
        int d = t * per;
        while (d < t * per + per - 1 && before + (long long)h[d] <= want) { before += h[d]; ++d; }
        st.prefix[i] |= (unsigned long long)d << lo_bit;
        st.remaining[i] = want - before;
        st.count[i] = h[d];
    }
}

// lead[cols][R]: 1 if this rank owns its bucket's buffer (ranks in the same bucket share the
// first one's); offset[cols][R]: start of the buffer; cursor[cols][R]: append position.
// A sample is in play if its first digit is that of a bucket in play (every key of a column shares the bits above
// top[col], so the digit identifies the bucket): a byte per digit and column in shared memory holds the owning
// rank + 1, or 0 for the ~90-99 % of the samples that are in no bucket.  (Comparing every key with up to 12
// prefixes made this pass instruction bound.)
__global__ void __launch_bounds__(256) gather_kernel(ChainView v, SelectState st, const unsigned char *__restrict__ lead,
                                                     const long long *__restrict__ offset, unsigned int *__restrict__ cursor,
                                                     unsigned long long *__restrict__ buf) {
    extern __shared__ unsigned char s_slot[];            // [L][4096]
    const int l = threadIdx.x, bin = blockIdx.x;
    const int tid = threadIdx.y * blockDim.x + threadIdx.x, nthr = blockDim.x * blockDim.y;
    const size_t col = (size_t)bin * v.n_l + l;
    const int hi_bit = st.top[col];
    const bool active = hi_bit > kDigitBits;             // else nothing is left to decide below the first digit
    const int lo_bit = active ? hi_bit - kDigitBits : 0;
    for (int i = tid; i < v.n_l * (kDigits / 4); i += nthr) reinterpret_cast<unsigned int *>(s_slot)[i] = 0u;
    __syncthreads();
    const size_t base = col * st.n_ranks;
    unsigned char *slot = s_slot + (size_t)l * kDigits;
    if (active && threadIdx.y == 0)
        for (int r = 0; r < st.n_ranks; ++r)
            if (lead[base + r]) slot[(unsigned int)(st.prefix[base + r] >> lo_bit) & (kDigits - 1)] = (unsigned char)(r + 1);
    __syncthreads();
    if (!active) return;
    unsigned int *cur = cursor + base;
    const long long *off = offset + base;
    for_each_sample(v, bin, l, [&](double x) {
        const unsigned long long key = to_key(x);
        const unsigned int r1 = slot[(unsigned int)(key >> lo_bit) & (kDigits - 1)];
        if (r1) buf[off[r1 - 1] + atomicAdd(&cur[r1 - 1], 1u)] = key;
    });
}

__global__ void __launch_bounds__(256) finish_kernel(SelectState st, const long long *__restrict__ offset,
                                                     const unsigned long long *__restrict__ buf) {
    const size_t col = blockIdx.x;
    const size_t i = col * st.n_ranks + blockIdx.y;
    int hi = st.top[col] - kDigitBits;                   // bits [0, hi) are still undecided
    if (hi <= 0) return;
    const unsigned long long *b = buf + offset[i];
    const unsigned int n = st.count[i];
    __shared__ unsigned int h[256];
    __shared__ unsigned long long s_pre;
    __shared__ long long s_want;
    if (threadIdx.x == 0) { s_pre = st.prefix[i]; s_want = st.remaining[i]; }
    __syncthreads();
    for (; hi > 0; hi -= 8) {
        const int lo = hi > 8 ? hi - 8 : 0;
        const unsigned long long mask = (1ull << (hi - lo)) - 1ull;
        const unsigned long long above = ~0ull << hi;
        h[threadIdx.x] = 0u;
        __syncthreads();
        const unsigned long long pre = s_pre;
        for (unsigned int j = threadIdx.x; j < n; j += blockDim.x) {
            const unsigned long long key = b[j];
            if ((key & above) == (pre & above)) atomicAdd(&h[(unsigned int)((key >> lo) & mask)], 1u);
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            long long before = 0;
            int d = 0;
            while (d < 255 && before + (long long)h[d] <= s_want) { before += h[d]; ++d; }
            s_pre = pre | ((unsigned long long)d << lo);
            s_want -= before;
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) st.prefix[i] = s_pre;
}

}  // namespace

static dim3 summary_block(int n_l, int threads = 256) {
    int ty = threads / n_l;
    if (ty < 1) ty = 1;
    return dim3((unsigned)n_l, (unsigned)ty);
}

// grid = (bins, chunks): about `per_sm` CTAs per SM and `rounds` rounds of them, so that the last round's
// partial occupancy costs little.
static dim3 summary_grid(int bins, long long kept, int sm_count, int per_sm = 8, int rounds = 2) {
    long long chunks = ((long long)per_sm * rounds * sm_count + bins - 1) / (bins > 0 ? bins : 1);
    if (chunks > kept) chunks = kept;
    if (chunks < 1) chunks = 1;
    if (chunks > 65535) chunks = 65535;
    return dim3((unsigned)bins, (unsigned)chunks);
}

// Dynamic shared memory a kernel may use (opt-in limit minus its static part), raised once per kernel.
template <typename K>
static size_t smem_room(K kernel) {
    int dev = 0, optin = 0;
    cudaFuncAttributes fa;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev) != cudaSuccess ||
        cudaFuncGetAttributes(&fa, reinterpret_cast<const void *>(kernel)) != cudaSuccess)
        return 0;
    const size_t room = (size_t)optin - fa.sharedSizeBytes;
    if (cudaFuncSetAttribute(reinterpret_cast<const void *>(kernel), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)room) != cudaSuccess)
        return 0;
    return room;
}

cudaError_t launch_chain_minmax(const double *chain, long long k0, long long k1, int bins, int walkers, int n_l,
                                unsigned long long *minmax, int sm_count, cudaStream_t stream) {
    if (bins <= 0 || k1 <= k0) return cudaSuccess;
    if (n_l > 32) return cudaErrorInvalidValue;
    ChainView v{chain, k0, k1, bins, walkers, n_l};
    minmax_kernel<<<summary_grid(bins, k1 - k0, sm_count), summary_block(n_l), 0, stream>>>(v, minmax);
    return cudaGetLastError();
}

cudaError_t launch_chain_hist(const double *chain, long long k0, long long k1, int bins, int walkers, int n_l,
                              const unsigned long long *minmax, double *edges, int n_bins, unsigned int *hist,
                              int *peak_bin, long long *peak_cum, int sm_count, cudaStream_t stream) {
    if (bins <= 0 || k1 <= k0) return cudaSuccess;
    if (n_l > 32) return cudaErrorInvalidValue;
    ChainView v{chain, k0, k1, bins, walkers, n_l};
    const size_t cols = (size_t)bins * n_l;
    cudaError_t e = cudaMemsetAsync(hist, 0, sizeof(unsigned int) * cols * (size_t)n_bins, stream);
    if (e != cudaSuccess) return e;
    edges_kernel<<<(unsigned)cols, 256, 0, stream>>>(minmax, n_bins, edges);
    static const size_t room = smem_room(hist_kernel<true>);
    const size_t need = sizeof(unsigned int) * (size_t)n_l * (size_t)n_bins;
    if (need <= room && need <= 56 * 1024)               // four or more CTAs per SM next to the private counts
        hist_kernel<true><<<summary_grid(bins, k1 - k0, sm_count, 4, 3), summary_block(n_l), need, stream>>>(v, edges, n_bins, hist);
    else
        hist_kernel<false><<<summary_grid(bins, k1 - k0, sm_count), summary_block(n_l), 0, stream>>>(v, edges, n_bins, hist);
    peak_kernel<<<(unsigned)cols, 256, 0, stream>>>(hist, n_bins, peak_bin, peak_cum);
    return cudaGetLastError();
}

cudaError_t launch_select_first(const double *chain, long long k0, long long k1, int bins, int walkers, int n_l,
                                unsigned long long *prefix, long long *remaining, unsigned int *count, const int *top,
                                unsigned int *hist, int n_ranks, int sm_count, cudaStream_t stream) {
    if (bins <= 0 || k1 <= k0 || n_ranks <= 0) return cudaSuccess;
    if (n_l > 32 || n_ranks > kMaxRanks) return cudaErrorInvalidValue;
    ChainView v{chain, k0, k1, bins, walkers, n_l};
    SelectState st{prefix, remaining, count, top, n_ranks};
    const size_t cols = (size_t)bins * n_l;
    cudaError_t e = cudaMemsetAsync(hist, 0, sizeof(unsigned int) * cols * kDigits, stream);
    if (e != cudaSuccess) return e;
    static const size_t room = smem_room(first_hist_kernel<true>);
    const size_t need = sizeof(unsigned int) * (size_t)n_l * kDigits;
    if (need <= room)                                    // one CTA of up to 1024 threads per SM
        first_hist_kernel<true><<<summary_grid(bins, k1 - k0, sm_count, 1, 8), summary_block(n_l, 1024), need, stream>>>(v, top, hist);
    else
        first_hist_kernel<false><<<summary_grid(bins, k1 - k0, sm_count), summary_block(n_l), 0, stream>>>(v, top, hist);
    first_scan_kernel<<<dim3((unsigned)cols, (unsigned)n_ranks), 256, 0, stream>>>(st, hist);
    return cudaGetLastError();
}

cudaError_t launch_select_finish(const double *chain, long long k0, long long k1, int bins, int walkers, int n_l,
                                 unsigned long long *prefix, long long *remaining, unsigned int *count, const int *top,
                                 const unsigned char *lead, const long long *offset, unsigned int *cursor,
                                 unsigned long long *buf, int n_ranks, int sm_count, cudaStream_t stream) {
    if (bins <= 0 || k1 <= k0 || n_ranks <= 0) return cudaSuccess;
    ChainView v{chain, k0, k1, bins, walkers, n_l};
    SelectState st{prefix, remaining, count, top, n_ranks};
    const size_t cols = (size_t)bins * n_l;
    cudaError_t e = cudaMemsetAsync(cursor, 0, sizeof(unsigned int) * cols * n_ranks, stream);
    if (e != cudaSuccess) return e;
    static const size_t room = smem_room(gather_kernel);
    const size_t need = (size_t)n_l * kDigits;
    if (need > room) return cudaErrorInvalidValue;       // n_l <= 32 is 128 KB
    gather_kernel<<<summary_grid(bins, k1 - k0, sm_count, 4, 3), summary_block(n_l), need, stream>>>(v, st, lead, offset, cursor, buf);
    static const bool timing = getenv("MDA_SUMMARY_TIMING") != nullptr;
    cudaEvent_t ev[3] = {nullptr, nullptr, nullptr};
    if (timing) { for (auto &x : ev) cudaEventCreate(&x); cudaEventRecord(ev[1], stream); }
    finish_kernel<<<dim3((unsigned)cols, (unsigned)n_ranks), 256, 0, stream>>>(st, offset, buf);
    if (timing) {
        cudaEventRecord(ev[2], stream);
        cudaEventSynchronize(ev[2]);
        float ms = 0.f;
        cudaEventElapsedTime(&ms, ev[1], ev[2]);
        fprintf(stderr, "[mda summary]     finish_kernel alone        %8.3f ms\n", ms);
        for (auto &x : ev) cudaEventDestroy(x);
    }
    return cudaGetLastError();
}

}  // namespace mda
