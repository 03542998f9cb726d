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

namespace mda {

namespace {

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

// Visit every sample of column (bin, l) owned by this thread: four independent loads are in
// flight per thread, enough bytes in flight per SM to run at HBM rate with ~1000 threads.
template <typename F>
__device__ __forceinline__ void for_each_sample(const ChainView &v, int bin, int l, F visit) {
    const int stride = blockDim.y;
    for (long long k = v.k0 + blockIdx.y; k < v.k1; k += gridDim.y) {
        const double *slab = v.chain + ((size_t)k * v.bins + bin) * v.walkers * v.n_l + l;
        int w = threadIdx.y;
        for (; w + 3 * stride < v.walkers; w += 4 * stride) {
            const double x0 = slab[(size_t)w * v.n_l];
            const double x1 = slab[(size_t)(w + stride) * v.n_l];
            const double x2 = slab[(size_t)(w + 2 * stride) * v.n_l];
            const double x3 = slab[(size_t)(w + 3 * stride) * v.n_l];
            visit(x0); visit(x1); visit(x2); visit(x3);
        }
        for (; w < v.walkers; w += stride) visit(slab[(size_t)w * v.n_l]);
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
__global__ void __launch_bounds__(256) hist_kernel(ChainView v, const double *__restrict__ edges, int n_bins,
                                                   unsigned int *__restrict__ hist) {
    const int l = threadIdx.x, bin = blockIdx.x;
    const size_t col = (size_t)bin * v.n_l + l;
    const double *e = edges + col * (size_t)(n_bins + 1);
    const double first = e[0], last = e[n_bins];
    const double norm = (double)n_bins / (last - first);
    unsigned int *h = hist + col * (size_t)n_bins;
    if (!(last > first)) return;
    for_each_sample(v, bin, l, [&](double x) {
        // any estimate within one bin of the truth will do: numpy corrects its own estimate
        // against the edges in exactly this way
        int idx = (int)((x - first) * norm);
        idx = idx < 0 ? 0 : (idx >= n_bins ? n_bins - 1 : idx);
        if (x < e[idx]) idx -= 1;
        else if (idx != n_bins - 1 && x >= e[idx + 1]) idx += 1;
        atomicAdd(&h[idx], 1u);
    });
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

constexpr int kDigitBits = 12;
constexpr int kDigits = 1 << kDigitBits;

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
__global__ void __launch_bounds__(256) first_hist_kernel(ChainView v, const int *__restrict__ top,
                                                         unsigned int *__restrict__ hist) {
    const int l = threadIdx.x, bin = blockIdx.x;
    const size_t col = (size_t)bin * v.n_l + l;
    const int hi_bit = top[col];
    if (hi_bit <= 0) return;
    const int lo_bit = hi_bit > kDigitBits ? hi_bit - kDigitBits : 0;
    const unsigned long long digit_mask = (1ull << (hi_bit - lo_bit)) - 1ull;
    unsigned int *h = hist + col * kDigits;
    for_each_sample(v, bin, l, [&](double x) {
        atomicAdd(&h[(unsigned int)((to_key(x) >> lo_bit) & digit_mask)], 1u);
    });
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
template <int NR>
__global__ void __launch_bounds__(256) gather_kernel(ChainView v, SelectState st, const unsigned char *__restrict__ lead,
                                                     const long long *__restrict__ offset, unsigned int *__restrict__ cursor,
                                                     unsigned long long *__restrict__ buf) {
    const int l = threadIdx.x, bin = blockIdx.x;
    const size_t col = (size_t)bin * v.n_l + l;
    const int hi_bit = st.top[col];
    if (hi_bit <= kDigitBits) return;                    // nothing left to decide below the first digit
    const unsigned long long above = ~0ull << (hi_bit - kDigitBits);
    unsigned long long pre[NR];
    long long off[NR];
    unsigned int is_lead = 0;
    unsigned long long pre_min = ~0ull, pre_max = 0ull;
#pragma unroll
    for (int r = 0; r < NR; ++r) {
        const bool on = r < st.n_ranks && lead[col * st.n_ranks + (r < st.n_ranks ? r : 0)];
        pre[r] = on ? st.prefix[col * st.n_ranks + r] : ~0ull;
        off[r] = on ? offset[col * st.n_ranks + r] : 0;
        if (on) {
            is_lead |= 1u << r;
            pre_min = pre[r] < pre_min ? pre[r] : pre_min;
            pre_max = pre[r] > pre_max ? pre[r] : pre_max;
        }
    }
    unsigned int *cur = cursor + col * st.n_ranks;
    for_each_sample(v, bin, l, [&](double x) {
        const unsigned long long key = to_key(x);
        const unsigned long long kp = key & above;
        if (kp < pre_min || kp > pre_max) return;        // outside every bucket in play
#pragma unroll
        for (int r = 0; r < NR; ++r)
            if ((is_lead >> r & 1u) && kp == pre[r]) buf[off[r] + atomicAdd(&cur[r], 1u)] = key;
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

static dim3 summary_block(int n_l) {
    int ty = 256 / n_l;
    if (ty < 1) ty = 1;
    return dim3((unsigned)n_l, (unsigned)ty);
}

static dim3 summary_grid(int bins, long long kept, int sm_count) {
    long long chunks = (4LL * sm_count + bins - 1) / (bins > 0 ? bins : 1);
    if (chunks > kept) chunks = kept;
    if (chunks < 1) chunks = 1;
    if (chunks > 65535) chunks = 65535;
    return dim3((unsigned)bins, (unsigned)chunks);
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
    hist_kernel<<<summary_grid(bins, k1 - k0, sm_count), summary_block(n_l), 0, stream>>>(v, edges, n_bins, hist);
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
    first_hist_kernel<<<summary_grid(bins, k1 - k0, sm_count), summary_block(n_l), 0, stream>>>(v, top, hist);
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
    if (n_ranks <= 6)
        gather_kernel<6><<<summary_grid(bins, k1 - k0, sm_count), summary_block(n_l), 0, stream>>>(v, st, lead, offset, cursor, buf);
    else
        gather_kernel<kMaxRanks><<<summary_grid(bins, k1 - k0, sm_count), summary_block(n_l), 0, stream>>>(v, st, lead, offset, cursor, buf);
    finish_kernel<<<dim3((unsigned)cols, (unsigned)n_ranks), 256, 0, stream>>>(st, offset, buf);
    return cudaGetLastError();
}

}  // namespace mda
