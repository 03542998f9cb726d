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
//   select_*          exact order statistics by most-significant-digit radix selection,
//                     12 bits per pass, only below the bits min and max share; up to
//                     kMaxRanks ranks per column advance together
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

constexpr int kDigitBits = 12;
constexpr int kDigits = 1 << kDigitBits;

// State of one radix selection: per (column, rank) the key bits fixed so far and the rank
// that remains inside that prefix.  `top` is the number of low key bits still undecided
// for the column (bits above are shared by min and max); pass p fixes bits
// [top - 12(p+1), top - 12p).
struct SelectState {
    unsigned long long *prefix;     // [cols][R]  key with the decided bits, undecided bits zero
    long long *remaining;           // [cols][R]
    const int *top;                 // [cols]
    unsigned int *hist;             // [cols][R][kDigits]
    int n_ranks;
};

template <int NR>
__global__ void __launch_bounds__(256) select_hist_kernel(ChainView v, SelectState st, int pass) {
    const int l = threadIdx.x, bin = blockIdx.x;
    const size_t col = (size_t)bin * v.n_l + l;
    const int hi_bit = st.top[col] - kDigitBits * pass;         // bits [lo_bit, hi_bit) are decided in this pass
    if (hi_bit <= 0) return;
    const int lo_bit = hi_bit > kDigitBits ? hi_bit - kDigitBits : 0;
    const unsigned long long digit_mask = (hi_bit - lo_bit >= 64) ? ~0ull : ((1ull << (hi_bit - lo_bit)) - 1ull);
    const unsigned long long above = hi_bit >= 64 ? 0ull : (~0ull << hi_bit);
    // the ranks' prefixes live in registers (NR is a compile-time bound, loops fully unrolled);
    // a rank whose prefix equals an earlier rank's shares that rank's histogram
    unsigned long long pre[NR];
    unsigned int lead = 0;
    unsigned long long pre_min = ~0ull, pre_max = 0ull;
#pragma unroll
    for (int r = 0; r < NR; ++r) {
        pre[r] = r < st.n_ranks ? (st.prefix[col * st.n_ranks + r] & above) : ~0ull;
        bool is_lead = r < st.n_ranks;
#pragma unroll
        for (int q = 0; q < r; ++q) is_lead = is_lead && (pre[q] != pre[r]);
        if (is_lead) {
            lead |= 1u << r;
            pre_min = pre[r] < pre_min ? pre[r] : pre_min;
            pre_max = pre[r] > pre_max ? pre[r] : pre_max;
        }
    }
    unsigned int *h = st.hist + col * st.n_ranks * kDigits;
    for_each_sample(v, bin, l, [&](double x) {
        const unsigned long long key = to_key(x);
        const unsigned long long kp = key & above;
        if (kp < pre_min || kp > pre_max) return;                // outside every bucket still in play
        const unsigned int digit = (unsigned int)((key >> lo_bit) & digit_mask);
#pragma unroll
        for (int r = 0; r < NR; ++r)
            if ((lead >> r & 1u) && kp == pre[r]) atomicAdd(&h[(size_t)r * kDigits + digit], 1u);
    });
}

// One CTA per (column, rank): locate the digit whose cumulative count passes the remaining
// rank, fix it in the prefix, reduce the rank, clear the histogram for the next pass.
__global__ void __launch_bounds__(256) select_scan_kernel(SelectState st, int pass, int n_l) {
    const size_t col = blockIdx.x;
    const int r = blockIdx.y;
    const int hi_bit = st.top[col] - kDigitBits * pass;
    if (hi_bit <= 0) return;
    const int lo_bit = hi_bit > kDigitBits ? hi_bit - kDigitBits : 0;
    const unsigned long long mine = st.prefix[col * st.n_ranks + r];
    int leader = r;
    for (int q = 0; q < r; ++q)
        if (st.prefix[col * st.n_ranks + q] == mine) { leader = q; break; }
    const unsigned int *h = st.hist + (col * st.n_ranks + leader) * kDigits;
    __shared__ unsigned int part[256];
    __shared__ int found_digit;
    __shared__ long long found_before;
    const int per = kDigits / 256;
    unsigned int sum = 0;
    for (int i = 0; i < per; ++i) sum += h[threadIdx.x * per + i];
    part[threadIdx.x] = sum;
    __syncthreads();
    if (threadIdx.x == 0) {
        const long long want = st.remaining[col * st.n_ranks + r];
        long long before = 0;
        int t = 0;
        while (t < 255 && before + (long long)part[t] <= want) { before += part[t]; ++t; }
        int d = t * per;
        while (d < t * per + per - 1 && before + (long long)h[d] <= want) { before += h[d]; ++d; }
        found_digit = d;
        found_before = before;
    }
    __syncthreads();
    (void)n_l;
    // every rank of the column must have read the shared leader histogram before anyone
    // rewrites prefixes: the new state goes to a shadow slot, committed by select_commit_kernel
    if (threadIdx.x == 0) {
        st.prefix[(size_t)gridDim.x * st.n_ranks + col * st.n_ranks + r] = mine | ((unsigned long long)found_digit << lo_bit);
        st.remaining[(size_t)gridDim.x * st.n_ranks + col * st.n_ranks + r] = st.remaining[col * st.n_ranks + r] - found_before;
    }
}

__global__ void select_commit_kernel(SelectState st, size_t n_state, int pass, int n_cols) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n_state) {
        const size_t col = i / st.n_ranks;
        if (st.top[col] - kDigitBits * pass > 0) {
            st.prefix[i] = st.prefix[n_state + i];
            st.remaining[i] = st.remaining[n_state + i];
        }
    }
    (void)n_cols;
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
                              const double *edges, int n_bins, unsigned int *hist, int sm_count, cudaStream_t stream) {
    if (bins <= 0 || k1 <= k0) return cudaSuccess;
    if (n_l > 32) return cudaErrorInvalidValue;
    ChainView v{chain, k0, k1, bins, walkers, n_l};
    hist_kernel<<<summary_grid(bins, k1 - k0, sm_count), summary_block(n_l), 0, stream>>>(v, edges, n_bins, hist);
    return cudaGetLastError();
}

cudaError_t launch_chain_select(const double *chain, long long k0, long long k1, int bins, int walkers, int n_l,
                                unsigned long long *prefix, long long *remaining, const int *top, unsigned int *hist,
                                int n_ranks, int max_top, int sm_count, cudaStream_t stream) {
    if (bins <= 0 || k1 <= k0 || n_ranks <= 0) return cudaSuccess;
    if (n_l > 32 || n_ranks > kMaxRanks) return cudaErrorInvalidValue;
    ChainView v{chain, k0, k1, bins, walkers, n_l};
    SelectState st{prefix, remaining, top, hist, n_ranks};
    const size_t cols = (size_t)bins * n_l, n_state = cols * n_ranks;
    const int passes = (max_top + kDigitBits - 1) / kDigitBits;
    for (int p = 0; p < passes; ++p) {
        cudaError_t e = cudaMemsetAsync(hist, 0, sizeof(unsigned int) * n_state * kDigits, stream);
        if (e != cudaSuccess) return e;
        if (n_ranks <= 6) select_hist_kernel<6><<<summary_grid(bins, k1 - k0, sm_count), summary_block(n_l), 0, stream>>>(v, st, p);
        else select_hist_kernel<kMaxRanks><<<summary_grid(bins, k1 - k0, sm_count), summary_block(n_l), 0, stream>>>(v, st, p);
        select_scan_kernel<<<dim3((unsigned)cols, (unsigned)n_ranks), 256, 0, stream>>>(st, p, n_l);
        select_commit_kernel<<<(unsigned)((n_state + 255) / 256), 256, 0, stream>>>(st, n_state, p, (int)cols);
    }
    return cudaGetLastError();
}

}  // namespace mda
