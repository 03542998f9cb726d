// chisq_abi.cu -- the C ABI of include/mda_chisq.h on top of the sm_100a kernels.
//
// Host-side runtime: handles, HBM store, pinned staging, per-shape CUDA graphs,
// error side channel.  No torch types, no CPU arithmetic path: every evaluation is a
// kernel launch, and without a usable device every entry point fails loudly.
#include "../../include/mda_chisq.h"
#include "mda_kernels.cuh"
#include "philox.cuh"

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <map>
#include <mutex>
#include <string>
#include <vector>

namespace {

using namespace mda;

// ---- error side channel -------------------------------------------------------------
thread_local int t_err_code = MDA_OK;
thread_local char t_err_msg[512] = "no error";

int set_error(int code, const char *fmt, ...) {
    t_err_code = code;
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(t_err_msg, sizeof(t_err_msg), fmt, ap);
    va_end(ap);
    if (getenv("MDA_VERBOSE")) fprintf(stderr, "[mda_chisq] error %d: %s\n", code, t_err_msg);
    return code;
}

int cuda_fail(cudaError_t e, const char *what) {
    const int code = (e == cudaErrorMemoryAllocation) ? MDA_ERR_OUT_OF_MEMORY
                     : (e == cudaErrorNoDevice || e == cudaErrorInsufficientDriver || e == cudaErrorInvalidDevice)
                         ? MDA_ERR_NO_DEVICE
                         : MDA_ERR_CUDA;
    return set_error(code, "%s: %s (%s)", what, cudaGetErrorString(e), cudaGetErrorName(e));
}

#define MDA_CUDA(call, what)                                   \
    do {                                                       \
        cudaError_t e__ = (call);                              \
        if (e__ != cudaSuccess) return cuda_fail(e__, what);   \
    } while (0)

std::atomic<unsigned long long> g_launches{0};

// ---- per-process runtime ------------------------------------------------------------
struct Runtime {
    std::mutex mu;
    bool ready = false;
    int requested = -1;      // mdaSetDevice
    int device = -1;
    int sm_count = 0;
    int cc = 0;
    size_t smem_optin = 0;
    char name[128] = {0};
};
Runtime g_rt;

int env_int(const char *name, int fallback) {
    const char *v = getenv(name);
    if (!v || !*v) return fallback;
    char *end = nullptr;
    long x = strtol(v, &end, 10);
    return (end && *end == 0) ? (int)x : fallback;
}

// Brings CUDA up on the chosen device the first time any handle is made -- never at
// library load (workers dlopen after fork, mda.py:1109).
int ensure_runtime() {
    std::lock_guard<std::mutex> lock(g_rt.mu);
    if (g_rt.ready) {
        cudaError_t e = cudaSetDevice(g_rt.device);
        if (e != cudaSuccess) return cuda_fail(e, "cudaSetDevice");
        return MDA_OK;
    }
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess) {
        cuda_fail(e, "cudaGetDeviceCount (this library has no CPU fallback)");
        return t_err_code = MDA_ERR_NO_DEVICE;
    }
    if (count <= 0) return set_error(MDA_ERR_NO_DEVICE, "no CUDA device visible; this library has no CPU fallback");
    int dev = g_rt.requested;
    if (dev < 0) dev = env_int("MDA_DEVICE", -1);
    if (dev < 0) {
        const int lr = env_int("LOCAL_RANK", -1);
        dev = lr >= 0 ? lr % count : 0;
    }
    if (dev >= count) return set_error(MDA_ERR_NO_DEVICE, "device %d requested but only %d visible", dev, count);
    MDA_CUDA(cudaSetDevice(dev), "cudaSetDevice");
    cudaDeviceProp prop;
    MDA_CUDA(cudaGetDeviceProperties(&prop, dev), "cudaGetDeviceProperties");
    if (prop.major < 10)
        return set_error(MDA_ERR_NO_DEVICE, "device %d (%s, sm_%d%d) is not Blackwell; kernels are built for sm_100a only",
                         dev, prop.name, prop.major, prop.minor);
    g_rt.device = dev;
    g_rt.sm_count = prop.multiProcessorCount;
    g_rt.cc = prop.major * 10 + prop.minor;
    // dynamic limit: the opt-in maximum minus room for the kernels' static shared memory (mbarrier)
    g_rt.smem_optin = prop.sharedMemPerBlockOptin > 1024 ? prop.sharedMemPerBlockOptin - 1024 : 0;
    snprintf(g_rt.name, sizeof(g_rt.name), "%s", prop.name);
    MDA_CUDA(configure_kernels(g_rt.smem_optin), "cudaFuncSetAttribute(max dynamic smem)");
    MDA_CUDA(configure_ensemble_kernels(g_rt.smem_optin), "cudaFuncSetAttribute(max dynamic smem, ensemble)");
    MDA_CUDA(configure_lnpost_mma_kernels(g_rt.smem_optin), "cudaFuncSetAttribute(max dynamic smem, DMMA batch)");
    MDA_CUDA(configure_ensemble_mma_kernels(g_rt.smem_optin), "cudaFuncSetAttribute(max dynamic smem, DMMA ensemble)");
    MDA_CUDA(configure_ensemble_ws_kernels(g_rt.smem_optin), "cudaFuncSetAttribute(max dynamic smem, warp-specialised DMMA ensemble)");
    MDA_CUDA(configure_ensemble_quad_kernels(g_rt.smem_optin), "cudaFuncSetAttribute(max dynamic smem, triangular-factor ensemble)");
    MDA_CUDA(configure_ensemble_qws_kernels(g_rt.smem_optin), "cudaFuncSetAttribute(max dynamic smem, warp-specialised triangular-factor ensemble)");
    g_rt.ready = true;
    return MDA_OK;
}

// ---- legacy single-bin handle ---------------------------------------------------------
constexpr uint32_t kStructMagic = 0x4D444131u;  // "MDA1"
constexpr uint32_t kStoreMagic = 0x4D444153u;   // "MDAS"
constexpr uint32_t kEnsembleMagic = 0x4D444145u; // "MDAE"

struct MdaHandle {
    uint32_t magic = kStructMagic;
    int n = 0, n_l = 0;
    std::vector<double> host;        // [1+L][n]: row 0 data, rows 1..L dists (reference row layout)
    double *d_consts = nullptr;      // same layout in HBM
    double *d_resid = nullptr;       // [n]
    double *h_io = nullptr;          // mapped pinned: [0..L) params, [L] result, [L+1] ticket of the last finished call
    unsigned long long ticket = 0;   // calls made on this handle
    double *d_io = nullptr;          // device alias of h_io
    cudaStream_t stream = nullptr;
    bool dirty = true;
};

MdaHandle *as_struct(void *p, const char *fn) {
    MdaHandle *h = static_cast<MdaHandle *>(p);
    if (!h || h->magic != kStructMagic) {
        set_error(MDA_ERR_BAD_ARGUMENT, "%s: not a handle from makeMdaStruct", fn);
        return nullptr;
    }
    return h;
}

void destroy_struct(MdaHandle *h) {
    if (h->d_consts) cudaFree(h->d_consts);
    if (h->d_resid) cudaFree(h->d_resid);
    if (h->h_io) cudaFreeHost(h->h_io);
    if (h->stream) cudaStreamDestroy(h->stream);
    h->magic = 0;
    delete h;
}

double scalar_eval(void *p, double *params, double *resid_out, int mode, const char *fn) {
    MdaHandle *h = as_struct(p, fn);
    if (!h) return NAN;
    if (!params) { set_error(MDA_ERR_BAD_ARGUMENT, "%s: params is NULL", fn); return NAN; }
    if (ensure_runtime() != MDA_OK) return NAN;
    cudaError_t e;
    if (h->dirty) {
        e = cudaMemcpyAsync(h->d_consts, h->host.data(), sizeof(double) * h->host.size(), cudaMemcpyHostToDevice,
                            h->stream);
        if (e != cudaSuccess) { cuda_fail(e, "upload of struct constants"); return NAN; }
        h->dirty = false;
    }
    memcpy(h->h_io, params, sizeof(double) * (size_t)h->n_l);
    h->h_io[h->n_l] = NAN;
    volatile unsigned long long *done = reinterpret_cast<volatile unsigned long long *>(h->h_io + h->n_l + 1);
    const unsigned long long ticket = ++h->ticket;
    e = launch_scalar(h->d_consts, h->n, h->n_l, h->h_io, h->d_io, h->d_io + h->n_l, h->d_resid, mode,
                      reinterpret_cast<unsigned long long *>(h->d_io + h->n_l + 1), ticket, h->stream);
    if (e != cudaSuccess) { cuda_fail(e, "scalar kernel launch"); return NAN; }
    g_launches.fetch_add(1, std::memory_order_relaxed);
    if (resid_out && h->n > 0) {                          // (the tester's entry point: a copy follows, so synchronise)
        e = cudaMemcpyAsync(resid_out, h->d_resid, sizeof(double) * (size_t)h->n, cudaMemcpyDeviceToHost, h->stream);
        if (e != cudaSuccess) { cuda_fail(e, "residual read-back"); return NAN; }
        e = cudaStreamSynchronize(h->stream);
        if (e != cudaSuccess) { cuda_fail(e, fn); return NAN; }
        return h->h_io[h->n_l];
    }
    // The kernel writes the result, fences at system scope and then writes this call's ticket into mapped memory: spinning on
    // it costs a PCIe write's latency, a stream synchronisation several microseconds of driver wake-up.  If the ticket does not
    // show up within 2 ms (a fault, a device busy with something long) the synchronisation decides and reports.
    const auto t0 = std::chrono::steady_clock::now();
    for (unsigned spins = 0; *done != ticket; ++spins) {
        if ((spins & 1023u) == 1023u && std::chrono::steady_clock::now() - t0 > std::chrono::milliseconds(2)) {
            e = cudaStreamSynchronize(h->stream);
            if (e != cudaSuccess) { cuda_fail(e, fn); return NAN; }
            break;
        }
    }
    std::atomic_thread_fence(std::memory_order_acquire);
    return h->h_io[h->n_l];
}

// ---- multi-bin device store ---------------------------------------------------------
struct GraphKey {
    int walkers, mode;
    bool operator<(const GraphKey &o) const { return walkers != o.walkers ? walkers < o.walkers : mode < o.mode; }
};

struct MdaStore {
    uint32_t magic = kStoreMagic;
    int bins = 0, n_l = 0, max_pts = 0, n_pad = 0;
    std::vector<double> host;        // [bins][n_pad/2][1+L][2]  (the HBM layout)
    std::vector<double> host_mma;    // [bins][n_tiles][8 (1+L)]  fragment-major copy for the DMMA stepper
    int n_tiles = 0;                 // angle tiles of 8
    std::vector<double> host_quad;   // [bins][quad_stride(L)]  packed triangular factor of [D^T | d] (see factor_bin)
    std::vector<double> kappa;       // [bins]  |d| / sqrt(min chi^2): how much rounding the triangular-factor chi^2 amplifies (see factor_bin)
    bool factors_stale = true;       // host_quad / kappa do not reflect the host mirror yet
    unsigned long long generation = 1;   // bumped by every change of a bin or of the bounds: ensembles remember the one their state belongs to
    int live_ensembles = 0;          // ensembles created on this store and not yet freed
    bool closing = false;            // mdaStoreFree was called while ensembles were alive: freed with the last of them
    double *d_quad = nullptr;
    std::vector<int> n_pts;          // [bins]
    std::vector<double> lo, hi;      // [L]
    double *d_consts = nullptr;
    double *d_consts_mma = nullptr;
    int *d_npts = nullptr;
    double *d_lo = nullptr, *d_hi = nullptr;
    bool dirty = true;
    cudaStream_t stream = nullptr;
    int force_threads = 0, force_wpt = 0;
    // staging for the host-pointer and staged entry points
    size_t cap_walkers = 0;          // walkers per bin the buffers below can hold
    double *h_params = nullptr, *h_out = nullptr;   // pinned
    double *d_params = nullptr, *d_out = nullptr;
    std::map<GraphKey, cudaGraphExec_t> graphs;
};

MdaStore *as_store(void *p, const char *fn) {
    MdaStore *s = static_cast<MdaStore *>(p);
    if (!s || s->magic != kStoreMagic || s->closing) {
        set_error(MDA_ERR_BAD_ARGUMENT, "%s: not a (live) handle from mdaStoreCreate", fn);
        return nullptr;
    }
    return s;
}

void drop_graphs(MdaStore *s) {
    for (auto &kv : s->graphs) cudaGraphExecDestroy(kv.second);
    s->graphs.clear();
}

size_t bin_stride(const MdaStore *s) { return (size_t)s->n_pad * (size_t)(s->n_l + 1); }

void write_bin(MdaStore *s, int bin, int n, const double *data, const double *dists, size_t dist_stride) {
    double *blk = s->host.data() + (size_t)bin * bin_stride(s);
    std::fill(blk, blk + bin_stride(s), 0.0);
    const int row = s->n_l + 1;
    for (int j = 0; j < n; ++j) {
        double *cell = blk + ((size_t)(j >> 1) * row) * 2 + (j & 1);
        cell[0] = data[j];
        for (int l = 0; l < s->n_l; ++l) cell[(size_t)(1 + l) * 2] = dists[(size_t)l * dist_stride + j];
    }
    // fragment-major copy (mda_kernels.cuh): per angle tile the rows the DFMA prologue reads, then
    // the B fragments of mma.m8n8k4 (lane (g, tig) <- dist R + 4 ks + tig at angle 8 t + g); the
    // distributions are negated here so the device only multiply-accumulates
    const int n_l = s->n_l, r_lead = n_l % 4, k_steps = n_l / 4, ts = 8 * (n_l + 1);
    double *mblk = s->host_mma.data() + (size_t)bin * (size_t)s->n_tiles * ts;
    std::fill(mblk, mblk + (size_t)s->n_tiles * ts, 0.0);
    for (int j = 0; j < n; ++j) {
        double *tile = mblk + (size_t)(j >> 3) * ts;
        const int c = j & 7;
        tile[c] = data[j];
        for (int l = 0; l < r_lead; ++l) tile[8 * (1 + l) + c] = -dists[(size_t)l * dist_stride + j];
        for (int ks = 0; ks < k_steps; ++ks)
            for (int tig = 0; tig < 4; ++tig)
                tile[8 * (1 + r_lead) + 32 * ks + 4 * c + tig] = -dists[(size_t)(r_lead + 4 * ks + tig) * dist_stride + j];
    }
    s->n_pts[bin] = n;
    s->dirty = true;
    s->factors_stale = true;
    ++s->generation;
}

// chi^2(a) = |d - D^T a|^2 (C/ChiSquare.c:188-210) is |M [a; -1]|^2 with M = [D^T | d] (N x (L+1)), and M = Q T with
// T upper triangular gives chi^2(a) = |T [a; -1]|^2: (L+1)(L+2)/2 multiply-adds per evaluation instead of N (L+1),
// whatever N is.  No inverse is taken (rank-deficient bins are fine) and every term is a square, so there is no
// cancellation beyond the one inside a row, which is that of the residuals themselves.  T comes from Householder
// reflections in long double (64-bit mantissa) and is rounded to double once; packed by rows: T[i][i..L].
//
// What it costs in accuracy.  Both this form and the reference's angle-by-angle sum round every partial result of a
// residual at the size of the DATA (eps |d_j|), while the residual that survives is smaller by the relative error of the
// measurement; chi^2 computed either way is therefore only good to about eps * kappa relative, kappa = |d| / sqrt(chi^2),
// and two different orders of the same arithmetic differ from each other by that much (measured against the unmodified
// reference library: 1.0 eps_machine * kappa -- 5e-15 at 5 % errors, 2e-13 at 0.1 %, 2e-12 at 0.01 %).  The angle-by-angle
// kernels in the reference's operation order reproduce its rounding itself and are immune.  Returned: kappa at the
// unconstrained minimum, |d| / |T[L][L]| (the largest it gets; infinite when the bin can be fitted exactly).  Stores with
// a bin beyond kQuadKappaMax are stepped by the reference-order stepper instead (ensemble_shape_for).
constexpr double kQuadKappaMax = 1e-13 / (4.0 * 1.1102230246251565e-16);      // 4 eps kappa <= 1e-13: a tenth of north_star's 1e-12
double factor_bin(const MdaStore *s, int bin, double *packed) {
    const int n = s->n_pts[bin], L = s->n_l, C = L + 1, row = L + 1;
    const double *blk = s->host.data() + (size_t)bin * bin_stride(s);
    const int rows = std::max(n, C);
    std::vector<long double> m((size_t)rows * C, 0.0L);
    for (int j = 0; j < n; ++j) {
        const double *cell = blk + ((size_t)(j >> 1) * row) * 2 + (j & 1);
        for (int l = 0; l < L; ++l) m[(size_t)j * C + l] = cell[(size_t)(1 + l) * 2];
        m[(size_t)j * C + L] = cell[0];
    }
    for (int k = 0; k < C; ++k) {                                    // Householder reflection k zeroes column k below the diagonal
        long double norm2 = 0.0L;
        for (int j = k; j < rows; ++j) norm2 += m[(size_t)j * C + k] * m[(size_t)j * C + k];
        if (norm2 == 0.0L) continue;
        const long double norm = sqrtl(norm2), x0 = m[(size_t)k * C + k];
        const long double alpha = x0 > 0.0L ? -norm : norm;
        std::vector<long double> v((size_t)rows - k);
        for (int j = k; j < rows; ++j) v[j - k] = m[(size_t)j * C + k];
        v[0] -= alpha;
        long double vv = 0.0L;
        for (long double x : v) vv += x * x;
        if (vv == 0.0L) continue;
        for (int c = k; c < C; ++c) {
            long double dot = 0.0L;
            for (int j = k; j < rows; ++j) dot += v[j - k] * m[(size_t)j * C + c];
            const long double f = 2.0L * dot / vv;
            for (int j = k; j < rows; ++j) m[(size_t)j * C + c] -= f * v[j - k];
        }
    }
    size_t o = 0;
    for (int i = 0; i < C; ++i)
        for (int c = i; c < C; ++c) packed[o++] = (double)m[(size_t)i * C + c];
    long double d2 = 0.0L;                                           // |Q^T d|^2 = |d|^2: the last column of T
    for (int i = 0; i < C; ++i) d2 += m[(size_t)i * C + L] * m[(size_t)i * C + L];
    const long double floor_chi = fabsl(m[(size_t)L * C + L]);       // sqrt of the unconstrained minimum of chi^2
    if (d2 == 0.0L) return 1.0;                                      // an empty bin: chi^2 is exactly 0 everywhere
    return floor_chi > 0.0L ? (double)(sqrtl(d2) / floor_chi) : std::numeric_limits<double>::infinity();
}

// host side of upload_store: triangular factors and their conditioning, from the host mirror
void ensure_factors(MdaStore *s) {
    if (!s->factors_stale) return;
    s->kappa.assign((size_t)s->bins, 1.0);
    for (int b = 0; b < s->bins; ++b) s->kappa[(size_t)b] = factor_bin(s, b, s->host_quad.data() + (size_t)b * (size_t)quad_stride(s->n_l));
    s->factors_stale = false;
}

// every bin's chi^2 from the triangular factor stays within 4 eps kappa <= 1e-13 of the reference's arithmetic
bool quad_is_safe(MdaStore *s) {
    ensure_factors(s);
    for (double k : s->kappa) if (!(k <= kQuadKappaMax)) return false;
    return true;
}

int upload_store(MdaStore *s) {
    if (!s->dirty) return MDA_OK;
    ensure_factors(s);
    MDA_CUDA(cudaMemcpyAsync(s->d_quad, s->host_quad.data(), sizeof(double) * s->host_quad.size(), cudaMemcpyHostToDevice,
                             s->stream), "store upload (triangular factors)");
    MDA_CUDA(cudaMemcpyAsync(s->d_consts, s->host.data(), sizeof(double) * s->host.size(), cudaMemcpyHostToDevice,
                             s->stream), "store upload (constants)");
    MDA_CUDA(cudaMemcpyAsync(s->d_consts_mma, s->host_mma.data(), sizeof(double) * s->host_mma.size(), cudaMemcpyHostToDevice,
                             s->stream), "store upload (fragment-major constants)");
    MDA_CUDA(cudaMemcpyAsync(s->d_npts, s->n_pts.data(), sizeof(int) * s->n_pts.size(), cudaMemcpyHostToDevice,
                             s->stream), "store upload (angle counts)");
    MDA_CUDA(cudaMemcpyAsync(s->d_lo, s->lo.data(), sizeof(double) * s->lo.size(), cudaMemcpyHostToDevice, s->stream),
             "store upload (bounds)");
    MDA_CUDA(cudaMemcpyAsync(s->d_hi, s->hi.data(), sizeof(double) * s->hi.size(), cudaMemcpyHostToDevice, s->stream),
             "store upload (bounds)");
    // the host mirror is pageable: make sure the DMA engine is done with it before it can change
    MDA_CUDA(cudaStreamSynchronize(s->stream), "store upload");
    s->dirty = false;
    return MDA_OK;
}

int ensure_staging(MdaStore *s, int walkers) {
    if ((size_t)walkers <= s->cap_walkers) return MDA_OK;
    MDA_CUDA(cudaStreamSynchronize(s->stream), "staging resize");
    drop_graphs(s);
    if (s->h_params) cudaFreeHost(s->h_params);
    if (s->h_out) cudaFreeHost(s->h_out);
    if (s->d_params) cudaFree(s->d_params);
    if (s->d_out) cudaFree(s->d_out);
    s->h_params = s->h_out = s->d_params = s->d_out = nullptr;
    s->cap_walkers = 0;
    const size_t n_eval = (size_t)s->bins * (size_t)walkers;
    const size_t pbytes = sizeof(double) * n_eval * (size_t)s->n_l, obytes = sizeof(double) * n_eval;
    MDA_CUDA(cudaHostAlloc((void **)&s->h_params, pbytes ? pbytes : 8, cudaHostAllocDefault), "pinned params staging");
    MDA_CUDA(cudaHostAlloc((void **)&s->h_out, obytes ? obytes : 8, cudaHostAllocDefault), "pinned result staging");
    MDA_CUDA(cudaMalloc((void **)&s->d_params, pbytes ? pbytes : 8), "device params block");
    MDA_CUDA(cudaMalloc((void **)&s->d_out, obytes ? obytes : 8), "device result block");
    s->cap_walkers = (size_t)walkers;
    return MDA_OK;
}

TileShape tile_for(const MdaStore *s, int walkers) {
    return choose_tile(s->bins, walkers, s->n_l, s->n_pad, g_rt.sm_count, g_rt.smem_optin, s->force_threads,
                       s->force_wpt);
}

// The batched evaluation runs on the fp64 tensor path (lnpost_mma_kernel) where that is the faster kernel: from 12
// angle tiles (more than 88 angles) on -- measured at 200 bins x 8192 walkers: 80 vs 70 % of the fp64 peak at 256
// angles, 59 vs 65 % at 60, 39 vs 52 % at 30 (profiles/r01_sweep_big_dmma_vs_dfma.txt).  A forced tiling
// (mdaStoreSetTile) selects the register-tiled DFMA kernel, MDA_BATCH_IMPL = mma | dfma either one.
bool use_batch_mma(const MdaStore *s) {
    if (s->force_threads > 0 || s->force_wpt > 0) return false;
    if (!lnpost_mma_supported(s->n_l, s->n_tiles, g_rt.smem_optin)) return false;
    const char *impl = getenv("MDA_BATCH_IMPL");
    if (impl && strcmp(impl, "dfma") == 0) return false;
    if (impl && strcmp(impl, "mma") == 0) return true;
    return s->n_tiles >= 12;
}

// The angle-split kernel (G lanes per walker, chi^2 handed from lane to lane in the reference's order: the same bits as
// the tile kernel, a G-th of the latency chain) was built for launches too small to give every SM a few hundred walkers --
// the half-step of a host-driven sampler on BASELINE.json's shapes.  Measured, it loses at every size (C4 half-step
// 14.4 us against 8.3; profiles/experiments/r02_split_kernel.md): G times the threads means several waves of CTAs, each
// paying the ~2 us staging latency again, while the tile kernel's launch is already within ~4 us of an EMPTY launch
// (4.2 us per launch through this path).  Kept selectable: MDA_BATCH_IMPL = split.
bool use_batch_split(const MdaStore *s, int) {
    if (s->force_threads > 0 || s->force_wpt > 0) return false;
    if (!lnpost_split_lanes(s->n_l, s->n_pad)) return false;
    const char *impl = getenv("MDA_BATCH_IMPL");
    return impl && strcmp(impl, "split") == 0;
}

// Enqueue one evaluation of device-resident params on the store's stream.
int enqueue_eval(MdaStore *s, int walkers, const double *d_params, double *d_out, int mode) {
    BatchArgs args;
    memset(&args, 0, sizeof(args));
    args.consts = s->d_consts;
    args.n_pts = s->d_npts;
    args.params = d_params;
    args.out = d_out;
    args.bins = s->bins;
    args.walkers = walkers;
    args.n_l = s->n_l;
    args.n_pad = s->n_pad;
    args.mode = mode;
    args.params_bulk = (reinterpret_cast<uintptr_t>(d_params) & 15u) == 0 && !getenv("MDA_BATCH_NO_BULK_PARAMS");
    if (s->n_l <= kMaxRegL)
        for (int l = 0; l < s->n_l; ++l) { args.lo[l] = s->lo[l]; args.hi[l] = s->hi[l]; }
    const char *batch_impl = getenv("MDA_BATCH_IMPL");
    if (batch_impl && strcmp(batch_impl, "quad") == 0 && lnpost_quad_supported(s->n_l) && s->force_threads == 0 && s->force_wpt == 0) {
        args.quad = s->d_quad;
        MDA_CUDA(launch_batch_quad(args, s->stream), "batched likelihood launch (triangular factor)");
        g_launches.fetch_add(1, std::memory_order_relaxed);
        return MDA_OK;
    }
    if (use_batch_split(s, walkers)) {
        MDA_CUDA(launch_batch_split(args, s->stream, nullptr), "batched likelihood launch (angle-split)");
        g_launches.fetch_add(1, std::memory_order_relaxed);
        return MDA_OK;
    }
    if (use_batch_mma(s)) {
        args.consts_mma = s->d_consts_mma;
        args.n_tiles = s->n_tiles;
        MDA_CUDA(launch_batch_mma(args, g_rt.sm_count, s->stream, nullptr, nullptr), "batched likelihood launch (DMMA)");
        g_launches.fetch_add(1, std::memory_order_relaxed);
        return MDA_OK;
    }
    const TileShape tile = tile_for(s, walkers);
    if (tile.smem_bytes > g_rt.smem_optin)
        return set_error(MDA_ERR_BAD_ARGUMENT, "tile needs %zu bytes of shared memory, device offers %zu",
                         tile.smem_bytes, g_rt.smem_optin);
    MDA_CUDA(launch_batch(args, tile, s->stream, s->d_lo, s->d_hi), "batched likelihood launch");
    g_launches.fetch_add(1, std::memory_order_relaxed);
    return MDA_OK;
}

int check_eval_args(MdaStore *s, int walkers, const void *a, const void *b, const char *fn) {
    if (walkers < 0) return set_error(MDA_ERR_BAD_ARGUMENT, "%s: walkersPerBin < 0", fn);
    if (walkers > 0 && s->bins > 0 && (!a || !b)) return set_error(MDA_ERR_BAD_ARGUMENT, "%s: NULL buffer", fn);
    return MDA_OK;
}

int eval_host(void *store, int walkers, const double *params, double *out, int mode, const char *fn) {
    MdaStore *s = as_store(store, fn);
    if (!s) return MDA_ERR_BAD_ARGUMENT;
    int rc = check_eval_args(s, walkers, params, out, fn);
    if (rc) return rc;
    if ((rc = ensure_runtime())) return rc;
    if (walkers == 0 || s->bins == 0) return MDA_OK;
    if ((rc = upload_store(s))) return rc;
    if ((rc = ensure_staging(s, walkers))) return rc;
    const size_t n_eval = (size_t)s->bins * (size_t)walkers;
    MDA_CUDA(cudaMemcpyAsync(s->d_params, params, sizeof(double) * n_eval * s->n_l, cudaMemcpyHostToDevice, s->stream),
             "params H2D");
    if ((rc = enqueue_eval(s, walkers, s->d_params, s->d_out, mode))) return rc;
    MDA_CUDA(cudaMemcpyAsync(out, s->d_out, sizeof(double) * n_eval, cudaMemcpyDeviceToHost, s->stream), "results D2H");
    MDA_CUDA(cudaStreamSynchronize(s->stream), fn);
    return MDA_OK;
}

// {H2D params, kernel, D2H results} captured once per (walkers, mode), replayed per call.
int eval_staged(void *store, int walkers, int mode, const char *fn) {
    MdaStore *s = as_store(store, fn);
    if (!s) return MDA_ERR_BAD_ARGUMENT;
    if (walkers < 0) return set_error(MDA_ERR_BAD_ARGUMENT, "%s: walkersPerBin < 0", fn);
    int rc;
    if ((rc = ensure_runtime())) return rc;
    if (walkers == 0 || s->bins == 0) return MDA_OK;
    if ((size_t)walkers > s->cap_walkers)
        return set_error(MDA_ERR_BAD_STATE, "%s: call mdaStoreHostParams(store, %d) and fill it first", fn, walkers);
    if ((rc = upload_store(s))) return rc;
    const GraphKey key{walkers, mode};
    auto it = s->graphs.find(key);
    if (it == s->graphs.end()) {
        const size_t n_eval = (size_t)s->bins * (size_t)walkers;
        cudaGraph_t graph = nullptr;
        MDA_CUDA(cudaStreamBeginCapture(s->stream, cudaStreamCaptureModeThreadLocal), "graph capture begin");
        cudaError_t e = cudaMemcpyAsync(s->d_params, s->h_params, sizeof(double) * n_eval * s->n_l,
                                        cudaMemcpyHostToDevice, s->stream);
        if (e == cudaSuccess) rc = enqueue_eval(s, walkers, s->d_params, s->d_out, mode);
        if (e == cudaSuccess && rc == MDA_OK)
            e = cudaMemcpyAsync(s->h_out, s->d_out, sizeof(double) * n_eval, cudaMemcpyDeviceToHost, s->stream);
        cudaError_t e2 = cudaStreamEndCapture(s->stream, &graph);
        if (rc != MDA_OK) { if (graph) cudaGraphDestroy(graph); return rc; }
        if (e != cudaSuccess) { if (graph) cudaGraphDestroy(graph); return cuda_fail(e, "graph capture"); }
        if (e2 != cudaSuccess) return cuda_fail(e2, "graph capture end");
        g_launches.fetch_sub(1, std::memory_order_relaxed);   // the captured launch did not run
        cudaGraphExec_t exec = nullptr;
        e = cudaGraphInstantiate(&exec, graph, 0);
        cudaGraphDestroy(graph);
        if (e != cudaSuccess) return cuda_fail(e, "graph instantiate");
        it = s->graphs.emplace(key, exec).first;
    }
    MDA_CUDA(cudaGraphLaunch(it->second, s->stream), "graph launch");
    g_launches.fetch_add(1, std::memory_order_relaxed);
    MDA_CUDA(cudaStreamSynchronize(s->stream), fn);
    return MDA_OK;
}

int eval_device(void *store, int walkers, const double *d_params, double *d_out, int mode, const char *fn) {
    MdaStore *s = as_store(store, fn);
    if (!s) return MDA_ERR_BAD_ARGUMENT;
    int rc = check_eval_args(s, walkers, d_params, d_out, fn);
    if (rc) return rc;
    if ((rc = ensure_runtime())) return rc;
    if (walkers == 0 || s->bins == 0) return MDA_OK;
    if ((rc = upload_store(s))) return rc;
    return enqueue_eval(s, walkers, d_params, d_out, mode);
}

// ---- device-resident ensemble -----------------------------------------------------------
struct MdaEnsemble {
    uint32_t magic = kEnsembleMagic;
    MdaStore *store = nullptr;
    int walkers = 0;
    unsigned long long seed = 0;
    double a = 2.0;
    double *d_pos = nullptr, *d_lnp = nullptr;
    unsigned int *d_nacc = nullptr;
    int *d_bin_ids = nullptr, *d_status = nullptr;
    unsigned long long *d_progress = nullptr;   // [bins] hand-over words of the DMMA stepper's (chunk, bin) units
    unsigned long long launches = 0;            // stepping launches so far (the words' high half)
    double *d_chain = nullptr, *d_lnp_chain = nullptr;
    long long keep_cap = 0, kept = 0, steps = 0;
    long long alloc_cap = 0;         // kept samples the chain allocations can hold (>= keep_cap: a smaller reservation reuses them)
    bool has_state = false;
    unsigned long long state_generation = 0;   // the store's generation when the state (and its log-probabilities) was set
    unsigned long long epoch = 0;    // mdaEnsembleSetState calls before the current one: folded into the Philox key, so that a run
                                     // restarted on the same ensemble (burn-in, then production) does not replay the first run's draws
    long long *d_prof = nullptr;     // [bins][8] section cycle counters of the last launch (debug)
    void *scratch = nullptr;         // device scratch of the posterior reductions, grown on demand
    size_t scratch_bytes = 0;
    void *buckets = nullptr;         // keys gathered for the order-statistic selection
    size_t bucket_bytes = 0;
    cudaStream_t copy_stream = nullptr;
    cudaEvent_t chunk_done[2] = {nullptr, nullptr}, copy_done[2] = {nullptr, nullptr};
};

// Which stepper runs an ensemble.  DMMA wherever it can: with two or more bins per SM the homogeneous 8-warp kernel
// (ensemble_mma_kernel, two CTAs -- two bins -- per SM), below that the warp-specialised 16-warp kernel
// (ensemble_ws_kernel, one bin per SM at a time).  The register-tiled DFMA stepper carries numL < 4, ensembles of
// more than 512 walkers and ensembles larger than shared memory.  MDA_ENS_IMPL = quad | mma | ws | dfma forces one,
// direct = the angle-by-angle steppers picked by shape as described.
EnsembleShape ensemble_shape_for(MdaStore *s, int walkers) {
    const char *impl = getenv("MDA_ENS_IMPL");
    // A store with an ill-conditioned bin (see factor_bin) is stepped in the reference's operation order, whatever it costs;
    // an explicit MDA_ENS_IMPL is honoured (experiments).
    if (!impl && !quad_is_safe(s))
        return choose_ensemble_shape(s->bins, walkers, s->n_l, s->n_pad, g_rt.sm_count, g_rt.smem_optin, env_int("MDA_ENS_GROUPS", 0), env_int("MDA_ENS_WPT", 0));
    const bool want_dfma = impl && strcmp(impl, "dfma") == 0, want_ws = impl && strcmp(impl, "ws") == 0,
               want_mma = impl && strcmp(impl, "mma") == 0, want_quad = impl && strcmp(impl, "quad") == 0,
               want_direct = impl && strcmp(impl, "direct") == 0,       // angle-by-angle evaluation, stepper picked by shape
               want_qws = impl && strcmp(impl, "qws") == 0;
    EnsembleShape shape{};
    // default: the triangular-factor stepper (fewest operations per evaluation, same chains); the others on request.
    // Warp-specialised (ensemble_qws_kernel): at any number of bins where T fits the walker threads' registers (one CTA per SM
    // steps its share of the bins one after the other), else below two bins per SM; a plain CTA per bin and several per SM
    // (ensemble_quad_kernel) from there on.
    if ((want_qws || !impl) && choose_ensemble_qws_shape(s->bins, walkers, s->n_l, g_rt.sm_count, g_rt.smem_optin, &shape) &&
        (want_qws || shape.treg || s->bins < 2 * g_rt.sm_count)) return shape;
    shape = EnsembleShape{};
    if ((want_quad || want_qws || !(want_dfma || want_ws || want_mma || want_direct)) && choose_ensemble_quad_shape(walkers, s->n_l, g_rt.smem_optin, &shape)) return shape;
    if (!want_dfma) {
        const bool prefer_mma = want_mma || (!want_ws && s->bins >= 2 * g_rt.sm_count);
        const int wt = env_int("MDA_ENS_WT", 0);
        if (prefer_mma && choose_ensemble_mma_shape(s->bins, walkers, s->n_l, s->n_tiles, g_rt.sm_count, g_rt.smem_optin, wt, env_int("MDA_ENS_DRAWW", 0), &shape)) return shape;
        if (choose_ensemble_ws_shape(s->bins, walkers, s->n_l, s->n_tiles, g_rt.sm_count, g_rt.smem_optin, wt, &shape)) return shape;
        if (choose_ensemble_mma_shape(s->bins, walkers, s->n_l, s->n_tiles, g_rt.sm_count, g_rt.smem_optin, wt, env_int("MDA_ENS_DRAWW", 0), &shape)) return shape;
    }
    return choose_ensemble_shape(s->bins, walkers, s->n_l, s->n_pad, g_rt.sm_count, g_rt.smem_optin, env_int("MDA_ENS_GROUPS", 0), env_int("MDA_ENS_WPT", 0));
}

// Schedule of one launch of the DMMA stepper: (chunk of steps, bin) units, chunk-major, unit u on CTA
// u mod grid.  Up to one bin per SM every bin simply has its own CTA.  With more bins than SMs the resident CTAs
// walk the list, so that every SM carries the same share of the run (C4: 200 bins on 148 SMs; from two bins per
// SM on, the 8-warp kernel keeps two CTAs per SM).  MDA_ENS_GRID / MDA_ENS_CHUNK override, for experiments and tests.
struct SteppingPlan { int grid, chunk_steps, n_chunks; };
constexpr int kQuadWalkPerSm = 1;      // CTAs per SM when the triangular-factor stepper walks units
SteppingPlan plan_stepping(const MdaStore *s, const EnsembleShape &shape, int nSteps) {
    SteppingPlan p{0, nSteps, 1};
    const int capacity = shape.ws ? ensemble_ws_capacity(s->n_l, shape, g_rt.sm_count) : ensemble_mma_capacity(s->n_l, shape, g_rt.sm_count);
    if (capacity < 1) return p;
    // between one and two bins per SM a second CTA per SM does not pay (measured: the units that wait for their
    // predecessor leave SMs idle), so the grid is one CTA per SM there
    int grid = s->bins <= g_rt.sm_count ? s->bins : (s->bins < capacity ? g_rt.sm_count : capacity);
    const int force_grid = env_int("MDA_ENS_GRID", 0);
    if (force_grid > 0) grid = force_grid;
    grid = std::max(1, std::min(grid, capacity));
    int chunk_steps = nSteps;
    const int min_chunk = 8, max_chunks = std::max(1, std::min(96, nSteps / min_chunk));
    if (grid < s->bins) {
        // Units are walked in rounds of `grid`; the last round is only partly filled.  Among the chunk counts that keep a
        // chunk at 8 steps or more, take the one that wastes the smallest share of CTA time in that last round; ties go
        // to the fewer hand-overs.
        double best_waste = 1e30;
        int best = 1;
        for (int n = std::min(4, max_chunks); n <= max_chunks; ++n) {
            const int cs = (nSteps + n - 1) / n, real_n = (nSteps + cs - 1) / cs;
            const long long units = (long long)real_n * s->bins, rounds = (units + grid - 1) / grid;
            const double waste = (double)(rounds * grid) / (double)units + 0.0005 * real_n;   // each hand-over costs ~0.05 % of a chunk
            if (waste < best_waste - 1e-12) { best_waste = waste; best = real_n; }
        }
        chunk_steps = (nSteps + best - 1) / best;
    }
    const int force_chunk = env_int("MDA_ENS_CHUNK", 0);
    if (force_chunk > 0) chunk_steps = std::min(force_chunk, nSteps);
    p.grid = grid;
    p.chunk_steps = std::max(1, chunk_steps);
    p.n_chunks = (nSteps + p.chunk_steps - 1) / p.chunk_steps;
    return p;
}

// Schedule of the triangular-factor stepper: a CTA per bin for the whole launch, unless a grid smaller than the number of
// bins pays (just above one bin per SM a few SMs would carry two bins for the whole run and set its length): then one CTA
// per SM walks (chunk, bin) units as above.  MDA_ENS_GRID / MDA_ENS_CHUNK / MDA_ENS_QUAD_WALK override.
SteppingPlan plan_quad_stepping(const MdaStore *s, const EnsembleShape &shape, int nSteps) {
    SteppingPlan p{s->bins, nSteps, 1};
    int grid = s->bins;
    const int forced_walk = env_int("MDA_ENS_QUAD_WALK", -1);                // experiments: CTAs per SM of the walk (0: off)
    const int walk_per_sm = forced_walk >= 0 ? forced_walk : kQuadWalkPerSm;
    // measured at 512 x 9: one bin per SM 2.57 us per step, two bins per SM 3.83 -- a CTA per bin costs the two-bin time as soon
    // as one SM carries two, walking with one CTA per SM costs bins / SMs x 2.6: it pays up to 1.4 bins per SM (C4: 3.83 -> 3.57)
    // ... for launches long enough that a unit is ~100 steps (a hand-over moves 40 KB of state and the constants; with 8-step
    // units a 100-step launch ran at 4.5 us per step under ncu)
    const bool walk_pays = forced_walk >= 0 ? s->bins < 3 * g_rt.sm_count : (10LL * s->bins <= 14LL * g_rt.sm_count && nSteps >= 512);
    if (walk_per_sm > 0 && s->bins > g_rt.sm_count && walk_pays) grid = walk_per_sm * g_rt.sm_count;
    const int force_grid = env_int("MDA_ENS_GRID", 0);
    if (force_grid > 0) grid = force_grid;
    if ((grid == s->bins || (grid > s->bins && force_grid <= 0)) && env_int("MDA_ENS_CHUNK", 0) <= 0) return p;   // every bin has its own CTA
    const int capacity = ensemble_quad_capacity(s->n_l, shape, grid, g_rt.sm_count);
    if (capacity < 1) { p.grid = 0; return p; }
    grid = std::max(1, std::min(grid, capacity));          // (a forced grid may exceed the bins: more CTAs than active units)
    if (force_grid <= 0) grid = std::min(grid, s->bins);
    int chunk_steps = nSteps;
    if (grid != s->bins) {
        const int min_chunk = 8, max_chunks = std::max(1, std::min(96, nSteps / min_chunk));
        double best_waste = 1e30;
        int best = 1;
        for (int n = std::min(4, max_chunks); n <= max_chunks; ++n) {
            const int cs = (nSteps + n - 1) / n, real_n = (nSteps + cs - 1) / cs;
            const long long units = (long long)real_n * s->bins, rounds = (units + grid - 1) / grid;
            const double waste = (double)(rounds * grid) / (double)units + 0.0005 * real_n;
            if (waste < best_waste - 1e-12) { best_waste = waste; best = real_n; }
        }
        chunk_steps = (nSteps + best - 1) / best;
    }
    const int force_chunk = env_int("MDA_ENS_CHUNK", 0);
    if (force_chunk > 0) chunk_steps = std::min(force_chunk, nSteps);
    p.grid = grid;
    p.chunk_steps = std::max(1, chunk_steps);
    p.n_chunks = (nSteps + p.chunk_steps - 1) / p.chunk_steps;
    return p;
}

// Grid of the warp-specialised triangular-factor stepper: a CTA per bin while they are all resident at once, else as many
// CTAs as are resident, each carrying an equal share of the bins' steps (see qws_next_unit).  MDA_ENS_GRID overrides (tests).
int qws_grid(const MdaStore *s, const EnsembleShape &shape) {
    const int capacity = ensemble_qws_capacity(s->n_l, shape, g_rt.sm_count);
    if (capacity < 1) return 0;
    int grid = std::min(s->bins, capacity);
    const int force_grid = env_int("MDA_ENS_GRID", 0);
    if (force_grid > 0) grid = std::max(1, std::min(force_grid, grid));
    return grid;
}

// what the last mdaBatchFit of this thread met (mdaBatchFitStats)
struct FitStats { long long unconverged = 0, null_column = 0; int max_iterations = 0; };
thread_local FitStats t_fit_stats;

// MDA_E2E_TIMING=1: wall-clock laps of the host-buffer calls on stderr (where a millisecond-long e2e region goes)
struct LapTimer {
    bool on = getenv("MDA_E2E_TIMING") != nullptr;
    std::chrono::steady_clock::time_point t = std::chrono::steady_clock::now();
    void lap(const char *what) {
        if (!on) return;
        const auto now = std::chrono::steady_clock::now();
        fprintf(stderr, "[mda e2e] %-40s %8.3f ms\n", what, std::chrono::duration<double, std::milli>(now - t).count());
        t = now;
    }
};

MdaEnsemble *as_ensemble(void *p, const char *fn) {
    MdaEnsemble *e = static_cast<MdaEnsemble *>(p);
    if (!e || e->magic != kEnsembleMagic || !e->store || e->store->magic != kStoreMagic) {      // (a store freed before its ensembles stays alive until the last of them goes)
        set_error(MDA_ERR_BAD_ARGUMENT, "%s: not a live handle from mdaEnsembleCreate", fn);
        return nullptr;
    }
    return e;
}

}  // namespace

// =====================================================================================
// exported C ABI
// =====================================================================================
extern "C" {

// ---- the reference's seven symbols ---------------------------------------------------
void *makeMdaStruct(int numPts, int numL) {
    if (numPts < 0 || numL < 0) { set_error(MDA_ERR_BAD_ARGUMENT, "makeMdaStruct(%d, %d): negative size", numPts, numL); return nullptr; }
    if (ensure_runtime() != MDA_OK) return nullptr;
    MdaHandle *h = new MdaHandle();
    h->n = numPts;
    h->n_l = numL;
    h->host.assign((size_t)(1 + numL) * (size_t)numPts, 0.0);
    cudaError_t e = cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaMalloc((void **)&h->d_consts, sizeof(double) * (h->host.size() ? h->host.size() : 1));
    if (e == cudaSuccess) e = cudaMalloc((void **)&h->d_resid, sizeof(double) * (size_t)(numPts > 0 ? numPts : 1));
    if (e == cudaSuccess) e = cudaHostAlloc((void **)&h->h_io, sizeof(double) * (size_t)(numL + 2), cudaHostAllocMapped);
    if (e == cudaSuccess) e = cudaHostGetDevicePointer((void **)&h->d_io, h->h_io, 0);
    if (e != cudaSuccess) { cuda_fail(e, "makeMdaStruct"); destroy_struct(h); return nullptr; }
    memset(h->h_io, 0, sizeof(double) * (size_t)(numL + 2));       // the ticket word starts at 0: no call has finished
    return h;
}

void freeMdaStruct(void *strPtr) {
    MdaHandle *h = as_struct(strPtr, "freeMdaStruct");
    if (!h) return;
    if (g_rt.ready) cudaSetDevice(g_rt.device);
    destroy_struct(h);
}

void setMdaData(void *strPtr, double *divData) {
    MdaHandle *h = as_struct(strPtr, "setMdaData");
    if (!h) return;
    if (!divData) { set_error(MDA_ERR_BAD_ARGUMENT, "setMdaData: divData is NULL"); return; }
    memcpy(h->host.data(), divData, sizeof(double) * (size_t)h->n);
    h->dirty = true;
}

void setMdaDist(void *strPtr, int distIndex, double *dist) {
    MdaHandle *h = as_struct(strPtr, "setMdaDist");
    if (!h) return;
    if (distIndex >= h->n_l) return;                       // silently ignored, C/ChiSquare.c:88-89
    if (distIndex < 0) { set_error(MDA_ERR_BAD_ARGUMENT, "setMdaDist: negative distIndex %d ignored", distIndex); return; }
    if (!dist) { set_error(MDA_ERR_BAD_ARGUMENT, "setMdaDist: dist is NULL"); return; }
    memcpy(h->host.data() + (size_t)(1 + distIndex) * (size_t)h->n, dist, sizeof(double) * (size_t)h->n);
    h->dirty = true;
}

double calculateChi(void *strPtr, double *params) { return scalar_eval(strPtr, params, nullptr, kChi, "calculateChi"); }

double calculateLnLiklihood(void *strPtr, double *params) {
    return scalar_eval(strPtr, params, nullptr, kLnLike, "calculateLnLiklihood");
}

double calculateLnLiklihoodResids(void *strPtr, double *params, double *residArray) {
    if (!residArray) { set_error(MDA_ERR_BAD_ARGUMENT, "calculateLnLiklihoodResids: residArray is NULL"); return NAN; }
    return scalar_eval(strPtr, params, residArray, kLnLike, "calculateLnLiklihoodResids");
}

// ---- error channel -------------------------------------------------------------------
int mdaLastError(void) { return t_err_code; }
const char *mdaLastErrorString(void) { return t_err_msg; }
void mdaClearError(void) { t_err_code = MDA_OK; snprintf(t_err_msg, sizeof(t_err_msg), "no error"); }
const char *mdaVersion(void) { return "mimic_mda_b200 0.1 (sm_100a, fp64 DFMA path)"; }

// ---- device and memory helpers ---------------------------------------------------------
int mdaDeviceCount(void) {
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess) { cudaGetLastError(); return 0; }
    return count;
}

int mdaSetDevice(int device) {
    std::lock_guard<std::mutex> lock(g_rt.mu);
    if (g_rt.ready && device != g_rt.device)
        return set_error(MDA_ERR_BAD_STATE, "mdaSetDevice(%d): device %d already in use by this process", device, g_rt.device);
    if (device < 0) return set_error(MDA_ERR_BAD_ARGUMENT, "mdaSetDevice(%d)", device);
    g_rt.requested = device;
    return MDA_OK;
}

int mdaGetDevice(void) { return g_rt.ready ? g_rt.device : g_rt.requested; }

int mdaDeviceInfo(char *name, int nameLen, int *smCount, int *cc) {
    int rc = ensure_runtime();
    if (rc) return rc;
    if (name && nameLen > 0) snprintf(name, (size_t)nameLen, "%s", g_rt.name);
    if (smCount) *smCount = g_rt.sm_count;
    if (cc) *cc = g_rt.cc;
    return MDA_OK;
}

void *mdaHostAlloc(size_t bytes) {
    if (ensure_runtime() != MDA_OK) return nullptr;
    void *p = nullptr;
    cudaError_t e = cudaHostAlloc(&p, bytes ? bytes : 8, cudaHostAllocDefault);
    if (e != cudaSuccess) { cuda_fail(e, "mdaHostAlloc"); return nullptr; }
    return p;
}
void mdaHostFree(void *p) { if (p) cudaFreeHost(p); }

void *mdaDeviceAlloc(size_t bytes) {
    if (ensure_runtime() != MDA_OK) return nullptr;
    void *p = nullptr;
    cudaError_t e = cudaMalloc(&p, bytes ? bytes : 8);
    if (e != cudaSuccess) { cuda_fail(e, "mdaDeviceAlloc"); return nullptr; }
    return p;
}
void mdaDeviceFree(void *p) { if (p) cudaFree(p); }

int mdaMemcpyH2D(void *dst, const void *src, size_t bytes) {
    int rc = ensure_runtime();
    if (rc) return rc;
    MDA_CUDA(cudaMemcpy(dst, src, bytes, cudaMemcpyHostToDevice), "mdaMemcpyH2D");
    return MDA_OK;
}
int mdaMemcpyD2H(void *dst, const void *src, size_t bytes) {
    int rc = ensure_runtime();
    if (rc) return rc;
    MDA_CUDA(cudaMemcpy(dst, src, bytes, cudaMemcpyDeviceToHost), "mdaMemcpyD2H");
    return MDA_OK;
}
int mdaDeviceSynchronize(void) {
    int rc = ensure_runtime();
    if (rc) return rc;
    MDA_CUDA(cudaDeviceSynchronize(), "mdaDeviceSynchronize");
    return MDA_OK;
}

// ---- store ---------------------------------------------------------------------------
void *mdaStoreCreate(int numBins, int numL, int maxPts) {
    if (numBins < 0 || numL < 1 || maxPts < 0) {
        set_error(MDA_ERR_BAD_ARGUMENT, "mdaStoreCreate(%d, %d, %d): need numBins >= 0, numL >= 1, maxPts >= 0", numBins, numL, maxPts);
        return nullptr;
    }
    if (ensure_runtime() != MDA_OK) return nullptr;
    MdaStore *s = new MdaStore();
    s->bins = numBins;
    s->n_l = numL;
    s->max_pts = maxPts;
    s->n_pad = maxPts + (maxPts & 1);
    if (s->n_pad == 0) s->n_pad = 2;
    s->host.assign((size_t)numBins * bin_stride(s), 0.0);
    s->n_tiles = (s->n_pad + 7) / 8;
    s->host_mma.assign((size_t)numBins * (size_t)s->n_tiles * 8 * (size_t)(numL + 1), 0.0);
    s->host_quad.assign((size_t)numBins * (size_t)quad_stride(numL), 0.0);
    s->n_pts.assign((size_t)numBins, 0);
    s->lo.assign((size_t)numL, 0.0);
    s->hi.assign((size_t)numL, 1.0);
    cudaError_t e = cudaStreamCreateWithFlags(&s->stream, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaMalloc((void **)&s->d_consts, sizeof(double) * (s->host.size() ? s->host.size() : 1));
    if (e == cudaSuccess) e = cudaMalloc((void **)&s->d_consts_mma, sizeof(double) * (s->host_mma.size() ? s->host_mma.size() : 1));
    if (e == cudaSuccess) e = cudaMalloc((void **)&s->d_quad, sizeof(double) * (s->host_quad.size() ? s->host_quad.size() : 1));
    if (e == cudaSuccess) e = cudaMalloc((void **)&s->d_npts, sizeof(int) * (size_t)(numBins ? numBins : 1));
    if (e == cudaSuccess) e = cudaMalloc((void **)&s->d_lo, sizeof(double) * (size_t)numL);
    if (e == cudaSuccess) e = cudaMalloc((void **)&s->d_hi, sizeof(double) * (size_t)numL);
    if (e != cudaSuccess) { cuda_fail(e, "mdaStoreCreate"); mdaStoreFree(s); return nullptr; }
    return s;
}

}  // extern "C"

namespace {
void destroy_store(MdaStore *s);
}

extern "C" {

void mdaStoreFree(void *store) {
    MdaStore *s = as_store(store, "mdaStoreFree");
    if (!s) return;
    if (s->live_ensembles > 0) { s->closing = true; return; }      // its ensembles still point at it: freed with the last of them
    destroy_store(s);
}

}  // extern "C"

namespace {
void destroy_store(MdaStore *s) {
    if (g_rt.ready) cudaSetDevice(g_rt.device);
    if (s->stream) cudaStreamSynchronize(s->stream);
    drop_graphs(s);
    if (s->h_params) cudaFreeHost(s->h_params);
    if (s->h_out) cudaFreeHost(s->h_out);
    if (s->d_params) cudaFree(s->d_params);
    if (s->d_out) cudaFree(s->d_out);
    if (s->d_consts) cudaFree(s->d_consts);
    if (s->d_consts_mma) cudaFree(s->d_consts_mma);
    if (s->d_quad) cudaFree(s->d_quad);
    if (s->d_npts) cudaFree(s->d_npts);
    if (s->d_lo) cudaFree(s->d_lo);
    if (s->d_hi) cudaFree(s->d_hi);
    if (s->stream) cudaStreamDestroy(s->stream);
    s->magic = 0;
    delete s;
}
}  // namespace

extern "C" {

int mdaStoreSetBin(void *store, int bin, int numPts, const double *divData, const double *dists) {
    MdaStore *s = as_store(store, "mdaStoreSetBin");
    if (!s) return MDA_ERR_BAD_ARGUMENT;
    if (bin < 0 || bin >= s->bins) return set_error(MDA_ERR_BAD_ARGUMENT, "mdaStoreSetBin: bin %d outside [0, %d)", bin, s->bins);
    if (numPts < 0 || numPts > s->max_pts)
        return set_error(MDA_ERR_BAD_ARGUMENT, "mdaStoreSetBin: numPts %d outside [0, %d]", numPts, s->max_pts);
    if (numPts > 0 && (!divData || !dists)) return set_error(MDA_ERR_BAD_ARGUMENT, "mdaStoreSetBin: NULL buffer");
    write_bin(s, bin, numPts, divData, dists, (size_t)numPts);
    return MDA_OK;
}

int mdaStoreSetBinFromStruct(void *store, int bin, void *mdaStruct) {
    MdaStore *s = as_store(store, "mdaStoreSetBinFromStruct");
    MdaHandle *h = as_struct(mdaStruct, "mdaStoreSetBinFromStruct");
    if (!s || !h) return MDA_ERR_BAD_ARGUMENT;
    if (bin < 0 || bin >= s->bins) return set_error(MDA_ERR_BAD_ARGUMENT, "mdaStoreSetBinFromStruct: bin %d outside [0, %d)", bin, s->bins);
    if (h->n_l != s->n_l) return set_error(MDA_ERR_BAD_ARGUMENT, "mdaStoreSetBinFromStruct: struct has %d dists, store %d", h->n_l, s->n_l);
    if (h->n > s->max_pts) return set_error(MDA_ERR_BAD_ARGUMENT, "mdaStoreSetBinFromStruct: struct has %d angles, store holds %d", h->n, s->max_pts);
    write_bin(s, bin, h->n, h->host.data(), h->host.data() + h->n, (size_t)h->n);
    return MDA_OK;
}

int mdaStoreSetBounds(void *store, const double *lo, const double *hi) {
    MdaStore *s = as_store(store, "mdaStoreSetBounds");
    if (!s) return MDA_ERR_BAD_ARGUMENT;
    if (!lo || !hi) return set_error(MDA_ERR_BAD_ARGUMENT, "mdaStoreSetBounds: NULL buffer");
    s->lo.assign(lo, lo + s->n_l);
    s->hi.assign(hi, hi + s->n_l);
    s->dirty = true;
    ++s->generation;
    if (s->stream) cudaStreamSynchronize(s->stream);
    drop_graphs(s);                                       // bounds are baked into captured launches
    return MDA_OK;
}

int mdaStoreUpload(void *store) {
    MdaStore *s = as_store(store, "mdaStoreUpload");
    if (!s) return MDA_ERR_BAD_ARGUMENT;
    int rc = ensure_runtime();
    if (rc) return rc;
    return upload_store(s);
}

int mdaDeviceMemInfo(size_t *freeBytes, size_t *totalBytes) {
    int rc = ensure_runtime();
    if (rc) return rc;
    size_t f = 0, t = 0;
    MDA_CUDA(cudaMemGetInfo(&f, &t), "mdaDeviceMemInfo");
    if (freeBytes) *freeBytes = f;
    if (totalBytes) *totalBytes = t;
    return MDA_OK;
}

int mdaStoreConditioning(void *store, double *kappa, double *kappaMax, int *quadSafe) {
    MdaStore *s = as_store(store, "mdaStoreConditioning");
    if (!s) return MDA_ERR_BAD_ARGUMENT;
    const bool safe = quad_is_safe(s);                   // (host arithmetic only: no device is touched)
    if (kappa) for (int b = 0; b < s->bins; ++b) kappa[b] = s->kappa[(size_t)b];
    if (kappaMax) *kappaMax = kQuadKappaMax;
    if (quadSafe) *quadSafe = safe ? 1 : 0;
    return MDA_OK;
}

int mdaStoreNumBins(void *store) { MdaStore *s = as_store(store, "mdaStoreNumBins"); return s ? s->bins : -1; }
int mdaStoreNumL(void *store) { MdaStore *s = as_store(store, "mdaStoreNumL"); return s ? s->n_l : -1; }
int mdaStoreMaxPts(void *store) { MdaStore *s = as_store(store, "mdaStoreMaxPts"); return s ? s->max_pts : -1; }

int mdaStoreSetTile(void *store, int threads, int walkersPerThread) {
    MdaStore *s = as_store(store, "mdaStoreSetTile");
    if (!s) return MDA_ERR_BAD_ARGUMENT;
    const bool t_ok = threads == 0 || threads == 32 || threads == 64 || threads == 128 || threads == 256;
    const bool w_ok = walkersPerThread == 0 || walkersPerThread == 1 || walkersPerThread == 2 || walkersPerThread == 4;
    if (!t_ok || !w_ok) return set_error(MDA_ERR_BAD_ARGUMENT, "mdaStoreSetTile(%d, %d): unsupported tiling", threads, walkersPerThread);
    s->force_threads = threads;
    s->force_wpt = walkersPerThread;
    if (s->stream) cudaStreamSynchronize(s->stream);
    drop_graphs(s);
    return MDA_OK;
}

int mdaStoreGetTile(void *store, int walkersPerBin, int *threads, int *walkersPerThread, int *gridSize) {
    MdaStore *s = as_store(store, "mdaStoreGetTile");
    if (!s) return MDA_ERR_BAD_ARGUMENT;
    int rc = ensure_runtime();
    if (rc) return rc;
    if (use_batch_split(s, walkersPerBin > 0 ? walkersPerBin : 1)) {   // the angle-split kernel: 256 threads, G lanes per walker -> reported as -G
        const int g = lnpost_split_lanes(s->n_l, s->n_pad), wpc = 256 / g;
        if (threads) *threads = 256;
        if (walkersPerThread) *walkersPerThread = -g;
        if (gridSize) *gridSize = s->bins * (((walkersPerBin > 0 ? walkersPerBin : 1) + wpc - 1) / wpc);
        return MDA_OK;
    }
    if (use_batch_mma(s)) {                            // the DMMA kernel: 128 threads, a warp per 32 walkers -> reported as 0 walkers per thread
        const long n_items = ((walkersPerBin > 0 ? walkersPerBin : 1) + 31) / 32, total = n_items * s->bins;
        long passes = (total + 4 * 8L * g_rt.sm_count - 1) / (4 * 8L * g_rt.sm_count);
        if (passes < 1) passes = 1;
        if (threads) *threads = 128;
        if (walkersPerThread) *walkersPerThread = 0;
        if (gridSize) *gridSize = (int)(s->bins * ((n_items + 4 * passes - 1) / (4 * passes)));
        return MDA_OK;
    }
    const TileShape t = tile_for(s, walkersPerBin > 0 ? walkersPerBin : 1);
    if (threads) *threads = t.threads;
    if (walkersPerThread) *walkersPerThread = t.wpt;
    if (gridSize) *gridSize = t.grid;
    return MDA_OK;
}

// ---- batched evaluation -----------------------------------------------------------------
int mdaBatchLnPost(void *store, int walkersPerBin, const double *params, double *lnPost) {
    return eval_host(store, walkersPerBin, params, lnPost, kLnPost, "mdaBatchLnPost");
}
int mdaBatchChi(void *store, int walkersPerBin, const double *params, double *chi) {
    return eval_host(store, walkersPerBin, params, chi, kChi, "mdaBatchChi");
}

double *mdaStoreHostParams(void *store, int walkersPerBin) {
    MdaStore *s = as_store(store, "mdaStoreHostParams");
    if (!s || walkersPerBin < 0) return nullptr;
    if (ensure_runtime() != MDA_OK) return nullptr;
    if (ensure_staging(s, walkersPerBin > 0 ? walkersPerBin : 1) != MDA_OK) return nullptr;
    return s->h_params;
}
double *mdaStoreHostOut(void *store, int walkersPerBin) {
    MdaStore *s = as_store(store, "mdaStoreHostOut");
    if (!s || walkersPerBin < 0) return nullptr;
    if (ensure_runtime() != MDA_OK) return nullptr;
    if (ensure_staging(s, walkersPerBin > 0 ? walkersPerBin : 1) != MDA_OK) return nullptr;
    return s->h_out;
}
int mdaBatchLnPostStaged(void *store, int walkersPerBin) { return eval_staged(store, walkersPerBin, kLnPost, "mdaBatchLnPostStaged"); }
int mdaBatchChiStaged(void *store, int walkersPerBin) { return eval_staged(store, walkersPerBin, kChi, "mdaBatchChiStaged"); }

int mdaBatchLnPostDevice(void *store, int walkersPerBin, const double *dParams, double *dLnPost) {
    return eval_device(store, walkersPerBin, dParams, dLnPost, kLnPost, "mdaBatchLnPostDevice");
}
int mdaBatchChiDevice(void *store, int walkersPerBin, const double *dParams, double *dChi) {
    return eval_device(store, walkersPerBin, dParams, dChi, kChi, "mdaBatchChiDevice");
}
int mdaBatchLnPostDeviceMany(void *store, int walkersPerBin, const double *const *dParams, double *const *dOut,
                             int nBlocks, int count) {
    MdaStore *s = as_store(store, "mdaBatchLnPostDeviceMany");
    if (!s) return MDA_ERR_BAD_ARGUMENT;
    if (nBlocks <= 0 || count < 0 || !dParams || !dOut) return set_error(MDA_ERR_BAD_ARGUMENT, "mdaBatchLnPostDeviceMany: bad block list");
    int rc = ensure_runtime();
    if (rc) return rc;
    if (walkersPerBin <= 0 || s->bins == 0) return MDA_OK;
    if ((rc = upload_store(s))) return rc;
    for (int i = 0; i < count; ++i)
        if ((rc = enqueue_eval(s, walkersPerBin, dParams[i % nBlocks], dOut[i % nBlocks], kLnPost))) return rc;
    return MDA_OK;
}

int mdaStoreSynchronize(void *store) {
    MdaStore *s = as_store(store, "mdaStoreSynchronize");
    if (!s) return MDA_ERR_BAD_ARGUMENT;
    int rc = ensure_runtime();
    if (rc) return rc;
    MDA_CUDA(cudaStreamSynchronize(s->stream), "mdaStoreSynchronize");
    return MDA_OK;
}

// ---- events -------------------------------------------------------------------------------
void *mdaEventCreate(void) {
    if (ensure_runtime() != MDA_OK) return nullptr;
    cudaEvent_t ev = nullptr;
    cudaError_t e = cudaEventCreate(&ev);
    if (e != cudaSuccess) { cuda_fail(e, "mdaEventCreate"); return nullptr; }
    return ev;
}
void mdaEventDestroy(void *event) { if (event) cudaEventDestroy(static_cast<cudaEvent_t>(event)); }
int mdaEventRecord(void *event, void *store) {
    MdaStore *s = as_store(store, "mdaEventRecord");
    if (!s || !event) return set_error(MDA_ERR_BAD_ARGUMENT, "mdaEventRecord: NULL argument");
    MDA_CUDA(cudaEventRecord(static_cast<cudaEvent_t>(event), s->stream), "mdaEventRecord");
    return MDA_OK;
}
int mdaEventSynchronize(void *event) {
    if (!event) return set_error(MDA_ERR_BAD_ARGUMENT, "mdaEventSynchronize: NULL event");
    MDA_CUDA(cudaEventSynchronize(static_cast<cudaEvent_t>(event)), "mdaEventSynchronize");
    return MDA_OK;
}
int mdaEventElapsedMs(void *start, void *stop, float *ms) {
    if (!start || !stop || !ms) return set_error(MDA_ERR_BAD_ARGUMENT, "mdaEventElapsedMs: NULL argument");
    MDA_CUDA(cudaEventElapsedTime(ms, static_cast<cudaEvent_t>(start), static_cast<cudaEvent_t>(stop)), "mdaEventElapsedMs");
    return MDA_OK;
}
unsigned long long mdaKernelLaunchCount(void) { return g_launches.load(std::memory_order_relaxed); }

// ---- batched initial fits -----------------------------------------------------------------------
int mdaBatchFit(void *store, int startsPerBin, int sharedStarts, const double *starts, double *params, double *chi) {
    MdaStore *s = as_store(store, "mdaBatchFit");
    if (!s) return MDA_ERR_BAD_ARGUMENT;
    if (startsPerBin < 0 || (startsPerBin > 0 && !starts)) return set_error(MDA_ERR_BAD_ARGUMENT, "mdaBatchFit: bad start list");
    if (s->n_l > kMaxRegL) return set_error(MDA_ERR_BAD_ARGUMENT, "mdaBatchFit: numL %d > %d", s->n_l, kMaxRegL);
    int rc = ensure_runtime();
    if (rc) return rc;
    if (startsPerBin == 0 || s->bins == 0) return MDA_OK;
    if ((rc = upload_store(s))) return rc;
    const size_t L = (size_t)s->n_l, B = (size_t)s->bins, S = (size_t)startsPerBin;
    const size_t n_start = (sharedStarts ? 1 : B) * S * L;
    double *d_starts = nullptr, *d_fit = nullptr, *d_chi = nullptr;
    int *d_iters = nullptr;
    std::vector<int> h_iters(B * S);
    cudaError_t e = cudaMalloc((void **)&d_starts, sizeof(double) * n_start);
    if (e == cudaSuccess) e = cudaMalloc((void **)&d_fit, sizeof(double) * B * S * L);
    if (e == cudaSuccess) e = cudaMalloc((void **)&d_chi, sizeof(double) * B * S);
    if (e == cudaSuccess) e = cudaMalloc((void **)&d_iters, sizeof(int) * B * S);
    if (e == cudaSuccess) e = cudaMemcpyAsync(d_starts, starts, sizeof(double) * n_start, cudaMemcpyHostToDevice, s->stream);
    // the active-set solves run on the bins' triangular factors (no normal equations: fit_kernel.cu)
    if (e == cudaSuccess) e = launch_bvls(s->d_quad, s->d_lo, s->d_hi, d_starts, s->bins, s->n_l, startsPerBin, sharedStarts != 0, d_fit, d_iters, s->stream);
    if (e == cudaSuccess) g_launches.fetch_add(1, std::memory_order_relaxed);
    if (e == cudaSuccess) rc = enqueue_eval(s, startsPerBin, d_fit, d_chi, kChi);
    if (e == cudaSuccess && rc == MDA_OK && params) e = cudaMemcpyAsync(params, d_fit, sizeof(double) * B * S * L, cudaMemcpyDeviceToHost, s->stream);
    if (e == cudaSuccess && rc == MDA_OK && chi) e = cudaMemcpyAsync(chi, d_chi, sizeof(double) * B * S, cudaMemcpyDeviceToHost, s->stream);
    if (e == cudaSuccess && rc == MDA_OK) e = cudaMemcpyAsync(h_iters.data(), d_iters, sizeof(int) * B * S, cudaMemcpyDeviceToHost, s->stream);
    cudaError_t e2 = cudaStreamSynchronize(s->stream);
    cudaFree(d_starts); cudaFree(d_fit); cudaFree(d_chi); cudaFree(d_iters);
    if (e != cudaSuccess) return cuda_fail(e, "mdaBatchFit");
    if (rc != MDA_OK) return rc;
    if (e2 != cudaSuccess) return cuda_fail(e2, "mdaBatchFit");
    t_fit_stats = FitStats{};
    for (int v : h_iters) {
        if (v < 0) { ++t_fit_stats.unconverged; t_fit_stats.max_iterations = std::max(t_fit_stats.max_iterations, -v); continue; }
        if (v & (1 << 16)) ++t_fit_stats.null_column;
        t_fit_stats.max_iterations = std::max(t_fit_stats.max_iterations, v & 0xFFFF);
    }
    return MDA_OK;
}

int mdaBatchFitStats(long long *unconverged, long long *nullColumn, int *maxIterations) {
    if (unconverged) *unconverged = t_fit_stats.unconverged;
    if (nullColumn) *nullColumn = t_fit_stats.null_column;
    if (maxIterations) *maxIterations = t_fit_stats.max_iterations;
    return MDA_OK;
}

// ---- device-resident ensemble sampling --------------------------------------------------------
void *mdaEnsembleCreate(void *store, int nWalkers, unsigned long long seed, double stretchA) {
    MdaStore *s = as_store(store, "mdaEnsembleCreate");
    if (!s) return nullptr;
    if (nWalkers < 2 || (nWalkers & 1)) { set_error(MDA_ERR_BAD_ARGUMENT, "mdaEnsembleCreate: the number of walkers must be even and >= 2 (got %d)", nWalkers); return nullptr; }
    if (nWalkers < 2 * s->n_l) { set_error(MDA_ERR_BAD_ARGUMENT, "mdaEnsembleCreate: need at least twice as many walkers (%d) as dimensions (%d)", nWalkers, s->n_l); return nullptr; }
    if (!(stretchA > 1.0)) { set_error(MDA_ERR_BAD_ARGUMENT, "mdaEnsembleCreate: stretch scale must be > 1"); return nullptr; }
    if (ensure_runtime() != MDA_OK) return nullptr;
    const EnsembleShape shape = ensemble_shape_for(s, nWalkers);
    if (!shape.fits) {
        set_error(MDA_ERR_BAD_ARGUMENT, "mdaEnsembleCreate: %d walkers x %d dists x %d angles need %zu bytes of shared memory (limit %zu, numL <= %d)",
                  nWalkers, s->n_l, s->n_pad, shape.smem_bytes, g_rt.smem_optin, kMaxRegL);
        return nullptr;
    }
    MdaEnsemble *e = new MdaEnsemble();
    e->store = s;
    ++s->live_ensembles;
    e->walkers = nWalkers;
    e->seed = seed;
    e->a = stretchA;
    const size_t n_eval = (size_t)s->bins * (size_t)nWalkers;
    cudaError_t err = cudaMalloc((void **)&e->d_pos, sizeof(double) * (n_eval * s->n_l ? n_eval * s->n_l : 1));
    if (err == cudaSuccess) err = cudaMalloc((void **)&e->d_lnp, sizeof(double) * (n_eval ? n_eval : 1));
    if (err == cudaSuccess) err = cudaMalloc((void **)&e->d_nacc, sizeof(unsigned int) * (n_eval ? n_eval : 1));
    if (err == cudaSuccess) err = cudaMalloc((void **)&e->d_bin_ids, sizeof(int) * (size_t)(s->bins ? s->bins : 1));
    if (err == cudaSuccess) err = cudaMalloc((void **)&e->d_status, sizeof(int));
    if (err == cudaSuccess) err = cudaMemset(e->d_status, 0, sizeof(int));
    if (err == cudaSuccess) err = cudaMalloc((void **)&e->d_progress, sizeof(unsigned long long) * (size_t)(s->bins ? s->bins : 1));
    if (err == cudaSuccess) err = cudaMemset(e->d_progress, 0, sizeof(unsigned long long) * (size_t)(s->bins ? s->bins : 1));
    if (err == cudaSuccess) {
        std::vector<int> ids((size_t)s->bins);
        for (int b = 0; b < s->bins; ++b) ids[(size_t)b] = b;
        err = cudaMemcpy(e->d_bin_ids, ids.data(), sizeof(int) * ids.size(), cudaMemcpyHostToDevice);
    }
    if (err != cudaSuccess) { cuda_fail(err, "mdaEnsembleCreate"); mdaEnsembleFree(e); return nullptr; }
    return e;
}

void mdaEnsembleFree(void *ens) {
    MdaEnsemble *e = static_cast<MdaEnsemble *>(ens);
    if (!e || e->magic != kEnsembleMagic) { set_error(MDA_ERR_BAD_ARGUMENT, "mdaEnsembleFree: not a handle from mdaEnsembleCreate"); return; }
    if (g_rt.ready) cudaSetDevice(g_rt.device);
    MdaStore *owner = (e->store && e->store->magic == kStoreMagic) ? e->store : nullptr;
    if (owner && owner->stream) cudaStreamSynchronize(owner->stream);
    if (e->d_pos) cudaFree(e->d_pos);
    if (e->d_lnp) cudaFree(e->d_lnp);
    if (e->d_nacc) cudaFree(e->d_nacc);
    if (e->d_bin_ids) cudaFree(e->d_bin_ids);
    if (e->d_status) cudaFree(e->d_status);
    if (e->d_progress) cudaFree(e->d_progress);
    if (e->d_chain) cudaFree(e->d_chain);
    if (e->d_lnp_chain) cudaFree(e->d_lnp_chain);
    if (e->d_prof) cudaFree(e->d_prof);
    if (e->scratch) cudaFree(e->scratch);
    if (e->buckets) cudaFree(e->buckets);
    if (e->copy_stream) cudaStreamDestroy(e->copy_stream);
    for (int i = 0; i < 2; ++i) {
        if (e->chunk_done[i]) cudaEventDestroy(e->chunk_done[i]);
        if (e->copy_done[i]) cudaEventDestroy(e->copy_done[i]);
    }
    e->magic = 0;
    delete e;
    if (owner && --owner->live_ensembles == 0 && owner->closing) destroy_store(owner);   // mdaStoreFree came first
}

int mdaEnsembleSetBinIds(void *ens, const int *binIds) {
    MdaEnsemble *e = as_ensemble(ens, "mdaEnsembleSetBinIds");
    if (!e) return MDA_ERR_BAD_ARGUMENT;
    if (!binIds) return set_error(MDA_ERR_BAD_ARGUMENT, "mdaEnsembleSetBinIds: NULL");
    int rc = ensure_runtime();
    if (rc) return rc;
    MDA_CUDA(cudaStreamSynchronize(e->store->stream), "mdaEnsembleSetBinIds");
    MDA_CUDA(cudaMemcpy(e->d_bin_ids, binIds, sizeof(int) * (size_t)e->store->bins, cudaMemcpyHostToDevice), "mdaEnsembleSetBinIds");
    return MDA_OK;
}

int mdaEnsembleSetState(void *ens, const double *p0) {
    MdaEnsemble *e = as_ensemble(ens, "mdaEnsembleSetState");
    if (!e) return MDA_ERR_BAD_ARGUMENT;
    if (!p0) return set_error(MDA_ERR_BAD_ARGUMENT, "mdaEnsembleSetState: p0 is NULL");
    MdaStore *s = e->store;
    const size_t n_eval = (size_t)s->bins * (size_t)e->walkers;
    LapTimer laps;
    int rc = ensure_runtime();
    if (rc) return rc;
    if ((rc = upload_store(s))) return rc;
    laps.lap("SetState: store upload (if dirty)");
    e->has_state = false;                                  // until everything below has succeeded
    MDA_CUDA(cudaMemcpyAsync(e->d_pos, p0, sizeof(double) * n_eval * s->n_l, cudaMemcpyHostToDevice, s->stream), "ensemble p0 H2D");
    MDA_CUDA(cudaMemsetAsync(e->d_nacc, 0, sizeof(unsigned int) * n_eval, s->stream), "ensemble counters");
    MDA_CUDA(cudaMemsetAsync(e->d_status, 0, sizeof(int), s->stream), "ensemble status");
    if (n_eval && (rc = enqueue_eval(s, e->walkers, e->d_pos, e->d_lnp, kLnPost))) return rc;
    // While the copy and the first log-probabilities are under way: every start position must be finite (emcee raises on
    // NaN / inf positions).  x * 0 is 0 for finite x and NaN otherwise -- one vectorisable pass over the block instead of a
    // million classified compares (this call is inside the e2e region; the pass is a read of pinned memory).
    bool finite = true;
    {
        const size_t count = n_eval * (size_t)s->n_l;
        double acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        size_t i = 0;
        for (; i + 8 <= count; i += 8)
            for (int k = 0; k < 8; ++k) acc[k] += p0[i + k] * 0.0;
        for (; i < count; ++i) acc[0] += p0[i] * 0.0;
        double total = 0.0;
        for (int k = 0; k < 8; ++k) total += acc[k];
        finite = total == 0.0;
    }
    laps.lap("SetState: finiteness of p0 (under the H2D copy)");
    MDA_CUDA(cudaStreamSynchronize(s->stream), "mdaEnsembleSetState");
    laps.lap("SetState: p0 H2D + log-probabilities");
    if (!finite) {
        const size_t count = n_eval * (size_t)s->n_l;
        for (size_t i = 0; i < count; ++i)
            if (!std::isfinite(p0[i]))
                return set_error(MDA_ERR_BAD_ARGUMENT, "mdaEnsembleSetState: at least one parameter value was %s", std::isnan(p0[i]) ? "NaN" : "infinite");
    }
    if (e->state_generation != 0) ++e->epoch;             // a restart on the same ensemble draws from a fresh stream
    e->state_generation = s->generation;
    e->steps = 0;
    e->kept = 0;
    e->has_state = true;
    return MDA_OK;
}

int mdaEnsembleReserveChain(void *ens, long long capacity) {
    MdaEnsemble *e = as_ensemble(ens, "mdaEnsembleReserveChain");
    if (!e) return MDA_ERR_BAD_ARGUMENT;
    if (capacity < 0) return set_error(MDA_ERR_BAD_ARGUMENT, "mdaEnsembleReserveChain: negative capacity");
    int rc = ensure_runtime();
    if (rc) return rc;
    MdaStore *s = e->store;
    MDA_CUDA(cudaStreamSynchronize(s->stream), "mdaEnsembleReserveChain");
    e->kept = 0;
    if (capacity > 0 && capacity <= e->alloc_cap && e->d_chain && e->d_lnp_chain) {
        e->keep_cap = capacity;                            // the chain is step-major: a shorter one is a prefix of the allocation
        return MDA_OK;                                     // (gigabytes of cudaFree + cudaMalloc cost tens of milliseconds)
    }
    if (e->d_chain) cudaFree(e->d_chain);
    if (e->d_lnp_chain) cudaFree(e->d_lnp_chain);
    e->d_chain = e->d_lnp_chain = nullptr;
    e->keep_cap = 0;
    e->alloc_cap = 0;
    if (capacity == 0) return MDA_OK;
    const size_t per_keep = (size_t)s->bins * (size_t)e->walkers;
    MDA_CUDA(cudaMalloc((void **)&e->d_chain, sizeof(double) * per_keep * (size_t)s->n_l * (size_t)capacity), "chain allocation");
    MDA_CUDA(cudaMalloc((void **)&e->d_lnp_chain, sizeof(double) * per_keep * (size_t)capacity), "lnprob chain allocation");
    e->keep_cap = capacity;
    e->alloc_cap = capacity;
    return MDA_OK;
}

int mdaEnsembleAdvance(void *ens, int nSteps, int thin) {
    MdaEnsemble *e = as_ensemble(ens, "mdaEnsembleAdvance");
    if (!e) return MDA_ERR_BAD_ARGUMENT;
    if (nSteps < 0 || thin < 1) return set_error(MDA_ERR_BAD_ARGUMENT, "mdaEnsembleAdvance: nSteps >= 0 and thin >= 1 required");
    if (!e->has_state) return set_error(MDA_ERR_BAD_STATE, "mdaEnsembleAdvance: call mdaEnsembleSetState first");
    if (e->state_generation != e->store->generation)
        return set_error(MDA_ERR_BAD_STATE, "mdaEnsembleAdvance: the store's bins or bounds changed after mdaEnsembleSetState: the stored "
                                            "log-probabilities belong to the old ones, call mdaEnsembleSetState again");
    if ((unsigned long long)(e->steps + nSteps) >= (1ull << 31)) return set_error(MDA_ERR_BAD_ARGUMENT, "mdaEnsembleAdvance: step counter would exceed 2^31");
    int rc = ensure_runtime();
    if (rc) return rc;
    MdaStore *s = e->store;
    if (nSteps == 0 || s->bins == 0) return MDA_OK;
    EnsembleArgs args;
    memset(&args, 0, sizeof(args));
    args.consts = s->d_consts;
    args.n_pts = s->d_npts;
    args.bin_ids = e->d_bin_ids;
    args.pos = e->d_pos;
    args.lnp = e->d_lnp;
    args.nacc = e->d_nacc;
    args.chain = e->d_chain;
    args.lnp_chain = e->d_lnp_chain;
    args.status = e->d_status;
    args.prof = e->d_prof;
    args.bins = s->bins;
    args.walkers = e->walkers;
    args.n_l = s->n_l;
    args.n_pad = s->n_pad;
    args.n_steps = nSteps;
    args.thin = thin;
    args.step0 = e->steps;
    args.keep0 = e->kept;
    args.keep_cap = e->keep_cap;
    const unsigned long long key = e->seed + e->epoch * 0x9E3779B97F4A7C15ull;     // epoch 0 (a fresh ensemble): the seed itself
    args.seed_lo = (unsigned int)(key & 0xFFFFFFFFull);
    args.seed_hi = (unsigned int)(key >> 32);
    args.a = e->a;
    {
        uint32_t rk[20];
        philox_round_keys(args.seed_lo, args.seed_hi, rk);
        for (int r = 0; r < 20; ++r) args.round_keys[r] = rk[r];
        int exponent = 0;
        args.a_inv = (std::frexp(e->a, &exponent) == 0.5) ? 1.0 / e->a : 0.0;      // a power of two: the division is a multiplication
    }
    for (int l = 0; l < s->n_l; ++l) { args.lo[l] = s->lo[l]; args.hi[l] = s->hi[l]; }
    args.consts_mma = s->d_consts_mma;
    args.n_tiles = s->n_tiles;
    args.quad = s->d_quad;
    const EnsembleShape shape = ensemble_shape_for(s, e->walkers);
    if (shape.qws) {
        const int grid = qws_grid(s, shape);
        if (grid < 1) return set_error(MDA_ERR_CUDA, "mdaEnsembleAdvance: the stepping kernel cannot be resident");
        args.chunk_steps = nSteps;
        args.n_chunks = 1;
        args.progress = e->d_progress;
        args.launch_serial = (++e->launches) << 32;
        MDA_CUDA(launch_ensemble_qws(args, shape, grid, s->stream), "ensemble stepping launch (triangular factor, warp-specialised)");
    } else if (shape.quad) {
        const SteppingPlan plan = plan_quad_stepping(s, shape, nSteps);
        if (plan.grid < 1) return set_error(MDA_ERR_CUDA, "mdaEnsembleAdvance: the stepping kernel cannot be resident");
        args.chunk_steps = plan.chunk_steps;
        args.n_chunks = plan.n_chunks;
        args.progress = e->d_progress;
        args.launch_serial = (++e->launches) << 32;
        MDA_CUDA(launch_ensemble_quad(args, shape, plan.grid, g_rt.sm_count, s->stream), "ensemble stepping launch (triangular factor)");
    } else if (shape.mma) {
        const SteppingPlan plan = plan_stepping(s, shape, nSteps);
        if (plan.grid < 1) return set_error(MDA_ERR_CUDA, "mdaEnsembleAdvance: the stepping kernel cannot be resident");
        args.chunk_steps = plan.chunk_steps;
        args.n_chunks = plan.n_chunks;
        args.progress = e->d_progress;
        args.launch_serial = (++e->launches) << 32;
        if (shape.ws) MDA_CUDA(launch_ensemble_ws(args, shape, plan.grid, s->stream), "ensemble stepping launch (DMMA, warp-specialised)");
        else MDA_CUDA(launch_ensemble_mma(args, shape, plan.grid, s->stream), "ensemble stepping launch (DMMA)");
    } else {
        MDA_CUDA(launch_ensemble(args, shape, s->stream), "ensemble stepping launch");
    }
    g_launches.fetch_add(1, std::memory_order_relaxed);
    const long long new_keeps = (e->steps + nSteps) / thin - e->steps / thin;
    e->steps += nSteps;
    if (e->d_chain) e->kept = std::min(e->keep_cap, e->kept + new_keeps);
    return MDA_OK;
}

int mdaEnsembleDescribe(void *ens, int nSteps, char *text, int textLen) {
    MdaEnsemble *e = as_ensemble(ens, "mdaEnsembleDescribe");
    if (!e) return MDA_ERR_BAD_ARGUMENT;
    if (!text || textLen < 1 || nSteps < 1) return set_error(MDA_ERR_BAD_ARGUMENT, "mdaEnsembleDescribe: a buffer and nSteps >= 1 required");
    int rc = ensure_runtime();
    if (rc) return rc;
    MdaStore *s = e->store;
    const EnsembleShape shape = ensemble_shape_for(s, e->walkers);
    if (shape.qws) {
        snprintf(text, (size_t)textLen,
                 "ensemble_qws_kernel<L=%d,%s>: chi^2 = |T [a; -1]|^2 from the bin's triangular factor (%d multiply-adds per evaluation), 8 walker warps "
                 "(a thread per walker of the active half) + 4 draw warps (Philox draws two half-steps ahead)%s, one barrier per half-step, "
                 "bulk-TMA state load and chain stores, %zu B shared memory; %d CTAs for %d bins%s",
                 s->n_l, shape.treg ? "T in registers via setmaxnreg 208/56/40" : "T in shared memory", quad_doubles(s->n_l),
                 shape.treg ? " + a copy warpgroup (bulk copies)" : " (one of their threads issues the bulk copies)", shape.smem_bytes,
                 qws_grid(s, shape), s->bins, qws_grid(s, shape) < s->bins ? " (every CTA carries the same number of bin-steps; cooperative launch)" : "");
    } else if (shape.quad) {
        snprintf(text, (size_t)textLen,
                 "ensemble_quad_kernel<L=%d>: chi^2 = |T [a; -1]|^2 from the bin's triangular factor (%d multiply-adds per evaluation), %d threads x %d walkers, "
                 "one barrier per half-step, bulk-TMA chain stores, %zu B shared memory; %d CTAs, %d bins x %d chunks of %d steps",
                 s->n_l, quad_doubles(s->n_l), shape.threads, shape.wpt, shape.smem_bytes, plan_quad_stepping(s, shape, nSteps).grid, s->bins,
                 plan_quad_stepping(s, shape, nSteps).n_chunks, plan_quad_stepping(s, shape, nSteps).chunk_steps);
    } else if (shape.mma) {
        const SteppingPlan plan = plan_stepping(s, shape, nSteps);
        if (shape.ws)
            snprintf(text, (size_t)textLen,
                     "ensemble_ws_kernel<L=%d,WT=%d>: fp64 DMMA.8x8x4 chi^2, %d walker + %d math warps per CTA meeting at named barriers, %zu B shared memory; "
                     "%d CTAs walk %d bins x %d chunks of %d steps",
                     s->n_l, shape.wpt, shape.groups, shape.threads / 32 - shape.groups, shape.smem_bytes, plan.grid, s->bins,
                     plan.n_chunks, plan.chunk_steps);
        else
            snprintf(text, (size_t)textLen,
                     "ensemble_mma_kernel<L=%d,WT=%d,%s>: fp64 DMMA.8x8x4 chi^2, %d warps per CTA (one item of %d walkers each, one barrier per half-step), %zu B shared memory; "
                     "%d CTAs walk %d bins x %d chunks of %d steps",
                     s->n_l, shape.wpt, shape.stage == 1 ? "draw warps" : "two CTAs per SM", shape.groups, 8 * shape.wpt, shape.smem_bytes, plan.grid, s->bins,
                     plan.n_chunks, plan.chunk_steps);
    } else {
        snprintf(text, (size_t)textLen, "ensemble_kernel<L=%d,WPT=%d,%s>: DFMA chi^2, %d threads x %d angle groups per CTA, %zu B shared memory; one CTA per bin (%d)",
                 s->n_l, shape.wpt, shape.global_state ? "state in HBM" : "state in shared memory", shape.threads, shape.groups, shape.smem_bytes, s->bins);
    }
    // the conditioning decision (factor_bin): which bin, if any, ruled the triangular-factor stepper out
    double worst = 0.0;
    int worst_bin = -1;
    ensure_factors(s);
    for (int b = 0; b < s->bins; ++b) if (!(s->kappa[(size_t)b] <= worst)) { worst = s->kappa[(size_t)b]; worst_bin = b; }
    const size_t used = strlen(text);
    if (used + 1 < (size_t)textLen)
        snprintf(text + used, (size_t)textLen - used, "; conditioning: max kappa = |d|/sqrt(min chi^2) = %.3g (bin %d), limit %.3g for the triangular factor -> %s",
                 worst, worst_bin, kQuadKappaMax,
                 shape.quad ? "chi^2 from the triangular factor (within 4 eps kappa of the reference's arithmetic)"
                            : (worst <= kQuadKappaMax ? "angle-by-angle stepper on request" : "reference-order stepper (angle by angle, C/ChiSquare.c:188-210)"));
    return MDA_OK;
}

int mdaEnsembleSynchronize(void *ens) {
    MdaEnsemble *e = as_ensemble(ens, "mdaEnsembleSynchronize");
    if (!e) return MDA_ERR_BAD_ARGUMENT;
    int rc = ensure_runtime();
    if (rc) return rc;
    int status = 0;
    MDA_CUDA(cudaMemcpyAsync(&status, e->d_status, sizeof(int), cudaMemcpyDeviceToHost, e->store->stream), "ensemble status read");
    MDA_CUDA(cudaStreamSynchronize(e->store->stream), "mdaEnsembleSynchronize");
    if (status & 4) {
        // bins were stopped at different steps: the ensemble is torn, nobody may continue from it
        e->has_state = false;
        e->kept = 0;
        cudaMemsetAsync(e->d_status, 0, sizeof(int), e->store->stream);
        return set_error(MDA_ERR_CUDA, "ensemble stepping gave up: a piece of work waited seconds for its predecessor (is another kernel holding SMs of "
                                       "this device?); the ensemble state is invalid, call mdaEnsembleSetState again");
    }
    if (status & 1) return set_error(MDA_ERR_BAD_STATE, "lnprob returned NaN during ensemble stepping");
    return MDA_OK;
}

long long mdaEnsembleSteps(void *ens) { MdaEnsemble *e = as_ensemble(ens, "mdaEnsembleSteps"); return e ? e->steps : -1; }
long long mdaEnsembleKept(void *ens) { MdaEnsemble *e = as_ensemble(ens, "mdaEnsembleKept"); return e ? e->kept : -1; }

int mdaEnsembleGetState(void *ens, double *pos, double *lnProb, long long *nAccepted) {
    MdaEnsemble *e = as_ensemble(ens, "mdaEnsembleGetState");
    if (!e) return MDA_ERR_BAD_ARGUMENT;
    int rc = mdaEnsembleSynchronize(ens);
    if (rc) return rc;
    MdaStore *s = e->store;
    const size_t n_eval = (size_t)s->bins * (size_t)e->walkers;
    if (pos) MDA_CUDA(cudaMemcpy(pos, e->d_pos, sizeof(double) * n_eval * s->n_l, cudaMemcpyDeviceToHost), "ensemble positions D2H");
    if (lnProb) MDA_CUDA(cudaMemcpy(lnProb, e->d_lnp, sizeof(double) * n_eval, cudaMemcpyDeviceToHost), "ensemble lnprob D2H");
    if (nAccepted) {
        std::vector<unsigned int> tmp(n_eval ? n_eval : 1);
        MDA_CUDA(cudaMemcpy(tmp.data(), e->d_nacc, sizeof(unsigned int) * n_eval, cudaMemcpyDeviceToHost), "ensemble counters D2H");
        for (size_t i = 0; i < n_eval; ++i) nAccepted[i] = (long long)tmp[i];
    }
    return MDA_OK;
}

int mdaEnsembleGetChain(void *ens, long long first, long long count, double *chain, double *lnProb) {
    MdaEnsemble *e = as_ensemble(ens, "mdaEnsembleGetChain");
    if (!e) return MDA_ERR_BAD_ARGUMENT;
    if (first < 0 || count < 0 || first + count > e->kept)
        return set_error(MDA_ERR_BAD_ARGUMENT, "mdaEnsembleGetChain: [%lld, %lld) outside the %lld kept samples", first, first + count, e->kept);
    int rc = mdaEnsembleSynchronize(ens);
    if (rc) return rc;
    if (count == 0) return MDA_OK;
    MdaStore *s = e->store;
    const size_t slab = (size_t)s->bins * (size_t)e->walkers * (size_t)s->n_l * sizeof(double);   // one kept sample, all bins
    const size_t lslab = (size_t)s->bins * (size_t)e->walkers * sizeof(double);
    if (chain)
        MDA_CUDA(cudaMemcpy(chain, reinterpret_cast<const char *>(e->d_chain) + slab * (size_t)first, slab * (size_t)count,
                            cudaMemcpyDeviceToHost), "chain D2H");
    if (lnProb)
        MDA_CUDA(cudaMemcpy(lnProb, reinterpret_cast<const char *>(e->d_lnp_chain) + lslab * (size_t)first, lslab * (size_t)count,
                            cudaMemcpyDeviceToHost), "lnprob chain D2H");
    return MDA_OK;
}

int mdaEnsembleProfile(void *ens, long long *cycles) {
    MdaEnsemble *e = as_ensemble(ens, "mdaEnsembleProfile");
    if (!e) return MDA_ERR_BAD_ARGUMENT;
    int rc = ensure_runtime();
    if (rc) return rc;
    const size_t bytes = sizeof(long long) * 128 * (size_t)(e->store->bins ? e->store->bins : 1);
    if (!e->d_prof) {                                   // first call switches the counters on
        MDA_CUDA(cudaStreamSynchronize(e->store->stream), "mdaEnsembleProfile");
        MDA_CUDA(cudaMalloc((void **)&e->d_prof, bytes), "profile counters");
        MDA_CUDA(cudaMemset(e->d_prof, 0, bytes), "profile counters");
    }
    MDA_CUDA(cudaStreamSynchronize(e->store->stream), "mdaEnsembleProfile");
    if (cycles) MDA_CUDA(cudaMemcpy(cycles, e->d_prof, bytes, cudaMemcpyDeviceToHost), "profile counters D2H");
    return MDA_OK;
}

int mdaEnsembleRunToHost(void *ens, int nSteps, int thin, int chunkSteps, double *hostChain, double *hostLnProb) {
    MdaEnsemble *e = as_ensemble(ens, "mdaEnsembleRunToHost");
    if (!e) return MDA_ERR_BAD_ARGUMENT;
    if (nSteps < 0 || thin < 1) return set_error(MDA_ERR_BAD_ARGUMENT, "mdaEnsembleRunToHost: nSteps >= 0 and thin >= 1 required");
    if (!e->has_state) return set_error(MDA_ERR_BAD_STATE, "mdaEnsembleRunToHost: call mdaEnsembleSetState first");
    if (e->state_generation != e->store->generation)
        return set_error(MDA_ERR_BAD_STATE, "mdaEnsembleRunToHost: the store changed after mdaEnsembleSetState, call it again");
    if (e->steps % thin != 0) return set_error(MDA_ERR_BAD_STATE, "mdaEnsembleRunToHost: %lld steps already taken is not a multiple of thin = %d", e->steps, thin);
    int rc = ensure_runtime();
    if (rc) return rc;
    MdaStore *s = e->store;
    if (nSteps == 0 || s->bins == 0) return MDA_OK;
    const size_t slab = (size_t)s->bins * (size_t)e->walkers * (size_t)s->n_l * sizeof(double);   // one kept sample, all bins
    const size_t lslab = (size_t)s->bins * (size_t)e->walkers * sizeof(double);
    if (chunkSteps <= 0) {                      // default: ~64 MB of chain per chunk, at least one kept sample
        const size_t per_keep = slab + lslab;
        long long keeps = (long long)((64u << 20) / (per_keep ? per_keep : 1));
        if (keeps < 1) keeps = 1;
        chunkSteps = (int)std::min<long long>(keeps * thin, 1 << 20);
    }
    chunkSteps = std::max(thin, chunkSteps / thin * thin);
    const long long chunk_keep = chunkSteps / thin, total_keep = nSteps / thin;
    LapTimer laps;
    if ((rc = mdaEnsembleReserveChain(ens, 2 * chunk_keep))) return rc;
    laps.lap("RunToHost: chain ring reservation");
    if (!e->copy_stream) {
        MDA_CUDA(cudaStreamCreateWithFlags(&e->copy_stream, cudaStreamNonBlocking), "copy stream");
        for (int i = 0; i < 2; ++i) {
            MDA_CUDA(cudaEventCreateWithFlags(&e->chunk_done[i], cudaEventDisableTiming), "chunk event");
            MDA_CUDA(cudaEventCreateWithFlags(&e->copy_done[i], cudaEventDisableTiming), "copy event");
        }
    }
    long long done_keep = 0;
    int remaining = nSteps;
    for (int c = 0; remaining > 0; ++c) {
        const int buf = c & 1;
        const int steps = std::min(remaining, chunkSteps);
        const long long keeps = (e->steps + steps) / thin - e->steps / thin;
        if (c >= 2) MDA_CUDA(cudaStreamWaitEvent(s->stream, e->copy_done[buf], 0), "wait for the chunk buffer to drain");
        e->kept = buf * chunk_keep;                                  // the kernel appends from this slot
        if ((rc = mdaEnsembleAdvance(ens, steps, thin))) return rc;
        MDA_CUDA(cudaEventRecord(e->chunk_done[buf], s->stream), "chunk event record");
        MDA_CUDA(cudaStreamWaitEvent(e->copy_stream, e->chunk_done[buf], 0), "copy waits for the chunk");
        if (keeps > 0 && done_keep + keeps <= total_keep) {
            if (hostChain)
                MDA_CUDA(cudaMemcpyAsync(reinterpret_cast<char *>(hostChain) + slab * (size_t)done_keep,
                                         reinterpret_cast<const char *>(e->d_chain) + slab * (size_t)(buf * chunk_keep),
                                         slab * (size_t)keeps, cudaMemcpyDeviceToHost, e->copy_stream), "chain chunk D2H");
            if (hostLnProb)
                MDA_CUDA(cudaMemcpyAsync(reinterpret_cast<char *>(hostLnProb) + lslab * (size_t)done_keep,
                                         reinterpret_cast<const char *>(e->d_lnp_chain) + lslab * (size_t)(buf * chunk_keep),
                                         lslab * (size_t)keeps, cudaMemcpyDeviceToHost, e->copy_stream), "lnprob chunk D2H");
        }
        MDA_CUDA(cudaEventRecord(e->copy_done[buf], e->copy_stream), "copy event record");
        done_keep += keeps;
        remaining -= steps;
    }
    laps.lap("RunToHost: launches and copies enqueued");
    MDA_CUDA(cudaStreamSynchronize(e->copy_stream), "mdaEnsembleRunToHost (copies)");
    laps.lap("RunToHost: copies drained");
    e->kept = 0;                                                     // the device chain was only a staging ring
    rc = mdaEnsembleSynchronize(ens);
    laps.lap("RunToHost: final synchronise");
    return rc;
}

namespace {
struct StageTimer {
    bool on = getenv("MDA_SUMMARY_TIMING") != nullptr;
    std::chrono::steady_clock::time_point t = std::chrono::steady_clock::now();
    void lap(const char *what) {
        if (!on) return;
        const auto now = std::chrono::steady_clock::now();
        fprintf(stderr, "[mda summary] %-28s %8.3f ms\n", what, std::chrono::duration<double, std::milli>(now - t).count());
        t = now;
    }
};
}  // namespace

namespace {

// Device scratch owned by the ensemble: one allocation reused by every summary call.
int ensure_scratch(MdaEnsemble *e, size_t bytes) {
    if (bytes <= e->scratch_bytes) return MDA_OK;
    MDA_CUDA(cudaStreamSynchronize(e->store->stream), "summary scratch");
    if (e->scratch) cudaFree(e->scratch);
    if (e->buckets) cudaFree(e->buckets);                 // (released too, to make room; re-allocated on demand)
    e->scratch = nullptr;
    e->scratch_bytes = 0;
    e->buckets = nullptr;
    e->bucket_bytes = 0;
    MDA_CUDA(cudaMalloc(&e->scratch, bytes), "summary scratch");
    e->scratch_bytes = bytes;
    return MDA_OK;
}

double key_to_double(unsigned long long k) {
    const unsigned long long u = (k >> 63) ? (k & 0x7FFFFFFFFFFFFFFFull) : ~k;
    double x;
    memcpy(&x, &u, sizeof(x));
    return x;
}

// numpy.percentile(..., method="linear") pieces (numpy/lib/_function_base_impl.py, numpy 2.x):
// virtual index (n - 1) * q with q = percent / 100, neighbours floor / floor + 1, _lerp.
struct RankPair { long long lower, upper; double gamma; };
RankPair percentile_ranks(long long n, double percent) {
    const double q = percent / 100.0;
    const double vi = (double)(n - 1) * q;
    RankPair r;
    if (vi >= (double)(n - 1)) { r.lower = r.upper = n - 1; r.gamma = 0.0; return r; }
    if (vi < 0.0) { r.lower = r.upper = 0; r.gamma = 0.0; return r; }
    const double fl = std::floor(vi);
    r.lower = (long long)fl;
    r.upper = r.lower + 1;
    r.gamma = vi - fl;
    return r;
}
double numpy_lerp(double a, double b, double t) {
    const double diff = b - a;
    double out = a + diff * t;
    if (t >= 0.5) out = b - diff * (1.0 - t);
    return out;
}

// Exact order statistics for `n_ranks` ranks per column (ranks[cols][n_ranks]) -> values.
int select_ranks(MdaEnsemble *e, long long k0, long long k1, const std::vector<unsigned long long> &mm,
                 const std::vector<long long> &ranks, int n_ranks, std::vector<double> &values, StageTimer &timer) {
    MdaStore *s = e->store;
    const size_t cols = (size_t)s->bins * (size_t)s->n_l, n_state = cols * (size_t)n_ranks;
    std::vector<int> top(cols);
    std::vector<unsigned long long> prefix(n_state, 0ull);
    std::vector<long long> remaining(n_state, 0);
    std::vector<unsigned int> count(n_state, 0u);
    for (size_t c = 0; c < cols; ++c) {
        const unsigned long long x = mm[2 * c] ^ mm[2 * c + 1];
        int t = 0;
        while (t < 64 && (x >> t) != 0ull) ++t;               // bits [0, t) may differ inside the column
        top[c] = t;
        const unsigned long long common = t >= 64 ? 0ull : (mm[2 * c] >> t) << t;
        for (int r = 0; r < n_ranks; ++r) {
            prefix[c * n_ranks + r] = common;
            remaining[c * n_ranks + r] = ranks[c * n_ranks + r];
        }
    }
    auto up = [](size_t x) { return (x + 255) & ~(size_t)255; };
    const size_t b_prefix = up(8 * n_state), b_rem = up(8 * n_state), b_count = up(4 * n_state), b_top = up(4 * cols),
                 b_hist = up(4 * cols * 4096), b_lead = up(n_state), b_off = up(8 * n_state), b_cur = up(4 * n_state);
    const size_t fixed = b_prefix + b_rem + b_count + b_top + b_hist + b_lead + b_off + b_cur;
    int rc = ensure_scratch(e, fixed);
    if (rc) return rc;
    auto carve = [&](char *base) {
        struct { unsigned long long *prefix; long long *rem; unsigned int *count; int *top; unsigned int *hist;
                 unsigned char *lead; long long *off; unsigned int *cur; char *end; } p;
        p.prefix = reinterpret_cast<unsigned long long *>(base); base += b_prefix;
        p.rem = reinterpret_cast<long long *>(base); base += b_rem;
        p.count = reinterpret_cast<unsigned int *>(base); base += b_count;
        p.top = reinterpret_cast<int *>(base); base += b_top;
        p.hist = reinterpret_cast<unsigned int *>(base); base += b_hist;
        p.lead = reinterpret_cast<unsigned char *>(base); base += b_lead;
        p.off = reinterpret_cast<long long *>(base); base += b_off;
        p.cur = reinterpret_cast<unsigned int *>(base); base += b_cur;
        p.end = base;
        return p;
    };
    auto d = carve(static_cast<char *>(e->scratch));
    cudaError_t err = cudaMemcpyAsync(d.prefix, prefix.data(), 8 * n_state, cudaMemcpyHostToDevice, s->stream);
    if (err == cudaSuccess) err = cudaMemcpyAsync(d.rem, remaining.data(), 8 * n_state, cudaMemcpyHostToDevice, s->stream);
    if (err == cudaSuccess) err = cudaMemcpyAsync(d.top, top.data(), 4 * cols, cudaMemcpyHostToDevice, s->stream);
    if (err == cudaSuccess)
        err = launch_select_first(e->d_chain, k0, k1, s->bins, e->walkers, s->n_l, d.prefix, d.rem, d.count, d.top, d.hist,
                                  n_ranks, g_rt.sm_count, s->stream);
    if (err == cudaSuccess) err = cudaMemcpyAsync(prefix.data(), d.prefix, 8 * n_state, cudaMemcpyDeviceToHost, s->stream);
    if (err == cudaSuccess) err = cudaMemcpyAsync(count.data(), d.count, 4 * n_state, cudaMemcpyDeviceToHost, s->stream);
    if (err == cudaSuccess) err = cudaStreamSynchronize(s->stream);
    if (err != cudaSuccess) return cuda_fail(err, "order-statistic selection (first pass)");
    timer.lap("  first digit pass");
    g_launches.fetch_add(2, std::memory_order_relaxed);
    // buckets in play: ranks of a column that fell into the same bucket share one buffer
    std::vector<unsigned char> lead(n_state, 0);
    std::vector<long long> offset(n_state, 0);
    long long total = 0;
    for (size_t c = 0; c < cols; ++c) {
        if (top[c] <= 12) continue;                            // decided by the first digit already
        for (int r = 0; r < n_ranks; ++r) {
            const size_t i = c * n_ranks + r;
            int owner = r;
            for (int q = 0; q < r; ++q)
                if (prefix[c * n_ranks + q] == prefix[i]) { owner = q; break; }
            if (owner == r) { lead[i] = 1; offset[i] = total; total += (long long)count[i]; }
            else offset[i] = offset[c * n_ranks + owner];
        }
    }
    // the bucket buffers have their own allocation, so growing them never moves the state above
    const size_t b_buf = up(8 * (size_t)(total > 0 ? total : 1));
    if (b_buf > e->bucket_bytes) {
        MDA_CUDA(cudaStreamSynchronize(s->stream), "bucket buffers");
        if (e->buckets) cudaFree(e->buckets);
        e->buckets = nullptr;
        e->bucket_bytes = 0;
        MDA_CUDA(cudaMalloc(&e->buckets, b_buf), "bucket buffers");
        e->bucket_bytes = b_buf;
    }
    unsigned long long *d_buf = static_cast<unsigned long long *>(e->buckets);
    err = cudaMemcpyAsync(d.lead, lead.data(), n_state, cudaMemcpyHostToDevice, s->stream);
    if (err == cudaSuccess) err = cudaMemcpyAsync(d.off, offset.data(), 8 * n_state, cudaMemcpyHostToDevice, s->stream);
    if (err == cudaSuccess)
        err = launch_select_finish(e->d_chain, k0, k1, s->bins, e->walkers, s->n_l, d.prefix, d.rem, d.count, d.top, d.lead,
                                   d.off, d.cur, d_buf, n_ranks, g_rt.sm_count, s->stream);
    if (err == cudaSuccess) err = cudaMemcpyAsync(prefix.data(), d.prefix, 8 * n_state, cudaMemcpyDeviceToHost, s->stream);
    if (err == cudaSuccess) err = cudaStreamSynchronize(s->stream);
    if (err != cudaSuccess) return cuda_fail(err, "order-statistic selection (gather / finish)");
    timer.lap("  gather + finish");
    g_launches.fetch_add(2, std::memory_order_relaxed);
    values.resize(n_state);
    for (size_t i = 0; i < n_state; ++i) values[i] = key_to_double(prefix[i]);
    return MDA_OK;
}

}  // namespace

int mdaEnsembleSummarize(void *ens, long long firstKept, double confidenceInterval, int numBins,
                         double *points, double *peaks, double *acceptFrac) {
    StageTimer timer;
    MdaEnsemble *e = as_ensemble(ens, "mdaEnsembleSummarize");
    if (!e) return MDA_ERR_BAD_ARGUMENT;
    if (firstKept < 0 || firstKept >= e->kept)
        return set_error(MDA_ERR_BAD_ARGUMENT, "mdaEnsembleSummarize: firstKept %lld outside the %lld kept samples", firstKept, e->kept);
    if (!(confidenceInterval > 0.0 && confidenceInterval <= 1.0) || numBins < 1)
        return set_error(MDA_ERR_BAD_ARGUMENT, "mdaEnsembleSummarize: confidenceInterval in (0, 1] and numBins >= 1 required");
    int rc = mdaEnsembleSynchronize(ens);
    if (rc) return rc;
    MdaStore *s = e->store;
    const size_t B = (size_t)s->bins, L = (size_t)s->n_l, cols = B * L;
    if (cols == 0) return MDA_OK;
    if (s->n_l > 32) return set_error(MDA_ERR_BAD_ARGUMENT, "mdaEnsembleSummarize: numL %d > 32", s->n_l);
    const long long k0 = firstKept, k1 = e->kept;
    const long long n = (k1 - k0) * (long long)e->walkers;            // samples per column

    // 1. smallest / largest sample per column
    std::vector<unsigned long long> mm(2 * cols);
    for (size_t c = 0; c < cols; ++c) { mm[2 * c] = ~0ull; mm[2 * c + 1] = 0ull; }
    if ((rc = ensure_scratch(e, sizeof(unsigned long long) * 2 * cols))) return rc;
    unsigned long long *d_mm = static_cast<unsigned long long *>(e->scratch);
    cudaError_t err = cudaMemcpyAsync(d_mm, mm.data(), sizeof(unsigned long long) * 2 * cols, cudaMemcpyHostToDevice, s->stream);
    if (err == cudaSuccess) err = launch_chain_minmax(e->d_chain, k0, k1, s->bins, e->walkers, s->n_l, d_mm, g_rt.sm_count, s->stream);
    if (err == cudaSuccess) err = cudaMemcpyAsync(mm.data(), d_mm, sizeof(unsigned long long) * 2 * cols, cudaMemcpyDeviceToHost, s->stream);
    if (err == cudaSuccess) err = cudaStreamSynchronize(s->stream);
    if (err != cudaSuccess) return cuda_fail(err, "summary min/max");
    g_launches.fetch_add(1, std::memory_order_relaxed);
    timer.lap("sync + min/max");

    // 2. the histogram of find_most_likely_values (mda.py:1466-1473) needs only min and max; edges,
    //    counts, argmax and the cumulative count up to the peak bin all stay on the device
    const size_t nb = (size_t)numBins;
    std::vector<int> peak_bin(cols, 0);
    std::vector<long long> peak_cum(cols, 0);
    if (peaks) {
        auto up = [](size_t x) { return (x + 255) & ~(size_t)255; };
        const size_t b_mm = up(16 * cols), b_edges = up(8 * cols * (nb + 1)), b_hist = up(4 * cols * nb), b_bin = up(4 * cols), b_cum = up(8 * cols);
        if ((rc = ensure_scratch(e, b_mm + b_edges + b_hist + b_bin + b_cum))) return rc;
        char *base = static_cast<char *>(e->scratch);
        unsigned long long *d_mm2 = reinterpret_cast<unsigned long long *>(base);
        double *d_edges = reinterpret_cast<double *>(base + b_mm);
        unsigned int *d_hist = reinterpret_cast<unsigned int *>(base + b_mm + b_edges);
        int *d_bin = reinterpret_cast<int *>(base + b_mm + b_edges + b_hist);
        long long *d_cum = reinterpret_cast<long long *>(base + b_mm + b_edges + b_hist + b_bin);
        err = cudaMemcpyAsync(d_mm2, mm.data(), 16 * cols, cudaMemcpyHostToDevice, s->stream);
        if (err == cudaSuccess)
            err = launch_chain_hist(e->d_chain, k0, k1, s->bins, e->walkers, s->n_l, d_mm2, d_edges, numBins, d_hist, d_bin, d_cum,
                                    g_rt.sm_count, s->stream);
        if (err == cudaSuccess) err = cudaMemcpyAsync(peak_bin.data(), d_bin, 4 * cols, cudaMemcpyDeviceToHost, s->stream);
        if (err == cudaSuccess) err = cudaMemcpyAsync(peak_cum.data(), d_cum, 8 * cols, cudaMemcpyDeviceToHost, s->stream);
        if (err == cudaSuccess) err = cudaStreamSynchronize(s->stream);
        if (err != cudaSuccess) return cuda_fail(err, "summary histogram");
        g_launches.fetch_add(3, std::memory_order_relaxed);
    }
    timer.lap("edges + histogram");

    // 3. all percentiles of both estimators in ONE selection: six ranks for calc_param_values
    //    (mda.py:1426-1431), six around the histogram peak (mda.py:1468-1489)
    const double pct[3] = {100.0 * (0.5 - confidenceInterval / 2.0), 100.0 * 0.5, 100.0 * (0.5 + confidenceInterval / 2.0)};
    const int n_ranks = peaks ? 12 : 6;
    std::vector<long long> ranks(cols * (size_t)n_ranks, 0);
    std::vector<RankPair> rp(cols * 6);
    std::vector<char> flat(cols, 0);
    for (size_t c = 0; c < cols; ++c) {
        for (int i = 0; i < 3; ++i) rp[c * 6 + i] = percentile_ranks(n, pct[i]);
        if (peaks) {
            const double first = key_to_double(mm[2 * c]), last = key_to_double(mm[2 * c + 1]);
            if (1.0e-7 > std::fabs(last - first)) {
                flat[c] = 1;
                for (int i = 3; i < 6; ++i) rp[c * 6 + i] = RankPair{0, 0, 0.0};
            } else {
                // quantile = sum(hist[0..argmax]) / len(values): integer counts sum exactly in double
                double quantile = (double)peak_cum[c];
                quantile /= (double)n;
                const double peak_percentile = 100.0 * quantile;
                double lo_p = peak_percentile - 100.0 * (confidenceInterval / 2.0);
                if (lo_p < 0.0) lo_p = 0.0;
                double hi_p = peak_percentile + 100.0 * (confidenceInterval / 2.0);
                if (hi_p > 100.0) hi_p = 100.0;
                const double p3[3] = {lo_p, peak_percentile, hi_p};
                for (int i = 0; i < 3; ++i) rp[c * 6 + 3 + i] = percentile_ranks(n, p3[i]);
            }
        }
        for (int i = 0; i < n_ranks / 2; ++i) {
            ranks[c * n_ranks + 2 * i] = rp[c * 6 + i].lower;
            ranks[c * n_ranks + 2 * i + 1] = rp[c * 6 + i].upper;
        }
    }
    std::vector<double> vals;
    if ((rc = select_ranks(e, k0, k1, mm, ranks, n_ranks, vals, timer))) return rc;
    for (size_t c = 0; c < cols; ++c) {
        double v[6];
        for (int i = 0; i < n_ranks / 2; ++i)
            v[i] = numpy_lerp(vals[c * n_ranks + 2 * i], vals[c * n_ranks + 2 * i + 1], rp[c * 6 + i].gamma);
        if (points) {
            points[c * 3 + 0] = v[1];
            points[c * 3 + 1] = v[2] - v[1];
            points[c * 3 + 2] = v[1] - v[0];
        }
        if (peaks) {
            if (flat[c]) { peaks[c * 3] = key_to_double(mm[2 * c]); peaks[c * 3 + 1] = 0.0; peaks[c * 3 + 2] = 0.0; }
            else {
                peaks[c * 3 + 0] = v[4];
                peaks[c * 3 + 1] = v[5] - v[4];
                peaks[c * 3 + 2] = v[4] - v[3];
            }
        }
    }
    timer.lap("order statistics + lerp");

    // 4. acceptance fraction (mda.py:1316: np.mean(sampler.acceptance_fraction))
    if (acceptFrac) {
        std::vector<unsigned int> acc(B * (size_t)e->walkers);
        MDA_CUDA(cudaMemcpy(acc.data(), e->d_nacc, sizeof(unsigned int) * acc.size(), cudaMemcpyDeviceToHost), "acceptance counters D2H");
        for (size_t b = 0; b < B; ++b) {
            double sum = 0.0;
            for (int w = 0; w < e->walkers; ++w) sum += (double)acc[b * e->walkers + w] / (double)(e->steps > 0 ? e->steps : 1);
            acceptFrac[b] = sum / (double)e->walkers;
        }
    }
    return MDA_OK;
}

// emcee 2.x's EnsembleSampler.get_autocorr_time -> autocorr.integrated_time(np.mean(chain, axis=0), axis=0, ...),
// as mda.py:1307-1308 calls it (emcee is not part of the reference tree: the published algorithm restated, see
// oracle/autocorr_oracle.py).  The ensemble mean and the autocorrelation sums come from the device chain; the
// window search below is emcee's loop over M.
int mdaEnsembleAutocorrTime(void *ens, int low, int high, int step, double c, int fast, double *tau, int *window) {
    MdaEnsemble *e = as_ensemble(ens, "mdaEnsembleAutocorrTime");
    if (!e) return MDA_ERR_BAD_ARGUMENT;
    if (!tau) return set_error(MDA_ERR_BAD_ARGUMENT, "mdaEnsembleAutocorrTime: NULL result");
    if (low < 1 || step < 1 || !(c > 0.0) || high < 0)
        return set_error(MDA_ERR_BAD_ARGUMENT, "mdaEnsembleAutocorrTime: low >= 1, step >= 1, c > 0 and high >= 0 (0: default) required");
    int rc = mdaEnsembleSynchronize(ens);
    if (rc) return rc;
    MdaStore *s = e->store;
    const size_t B = (size_t)s->bins, L = (size_t)s->n_l, cols = B * L;
    if (cols == 0) return MDA_OK;
    const long long N = e->kept;                                      // the whole kept chain, burn-in included (sampler.chain)
    if (N < 1 || !e->d_chain) return set_error(MDA_ERR_BAD_STATE, "mdaEnsembleAutocorrTime: no chain kept");
    if (N > 0x7fffffffLL) return set_error(MDA_ERR_BAD_ARGUMENT, "mdaEnsembleAutocorrTime: %lld kept samples", N);
    const double nan = std::numeric_limits<double>::quiet_NaN();
    auto fail_all = [&]() {
        for (size_t i = 0; i < cols; ++i) tau[i] = nan;
        if (window) for (size_t b = 0; b < B; ++b) window[b] = 0;
        return MDA_OK;
    };
    const double size = 0.5 * (double)N;
    if (c * (double)low >= size) return fail_all();                   // emcee: AutocorrError("The chain is too short")
    int n = (int)N;
    if (fast) { n = 1; while (2LL * n <= N) n *= 2; }                 // the largest power of two
    if (high == 0) high = (int)(size / c);
    if (high <= low) return fail_all();                               // no window to try
    const int n_lags = std::max(1, std::min(high - 1, n));            // f[1:M] with M < high never reads further

    auto up = [](size_t x) { return (x + 255) & ~(size_t)255; };
    const size_t b_series = up(8 * cols * (size_t)N), b_acf = up(8 * cols * (size_t)n_lags);
    if ((rc = ensure_scratch(e, 2 * b_series + b_acf))) return rc;
    char *base = static_cast<char *>(e->scratch);
    double *d_series = reinterpret_cast<double *>(base), *d_centred = reinterpret_cast<double *>(base + b_series),
           *d_acf = reinterpret_cast<double *>(base + 2 * b_series);
    std::vector<double> acf(cols * (size_t)n_lags);
    cudaError_t err = launch_ensemble_mean(e->d_chain, N, s->bins, e->walkers, s->n_l, d_series, N, s->stream);
    if (err == cudaSuccess) err = launch_series_acf(d_series, N, n, (int)cols, n_lags, d_centred, d_acf, s->stream);
    if (err == cudaSuccess) err = cudaMemcpyAsync(acf.data(), d_acf, 8 * acf.size(), cudaMemcpyDeviceToHost, s->stream);
    if (err == cudaSuccess) err = cudaStreamSynchronize(s->stream);
    if (err != cudaSuccess) return cuda_fail(err, "autocorrelation sums");
    g_launches.fetch_add(3, std::memory_order_relaxed);

    std::vector<double> run(L);                                       // sum(f[1:M]) per parameter, terms added in order
    for (size_t b = 0; b < B; ++b) {
        const double *a = acf.data() + b * L * (size_t)n_lags;
        std::fill(run.begin(), run.end(), 0.0);
        int next = 1, accepted = 0;
        for (long long M = low; M < high; M += step) {
            for (; next < M && next < n_lags; ++next)
                for (size_t l = 0; l < L; ++l) run[l] += a[l * n_lags + next] / a[l * n_lags];
            bool all_above = true, any_nan = false;
            double tmax = -std::numeric_limits<double>::infinity();
            for (size_t l = 0; l < L; ++l) {
                const double t = 1.0 + 2.0 * run[l];
                tau[b * L + l] = t;
                all_above = all_above && t > 1.0;
                any_nan = any_nan || t != t;
                if (t > tmax) tmax = t;
            }
            if (any_nan) tmax = nan;                                  // numpy's max propagates NaN; every test below is then false
            if (all_above && (double)M > c * tmax) { accepted = (int)M; break; }
            if (c * tmax >= size) break;                              // too long to be estimated from this chain
        }
        if (window) window[b] = accepted;
        if (!accepted) for (size_t l = 0; l < L; ++l) tau[b * L + l] = nan;
    }
    return MDA_OK;
}

const double *mdaEnsembleDeviceChain(void *ens) { MdaEnsemble *e = as_ensemble(ens, "mdaEnsembleDeviceChain"); return e ? e->d_chain : nullptr; }
const double *mdaEnsembleDevicePositions(void *ens) { MdaEnsemble *e = as_ensemble(ens, "mdaEnsembleDevicePositions"); return e ? e->d_pos : nullptr; }

void mdaPhilox4x32(const unsigned int *counter, const unsigned int *key, unsigned int *out) {
    const Philox4 r = philox4x32_10(Philox4{counter[0], counter[1], counter[2], counter[3]}, key[0], key[1]);
    out[0] = r.x; out[1] = r.y; out[2] = r.z; out[3] = r.w;
}

// ---- roofline denominators -------------------------------------------------------------------
int mdaMeasureFp64Peak(int repeats, double *tflops) {
    if (!tflops) return set_error(MDA_ERR_BAD_ARGUMENT, "mdaMeasureFp64Peak: NULL result");
    int rc = ensure_runtime();
    if (rc) return rc;
    MDA_CUDA(measure_fp64_peak(repeats, g_rt.sm_count, tflops), "fp64 peak microbenchmark");
    g_launches.fetch_add((unsigned long long)(repeats > 0 ? repeats : 1) + 1ull, std::memory_order_relaxed);
    return MDA_OK;
}
int mdaMeasureHbmCopy(size_t bytes, int repeats, double *gbs) {
    if (!gbs || bytes == 0) return set_error(MDA_ERR_BAD_ARGUMENT, "mdaMeasureHbmCopy: bad argument");
    int rc = ensure_runtime();
    if (rc) return rc;
    MDA_CUDA(measure_hbm_copy(bytes, repeats, gbs), "HBM copy microbenchmark");
    return MDA_OK;
}

}  // extern "C"
