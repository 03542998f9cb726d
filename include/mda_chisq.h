/*
 * mda_chisq.h -- C ABI of the B200-native chi-square log-likelihood library.
 *
 * Drop-in boundary: the first seven functions are exactly the exports of the
 * reference's libChiSq.so (reference C/ChiSquare.h:31-49; ctypes mirror in
 * mda.py:1161-1189 and C/so_py_tester.py:14-24), same names ("Liklihood"
 * spelling included), same argument meaning, same ownership rules.  mda.py
 * loads this library through CONFIG["Shared Lib Path"] (mda_config0.py:16)
 * without modification.  Everything after them is additive: an error side
 * channel (the reference defines none) and the batched, device-resident
 * surface the reference does not have.
 *
 * All arithmetic is IEEE fp64 on the GPU (sm_100a).  There is no CPU
 * fallback: without a usable CUDA device every compute entry point fails
 * loudly (NaN / non-zero status + mdaLastError()).
 *
 * Threading: as in the reference a handle is not re-entrant (reference
 * C/ChiSquare.c:121,183 share one scratch array per struct); different
 * handles may be used from different threads.  The error channel is
 * thread-local.
 */
#ifndef MDA_CHISQ_H
#define MDA_CHISQ_H

#include <stddef.h>

/* Exported with default visibility; everything else in the library is hidden. */
#if defined(__GNUC__)
#define MDA_API __attribute__((visibility("default")))
#else
#define MDA_API
#endif

#ifdef __cplusplus
extern "C" {
#endif

/* ------------------------------------------------------------------------
 * 1. The reference's seven symbols (one Ex bin per handle, scalar calls)
 * --------------------------------------------------------------------- */

/* replaces makeMdaStruct, C/ChiSquare.h:31, C/ChiSquare.c:18-44.
 * Allocates host mirror + device storage for numPts angles and numL
 * distributions.  CUDA is first touched here (never at dlopen), so forking
 * before the first call is safe (mda.py:186,1109).  NULL on failure. */
MDA_API void *makeMdaStruct(int numPts, int numL);

/* replaces freeMdaStruct, C/ChiSquare.h:34, C/ChiSquare.c:46-65. */
MDA_API void freeMdaStruct(void *strPtr);

/* replaces setMdaData, C/ChiSquare.h:37, C/ChiSquare.c:67-83.
 * Copies numPts doubles (data already divided by its error, mda.py:176). */
MDA_API void setMdaData(void *strPtr, double *divData);

/* replaces setMdaDist, C/ChiSquare.h:40, C/ChiSquare.c:85-112.
 * Copies numPts doubles into row distIndex; distIndex >= numL is silently
 * ignored exactly as in the reference (C/ChiSquare.c:88-89); negative
 * indices (undefined behaviour there) are ignored too and flagged in
 * mdaLastError(). */
MDA_API void setMdaDist(void *strPtr, int distIndex, double *dist);

/* replaces calculateChi, C/ChiSquare.h:43, C/ChiSquare.c:114-172.
 * chi^2 = sum_j (d_j - sum_l a_l D_lj)^2.  params: numL doubles, borrowed. */
MDA_API double calculateChi(void *strPtr, double *params);

/* replaces calculateLnLiklihood, C/ChiSquare.h:46, C/ChiSquare.c:176-213.
 * Returns chi^2 / (-2).  NaN + mdaLastError() on a CUDA failure. */
MDA_API double calculateLnLiklihood(void *strPtr, double *params);

/* replaces calculateLnLiklihoodResids, C/ChiSquare.h:49, C/ChiSquare.c:216-252.
 * As above; residArray (numPts doubles, caller-owned) receives the residuals. */
MDA_API double calculateLnLiklihoodResids(void *strPtr, double *params, double *residArray);

/* ------------------------------------------------------------------------
 * 2. Error side channel (additive; thread-local; sticky until cleared)
 * --------------------------------------------------------------------- */
#define MDA_OK                 0
#define MDA_ERR_CUDA           1   /* a CUDA runtime call failed            */
#define MDA_ERR_NO_DEVICE      2   /* no usable sm_100 device / bad index   */
#define MDA_ERR_BAD_ARGUMENT   3   /* NULL handle, negative size, ...       */
#define MDA_ERR_OUT_OF_MEMORY  4
#define MDA_ERR_BAD_STATE      5   /* e.g. evaluating an un-filled store    */

MDA_API int         mdaLastError(void);
MDA_API const char *mdaLastErrorString(void);
MDA_API void        mdaClearError(void);
MDA_API const char *mdaVersion(void);

/* ------------------------------------------------------------------------
 * 3. Device selection and raw memory helpers
 * --------------------------------------------------------------------- */

/* Number of CUDA devices (0 if none / driver missing). */
MDA_API int mdaDeviceCount(void);
/* Select the device used by handles created afterwards by this process.
 * Default: $MDA_DEVICE, else $LOCAL_RANK, else 0.  One process per GPU. */
MDA_API int mdaSetDevice(int device);
MDA_API int mdaGetDevice(void);
/* Fills name (<= nameLen bytes), SM count and compute capability major*10+minor. */
MDA_API int mdaDeviceInfo(char *name, int nameLen, int *smCount, int *cc);

/* Page-locked host memory: buffers handed to the batched calls DMA at full
 * PCIe rate only if they come from here (or are otherwise pinned). */
MDA_API void *mdaHostAlloc(size_t bytes);
MDA_API void  mdaHostFree(void *p);
MDA_API void *mdaDeviceAlloc(size_t bytes);
MDA_API void  mdaDeviceFree(void *p);
MDA_API int   mdaMemcpyH2D(void *dst, const void *src, size_t bytes);
MDA_API int   mdaMemcpyD2H(void *dst, const void *src, size_t bytes);
MDA_API int   mdaDeviceSynchronize(void);
/* Free and total HBM of the selected device: callers size groups of bins (chains are numWalkers x kept x (numL + 1)
 * doubles per bin) from it. */
MDA_API int   mdaDeviceMemInfo(size_t *freeBytes, size_t *totalBytes);

/* ------------------------------------------------------------------------
 * 4. Device-resident multi-bin store (B Ex bins, uniform numL, ragged numPts)
 *
 * Callers hand bins over in the reference's layout (data[numPts] and
 * dists[l*numPts + j], C/ChiSquare.c:97-102); the store keeps them in HBM
 * angle-pair-major, consts[bin][nPad/2][1 + numL][2] doubles (for each pair
 * of angles: the data pair, then the pair of every distribution), zero
 * padded to nPad = maxPts rounded up to even, so that one bulk copy brings a
 * bin into shared memory in the order the kernel consumes it.  The internal
 * layout is not part of the ABI.  Bounds are per distribution and shared
 * by all bins: the reference's (0, 1/EWSR_l) of mda.py:1124.
 * --------------------------------------------------------------------- */
MDA_API void *mdaStoreCreate(int numBins, int numL, int maxPts);
MDA_API void  mdaStoreFree(void *store);
/* divData[numPts], dists[numL][numPts] dense (reference layout). Copies. */
MDA_API int   mdaStoreSetBin(void *store, int bin, int numPts, const double *divData, const double *dists);
/* Fill bin from a handle made with makeMdaStruct/setMdaData/setMdaDist. */
MDA_API int   mdaStoreSetBinFromStruct(void *store, int bin, void *mdaStruct);
/* lo[numL], hi[numL], inclusive.  Default: lo = 0, hi = 1. */
MDA_API int   mdaStoreSetBounds(void *store, const double *lo, const double *hi);
/* Push the host mirror to HBM (done lazily by the first evaluation too). */
MDA_API int   mdaStoreUpload(void *store);
/* Conditioning of the bins for the triangular-factor likelihood of section 6: kappa[bin] = |data| / sqrt(min chi^2)
 * (numBins values, may be NULL), the limit it is compared with (4 eps kappa <= 1e-13) and whether every bin passes; a
 * store with a bin that does not is sampled by the reference-order stepper instead.  Host arithmetic only. */
MDA_API int   mdaStoreConditioning(void *store, double *kappa, double *kappaMax, int *quadSafe);
MDA_API int   mdaStoreNumBins(void *store);
MDA_API int   mdaStoreNumL(void *store);
MDA_API int   mdaStoreMaxPts(void *store);
/* Force the kernel tiling (threads per CTA in {32,64,128,256}, walkers per
 * thread in {1,2,4}); 0,0 restores the built-in heuristic. */
MDA_API int   mdaStoreSetTile(void *store, int threads, int walkersPerThread);
/* Reports the tiling the next call with walkersPerBin walkers would use.  walkersPerThread = -G: the angle-split
 * kernel (lnpost_split_kernel: 256 threads, G lanes per walker, chi^2 handed from lane to lane in the reference's
 * order; opt-in with MDA_BATCH_IMPL = split -- measured slower than the default).  walkersPerThread = 0: the
 * DMMA kernel (lnpost_mma_kernel: 128 threads, one warp per item of 32 walkers), the default for
 * numL 4 .. 16 with more than 88 angles; a forced tiling selects the register-tiled DFMA kernel,
 * MDA_BATCH_IMPL = split | mma | dfma any one.  The split and register-tiled kernels are bit-identical to each other
 * (the reference's operation order); the DMMA kernel sums a residual's terms in tensor-fragment order (<= 1e-14). */
MDA_API int   mdaStoreGetTile(void *store, int walkersPerBin, int *threads, int *walkersPerThread, int *gridSize);

/* ------------------------------------------------------------------------
 * 5. Batched evaluation: all bins x walkersPerBin walkers in ONE launch
 *
 * params[bin][walker][l] (C-contiguous, the [W, L] blocks emcee hands to a
 * vectorised log-prob function, one per bin), out[bin][walker].
 * LnPost = flat box prior of mda.py:1574-1576 (-inf outside, NaN falls
 * through) then calculateLnLiklihood; Chi = calculateChi, no prior.
 * Up to 88 angles per bin (the register-tiled kernel) the operation order per
 * (bin, walker) is the reference's (C/ChiSquare.c:188-212), bit for bit whatever
 * the tiling; above that the default is the fp64 tensor-path kernel, which sums
 * a residual's terms in fragment order (<= 1e-14 relative of the other; either
 * can be forced, MDA_BATCH_IMPL = dfma | mma).  With either kernel a result
 * does not depend on batch shape, tiling or GPU count.  All return MDA_OK or
 * an error code.
 * --------------------------------------------------------------------- */

/* Host pointers; copies in, launches, copies out, synchronises. */
MDA_API int mdaBatchLnPost(void *store, int walkersPerBin, const double *params, double *lnPost);
MDA_API int mdaBatchChi(void *store, int walkersPerBin, const double *params, double *chi);

/* Pinned staging owned by the store, valid until the next call with a larger
 * walkersPerBin or mdaStoreFree: fill HostParams, call Staged, read HostOut.
 * The per-step work {H2D params, kernel, D2H results} is one CUDA graph
 * captured once per walkersPerBin and replayed on every call. */
MDA_API double *mdaStoreHostParams(void *store, int walkersPerBin);
MDA_API double *mdaStoreHostOut(void *store, int walkersPerBin);
MDA_API int     mdaBatchLnPostStaged(void *store, int walkersPerBin);
MDA_API int     mdaBatchChiStaged(void *store, int walkersPerBin);

/* Device pointers, asynchronous on the store's stream (no copies, no sync). */
MDA_API int mdaBatchLnPostDevice(void *store, int walkersPerBin, const double *dParams, double *dLnPost);
MDA_API int mdaBatchChiDevice(void *store, int walkersPerBin, const double *dParams, double *dChi);
/* Enqueue `count` independent blocks: block i reads dParams[i % nBlocks] and
 * writes dOut[i % nBlocks]. */
MDA_API int mdaBatchLnPostDeviceMany(void *store, int walkersPerBin, const double *const *dParams,
                             double *const *dOut, int nBlocks, int count);
MDA_API int mdaStoreSynchronize(void *store);

/* CUDA events on the store's stream, for device-side timing. */
MDA_API void *mdaEventCreate(void);
MDA_API void  mdaEventDestroy(void *event);
MDA_API int   mdaEventRecord(void *event, void *store);
MDA_API int   mdaEventSynchronize(void *event);
MDA_API int   mdaEventElapsedMs(void *start, void *stop, float *ms);
/* Kernels launched by this process through this library so far. */
MDA_API unsigned long long mdaKernelLaunchCount(void);

/* ------------------------------------------------------------------------
 * 6. Device-resident ensemble sampling (replaces the per-step host loop of
 *    emcee.EnsembleSampler.run_mcmc + ln_post_prob, mda.py:1128-1131,1540-1578)
 *
 * One persistent launch advances every bin's ensemble of nWalkers walkers by
 * nSteps stretch-move steps (two half-ensemble updates per step, emcee 2.x
 * semantics); positions and log-probabilities stay in shared memory.  Default
 * stepper: chi^2(a) = |T [a; -1]|^2 from the triangular factor T of the bin's
 * [distributions | data] matrix (QR in long double when the store is uploaded):
 * (L+1)(L+2)/2 multiply-adds per evaluation whatever the number of angles, same
 * decisions and chains as the angle-by-angle steppers.  Its lnL is within
 * ~eps * kappa of C/ChiSquare.c:176-213, kappa = |data| / sqrt(min chi^2): ~1e-14
 * at 1 % cross-section errors; a store with a bin whose kappa exceeds the limit
 * of mdaStoreConditioning (errors below ~0.5 %) is stepped in the reference's
 * operation order instead, so the 1e-12 bound holds whatever the data.
 * MDA_ENS_IMPL=direct selects the angle-by-angle steppers on the fp64 tensor path
 * (DMMA.8x8x4; two schedules picked by bins per SM; they sum a
 * residual's terms in fragment order and share the sensitivity to kappa), dfma
 * the reference-order one.  Random numbers come from Philox4x32-10 with counter
 * (walker, bin id, 2*step+half, 0) and key = seed + restarts * 0x9E3779B97F4A7C15
 * (restarts = mdaEnsembleSetState calls on this ensemble before the current one),
 * so a chain depends only on (seed, bin id, start positions, restart count) --
 * not on launch shape, stepper or GPU count.  Kept samples are appended to a
 * chain held in HBM; thin keeps the steps whose 1-based index is a multiple of
 * it (emcee 2.x keeps 0-based multiples: shift by one step to compare).
 * --------------------------------------------------------------------- */
MDA_API void *mdaEnsembleCreate(void *store, int nWalkers, unsigned long long seed, double stretchA);
MDA_API void  mdaEnsembleFree(void *ens);
/* Global bin index of every local bin (default 0..bins-1): set it to the
 * bin's index in the whole spectrum when bins are sharded over GPUs. */
MDA_API int   mdaEnsembleSetBinIds(void *ens, const int *binIds);
/* p0[bin][walker][l] (host).  Evaluates the starting log-posteriors, zeroes
 * the acceptance counters and the step counter.  Rejects NaN/inf positions.
 * Must be called again after mdaStoreSetBin / mdaStoreSetBounds (Advance
 * refuses a state whose log-probabilities belong to the old store). */
MDA_API int   mdaEnsembleSetState(void *ens, const double *p0);
/* Room for `capacity` kept samples: chain[kept][bin][walker][l] and
 * lnprob[kept][bin][walker] in HBM (step-major, so a run streams out as
 * contiguous slabs).  Resets the kept count. */
MDA_API int   mdaEnsembleReserveChain(void *ens, long long capacity);
/* Asynchronous: advance nSteps steps, keeping every thin-th (if a chain is reserved). */
MDA_API int   mdaEnsembleAdvance(void *ens, int nSteps, int thin);
/* Waits; MDA_ERR_BAD_STATE if a log-probability was NaN (emcee raises there). */
MDA_API int   mdaEnsembleSynchronize(void *ens);
MDA_API long long mdaEnsembleSteps(void *ens);
MDA_API long long mdaEnsembleKept(void *ens);
/* Host copies; any pointer may be NULL.  nAccepted[bin][walker]. */
MDA_API int   mdaEnsembleGetState(void *ens, double *pos, double *lnProb, long long *nAccepted);
/* chain[count][bin][walker][l], lnProb[count][bin][walker] for kept samples first..first+count-1. */
MDA_API int   mdaEnsembleGetChain(void *ens, long long first, long long count, double *chain, double *lnProb);
/* The emcee call with host buffers: advance nSteps steps, keeping every thin-th, and deliver
 * the kept samples to hostChain[nSteps/thin][bin][walker][l] / hostLnProb[nSteps/thin][bin][walker]
 * (either may be NULL; pinned memory from mdaHostAlloc gives full PCIe rate).  The run is cut
 * into chunks of chunkSteps steps (0 = library default); chunk c+1 is computed while chunk c
 * crosses PCIe from a double-buffered device chain.  Synchronous. */
MDA_API int   mdaEnsembleRunToHost(void *ens, int nSteps, int thin, int chunkSteps, double *hostChain, double *hostLnProb);
/* Posterior summaries computed on the device-resident chain (replaces pulling the chain to
 * the host for calc_param_values / find_most_likely_values, mda.py:1253-1259,1399-1491).
 * Samples = kept samples firstKept.. of every walker, per bin and parameter:
 *   points[bin][l] = (v50, vhi - v50, v50 - vlo) from numpy.percentile at
 *                    100*(0.5 -+ confidenceInterval/2) and 50 (linear interpolation);
 *   peaks[bin][l]  = the same triple around the percentile of the fullest of numBins
 *                    equal-width histogram bins (mda.py:1436-1491), or (v, 0, 0) if the
 *                    samples span less than 1e-7;
 *   acceptFrac[bin] = mean acceptance fraction of the bin's walkers (mda.py:1316).
 * Order statistics are exact; interpolation follows numpy.  Any output may be NULL. */
MDA_API int   mdaEnsembleSummarize(void *ens, long long firstKept, double confidenceInterval, int numBins,
                                   double *points, double *peaks, double *acceptFrac);
/* Integrated autocorrelation time per bin and parameter, computed on the device-resident chain
 * (replaces sampler.get_autocorr_time(c=CONFIG["ACorr WindSize"], fast=CONFIG["ACorr Use FFT"]),
 * mda.py:1304-1308; the times go into the diagnostics file, mda.py:253-272).  emcee 2.x semantics:
 * the series is the mean over the walkers of every kept sample (burn-in included, as sampler.chain
 * is); f = its normalised linear autocorrelation function over all kept samples, or over the first
 * 2^floor(log2 kept) of them if `fast`; for M = low, low + step, ... < high (0: int(kept / 2 / c))
 * tau[l] = 1 + 2 sum(f[1:M]) is accepted once every tau > 1 and M > c max(tau), and abandoned once
 * c max(tau) >= kept / 2.  tau[bin][l]; window[bin] (may be NULL) = the accepted M, or 0 where emcee
 * raises AutocorrError ("the chain is too short"): that bin's tau are NaN.  emcee's defaults are
 * low = 10, high = 0, step = 1, c = 10, fast = 0. */
MDA_API int   mdaEnsembleAutocorrTime(void *ens, int low, int high, int step, double c, int fast,
                                      double *tau, int *window);
/* Device pointers of the chain / state, for device-side consumers. */
MDA_API const double *mdaEnsembleDeviceChain(void *ens);
MDA_API const double *mdaEnsembleDevicePositions(void *ens);
/* Which stepping kernel an Advance of nSteps would launch and how its work is laid out
 * (kernel variant, warps, shared memory, grid, chunking), as text.  For logs and benchmarks. */
MDA_API int mdaEnsembleDescribe(void *ens, int nSteps, char *text, int textLen);
/* Debug: cycles[bin][128] of the last Advance.  DMMA steppers (profiling build: numL = 9, 512 walkers): clock
 * stamps of one half-step, [warp < 16][8].  Warp-specialised kernel: walker warps {start, -, draw made,
 * chi^2 seen, accepted}, math warps {start, proposals made, -, chi^2 posted}.  8-warp kernel: {start, proposals
 * made, chi^2 reduced, rows updated, next operands fetched, -, accept estimate, accept bits}.  DFMA stepper:
 * the first 8 words are section cycle counters of thread 0 {proposal, likelihood, accept, chi^2 loop, draw}.
 * The first call (cycles may be NULL) switches the counters on.  Environment switches for experiments:
 * MDA_ENS_IMPL = ws | mma | dfma (stepper), MDA_ENS_WT (tiles per item), MDA_ENS_DRAWW = 1 | 2 (8-warp kernel
 * with | without draw warps), MDA_ENS_GRID / MDA_ENS_CHUNK (unit schedule). */
MDA_API int mdaEnsembleProfile(void *ens, long long *cycles);
/* The generator itself, evaluated on the host from the same source the kernels compile
 * (csrc/philox.cuh): out[4] = Philox4x32-10(counter[4], key[2]).  For known-answer tests. */
MDA_API void mdaPhilox4x32(const unsigned int *counter, const unsigned int *key, unsigned int *out);

/* ------------------------------------------------------------------------
 * 7. Batched initial fits (replaces the multi-start L-BFGS-B loop on
 *    calculateChi, mda.py:1114,1664-1709,1912-1957)
 *
 * For every bin and every start point: the minimiser of chi^2(a) over the box
 * lo <= a <= hi (the bounds do_init_fit passes, mda.py:1697) and its chi^2.
 * chi^2 is a convex quadratic, so the result is the exact constrained
 * least-squares solution that L-BFGS-B approaches from that start.
 * starts: [startsPerBin][numL] if sharedStarts (the reference's one grid for
 * all bins, mda.py:179-183), else [bin][startsPerBin][numL]; params
 * [bin][startsPerBin][numL]; chi [bin][startsPerBin] (either may be NULL).
 * numL <= 16.  Host pointers; synchronous.
 * --------------------------------------------------------------------- */
MDA_API int mdaBatchFit(void *store, int startsPerBin, int sharedStarts, const double *starts, double *params, double *chi);
/* What the calling thread's last mdaBatchFit met: fits that hit the iteration limit (the returned point is feasible but
 * need not be the minimiser), fits in which a free set had a numerically null column (a bin with fewer independent angles
 * than distributions: every value along it is a minimiser, the start's is kept), and the largest iteration count. */
MDA_API int mdaBatchFitStats(long long *unconverged, long long *nullColumn, int *maxIterations);

/* ------------------------------------------------------------------------
 * 8. Roofline denominators measured on the device this process uses
 * --------------------------------------------------------------------- */
/* Dependent-free DFMA stream on every SM; best of `repeats`. TFLOP/s (2/FMA). */
MDA_API int mdaMeasureFp64Peak(int repeats, double *tflops);
/* Device-to-device copy of `bytes` bytes; GB/s counting read + write. */
MDA_API int mdaMeasureHbmCopy(size_t bytes, int repeats, double *gbs);

#ifdef __cplusplus
}
#endif
#endif /* MDA_CHISQ_H */
