"""Host-side mirror of the reference's Python<->C boundary, over the CUDA library.

The first four functions keep the reference's names, argument meaning and call
sequence so that code written against ``mda.py`` reads the same here:

    prepare_shared_object   mda.py:1142-1190  (dlopen + ctypes signatures)
    make_calc_struct        mda.py:1742-1780  (one struct per Ex bin)
    ln_post_prob            mda.py:1540-1578  (flat box prior, then lnL)
    call_chi_sq             mda.py:1712-1739  (BFGS objective)

``DeviceStore`` is the batched surface the reference lacks: all Ex bins resident
in HBM, one launch per ensemble half-step (include/mda_chisq.h sections 4-5).

There is no CPU arithmetic in this module and no fallback: if libChiSq.so has
not been built, or no B200 is visible, the calls raise.
"""
import ctypes as ct
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
DEFAULT_LIB = os.path.join(HERE, "libChiSq.so")

_DP = ct.POINTER(ct.c_double)
_IP = ct.POINTER(ct.c_int)
_VP = ct.c_void_p


class MdaError(RuntimeError):
    """A C-ABI call reported a failure through the library's error side channel."""

    def __init__(self, code, message):
        super().__init__("libChiSq error %d: %s" % (code, message))
        self.code = code


def prepare_shared_object(path=None):
    """Load the shared library and declare argument / return types.

    Same contract as mda.py:1142-1190 (``path`` plays CONFIG["Shared Lib Path"]);
    the additive symbols of include/mda_chisq.h are declared as well.
    """
    if path is None:
        path = os.environ.get("MDA_LIB_PATH", DEFAULT_LIB)      # plays CONFIG["Shared Lib Path"]
    if not os.path.exists(path):
        raise FileNotFoundError(
            "%s not built: run `python -m mimic_mda_b200.build` (needs nvcc); "
            "there is no CPU fallback" % path)
    cs_lib = ct.cdll.LoadLibrary(path)
    # --- the reference's seven symbols, exactly as mda.py:1161-1189 declares them
    cs_lib.makeMdaStruct.restype = ct.c_void_p
    cs_lib.makeMdaStruct.argtypes = [ct.c_int, ct.c_int]
    cs_lib.freeMdaStruct.restype = None
    cs_lib.freeMdaStruct.argtypes = [ct.c_void_p]
    cs_lib.setMdaData.restype = None
    cs_lib.setMdaData.argtypes = [ct.c_void_p, _DP]
    cs_lib.setMdaDist.restype = None
    cs_lib.setMdaDist.argtypes = [ct.c_void_p, ct.c_int, _DP]
    cs_lib.calculateChi.restype = ct.c_double
    cs_lib.calculateChi.argtypes = [ct.c_void_p, _DP]
    cs_lib.calculateLnLiklihood.restype = ct.c_double
    cs_lib.calculateLnLiklihood.argtypes = [ct.c_void_p, _DP]
    cs_lib.calculateLnLiklihoodResids.restype = ct.c_double
    cs_lib.calculateLnLiklihoodResids.argtypes = [ct.c_void_p, _DP, _DP]
    # --- additive surface
    sig = {
        "mdaLastError": (ct.c_int, []),
        "mdaLastErrorString": (ct.c_char_p, []),
        "mdaClearError": (None, []),
        "mdaVersion": (ct.c_char_p, []),
        "mdaDeviceCount": (ct.c_int, []),
        "mdaSetDevice": (ct.c_int, [ct.c_int]),
        "mdaGetDevice": (ct.c_int, []),
        "mdaDeviceInfo": (ct.c_int, [ct.c_char_p, ct.c_int, _IP, _IP]),
        "mdaHostAlloc": (_VP, [ct.c_size_t]),
        "mdaHostFree": (None, [_VP]),
        "mdaDeviceAlloc": (_VP, [ct.c_size_t]),
        "mdaDeviceFree": (None, [_VP]),
        "mdaMemcpyH2D": (ct.c_int, [_VP, _VP, ct.c_size_t]),
        "mdaMemcpyD2H": (ct.c_int, [_VP, _VP, ct.c_size_t]),
        "mdaDeviceSynchronize": (ct.c_int, []),
        "mdaStoreCreate": (_VP, [ct.c_int, ct.c_int, ct.c_int]),
        "mdaStoreFree": (None, [_VP]),
        "mdaStoreSetBin": (ct.c_int, [_VP, ct.c_int, ct.c_int, _DP, _DP]),
        "mdaStoreSetBinFromStruct": (ct.c_int, [_VP, ct.c_int, _VP]),
        "mdaStoreSetBounds": (ct.c_int, [_VP, _DP, _DP]),
        "mdaStoreUpload": (ct.c_int, [_VP]),
        "mdaBatchFitStats": (ct.c_int, [ct.POINTER(ct.c_longlong), ct.POINTER(ct.c_longlong), ct.POINTER(ct.c_int)]),
        "mdaDeviceMemInfo": (ct.c_int, [ct.POINTER(ct.c_size_t), ct.POINTER(ct.c_size_t)]),
        "mdaStoreConditioning": (ct.c_int, [_VP, _DP, _DP, ct.POINTER(ct.c_int)]),
        "mdaStoreNumBins": (ct.c_int, [_VP]),
        "mdaStoreNumL": (ct.c_int, [_VP]),
        "mdaStoreMaxPts": (ct.c_int, [_VP]),
        "mdaStoreSetTile": (ct.c_int, [_VP, ct.c_int, ct.c_int]),
        "mdaStoreGetTile": (ct.c_int, [_VP, ct.c_int, _IP, _IP, _IP]),
        "mdaBatchLnPost": (ct.c_int, [_VP, ct.c_int, _VP, _VP]),
        "mdaBatchChi": (ct.c_int, [_VP, ct.c_int, _VP, _VP]),
        "mdaStoreHostParams": (_VP, [_VP, ct.c_int]),
        "mdaStoreHostOut": (_VP, [_VP, ct.c_int]),
        "mdaBatchLnPostStaged": (ct.c_int, [_VP, ct.c_int]),
        "mdaBatchChiStaged": (ct.c_int, [_VP, ct.c_int]),
        "mdaBatchLnPostDevice": (ct.c_int, [_VP, ct.c_int, _VP, _VP]),
        "mdaBatchChiDevice": (ct.c_int, [_VP, ct.c_int, _VP, _VP]),
        "mdaBatchLnPostDeviceMany": (ct.c_int, [_VP, ct.c_int, ct.POINTER(_VP), ct.POINTER(_VP), ct.c_int, ct.c_int]),
        "mdaStoreSynchronize": (ct.c_int, [_VP]),
        "mdaEventCreate": (_VP, []),
        "mdaEventDestroy": (None, [_VP]),
        "mdaEventRecord": (ct.c_int, [_VP, _VP]),
        "mdaEventSynchronize": (ct.c_int, [_VP]),
        "mdaEventElapsedMs": (ct.c_int, [_VP, _VP, ct.POINTER(ct.c_float)]),
        "mdaKernelLaunchCount": (ct.c_ulonglong, []),
        "mdaBatchFit": (ct.c_int, [_VP, ct.c_int, ct.c_int, _VP, _VP, _VP]),
        "mdaPhilox4x32": (None, [ct.POINTER(ct.c_uint), ct.POINTER(ct.c_uint), ct.POINTER(ct.c_uint)]),
        "mdaMeasureFp64Peak": (ct.c_int, [ct.c_int, _DP]),
        "mdaMeasureHbmCopy": (ct.c_int, [ct.c_size_t, ct.c_int, _DP]),
    }
    for name, (restype, argtypes) in sig.items():
        fn = getattr(cs_lib, name)
        fn.restype = restype
        fn.argtypes = argtypes
    return cs_lib


def exported_symbols():
    """Every function include/mda_chisq.h declares (used by the ABI tests)."""
    import re
    header = os.path.join(HERE, "..", "include", "mda_chisq.h")
    with open(header) as fh:
        text = fh.read()
    return re.findall(r"^MDA_API[^;(]*?\b(\w+)\s*\(", text, flags=re.M)


def check(cs_lib, status, what=""):
    """Raise MdaError if a status-returning C-ABI call failed."""
    if status != 0:
        msg = cs_lib.mdaLastErrorString()
        raise MdaError(status, (what + ": " if what else "") + (msg.decode() if msg else "unknown"))


def _nonnull(cs_lib, ptr, what):
    if not ptr:
        msg = cs_lib.mdaLastErrorString()
        raise MdaError(cs_lib.mdaLastError() or -1, what + ": " + (msg.decode() if msg else "NULL returned"))
    return ptr


def make_calc_struct(cs_lib, data, dists):
    """Create one struct and fill it with a bin's data and distributions
    (mda.py:1742-1780).  Raises instead of returning NULL when CUDA is unusable."""
    data = np.ascontiguousarray(data, dtype=np.float64)
    out_struct = _nonnull(cs_lib, cs_lib.makeMdaStruct(len(data), len(dists)), "makeMdaStruct")
    cs_lib.setMdaData(out_struct, data.ctypes.data_as(_DP))
    for i, dis in enumerate(dists):
        dis = np.ascontiguousarray(dis, dtype=np.float64)
        cs_lib.setMdaDist(out_struct, i, dis.ctypes.data_as(_DP))
    return out_struct


def ln_post_prob(params, cs_lib, struct, bounds):
    """Log of the posterior for one walker (mda.py:1540-1578): flat prior inside
    the inclusive bounds, -inf outside, otherwise calculateLnLiklihood."""
    for i, bnd in enumerate(bounds):
        if params[i] < bnd[0] or params[i] > bnd[1]:
            return -np.inf
    return cs_lib.calculateLnLiklihood(struct, params.ctypes.data_as(_DP))


def call_chi_sq(params, cs_lib, struct):
    """chi^2 of one parameter set (mda.py:1712-1739)."""
    return cs_lib.calculateChi(struct, params.ctypes.data_as(_DP))


class DeviceStore:
    """All Ex bins of a spectrum resident in HBM; batched log-posteriors.

    ``consts`` is ``[B, 1+L, n_pad]`` (row 0 data/error, rows 1..L
    distributions/error, zero padded past ``n_pts[b]``) -- the per-bin arrays
    ``fit_data[i]`` / ``interp_dists[i]`` of mda.py:173-176 stacked.  ``bounds``
    is the reference's ``[(0, 1/EWSR_l)]`` list (mda.py:1124) or ``(lo, hi)``.
    """

    def __init__(self, consts, n_pts=None, bounds=None, cs_lib=None):
        self.lib = prepare_shared_object() if cs_lib is None else cs_lib
        consts = np.ascontiguousarray(consts, dtype=np.float64)
        if consts.ndim != 3 or consts.shape[1] < 2:
            raise ValueError("consts must be [bins, 1+L, n_pad]")
        self.n_bins, rows, n_pad = consts.shape
        self.n_l = rows - 1
        if n_pts is None:
            n_pts = np.full(self.n_bins, n_pad, dtype=np.int32)
        self.n_pts = np.ascontiguousarray(n_pts, dtype=np.int32)
        self.max_pts = int(self.n_pts.max()) if self.n_bins else 0
        self.handle = _nonnull(self.lib, self.lib.mdaStoreCreate(self.n_bins, self.n_l, self.max_pts), "mdaStoreCreate")
        for b in range(self.n_bins):
            n = int(self.n_pts[b])
            data = np.ascontiguousarray(consts[b, 0, :n])
            dists = np.ascontiguousarray(consts[b, 1:, :n])
            check(self.lib, self.lib.mdaStoreSetBin(self.handle, b, n, data.ctypes.data_as(_DP),
                                                    dists.ctypes.data_as(_DP)), "mdaStoreSetBin")
        if bounds is not None:
            self.set_bounds(bounds)
        check(self.lib, self.lib.mdaStoreUpload(self.handle), "mdaStoreUpload")

    # -- construction helpers ------------------------------------------------------
    @classmethod
    def from_structs(cls, cs_lib, structs, n_l, max_pts, bounds=None):
        """Gather handles made with ``make_calc_struct`` (one per bin) into a store."""
        self = cls.__new__(cls)
        self.lib = cs_lib
        self.n_bins, self.n_l, self.max_pts = len(structs), n_l, max_pts
        self.n_pts = None
        self.handle = _nonnull(cs_lib, cs_lib.mdaStoreCreate(self.n_bins, n_l, max_pts), "mdaStoreCreate")
        for b, st in enumerate(structs):
            check(cs_lib, cs_lib.mdaStoreSetBinFromStruct(self.handle, b, st), "mdaStoreSetBinFromStruct")
        if bounds is not None:
            self.set_bounds(bounds)
        check(cs_lib, cs_lib.mdaStoreUpload(self.handle), "mdaStoreUpload")
        return self

    def set_bounds(self, bounds):
        if isinstance(bounds, tuple) and len(bounds) == 2 and np.ndim(bounds[0]) == 1:
            lo, hi = bounds
        else:
            lo = [b[0] for b in bounds]
            hi = [b[1] for b in bounds]
        lo = np.ascontiguousarray(lo, dtype=np.float64)
        hi = np.ascontiguousarray(hi, dtype=np.float64)
        if lo.shape != (self.n_l,) or hi.shape != (self.n_l,):
            raise ValueError("bounds must have one (lo, hi) pair per distribution")
        check(self.lib, self.lib.mdaStoreSetBounds(self.handle, lo.ctypes.data_as(_DP), hi.ctypes.data_as(_DP)),
              "mdaStoreSetBounds")

    def set_tile(self, threads=0, walkers_per_thread=0):
        check(self.lib, self.lib.mdaStoreSetTile(self.handle, threads, walkers_per_thread), "mdaStoreSetTile")

    def get_tile(self, walkers):
        t, w, g = ct.c_int(), ct.c_int(), ct.c_int()
        check(self.lib, self.lib.mdaStoreGetTile(self.handle, walkers, ct.byref(t), ct.byref(w), ct.byref(g)),
              "mdaStoreGetTile")
        return t.value, w.value, g.value

    # -- host-pointer evaluation -----------------------------------------------------
    def _params_block(self, params):
        params = np.ascontiguousarray(params, dtype=np.float64)
        if params.ndim == 2 and self.n_bins == 1:
            params = params[None]
        if params.ndim != 3 or params.shape[0] != self.n_bins or params.shape[2] != self.n_l:
            raise ValueError("params must be [bins=%d, walkers, L=%d], got %r" % (self.n_bins, self.n_l, params.shape))
        return params

    def lnpost(self, params):
        """[B, W, L] -> [B, W] log-posteriors (one launch; copies in and out)."""
        params = self._params_block(params)
        out = np.empty(params.shape[:2], dtype=np.float64)
        check(self.lib, self.lib.mdaBatchLnPost(self.handle, params.shape[1], params.ctypes.data, out.ctypes.data),
              "mdaBatchLnPost")
        return out

    def chi(self, params):
        """[B, W, L] -> [B, W] chi^2 values, no prior (batched calculateChi)."""
        params = self._params_block(params)
        out = np.empty(params.shape[:2], dtype=np.float64)
        check(self.lib, self.lib.mdaBatchChi(self.handle, params.shape[1], params.ctypes.data, out.ctypes.data),
              "mdaBatchChi")
        return out

    def fit(self, starts):
        """Batched initial fits (the multi-start L-BFGS-B loop of mda.py:1664-1709 in one call):
        ``starts`` is ``[S, L]`` (one grid shared by every bin, mda.py:179-183) or ``[B, S, L]``.
        Returns ``(params [B, S, L], chi2 [B, S])`` -- per bin and start the box-constrained
        minimiser of chi^2 and its value."""
        starts = np.ascontiguousarray(starts, dtype=np.float64)
        shared = starts.ndim == 2
        if starts.shape[-1] != self.n_l or (not shared and starts.shape[0] != self.n_bins):
            raise ValueError("starts must be [S, L] or [bins, S, L]")
        n_s = starts.shape[-2]
        params = np.empty((self.n_bins, n_s, self.n_l))
        chi = np.empty((self.n_bins, n_s))
        check(self.lib, self.lib.mdaBatchFit(self.handle, n_s, 1 if shared else 0, starts.ctypes.data, params.ctypes.data,
                                             chi.ctypes.data), "mdaBatchFit")
        return params, chi

    def fit_stats(self):
        """What this thread's last ``fit`` met: ``(unconverged, null_column, max_iterations)`` -- fits that hit the
        iteration limit, fits that met a numerically null column (rank-deficient bins), the largest iteration count."""
        unconv, null_col, max_it = ct.c_longlong(), ct.c_longlong(), ct.c_int()
        check(self.lib, self.lib.mdaBatchFitStats(ct.byref(unconv), ct.byref(null_col), ct.byref(max_it)), "mdaBatchFitStats")
        return unconv.value, null_col.value, max_it.value

    # -- pinned staging + CUDA graph ---------------------------------------------------
    def staging(self, walkers):
        """(params_view [B, W, L], out_view [B, W]) over the store's pinned buffers."""
        pp = _nonnull(self.lib, self.lib.mdaStoreHostParams(self.handle, walkers), "mdaStoreHostParams")
        op = _nonnull(self.lib, self.lib.mdaStoreHostOut(self.handle, walkers), "mdaStoreHostOut")
        n_eval = self.n_bins * walkers
        pbuf = (ct.c_double * (n_eval * self.n_l)).from_address(pp)
        obuf = (ct.c_double * n_eval).from_address(op)
        params = np.frombuffer(pbuf, dtype=np.float64).reshape(self.n_bins, walkers, self.n_l)
        out = np.frombuffer(obuf, dtype=np.float64).reshape(self.n_bins, walkers)
        return params, out

    def lnpost_staged(self, walkers):
        """Replay the captured {H2D, kernel, D2H} graph for ``walkers`` per bin."""
        check(self.lib, self.lib.mdaBatchLnPostStaged(self.handle, walkers), "mdaBatchLnPostStaged")

    def chi_staged(self, walkers):
        check(self.lib, self.lib.mdaBatchChiStaged(self.handle, walkers), "mdaBatchChiStaged")

    # -- device-pointer evaluation (asynchronous) -----------------------------------------
    def lnpost_device(self, walkers, d_params, d_out):
        check(self.lib, self.lib.mdaBatchLnPostDevice(self.handle, walkers, d_params, d_out), "mdaBatchLnPostDevice")

    def lnpost_device_many(self, walkers, d_params_list, d_out_list, count):
        n = len(d_params_list)
        pa = (_VP * n)(*d_params_list)
        oa = (_VP * n)(*d_out_list)
        check(self.lib, self.lib.mdaBatchLnPostDeviceMany(self.handle, walkers, pa, oa, n, count),
              "mdaBatchLnPostDeviceMany")

    def synchronize(self):
        check(self.lib, self.lib.mdaStoreSynchronize(self.handle), "mdaStoreSynchronize")

    def conditioning(self):
        """``(kappa [B], limit, quad_safe)``: per bin |data| / sqrt(min chi^2), the limit the triangular-factor
        likelihood of the sampler tolerates, and whether every bin is within it (else the ensemble is stepped in
        the reference's operation order)."""
        kappa = np.empty(self.n_bins, dtype=np.float64)
        limit = ct.c_double()
        safe = ct.c_int()
        check(self.lib, self.lib.mdaStoreConditioning(self.handle, kappa.ctypes.data_as(_DP), ct.byref(limit), ct.byref(safe)),
              "mdaStoreConditioning")
        return kappa, limit.value, bool(safe.value)

    def close(self):
        if getattr(self, "handle", None):
            self.lib.mdaStoreFree(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
