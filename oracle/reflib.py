"""ctypes access to the CHECKER libraries -- TEST INFRASTRUCTURE ONLY.

* ``build()``            runs oracle/Makefile (C restatement always; the
                         unmodified reference only where /root/reference exists).
* ``reference_lib_path`` picks oracle/_ref/libChiSq_native.so when a subprocess
                         probe shows it runs on this CPU, else the portable
                         oracle/_ref/libChiSq.so, else None.
* ``prepare_shared_object`` / ``make_calc_struct`` / ``ln_post_prob`` /
  ``call_chi_sq``        Python-3 restatements of mda.py:1142-1190, 1742-1780,
                         1540-1578 and 1712-1739: the exact call sequence the
                         reference makes against a 7-symbol library.
* ``RefLoop``            binding of oracle/ref_loop.c (tight C loop, threads
                         over bins) for pure-C baselines and bulk reference
                         outputs.
"""
import ctypes as ct
import os
import subprocess
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
BUILD_DIR = os.path.join(HERE, "_build")
REF_DIR = os.path.join(HERE, "_ref")
ORACLE_LIB = os.path.join(BUILD_DIR, "liboracle.so")
REF_LOOP_LIB = os.path.join(BUILD_DIR, "libref_loop.so")
REF_LIB_PORTABLE = os.path.join(REF_DIR, "libChiSq.so")
REF_LIB_NATIVE = os.path.join(REF_DIR, "libChiSq_native.so")
REF_TESTER = os.path.join(REF_DIR, "ref_tester")

_DP = ct.POINTER(ct.c_double)
_IP = ct.POINTER(ct.c_int)


def build(quiet=True):
    """Compile the checkers.  Building the checker is not using it."""
    res = subprocess.run(["make", "-C", HERE, "all"], capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("oracle build failed:\n" + res.stdout + res.stderr)
    if not quiet:
        print(res.stdout)
    return res.stdout


def _runs_here(path):
    """True if ``path`` loads and reproduces the known answer in a child process
    (a -march=native binary from another CPU would die with SIGILL there, not here)."""
    code = (
        "import ctypes as ct, sys\n"
        "lib = ct.cdll.LoadLibrary(sys.argv[1])\n"
        "lib.makeMdaStruct.restype = ct.c_void_p\n"
        "lib.makeMdaStruct.argtypes = [ct.c_int, ct.c_int]\n"
        "lib.setMdaData.argtypes = [ct.c_void_p, ct.POINTER(ct.c_double)]\n"
        "lib.setMdaDist.argtypes = [ct.c_void_p, ct.c_int, ct.POINTER(ct.c_double)]\n"
        "lib.calculateLnLiklihood.restype = ct.c_double\n"
        "lib.calculateLnLiklihood.argtypes = [ct.c_void_p, ct.POINTER(ct.c_double)]\n"
        "n = 64\n"
        "x = (ct.c_double * n)(*[0.1 * i for i in range(n)])\n"
        "a = (ct.c_double * 3)(0.2, 0.2, 0.2)\n"
        "s = lib.makeMdaStruct(n, 3)\n"
        "lib.setMdaData(s, x)\n"
        "[lib.setMdaDist(s, i, x) for i in range(3)]\n"
        "v = lib.calculateLnLiklihood(s, a)\n"
        "sys.exit(0 if v < 0 else 3)\n"
    )
    try:
        res = subprocess.run([sys.executable, "-c", code, path], capture_output=True, timeout=60)
    except Exception:
        return False
    return res.returncode == 0


_REF_PATH_CACHE = []


def reference_lib_path():
    """Path of a runnable build of the UNMODIFIED reference library, or None."""
    if _REF_PATH_CACHE:
        return _REF_PATH_CACHE[0]
    choice = None
    for cand in (REF_LIB_NATIVE, REF_LIB_PORTABLE):
        if os.path.exists(cand) and _runs_here(cand):
            choice = cand
            break
    _REF_PATH_CACHE.append(choice)
    return choice


def prepare_shared_object(path):
    """Load a 7-symbol library and declare its signatures (mda.py:1142-1190)."""
    cs_lib = ct.cdll.LoadLibrary(path)
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
    return cs_lib


def make_calc_struct(cs_lib, data, dists):
    """Create and fill one struct (mda.py:1742-1780)."""
    data = np.ascontiguousarray(data, dtype=np.float64)
    out_struct = cs_lib.makeMdaStruct(len(data), len(dists))
    cs_lib.setMdaData(out_struct, data.ctypes.data_as(_DP))
    for i, dis in enumerate(dists):
        dis = np.ascontiguousarray(dis, dtype=np.float64)
        cs_lib.setMdaDist(out_struct, i, dis.ctypes.data_as(_DP))
    return out_struct


def ln_post_prob(params, cs_lib, struct, bounds):
    """Flat box prior then calculateLnLiklihood (mda.py:1540-1578)."""
    for i, bnd in enumerate(bounds):
        if params[i] < bnd[0] or params[i] > bnd[1]:
            return -np.inf
    return cs_lib.calculateLnLiklihood(struct, params.ctypes.data_as(_DP))


def call_chi_sq(params, cs_lib, struct):
    """calculateChi through ctypes (mda.py:1712-1739)."""
    return cs_lib.calculateChi(struct, params.ctypes.data_as(_DP))


def lnpost_bins_scalar(cs_lib, consts, n_pts, lo, hi, params):
    """Evaluate [B, W] log-posteriors the way the reference does: one struct per
    bin, one ``ln_post_prob`` call per walker.  consts is [B, 1+L, n_pad]."""
    consts = np.asarray(consts, dtype=np.float64)
    params = np.ascontiguousarray(params, dtype=np.float64)
    n_bins, n_walk, n_l = params.shape
    bounds = [(float(lo[i]), float(hi[i])) for i in range(n_l)]
    out = np.empty((n_bins, n_walk), dtype=np.float64)
    for b in range(n_bins):
        n = int(n_pts[b])
        struct = make_calc_struct(cs_lib, consts[b, 0, :n], [consts[b, 1 + l, :n] for l in range(n_l)])
        for w in range(n_walk):
            out[b, w] = ln_post_prob(params[b, w], cs_lib, struct, bounds)
        cs_lib.freeMdaStruct(struct)
    return out


class OracleLib:
    """Batched entry points of our C restatement (oracle/chisq_oracle.c)."""

    def __init__(self, path=ORACLE_LIB):
        self.path = path
        lib = ct.cdll.LoadLibrary(path)
        lib.orc_lnpost_bins.restype = None
        lib.orc_lnpost_bins.argtypes = [ct.c_int, ct.c_int, ct.c_int, _IP, _DP, _DP, _DP, _DP, ct.c_int, _DP]
        lib.orc_lnlike_batch.restype = None
        lib.orc_lnlike_batch.argtypes = [ct.c_int, ct.c_int, _DP, _DP, _DP, ct.c_int, _DP]
        lib.orc_chi_batch.restype = None
        lib.orc_chi_batch.argtypes = [ct.c_int, ct.c_int, _DP, _DP, _DP, ct.c_int, _DP]
        self.lib = lib

    def lnpost_bins(self, consts, n_pts, lo, hi, params):
        consts = np.ascontiguousarray(consts, dtype=np.float64)
        params = np.ascontiguousarray(params, dtype=np.float64)
        n_pts = np.ascontiguousarray(n_pts, dtype=np.int32)
        lo = np.ascontiguousarray(lo, dtype=np.float64)
        hi = np.ascontiguousarray(hi, dtype=np.float64)
        n_bins, rows, n_pad = consts.shape
        n_l = rows - 1
        assert params.shape[0] == n_bins and params.shape[2] == n_l
        out = np.empty(params.shape[:2], dtype=np.float64)
        self.lib.orc_lnpost_bins(n_bins, n_l, n_pad, n_pts.ctypes.data_as(_IP),
                                 consts.ctypes.data_as(_DP), lo.ctypes.data_as(_DP),
                                 hi.ctypes.data_as(_DP), params.ctypes.data_as(_DP),
                                 params.shape[1], out.ctypes.data_as(_DP))
        return out

    def _one_bin(self, fn, data, dists, params):
        data = np.ascontiguousarray(data, dtype=np.float64)
        dists = np.ascontiguousarray(dists, dtype=np.float64)
        params = np.ascontiguousarray(np.atleast_2d(params), dtype=np.float64)
        out = np.empty(params.shape[0], dtype=np.float64)
        fn(len(data), dists.shape[0], data.ctypes.data_as(_DP), dists.ctypes.data_as(_DP),
           params.ctypes.data_as(_DP), params.shape[0], out.ctypes.data_as(_DP))
        return out

    def lnlike_batch(self, data, dists, params):
        return self._one_bin(self.lib.orc_lnlike_batch, data, dists, params)

    def chi_batch(self, data, dists, params):
        return self._one_bin(self.lib.orc_chi_batch, data, dists, params)


class RefLoop:
    """oracle/ref_loop.c: C loop over walkers through any 7-symbol library."""

    def __init__(self, lib_path, loop_path=REF_LOOP_LIB):
        self.lib_path = lib_path
        drv = ct.cdll.LoadLibrary(loop_path)
        drv.ref_loop_run.restype = ct.c_double
        drv.ref_loop_run.argtypes = [ct.c_char_p, ct.c_int, ct.c_int, ct.c_int, ct.c_int, _IP, _DP,
                                     _DP, _DP, _DP, ct.c_int, _DP, ct.c_int, ct.c_int]
        self.drv = drv

    def run(self, consts, n_pts, lo, hi, params, mode="lnpost", n_threads=1, repeats=1):
        """Returns (out [B, W], seconds for `repeats` passes)."""
        consts = np.ascontiguousarray(consts, dtype=np.float64)
        params = np.ascontiguousarray(params, dtype=np.float64)
        n_pts = np.ascontiguousarray(n_pts, dtype=np.int32)
        lo = np.ascontiguousarray(lo, dtype=np.float64)
        hi = np.ascontiguousarray(hi, dtype=np.float64)
        n_bins, rows, n_pad = consts.shape
        out = np.empty(params.shape[:2], dtype=np.float64)
        secs = self.drv.ref_loop_run(self.lib_path.encode(), 1 if mode == "chi" else 0,
                                     n_bins, rows - 1, n_pad, n_pts.ctypes.data_as(_IP),
                                     consts.ctypes.data_as(_DP), lo.ctypes.data_as(_DP),
                                     hi.ctypes.data_as(_DP), params.ctypes.data_as(_DP),
                                     params.shape[1], out.ctypes.data_as(_DP), n_threads, repeats)
        if secs < 0:
            raise RuntimeError("ref_loop_run failed (%g) for %s" % (secs, self.lib_path))
        return out, secs
