"""Builds mimic_mda_b200/libChiSq.so in-tree with nvcc for sm_100a only.

The output is a plain C-ABI shared library (include/mda_chisq.h) with the CUDA
runtime linked statically: ``ctypes.cdll.LoadLibrary`` (mda.py:1161) needs
nothing else on the library path.  It is git-ignored but travels to the GPU box
with the repo snapshot.

    python -m mimic_mda_b200.build [--force] [--verbose]
"""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB_PATH = os.path.join(HERE, "libChiSq.so")
SOURCES = ["chisq_abi.cu", "lnpost_kernel.cu", "lnpost_mma_kernel.cu", "lnpost_quad_kernel.cu", "lnpost_split_kernel.cu", "ensemble_kernel.cu", "ensemble_mma_kernel.cu", "ensemble_ws_kernel.cu", "ensemble_quad_kernel.cu", "ensemble_qws_kernel.cu", "fit_kernel.cu", "summary_kernel.cu", "autocorr_kernel.cu", "scalar_kernel.cu", "peaks.cu"]
HEADERS = [os.path.join(CSRC, "mda_kernels.cuh"), os.path.join(CSRC, "likelihood.cuh"), os.path.join(CSRC, "philox.cuh"), os.path.join(CSRC, "mma_chi.cuh"), os.path.join(CSRC, "quad_chi.cuh"),
           os.path.join(HERE, "..", "include", "mda_chisq.h")]

NVCC_FLAGS = [
    "-O3", "-std=c++17",
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo",
    "-Xcompiler", "-fPIC,-O3,-fvisibility=hidden",
    "-fmad=true",
    "-shared", "-cudart", "static",
] + (os.environ.get("MDA_NVCC_EXTRA", "").split())


def nvcc_path():
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found; cannot build libChiSq.so")


def is_stale():
    if not os.path.exists(LIB_PATH):
        return True
    built = os.path.getmtime(LIB_PATH)
    deps = [os.path.join(CSRC, s) for s in SOURCES] + HEADERS + [os.path.abspath(__file__)]
    return any(os.path.getmtime(d) > built for d in deps)


def build(force=False, verbose=False):
    """Compile every CUDA source (one nvcc per file, in parallel) and link libChiSq.so;
    a no-op when the library is newer than its sources."""
    if not force and not is_stale():
        return LIB_PATH
    from concurrent.futures import ThreadPoolExecutor
    nvcc = nvcc_path()
    obj_dir = os.path.join(HERE, "build")
    os.makedirs(obj_dir, exist_ok=True)
    compile_flags = [f for f in NVCC_FLAGS if f not in ("-shared",)]
    for drop in ("-cudart", "static"):
        compile_flags.remove(drop)

    def compile_one(src):
        obj = os.path.join(obj_dir, os.path.splitext(src)[0] + ".o")
        cmd = [nvcc] + compile_flags + (["-Xptxas", "-v"] if verbose else []) + ["-c", os.path.join(CSRC, src), "-o", obj]
        res = subprocess.run(cmd, capture_output=True, text=True)
        if res.returncode != 0:
            raise RuntimeError("nvcc failed:\n" + " ".join(cmd) + "\n" + res.stdout + res.stderr)
        return obj, res.stdout + res.stderr

    with ThreadPoolExecutor(max_workers=min(len(SOURCES), os.cpu_count() or 1)) as pool:
        results = list(pool.map(compile_one, SOURCES))
    link = [nvcc, "-shared", "-cudart", "static", "-gencode", "arch=compute_100a,code=sm_100a",
            "-Xcompiler", "-fPIC"] + [obj for obj, _ in results] + ["-o", LIB_PATH + ".tmp"]
    res = subprocess.run(link, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("link failed:\n" + " ".join(link) + "\n" + res.stdout + res.stderr)
    os.replace(LIB_PATH + ".tmp", LIB_PATH)
    if verbose:
        print("".join(log for _, log in results))
    return LIB_PATH


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
