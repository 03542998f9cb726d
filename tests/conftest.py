"""Shared fixtures.  Tests that need a B200 are marked ``gpu``; everything else runs on CPU."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200, sm_100a)")


class Golden:
    """tests/golden/golden_lnl.npz: inputs + outputs of the unmodified reference library."""

    def __init__(self, path):
        self._z = np.load(path)
        self.cases = sorted({k.split("/")[0] for k in self._z.files} - {"kat"})

    def case(self, name):
        return {k.split("/", 1)[1]: self._z[k] for k in self._z.files if k.startswith(name + "/")}


@pytest.fixture(scope="session")
def golden():
    return Golden(os.path.join(ROOT, "tests", "golden", "golden_lnl.npz"))


@pytest.fixture(scope="session")
def oracle_lib():
    """Our C restatement, built on demand (gcc only)."""
    from oracle import reflib
    if not (os.path.exists(reflib.ORACLE_LIB) and os.path.exists(reflib.REF_LOOP_LIB)):
        reflib.build()
    return reflib.OracleLib()


@pytest.fixture(scope="session")
def cuda_lib_path():
    """libChiSq.so, built on demand with nvcc (cross-compiles without a GPU)."""
    from mimic_mda_b200 import build
    return build.build()


@pytest.fixture(scope="session")
def cs_lib(cuda_lib_path):
    from mimic_mda_b200 import chisq
    return chisq.prepare_shared_object(cuda_lib_path)


@pytest.fixture(autouse=True)
def _clear_sticky_error(request):
    """The library's error channel is sticky until cleared: start every GPU test clean."""
    if "gpu" in request.keywords and "cs_lib" in request.fixturenames:
        request.getfixturevalue("cs_lib").mdaClearError()
    yield


def rel_err(got, want):
    """max |got - want| / |want| over finite entries (0/0 counts as 0)."""
    got = np.asarray(got, dtype=np.float64)
    want = np.asarray(want, dtype=np.float64)
    fin = np.isfinite(want)
    if not fin.any():
        return 0.0
    denom = np.where(np.abs(want[fin]) > 0, np.abs(want[fin]), 1.0)
    return float(np.max(np.abs(got[fin] - want[fin]) / denom))


def same_special(got, want):
    """-inf and NaN positions identical."""
    got = np.asarray(got)
    want = np.asarray(want)
    return bool(np.array_equal(np.isneginf(got), np.isneginf(want)) and np.array_equal(np.isnan(got), np.isnan(want)))
