"""The drop-in boundary: the C-ABI library loads through ctypes exactly as mda.py loads
libChiSq.so and exports every symbol include/mda_chisq.h declares.  No compute calls
are made here (this file runs without a GPU)."""
import ctypes as ct
import os
import re
import subprocess

import numpy as np
import pytest

from mimic_mda_b200 import chisq

REFERENCE_SYMBOLS = ["makeMdaStruct", "freeMdaStruct", "setMdaData", "setMdaDist", "calculateChi",
                     "calculateLnLiklihood", "calculateLnLiklihoodResids"]   # C/ChiSquare.h:31-49


def test_header_declares_reference_symbols():
    declared = chisq.exported_symbols()
    for name in REFERENCE_SYMBOLS:
        assert name in declared
    assert len(declared) == len(set(declared)) and len(declared) > 40


def test_library_exports_every_declared_symbol(cuda_lib_path):
    lib = ct.cdll.LoadLibrary(cuda_lib_path)          # mda.py:1161
    for name in chisq.exported_symbols():
        assert hasattr(lib, name), "libChiSq.so lacks %s" % name


def test_library_exports_nothing_undeclared(cuda_lib_path):
    out = subprocess.run(["nm", "-D", "--defined-only", cuda_lib_path], capture_output=True, text=True).stdout
    exported = {line.split()[-1] for line in out.splitlines() if " T " in line}
    extra = exported - set(chisq.exported_symbols()) - {"_init", "_fini"}
    assert not extra, "undeclared exports: %r" % sorted(extra)


def test_signatures_match_reference_ctypes_block(cs_lib):
    # the block mda.py:1165-1189 assigns; our loader must assign the same types
    assert cs_lib.makeMdaStruct.restype is ct.c_void_p
    assert cs_lib.makeMdaStruct.argtypes == [ct.c_int, ct.c_int]
    assert cs_lib.calculateLnLiklihood.restype is ct.c_double
    assert cs_lib.calculateLnLiklihoodResids.argtypes[1:] == [ct.POINTER(ct.c_double)] * 2


def test_no_cuda_at_load_and_version_string(cs_lib):
    assert b"sm_100a" in cs_lib.mdaVersion()
    assert cs_lib.mdaLastError() in (0, 1, 2, 3, 4, 5)


def test_fails_loudly_without_a_device(cs_lib):
    """No CPU fallback: with no GPU every entry point reports an error instead of computing."""
    if cs_lib.mdaDeviceCount() > 0:
        pytest.skip("a CUDA device is visible")
    cs_lib.mdaClearError()
    assert not cs_lib.makeMdaStruct(10, 3)
    assert cs_lib.mdaLastError() == 2                    # MDA_ERR_NO_DEVICE
    assert b"no CPU fallback" in cs_lib.mdaLastErrorString()
    cs_lib.mdaClearError()
    assert not cs_lib.mdaStoreCreate(2, 3, 10)
    assert cs_lib.mdaLastError() == 2
    with pytest.raises(chisq.MdaError):
        chisq.DeviceStore(np.zeros((2, 4, 10)), cs_lib=cs_lib)
    with pytest.raises(chisq.MdaError):
        chisq.make_calc_struct(cs_lib, np.zeros(4), [np.zeros(4)])
    bogus = ct.c_double(0.0)
    val = cs_lib.calculateLnLiklihood(None, ct.byref(bogus))
    assert np.isnan(val) and cs_lib.mdaLastError() == 3   # MDA_ERR_BAD_ARGUMENT


def test_missing_library_is_an_error(tmp_path):
    with pytest.raises(FileNotFoundError):
        chisq.prepare_shared_object(str(tmp_path / "libChiSq.so"))


def test_product_package_never_touches_the_oracle():
    """The oracle is test infrastructure: nothing under mimic_mda_b200/ may import,
    link or execute it (and bench.py only in its cpu_baseline / reference legs)."""
    root = os.path.join(os.path.dirname(os.path.abspath(chisq.__file__)))
    for dirpath, _, files in os.walk(root):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, fn)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M), fn
                assert "liboracle" not in text and "oracle/_ref" not in text and "chisq_oracle" not in text, fn
