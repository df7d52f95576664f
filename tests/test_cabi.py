"""CPU tests of the boundary: the C-ABI library loads and exports every symbol declared in
include/pwt_b200.h, host-only entry points work, and the product fails loudly without a GPU
(no CPU fallback)."""
import ctypes
import os
import re

import numpy as np
import pytest

from conftest import ROOT
from oracle import pdwt_oracle as O

LIB = os.path.join(ROOT, "pypwt_b200", "libpwt_b200.so")
HDR = os.path.join(ROOT, "include", "pwt_b200.h")


def _declared():
    text = open(HDR).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(pwt_[a-z0-9_]+)\s*\(", text)))


@pytest.fixture(scope="module")
def lib():
    if not os.path.exists(LIB):
        pytest.fail("libpwt_b200.so is not built: run `python pypwt_b200/_build.py`")
    return ctypes.CDLL(LIB)


def test_exports_every_declared_symbol(lib):
    names = _declared()
    assert len(names) >= 40
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing


def test_version_and_filters_host_side(lib):
    lib.pwt_version.restype = ctypes.c_char_p
    assert lib.pwt_version() == b"1.0.3"
    buf = [(ctypes.c_float * 40)() for _ in range(4)]
    for w in O.WAVELET_NAMES + ["db1", "bior1.1", "rbior1.1", "HAAR", "Sym8"]:
        n = lib.pwt_lookup_filters(w.encode(), *buf)
        ref = O.filters(w, np.float32)
        assert n == ref[0].size
        for b, r in zip(buf, ref):
            assert np.array_equal(np.array(b[:n], np.float32), r), w
    assert lib.pwt_lookup_filters(b"nosuch", *buf) == -2


def test_python_module_surface():
    import pycudwt
    import pypwt
    assert pypwt.Wavelets is pycudwt.Wavelets
    assert pycudwt.Wavelets.version() == "1.0.3"
    for name in ("forward", "inverse", "coeffs", "coeff_only", "image", "set_image", "soft_threshold",
                 "hard_threshold", "shrink", "norm1", "norm2sq", "add_wavelet", "set_coeff",
                 "set_wavelets_filters", "image_int_ptr", "coeff_int_ptr", "info", "version",
                 "Nr", "Nc", "sizes", "wname", "levels", "do_cycle_spinning", "do_swt", "do_separable",
                 "ndim", "batched1d"):
        assert hasattr(pycudwt.Wavelets, name), name
    assert pycudwt.Wavelets.div2(7) == 4


def test_fails_loudly_without_gpu(lib):
    """No CPU fallback: on a box without a CUDA device construction raises instead of computing."""
    import pycudwt
    if pycudwt.device_count() > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(RuntimeError, match="no CUDA device"):
        pycudwt.Wavelets(np.zeros((8, 8), np.float32), "haar", 1)
    h = ctypes.c_void_p()
    rc = lib.pwt_create(ctypes.byref(h), None, 8, 8, b"haar", 1, 1, 1, 0, 0, 2)
    assert rc == -3 and not h.value
    lib.pwt_last_error.restype = ctypes.c_char_p
    assert b"no CPU fallback" in lib.pwt_last_error()


def test_product_never_imports_the_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "pypwt_b200")):
        for f in files:
            if f.endswith((".py", ".pyx", ".cu", ".cpp", ".h")):
                src = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "import oracle" not in src and "from oracle" not in src and "oracle/" not in src.replace("oracle/pdwt_oracle.py", ""), f
