"""CPU: the C-ABI library loads, exports every symbol include/gficf_cuda.h declares, and the
argument checks that need no GPU behave.  No compute calls."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import gficf_b200
from gficf_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_functions():
    text = open(os.path.join(ROOT, "include", "gficf_cuda.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(gficf_cuda_\w+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    names = header_functions()
    assert len(names) >= 15
    L = C.CDLL(gficf_b200.library_path())
    for nm in names:
        assert hasattr(L, nm), nm
    assert sorted(_lib.PROTOTYPES) == names  # the Python binding covers the header, nothing else


def test_version_and_row_stride():
    L = gficf_b200.lib()
    assert b"sm_100a" in L.gficf_cuda_version()
    assert [L.gficf_cuda_row_stride(k) for k in (1, 4, 5, 8, 15, 16, 30, 32, 33, 100, 128)] == \
        [4, 4, 8, 8, 16, 16, 32, 32, 48, 112, 128]
    assert L.gficf_cuda_expand_scratch_bytes(0) >= 8


def test_argument_errors_without_gpu():
    L = gficf_b200.lib()
    err = C.create_string_buffer(256)
    a = np.ones((4, 2), order="F")
    out = np.empty((8, 3), order="F")
    assert L.gficf_cuda_jaccard(a.ctypes.data, -1, 2, out.ctypes.data, 1, 0, None, err, 256) == 1
    assert L.gficf_cuda_jaccard(a.ctypes.data, 4, 2, out.ctypes.data, 1, 7, None, err, 256) == 1
    assert b"mode" in err.value
    assert L.gficf_cuda_jaccard(None, 4, 2, out.ctypes.data, 1, 0, None, err, 256) == 1
    # empty matrix: nothing to do, no device needed
    assert L.gficf_cuda_jaccard(a.ctypes.data, 0, 2, out.ctypes.data, 1, 0, None, err, 256) == 0
    assert L.gficf_cuda_jaccard(a.ctypes.data, 4, 0, out.ctypes.data, 1, 0, None, err, 256) == 0
    # n*k beyond the reference's int range
    assert L.gficf_cuda_jaccard(a.ctypes.data, 1 << 30, 4, out.ctypes.data, 1, 0, None, err, 256) == 5
    assert L.gficf_cuda_set_devices(0) == 1
    assert L.gficf_cuda_set_devices(1) == 0 and L.gficf_cuda_get_devices() == 1


def test_no_cpu_fallback():
    """Without a GPU the product path must fail loudly, never compute on the host."""
    L = gficf_b200.lib()
    if L.gficf_cuda_device_count() > 0:
        pytest.skip("a GPU is visible")
    with pytest.raises(gficf_b200.GficfCudaError) as e:
        gficf_b200.rcpp_parallel_jaccard_coef(np.array([[2.0, 3.0], [1.0, 3.0], [1.0, 2.0]]))
    assert e.value.code == 3
    with pytest.raises(TypeError):
        gficf_b200.rcpp_parallel_jaccard_coef(np.array([["a", "b"]]))


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "gficf_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M), f
                assert "libgficf_oracle" not in text and "libgficf_ref" not in text, f
