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


def test_host_expand_writes_the_reference_rows(oracle):
    """The host half of the counts-over-PCIe output mode (gficf_cuda_expand_host, csrc/host_expand.cpp)
    needs no device: given the reference's own intersection counts it must write the reference's
    bytes -- all three columns, f64 and int32 input, partial tiles, u == 0 rows, every k <= 255."""
    import ctypes as C

    import gficf_b200
    from tests.conftest import random_knn

    L = gficf_b200.lib()
    rng = np.random.default_rng(21)
    for n, k, distinct in ((5000, 30, True), (700, 15, False), (333, 100, True), (300, 255, True), (64, 1, True),
                           (4099, 7, True)):
        r = random_knn(rng, n, k, distinct=distinct, with_self=not distinct)
        want = oracle.parallel(r)
        lut = np.array([u / (2.0 * k - u) for u in range(k + 1)])
        u = np.searchsorted(lut, want[:, 2]).astype(np.uint8)  # the reference's counts, from its weights
        assert np.array_equal(lut[u], want[:, 2])
        for as_int in (False, True):
            src = np.asfortranarray(r.astype(np.int32)) if as_int else r
            for threads in (1, 3):
                out = np.full((n * k, 3), -3.0, dtype=np.float64, order="F")
                rc = L.gficf_cuda_expand_host(src.ctypes.data, 4 if as_int else 8, n, k, u.ctypes.data, 0, n,
                                              out.ctypes.data, threads)
                assert rc == 0 and np.array_equal(out, want), (n, k, as_int, threads)
        # a row range in the middle: only those rows are written, counts are slab-relative
        lo, hi = n // 3, n // 3 + max(1, n // 5)
        out = np.full((n * k, 3), -3.0, dtype=np.float64, order="F")
        sub = np.ascontiguousarray(u[lo * k:hi * k])
        assert L.gficf_cuda_expand_host(r.ctypes.data, 8, n, k, sub.ctypes.data, lo, hi, out.ctypes.data, 2) == 0
        assert np.array_equal(out[lo * k:hi * k], want[lo * k:hi * k])
        assert (out[:lo * k] == -3.0).all() and (out[hi * k:] == -3.0).all()
    bad = np.zeros(4)
    assert L.gficf_cuda_expand_host(bad.ctypes.data, 2, 2, 2, bad.ctypes.data, 0, 2, bad.ctypes.data, 1) == 1


def test_product_build_never_enables_the_emulation():
    """GFICF_CUDA_EMU (plain-C++ bodies of the kernels for tests/cuda_emu) is a test-only switch: no build
    recipe of the product defines it, no product source includes the emulation header outside that switch,
    and the shipped library carries nothing of it."""
    import subprocess

    recipes = [os.path.join(ROOT, "gficf_b200", "csrc", "Makefile"), os.path.join(ROOT, "gficf_b200", "rpkg", "src", "Makevars.cuda"),
               os.path.join(ROOT, "tools", "stage_rpkg.sh"), os.path.join(ROOT, "__graft_entry__.py"),
               os.path.join(ROOT, "gficf_b200", "_lib.py")]
    for p in recipes:
        assert "GFICF_CUDA_EMU" not in open(p).read(), p
        assert "cuda_emu" not in open(p).read(), p
    csrc = os.path.join(ROOT, "gficf_b200", "csrc")
    for name in os.listdir(csrc):
        if name.endswith((".cu", ".cuh", ".h", ".cpp")):
            lines = open(os.path.join(csrc, name)).read().splitlines()
            for i, line in enumerate(lines):
                if '#include "cuda_emu.h"' in line:
                    assert "#ifdef GFICF_CUDA_EMU" in lines[i - 1], (name, i)
    syms = subprocess.run(["nm", "-D", "--defined-only", gficf_b200.library_path()], capture_output=True, text=True).stdout
    assert "cuda_emu" not in syms and "emu_" not in syms
