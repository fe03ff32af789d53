"""Host-buffer operators: the same names, argument meaning and error behaviour as the
reference's R-visible functions for the path, on numpy matrices.

Reference interface mirrored here:
  rcpp_parallel_jaccard_coef(mat, printOutput)   R/RcppExports.R:16-18
  jaccard_coeff(idx, printOutput)                R/RcppExports.R:8-10
  neigh[,-1] -> Jaccard -> relations[,3] > 0     R/clustCells.R:63-66
"""
from __future__ import annotations

import ctypes as C
import sys

import numpy as np

from . import _lib

MODE_PARALLEL = 0
MODE_SERIAL = 1

_BANNER_PARALLEL = "Running Parallell Jaccard Coefficient Estimation...\n"  # rcpp_parallel_jaccard_coeff.cpp:63
_BANNER_DONE = "Done!!\n"  # :77
_BANNER_SERIAL = "Running Jaccard Coefficient Estimation...\n"  # jaccard_coeff.cpp:25


def _as_numeric_matrix(mat) -> np.ndarray:
    """What Rcpp's input_parameter<NumericMatrix> does (src/RcppExports.cpp:40,65): any numeric
    matrix becomes column-major float64 (no copy when it already is) -- except an int32 matrix
    (R's INTSXP, what uwot returns), which keeps its type and takes the integer entry point:
    no coercion copy, half the H2D bytes."""
    a = np.asarray(mat)
    if a.ndim != 2:
        raise TypeError("a numeric matrix (2-D) is required")
    if a.dtype.kind not in "iufb":
        raise TypeError("not a numeric matrix")
    if a.dtype == np.int32:
        return np.asfortranarray(a)
    return np.asfortranarray(a, dtype=np.float64)


class _PinnedOwner:
    def __init__(self, ptr):
        self.ptr = ptr

    def __del__(self):
        try:
            _lib.lib().gficf_cuda_host_free(self.ptr)
        except Exception:
            pass


def pinned_empty(shape, dtype=np.float64, order="F") -> np.ndarray:
    """A page-locked numpy array (H2D/D2H at PCIe speed without staging)."""
    dtype = np.dtype(dtype)
    n = int(np.prod(shape)) * dtype.itemsize
    p = C.c_void_p()
    _lib.check(_lib.lib().gficf_cuda_host_alloc(C.byref(p), max(n, 1)))
    owner = _PinnedOwner(p)
    buf = (C.c_char * max(n, 1)).from_address(p.value)
    arr = np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape))).reshape(shape, order=order)
    arr = arr.view(_PinnedArray)
    arr._owner = owner  # keeps the allocation alive as long as any view of it
    return arr


class _PinnedArray(np.ndarray):
    _owner = None

    def __array_finalize__(self, obj):
        if obj is not None:
            self._owner = getattr(obj, "_owner", None)


def _call(a: np.ndarray, mode: int, n_devices: int, out=None):
    n, k = a.shape
    e = n * k
    if out is None:
        out = np.empty((e, 3), dtype=np.float64, order="F")
    else:
        if out.shape != (e, 3) or out.dtype != np.float64 or not out.flags.f_contiguous:
            raise ValueError("out must be a Fortran-ordered float64 (n*k, 3) matrix")
    err = C.create_string_buffer(512)
    nw = C.c_int64(0)
    entry = _lib.lib().gficf_cuda_jaccard_i32 if a.dtype == np.int32 else _lib.lib().gficf_cuda_jaccard
    rc = entry(a.ctypes.data, n, k, out.ctypes.data, int(n_devices), mode, C.byref(nw), err, 512)
    _lib.check(rc, err)
    return out, int(nw.value)


def rcpp_parallel_jaccard_coef(mat, printOutput: bool = False, n_devices: int = 0, out=None) -> np.ndarray:
    """Drop-in for the reference export (src/rcpp_parallel_jaccard_coeff.cpp:58-80).

    mat: n x k matrix of 1-based neighbour ids.  Returns the (n*k) x 3 column-major matrix
    (from, to, weight); row i*k+j stays zero when N(i) and N(mat[i,j]) do not intersect.
    n_devices=0 -> the value of set_devices() / GFICF_CUDA_DEVICES (default 1)."""
    a = _as_numeric_matrix(mat)
    if printOutput:
        sys.stdout.write(_BANNER_PARALLEL)
    res, _ = _call(a, MODE_PARALLEL, n_devices, out)
    if printOutput:
        sys.stdout.write(_BANNER_DONE)
    return res


def jaccard_coeff(idx, printOutput: bool = False, out=None) -> np.ndarray:
    """Drop-in for the reference's serial export (src/jaccard_coeff.cpp:19-45): rows compacted
    (only u>0 rows are emitted, in (i,j) order; the tail stays zero), unique-set intersection."""
    a = _as_numeric_matrix(idx)
    if printOutput:
        sys.stdout.write(_BANNER_SERIAL)
    res, _ = _call(a, MODE_SERIAL, 1, out)
    return res


def phenograph_edges(neigh, verbose: bool = False, n_gpu: int = 0) -> np.ndarray:
    """clustcells()'s graph-weighting step (R/clustCells.R:63-66): drop the self column of the
    kNN result, weight the edges, keep rows with weight > 0.  n_gpu is the new device-count
    option (R/clustCells.R gains `n.gpu`)."""
    neigh = np.asarray(neigh)
    relations = rcpp_parallel_jaccard_coef(neigh[:, 1:], verbose, n_gpu)
    return relations[relations[:, 2] > 0, :]


def set_devices(n: int) -> None:
    _lib.check(_lib.lib().gficf_cuda_set_devices(int(n)))


def last_timings() -> dict:
    buf = (C.c_double * 8)()
    _lib.check(_lib.lib().gficf_cuda_last_timings(buf))
    keys = ["h2d_ms", "layout_ms", "jaccard_ms", "d2h_ms", "wall_ms", "allgather_ms", "launches", "d2h_bytes"]
    return dict(zip(keys, list(buf)))


def last_output() -> dict:
    """How the last host-buffer call produced its output columns (include/gficf_cuda.h,
    gficf_cuda_last_output): mode dma / host / hybrid, share written by host threads, PCIe bytes."""
    mode, share, nbytes, hbytes = C.c_int32(0), C.c_double(0), C.c_double(0), C.c_double(0)
    _lib.check(_lib.lib().gficf_cuda_last_output(C.byref(mode), C.byref(share), C.byref(nbytes), C.byref(hbytes)))
    return {"mode": {1: "dma", 2: "host", 3: "hybrid"}.get(mode.value, str(mode.value)),
            "host_share": share.value, "d2h_bytes": nbytes.value, "h2d_bytes": hbytes.value}
