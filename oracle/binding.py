"""ctypes bindings for the CPU checkers (TEST INFRASTRUCTURE, NOT PRODUCT CODE).

``Oracle``    -> oracle/libgficf_oracle.so  (oracle/jaccard_oracle.c, our restatement of
                 /root/reference/src/rcpp_parallel_jaccard_coeff.cpp:24-55 and
                 /root/reference/src/jaccard_coeff.cpp:28-42)
``Reference`` -> oracle/_ref/libgficf_ref.so (those reference files themselves, built by
                 oracle/Makefile; absent if the container never built them)

All matrices use the reference's conventions: ``idx`` is n x k float64,
Fortran (column-major) order, 1-based; results are (n*k) x 3 float64 Fortran order.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE_SO = os.path.join(_HERE, "libgficf_oracle.so")
REF_SO = os.path.join(_HERE, "_ref", "libgficf_ref.so")
REF_WMU_SO = os.path.join(_HERE, "_ref", "libgficf_ref_wmu.so")
MODOPT_BIN = os.path.join(_HERE, "_ref", "modopt")

_dp = C.POINTER(C.c_double)


def build(verbose: bool = False) -> None:
    """Compile the checkers (the C restatement always; oracle/_ref when /root/reference exists)."""
    out = subprocess.run(["make", "-C", _HERE, "all"], capture_output=True, text=True)
    if verbose or out.returncode != 0:
        print(out.stdout, out.stderr)
    if out.returncode != 0:
        raise RuntimeError("oracle build failed")


def _as_idx(idx) -> np.ndarray:
    a = np.asarray(idx)
    if a.ndim != 2:
        raise ValueError("idx must be a 2-D matrix")
    return np.asfortranarray(a, dtype=np.float64)


def _ptr(a: np.ndarray):
    return a.ctypes.data_as(_dp)


class Oracle:
    """Our plain-C restatement."""

    def __init__(self):
        if not os.path.exists(ORACLE_SO):
            build()
        self.lib = C.CDLL(ORACLE_SO)
        self.lib.gficf_oracle_parallel_jaccard.argtypes = [_dp, C.c_int64, C.c_int32, _dp, C.c_int32]
        self.lib.gficf_oracle_parallel_jaccard.restype = None
        self.lib.gficf_oracle_parallel_jaccard_rows.argtypes = [
            _dp, C.c_int64, C.c_int32, C.c_int64, C.c_int64, _dp, C.c_int32]
        self.lib.gficf_oracle_parallel_jaccard_rows.restype = None
        self.lib.gficf_oracle_serial_jaccard.argtypes = [_dp, C.c_int64, C.c_int32, _dp]
        self.lib.gficf_oracle_serial_jaccard.restype = C.c_int64

    def parallel(self, idx, nthreads: int | None = None) -> np.ndarray:
        a = _as_idx(idx)
        n, k = a.shape
        out = np.empty((n * k, 3), dtype=np.float64, order="F")
        self.lib.gficf_oracle_parallel_jaccard(_ptr(a), n, k, _ptr(out), nthreads or os.cpu_count() or 1)
        return out

    def parallel_rows(self, idx, lo: int, hi: int, nthreads: int | None = None) -> np.ndarray:
        a = _as_idx(idx)
        n, k = a.shape
        out = np.zeros(((hi - lo) * k, 3), dtype=np.float64, order="F")
        self.lib.gficf_oracle_parallel_jaccard_rows(_ptr(a), n, k, lo, hi, _ptr(out),
                                                    nthreads or os.cpu_count() or 1)
        return out

    def serial(self, idx) -> np.ndarray:
        a = _as_idx(idx)
        n, k = a.shape
        out = np.empty((n * k, 3), dtype=np.float64, order="F")
        self.lib.gficf_oracle_serial_jaccard(_ptr(a), n, k, _ptr(out))
        return out


class Reference:
    """The reference's own Jaccard sources (oracle/_ref), R runtime stubbed."""

    @staticmethod
    def available() -> bool:
        return os.path.exists(REF_SO)

    def __init__(self):
        if not os.path.exists(REF_SO):
            raise FileNotFoundError(REF_SO + " (build it in the container: make -C oracle ref)")
        self.lib = C.CDLL(REF_SO)
        L = self.lib
        L.gficf_ref_parallel_jaccard.argtypes = [_dp, C.c_int64, C.c_int32, _dp, C.c_int32, C.c_int32]
        L.gficf_ref_parallel_jaccard.restype = C.c_double
        L.gficf_ref_parallel_jaccard_rows.argtypes = [
            _dp, C.c_int64, C.c_int32, C.c_int64, C.c_int64, _dp, C.c_int32]
        L.gficf_ref_parallel_jaccard_rows.restype = C.c_double
        L.gficf_ref_serial_jaccard.argtypes = [_dp, C.c_int64, C.c_int32, _dp, C.c_int32]
        L.gficf_ref_serial_jaccard.restype = C.c_double
        L.gficf_ref_capture_begin.argtypes = []
        L.gficf_ref_capture_begin.restype = None
        L.gficf_ref_capture_end.argtypes = [C.c_char_p, C.c_int64]
        L.gficf_ref_capture_end.restype = C.c_int64
        L.gficf_ref_hw_threads.argtypes = []
        L.gficf_ref_hw_threads.restype = C.c_int32
        self.last_seconds = 0.0

    def parallel(self, idx, print_output: bool = False, nthreads: int = 0) -> np.ndarray:
        a = _as_idx(idx)
        n, k = a.shape
        out = np.empty((n * k, 3), dtype=np.float64, order="F")
        self.last_seconds = self.lib.gficf_ref_parallel_jaccard(
            _ptr(a), n, k, _ptr(out), int(print_output), nthreads)
        return out

    def parallel_rows(self, idx, lo: int, hi: int, nthreads: int = 0) -> np.ndarray:
        a = _as_idx(idx)
        n, k = a.shape
        out = np.zeros(((hi - lo) * k, 3), dtype=np.float64, order="F")
        self.last_seconds = self.lib.gficf_ref_parallel_jaccard_rows(
            _ptr(a), n, k, lo, hi, _ptr(out), nthreads)
        return out

    def serial(self, idx, print_output: bool = False) -> np.ndarray:
        a = _as_idx(idx)
        n, k = a.shape
        out = np.empty((n * k, 3), dtype=np.float64, order="F")
        self.last_seconds = self.lib.gficf_ref_serial_jaccard(_ptr(a), n, k, _ptr(out), int(print_output))
        return out

    def captured(self, fn, *args, **kw):
        """Run fn and return (result, text the reference printed through Rprintf)."""
        self.lib.gficf_ref_capture_begin()
        try:
            res = fn(*args, **kw)
        finally:
            buf = C.create_string_buffer(4096)
            self.lib.gficf_ref_capture_end(buf, 4096)
        return res, buf.value.decode()

    def hw_threads(self) -> int:
        return int(self.lib.gficf_ref_hw_threads())


# ---------------------------------------------------------------------------------------------
# Mann-Whitney U per gene: matX n_genes x n1, matY n_genes x n2 (Fortran order
# float64, what R hands over) -> n_genes x 2 (p-value, log2 fold change), Fortran order.
# ---------------------------------------------------------------------------------------------
def _as_f64(m) -> np.ndarray:
    a = np.asarray(m)
    if a.ndim != 2:
        raise ValueError("a matrix is required")
    return np.asfortranarray(a, dtype=np.float64)


class WmuOracle:
    """oracle/wmu_oracle.c: our plain-C restatement of rcpp_parallel_mann_whitney.cpp:27-100."""

    def __init__(self):
        if not os.path.exists(ORACLE_SO):
            build()
        self.lib = C.CDLL(ORACLE_SO)
        self.lib.gficf_oracle_wmu.argtypes = [_dp, _dp, C.c_int64, C.c_int64, C.c_int64, _dp, C.c_int32]
        self.lib.gficf_oracle_wmu.restype = None
        self.lib.gsl_cdf_ugaussian_P.argtypes = [C.c_double]
        self.lib.gsl_cdf_ugaussian_P.restype = C.c_double
        self.lib.gsl_cdf_ugaussian_Q.argtypes = [C.c_double]
        self.lib.gsl_cdf_ugaussian_Q.restype = C.c_double

    def wmu(self, mat_x, mat_y, nthreads: int | None = None) -> np.ndarray:
        x, y = _as_f64(mat_x), _as_f64(mat_y)
        g = x.shape[0]
        if y.shape[0] != g:
            raise ValueError("matX and matY must have the same number of rows (genes)")
        out = np.empty((g, 2), dtype=np.float64, order="F")
        self.lib.gficf_oracle_wmu(_ptr(x), _ptr(y), g, x.shape[1], y.shape[1], _ptr(out),
                                  nthreads or os.cpu_count() or 1)
        return out


class WmuReference:
    """The reference's own mann_whitney.cpp + rcpp_parallel_mann_whitney.cpp (oracle/_ref), R runtime
    stubbed, GSL's normal cdf restated (parity unpinned on the cdf)."""

    @staticmethod
    def available() -> bool:
        return os.path.exists(REF_WMU_SO)

    def __init__(self):
        if not os.path.exists(REF_WMU_SO):
            raise FileNotFoundError(REF_WMU_SO + " (build it in the container: make -C oracle ref)")
        self.lib = C.CDLL(REF_WMU_SO)
        self.lib.gficf_ref_wmu.argtypes = [_dp, _dp, C.c_int64, C.c_int64, C.c_int64, _dp, C.c_int32, C.c_int32]
        self.lib.gficf_ref_wmu.restype = C.c_double
        self.last_seconds = 0.0

    def wmu(self, mat_x, mat_y, nthreads: int = 0) -> np.ndarray:
        x, y = _as_f64(mat_x), _as_f64(mat_y)
        g = x.shape[0]
        out = np.empty((g, 2), dtype=np.float64, order="F")
        self.last_seconds = self.lib.gficf_ref_wmu(_ptr(x), _ptr(y), g, x.shape[1], y.shape[1], _ptr(out), 0, nthreads)
        return out
