"""Independent brute-force restatement in pure Python/numpy (TEST INFRASTRUCTURE).

Written from the reference's semantics, not from oracle/jaccard_oracle.c, so the two
checkers can be cross-validated on small inputs:

* ``parallel_jaccard``  /root/reference/src/rcpp_parallel_jaccard_coeff.cpp:24-55
  (multiset intersection via std::set_intersection on sorted copies, fixed output slots)
* ``serial_jaccard``    /root/reference/src/jaccard_coeff.cpp:28-42
  (unique-set intersection via Rcpp::intersect, compacted rows)

Pure-Python loops: only for small n*k.
"""
from __future__ import annotations

from collections import Counter

import numpy as np


def _trunc_to_int(x: float) -> int:
    # `int k = mat(i,j)-1;` : double arithmetic, then C truncation toward zero
    return int(x - 1.0)


def parallel_jaccard(idx) -> np.ndarray:
    a = np.asarray(idx, dtype=np.float64)
    n, k = a.shape
    out = np.zeros((n * k, 3), dtype=np.float64, order="F")
    rows = [Counter(a[i].tolist()) for i in range(n)]
    for i in range(n):
        for j in range(k):
            t = _trunc_to_int(a[i, j])
            u = sum((rows[i] & rows[t]).values())  # min multiplicity per value
            if u > 0:
                r = i * k + j
                out[r, 0] = i + 1
                out[r, 1] = t + 1
                out[r, 2] = u / (2.0 * k - u)
    return out


def serial_jaccard(idx) -> np.ndarray:
    a = np.asarray(idx, dtype=np.float64)
    n, k = a.shape
    out = np.zeros((n * k, 3), dtype=np.float64, order="F")
    rows = [set(a[i].tolist()) for i in range(n)]
    r = 0
    for i in range(n):
        for j in range(k):
            t = _trunc_to_int(a[i, j])
            u = len(rows[i] & rows[t])
            if u > 0:
                out[r, 0] = i + 1
                out[r, 1] = t + 1
                out[r, 2] = u / (2.0 * k - u)
                r += 1
    return out


def weight_lut(k: int) -> np.ndarray:
    """w(u) = u / (2.0*k - u) for u = 0..k  (rcpp_parallel_jaccard_coeff.cpp:51)."""
    return np.array([u / (2.0 * k - u) for u in range(k + 1)], dtype=np.float64)
