"""Generates tests/golden/*.npz from the REFERENCE ITSELF (run in the build container only).

The reference's own tests pin nothing for this path (tests/testthat.R:4 is commented out), so
the golden vectors are outputs of the reference's unmodified sources
(/root/reference/src/rcpp_parallel_jaccard_coeff.cpp, src/jaccard_coeff.cpp) compiled into
oracle/_ref/libgficf_ref.so by oracle/Makefile (R runtime replaced by oracle/rshim/).

    python tests/golden/make_golden.py

Each file holds: idx (n x k float64, 1-based), parallel (E x 3), serial (E x 3).
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle.binding import Reference  # noqa: E402
from tests.conftest import random_knn  # noqa: E402


def cases():
    rng = np.random.default_rng(180582)
    # name, idx
    yield "kat_appendix_b", np.array([[2, 3], [1, 3], [1, 2], [1, 2]], dtype=np.float64)
    yield "n300_k15_distinct", random_knn(rng, 300, 15)
    yield "n200_k30_distinct", random_knn(rng, 200, 30)
    yield "n120_k30_with_self", random_knn(rng, 120, 30, with_self=True)
    yield "n60_k8_repeated_ids", random_knn(rng, 60, 8, distinct=False)
    yield "n40_k30_repeated_ids", random_knn(rng, 40, 30, distinct=False)
    yield "n150_k100_distinct", random_knn(rng, 150, 100)
    yield "n64_k33_repeated_ids", random_knn(rng, 64, 33, distinct=False)
    yield "n33_k1", random_knn(rng, 33, 1)
    yield "n500_k5_uniform_sparse", random_knn(rng, 500, 5)
    # identical lists (u == k -> w == 1.0) and disjoint lists (u == 0 -> zero / skipped row)
    a = np.array([[2, 3, 4], [1, 3, 4], [5, 6, 7], [5, 6, 7], [1, 2, 3], [1, 2, 3], [1, 2, 3]], dtype=np.float64)
    yield "identical_and_disjoint", a
    yield "dup_row_5_5", np.array([[2, 2], [2, 3], [1, 1]], dtype=np.float64)


def main():
    ref = Reference()
    for name, idx in cases():
        idx = np.asfortranarray(idx, dtype=np.float64)
        par = ref.parallel(idx, nthreads=1)
        ser = ref.serial(idx)
        np.savez_compressed(os.path.join(HERE, name + ".npz"), idx=idx, parallel=par, serial=ser)
        print(name, idx.shape, "nonzero rows:", int((par[:, 2] > 0).sum()), int((ser[:, 2] > 0).sum()))


if __name__ == "__main__":
    main()
