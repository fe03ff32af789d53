"""GPU (-m gpu): library lifecycle -- concurrent callers, release + reuse, alternating shapes."""
import threading

import numpy as np
import pytest

from tests.conftest import random_knn

pytestmark = pytest.mark.gpu


def test_concurrent_callers_are_serialised(cuda, oracle):
    rng = np.random.default_rng(0)
    cases = [random_knn(rng, 3000 + 500 * i, k) for i, k in enumerate((15, 30, 64, 7))]
    want = [oracle.parallel(c) for c in cases]
    got = [None] * len(cases)

    def run(i):
        for _ in range(3):
            got[i] = cuda.rcpp_parallel_jaccard_coef(cases[i])

    th = [threading.Thread(target=run, args=(i,)) for i in range(len(cases))]
    for t in th:
        t.start()
    for t in th:
        t.join()
    for g, w in zip(got, want):
        assert np.array_equal(g, w)


def test_release_and_reuse(cuda, oracle):
    rng = np.random.default_rng(1)
    a = random_knn(rng, 5000, 30)
    assert np.array_equal(cuda.rcpp_parallel_jaccard_coef(a), oracle.parallel(a))
    assert cuda.lib().gficf_cuda_release() == 0
    assert np.array_equal(cuda.rcpp_parallel_jaccard_coef(a), oracle.parallel(a))
    # shrinking and growing shapes reuse / grow the cached workspaces
    for n, k in ((100, 5), (20000, 30), (300, 100), (20001, 31)):
        b = random_knn(rng, n, k)
        assert np.array_equal(cuda.rcpp_parallel_jaccard_coef(b), oracle.parallel(b))
        assert np.array_equal(cuda.jaccard_coeff(b), oracle.serial(b))


def test_pageable_large_goes_through_staging_threads(cuda, oracle):
    """> 8 MiB of pageable memory takes the multi-threaded staged copies (both directions)."""
    from gficf_b200 import synth

    r = synth.to_r_matrix(synth.knn_index(200_000, 30, scramble=True))  # 48 MB in, 144 MB out
    got = cuda.rcpp_parallel_jaccard_coef(r)
    for lo, hi in ((0, 2000), (100_000, 102_000), (198_000, 200_000)):
        assert np.array_equal(got[lo * 30:hi * 30], oracle.parallel_rows(r, lo, hi))
    assert np.array_equal(cuda.jaccard_coeff(r)[: 30 * 1000], oracle.serial(r)[: 30 * 1000])


def test_narrowing_h2d_keeps_results_and_error_behaviour(cuda, oracle, monkeypatch):
    """An f64 matrix above 8 MiB is narrowed to int32 on the host while it streams to the device
    (half the PCIe bytes).  Same edges as sending the doubles, and every invalid id -- fractional,
    NaN, infinite, beyond int32, zero, negative, > n -- is still rejected with GFICF_E_RANGE."""
    from gficf_b200 import synth

    n, k = 45_000, 30  # 10.8 MB of doubles: above the small-copy threshold
    r = synth.to_r_matrix(synth.knn_index(n, k, scramble=True))
    want = oracle.parallel(r)
    for narrow, pinned in (("1", False), ("1", True), ("0", False), ("0", True)):
        monkeypatch.setenv("GFICF_CUDA_H2D_NARROW", narrow)
        src = r
        if pinned:
            src = cuda.pinned_empty((n, k))
            src[...] = r
        assert np.array_equal(cuda.rcpp_parallel_jaccard_coef(src), want), (narrow, pinned)
        assert cuda.last_output()["h2d_bytes"] == n * k * (4 if narrow == "1" else 8)
        assert np.array_equal(cuda.jaccard_coeff(src), oracle.serial(r))
    monkeypatch.setenv("GFICF_CUDA_H2D_NARROW", "1")
    for bad in (2.5, np.nan, np.inf, -np.inf, 3e9, -3e9, 0.0, -1.0, float(n + 1), 1e300, 0.999999):
        for pos in ((0, 0), (n - 1, k - 1), (n // 2, 7)):
            b = r.copy()
            b[pos] = bad
            with pytest.raises(cuda.GficfCudaError) as e:
                cuda.rcpp_parallel_jaccard_coef(b)
            assert e.value.code == 2, (bad, pos)
