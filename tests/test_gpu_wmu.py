"""GPU (-m gpu): Mann-Whitney U per gene through the C ABI (gficf_cuda_wmu_test) against the oracle
(oracle/wmu_oracle.c, pinned to the reference's own sources by tests/test_wmu_oracle.py).

Bit-exact: the device produces z and the mean ratio in integer / correctly-rounded arithmetic, the
host wrapper evaluates the normal cdf and log2 with the same libm the oracle uses.  (Against GSL
itself the cdf is unpinned -- GSL is absent here -- see oracle/gauss_cdf.c.)"""
import numpy as np
import pytest

from oracle.binding import WmuOracle
from tests.test_wmu_oracle import WMU_GOLDEN, sc_matrix

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("path", WMU_GOLDEN, ids=[__import__("os").path.basename(p)[:-4] for p in WMU_GOLDEN])
def test_wmu_golden_vectors(cuda, path):
    """Outputs of the reference itself (tests/golden/make_golden.py)."""
    g = np.load(path)
    assert np.array_equal(cuda.rcpp_parallel_WMU_test(g["x"], g["y"]), g["out"], equal_nan=True)


def _check(cuda, x, y):
    want = WmuOracle().wmu(x, y)
    got = cuda.rcpp_parallel_WMU_test(x, y)
    assert np.array_equal(got, want, equal_nan=True)
    return got


@pytest.mark.parametrize("genes,n1,n2,integer", [(64, 30, 200, False), (33, 1, 50, False), (40, 64, 64, True),
                                                  (20, 500, 3000, False), (12, 7, 5, True), (300, 513, 1024, False),
                                                  (5, 1, 1, False), (17, 2000, 30_000, True)])
def test_wmu_matches_oracle(cuda, genes, n1, n2, integer):
    rng = np.random.default_rng(genes * 7 + n1)
    m = sc_matrix(rng, genes, n1 + n2, integer=integer)
    m[0, :] = 3.0                      # one tie group: p = 1
    if genes > 4:
        m[1, :] = 0.0
        m[2, :n1] = 0.0                # complete separation
        m[2, n1:] = rng.random(n2) + 1.0
        m[3, :] = -m[3, :]             # negative values (log2 of a negative ratio: NaN in the reference too)
        m[4, ::2] = -0.0               # -0.0 ties with +0.0
    got = _check(cuda, np.asfortranarray(m[:, :n1]), np.asfortranarray(m[:, n1:]))
    assert got[0, 0] == 1.0


def test_wmu_dense_distinct_values_and_extremes(cuda):
    rng = np.random.default_rng(3)
    genes, n1, n2 = 50, 700, 900
    m = rng.normal(size=(genes, n1 + n2)) * 10.0 ** rng.integers(-300, 300, size=(genes, 1)).astype(np.float64)
    m[5, :10] = np.finfo(np.float64).max
    m[6, :10] = np.finfo(np.float64).tiny / 4  # subnormals
    m[7, :] = np.where(rng.random(n1 + n2) < 0.5, 1.0, np.nextafter(1.0, 2.0))  # neighbours in the last bit
    _check(cuda, np.asfortranarray(m[:, :n1]), np.asfortranarray(m[:, n1:]))


def test_wmu_large_n_ordered_tie_sum(cuda):
    """N >= 208064: the tie term no longer fits exact integer accumulation; the reference's sequential
    double sum is reproduced in sorted order (huge zero group first, then many small groups)."""
    rng = np.random.default_rng(9)
    genes, n1, n2 = 6, 20_000, 200_000
    m = np.floor(rng.gamma(2.0, 2.0, size=(genes, n1 + n2))) * (rng.random((genes, n1 + n2)) < 0.2)
    _check(cuda, np.asfortranarray(m[:, :n1]), np.asfortranarray(m[:, n1:]))


def test_wmu_many_genes_per_cta_and_staged_uploads(cuda):
    """More genes than resident CTAs (a CTA sorts several genes in turn) and matrices above the
    8 MiB small-copy threshold (both go through the pinned staging ring, back to back)."""
    rng = np.random.default_rng(11)
    genes, n1, n2 = 2500, 300, 2200   # 6 MB + 44 MB of pageable doubles
    m = sc_matrix(rng, genes, n1 + n2)
    _check(cuda, np.asfortranarray(m[:, :n1]), np.asfortranarray(m[:, n1:]))
    genes, n1, n2 = 120, 12_000, 30_000  # both above 8 MiB
    m = sc_matrix(rng, genes, n1 + n2)
    _check(cuda, np.asfortranarray(m[:, :n1]), np.asfortranarray(m[:, n1:]))


def test_wmu_argument_errors(cuda):
    x = np.asfortranarray(np.ones((4, 3)))
    with pytest.raises(ValueError):
        cuda.rcpp_parallel_WMU_test(x, np.ones((5, 3)))
    with pytest.raises(cuda.GficfCudaError) as e:
        cuda.rcpp_parallel_WMU_test(x, np.ones((4, 0)))
    assert e.value.code == 1
    assert cuda.rcpp_parallel_WMU_test(np.ones((0, 3)), np.ones((0, 2))).shape == (0, 2)
