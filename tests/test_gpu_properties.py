"""GPU (-m gpu): property-based parity -- arbitrary small neighbour matrices (repeated ids, self
ids, every k up to 40, tiny n) through the C ABI must equal the oracle bit for bit, in both
exports.  Exercises the fast kernels, the flag path and the exact kernels at random."""
import numpy as np
import pytest
from hypothesis import given, settings, strategies as st

pytestmark = pytest.mark.gpu


@st.composite
def knn_matrices(draw):
    n = draw(st.integers(1, 40))
    k = draw(st.integers(1, 40))
    distinct = draw(st.booleans())
    seed = draw(st.integers(0, 2**31 - 1))
    rng = np.random.default_rng(seed)
    if distinct and n >= k:
        a = np.stack([rng.choice(n, k, replace=False) for _ in range(n)])
    else:
        a = rng.integers(0, n, size=(n, k))
    return np.asfortranarray(a.astype(np.float64) + 1.0)


@settings(max_examples=120, deadline=None)
@given(knn_matrices())
def test_any_small_matrix(cuda, oracle, idx):
    assert np.array_equal(cuda.rcpp_parallel_jaccard_coef(idx), oracle.parallel(idx, nthreads=1))
    assert np.array_equal(cuda.jaccard_coeff(idx), oracle.serial(idx))
