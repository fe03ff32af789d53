"""CPU: property-based checks (hypothesis) of the oracle's semantics on arbitrary small
neighbour matrices, including repeated ids and self ids -- the invariants the GPU tests lean on."""
import numpy as np
from hypothesis import given, settings, strategies as st

from oracle import numpy_ref


@st.composite
def knn_matrices(draw):
    n = draw(st.integers(1, 12))
    k = draw(st.integers(1, 6))
    vals = draw(st.lists(st.integers(1, n), min_size=n * k, max_size=n * k))
    return np.asfortranarray(np.array(vals, dtype=np.float64).reshape(n, k))


@settings(max_examples=150, deadline=None)
@given(knn_matrices())
def test_oracle_equals_numpy_restatement(oracle, idx):
    assert np.array_equal(oracle.parallel(idx, nthreads=2), numpy_ref.parallel_jaccard(idx))
    assert np.array_equal(oracle.serial(idx), numpy_ref.serial_jaccard(idx))


@settings(max_examples=150, deadline=None)
@given(knn_matrices())
def test_invariants(oracle, idx):
    n, k = idx.shape
    par, ser = oracle.parallel(idx), oracle.serial(idx)
    lut = numpy_ref.weight_lut(k)
    # every weight is one of the k+1 legal doubles; zero rows are entirely zero
    assert np.isin(par[:, 2], lut).all() and np.isin(ser[:, 2], lut).all()
    z = par[:, 2] == 0
    assert (par[z] == 0).all() and (par[~z, 0] >= 1).all()
    # fixed slots: from == i+1, to == idx[i,j]
    rows = np.repeat(np.arange(1, n + 1), k).astype(np.float64)
    assert np.array_equal(par[~z, 0], rows[~z])
    assert np.array_equal(par[~z, 1], idx.reshape(-1)[~z])  # C-order flatten of (n,k) = row-major (i,j)
    # the serial export is compacted: non-zero rows first, in (i,j) order
    m = int((ser[:, 2] > 0).sum())
    assert (ser[:m, 2] > 0).all() and (ser[m:] == 0).all()
    # rows without repeated ids: both exports agree after the caller's w>0 filter (clustCells.R:66)
    if all(len(set(r)) == k for r in idx.tolist()) and all(
            len(set(idx[int(t) - 1].tolist())) == k for t in idx.reshape(-1)):
        assert np.array_equal(par[~z], ser[:m])
    # a cell that lists itself gets weight 1.0 on that edge when its list has no repeats
    for i in range(n):
        if len(set(idx[i].tolist())) == k:
            for j in range(k):
                if idx[i, j] == i + 1:
                    assert par[i * k + j, 2] == 1.0
