"""Shared inputs and comparisons of the network tests (emulated on the CPU, and on the GPU)."""
import numpy as np

def seq_sum(x, s0=0.0):
    """std::accumulate(first, last, s0): strictly left to right (np.cumsum on a 1-D array adds in order)."""
    x = np.asarray(x, dtype=np.float64)
    if x.size == 0:
        return float(s0)
    return float(np.cumsum(np.concatenate([[s0], x]))[-1])


def random_lower(rng, nv, m, k=20):
    """About m lower-triangle entries (node1 < node2) in column-major order with Jaccard-like weights
    u/(2k-u), some of them doubled like the mutual pairs of the SNN graph."""
    a = rng.integers(0, nv, 2 * m).astype(np.int64)
    b = rng.integers(0, nv, 2 * m).astype(np.int64)
    keep = a != b
    key = np.unique(np.minimum(a, b)[keep] * nv + np.maximum(a, b)[keep])  # sorted by (column, row)
    if key.size > m:
        key = np.sort(rng.choice(key, m, replace=False))
    u = rng.integers(1, k + 1, key.size)
    w = u / (2 * k - u) * rng.integers(1, 3, key.size)
    return (key // nv).astype(np.int32), (key % nv).astype(np.int32), w.astype(np.float64)


def to_csc(node1, node2, nv):
    colptr = np.zeros(nv + 1, np.int64)
    np.add.at(colptr, node1.astype(np.int64) + 1, 1)
    return np.cumsum(colptr), node2.copy()


def assert_same_network(got, want):
    """Every array and every scalar bit for bit."""
    assert got["n_nodes"] == want["n_nodes"]
    assert np.array_equal(got["first"].astype(np.int64), want["first"].astype(np.int64))
    assert np.array_equal(got["neighbor"], want["neighbor"])
    assert np.array_equal(got["edge_w"], want["edge_w"])
    assert np.array_equal(got["node_w"], want["node_w"])
    assert got["self_links"] == want["self_links"]
    assert got["total_w"] == want["total_w"]
