"""Shared inputs and comparisons of the network tests (emulated on the CPU, and on the GPU)."""
import numpy as np

def seq_sum_tol(n_terms):
    """Relative distance allowed between a fixed-tree sum (error ~ log2(n) u) and the reference's
    strictly sequential sum of n positive terms (error up to (n-1) u, u = 2^-53; with the few distinct
    Jaccard weights the rounding errors are correlated and the sequential sum really drifts that far)."""
    return max(n_terms, 64) * 2.0 ** -53


REL = seq_sum_tol(1 << 22)  # cases of up to ~4M terms


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


def assert_same_network(got, want, exact_totals=False):
    assert got["n_nodes"] == want["n_nodes"]
    assert np.array_equal(got["first"].astype(np.int64), want["first"].astype(np.int64))
    assert np.array_equal(got["neighbor"], want["neighbor"])
    assert np.array_equal(got["edge_w"], want["edge_w"])  # bit-exact
    assert np.array_equal(got["node_w"], want["node_w"])  # bit-exact
    assert abs(got["self_links"] - want["self_links"]) <= REL * max(1.0, abs(want["self_links"]))
    assert abs(got["total_w"] - want["total_w"]) <= REL * max(1.0, abs(want["total_w"]))
