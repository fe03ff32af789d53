"""Generates tests/golden/*.npz from the REFERENCE ITSELF (run in the build container only).

The reference's own tests pin nothing for this path (tests/testthat.R:4 is commented out), so
the golden vectors are outputs of the reference's unmodified sources
(/root/reference/src/rcpp_parallel_jaccard_coeff.cpp, src/jaccard_coeff.cpp) compiled into
oracle/_ref/libgficf_ref.so by oracle/Makefile (R runtime replaced by oracle/rshim/).

    python tests/golden/make_golden.py

Each file holds: idx (n x k float64, 1-based), parallel (E x 3), serial (E x 3).
wmu_*.npz: Mann-Whitney inputs and the reference worker's output (oracle/_ref/libgficf_ref_wmu.so).
net_*.npz: a lower-triangle edge list, a clustering and a second-level clustering, with what the
reference's own Network / VOSClusteringTechnique classes (oracle/_ref/libgficf_ref_modopt.so, built
from src/ModularityOptimizer.cpp unmodified) make of them: the network, its quality value, the
reduced network, and the same once more on the reduced network.
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


def wmu_cases():
    """Mann-Whitney fixtures: (name, matX, matY)."""
    rng = np.random.default_rng(180582)

    def sc(genes, cells, density, integer=False):
        m = rng.gamma(2.0, 50.0, size=(genes, cells)) * (rng.random((genes, cells)) < density)
        return np.floor(m / 40.0) if integer else m

    m = sc(24, 260, 0.2)
    m[0, :] = 3.0          # one tie group -> p = 1
    m[1, :] = 0.0
    m[2, :60] = 0.0        # complete separation
    m[2, 60:] = rng.random(200) + 1.0
    m[3, ::2] = -0.0       # -0.0 ties with +0.0
    yield "wmu_sparse_24x60_200", m[:, :60], m[:, 60:]
    m = sc(16, 96, 0.6, integer=True)
    yield "wmu_integer_ties_16x32_64", m[:, :32], m[:, 32:]
    m = rng.normal(size=(10, 41))
    yield "wmu_dense_negative_10x1_40", m[:, :1], m[:, 1:]


def net_cases():
    from oracle import louvain
    from gficf_b200 import synth
    from tests.network_cases import random_lower

    rng = np.random.default_rng(20261017)

    def clustering(n, nc):
        cl = rng.integers(0, nc, n).astype(np.int32)
        cl[rng.permutation(n)[:nc]] = np.arange(nc)
        return cl

    n1, n2, w = random_lower(rng, 400, 3000)
    yield "net_random_400_3000", n1, n2, w, clustering(int(max(n1.max(), n2.max())) + 1, 31), clustering(31, 5)
    n1, n2, w = random_lower(rng, 64, 1500)  # dense: long neighbour lists, heavy cluster pairs
    yield "net_dense_64_1500", n1, n2, w, clustering(int(max(n1.max(), n2.max())) + 1, 6), clustering(6, 2)
    # the real thing: SNN graph of a planted kNN matrix, clustering = the reference's own Louvain labels
    idx = synth.to_r_matrix(synth.knn_index(900, 12, family="planted", scramble=True))
    rel = Reference().parallel(idx, nthreads=1)
    names, cols, rows, data = louvain.lower_triangle_edges(rel)
    _, labels = louvain.louvain_labels(rel, n_start=2, n_iter=3)
    nc = int(labels.max()) + 1
    yield "net_snn_900_k12_louvain", cols.astype(np.int32), rows.astype(np.int32), data, labels.astype(np.int32), \
        clustering(nc, max(2, nc // 3))


def main():
    from oracle.binding import NetworkReference, WmuReference

    nref = NetworkReference()
    for name, n1, n2, w, cl, cl2 in net_cases():
        net = nref.network(n1, n2, w)
        res = 0.8 / (2 * net["total_w"] + net["self_links"])  # resolution2, RModularityOptimizer.cpp:101
        q = nref.quality(net, cl, res)
        red = nref.reduce(net, cl)
        q2 = nref.quality(red, cl2, res)
        red2 = nref.reduce(red, cl2)
        out = dict(node1=n1, node2=n2, w=w, cluster=cl, cluster2=cl2, resolution=res, quality=q, quality2=q2)
        for tag, x in (("net", net), ("red", red), ("red2", red2)):
            for key in ("first", "neighbor", "edge_w", "node_w", "total_w", "self_links"):
                out[tag + "_" + key] = x[key]
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
        print(name, "nodes", net["n_nodes"], "edges", net["neighbor"].size, "->", red["n_nodes"], red["neighbor"].size,
              "->", red2["n_nodes"], red2["neighbor"].size, "Q %.6f %.6f" % (q, q2))
        for x in (red2, red, net):
            nref.free(x)

    wref = WmuReference()
    for name, x, y in wmu_cases():
        x, y = np.asfortranarray(x, dtype=np.float64), np.asfortranarray(y, dtype=np.float64)
        out = wref.wmu(x, y, nthreads=1)
        np.savez_compressed(os.path.join(HERE, name + ".npz"), x=x, y=y, out=out)
        print(name, x.shape, y.shape, "p in [%.3g, %.3g]" % (np.nanmin(out[:, 0]), np.nanmax(out[:, 0])))
    ref = Reference()
    for name, idx in cases():
        idx = np.asfortranarray(idx, dtype=np.float64)
        par = ref.parallel(idx, nthreads=1)
        ser = ref.serial(idx)
        np.savez_compressed(os.path.join(HERE, name + ".npz"), idx=idx, parallel=par, serial=ser)
        print(name, idx.shape, "nonzero rows:", int((par[:, 2] > 0).sum()), int((ser[:, 2] > 0).sum()))


if __name__ == "__main__":
    main()
