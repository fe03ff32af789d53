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


class MirrorHooks:
    """The three bulk steps of the Louvain run served by gficf_b200.modularity (the product's Python
    mirror over the C ABI) for oracle.binding.NetworkReference.louvain: numpy in, numpy out.  `device`
    is "cuda" on the GPU box; the CPU suite passes "cpu" with the emulated library patched in."""

    def __init__(self, device):
        import torch

        from gficf_b200 import modularity

        self.torch, self.modularity, self.device = torch, modularity, device
        self.levels = 0

    def _t(self, a, dtype):
        return self.torch.as_tensor(np.ascontiguousarray(a), dtype=dtype).to(self.device)

    def _net(self, d):
        t = self.torch
        net = self.modularity.Network(d["n_nodes"], self._t(d["first"], t.int64), self._t(d["neighbor"], t.int32),
                                      self._t(d["edge_w"], t.float64), self._t(d["node_w"], t.float64), None,
                                      d["self_links"])
        return net

    @staticmethod
    def _dict(net):
        return dict(n_nodes=net.n_nodes, first=net.first_neighbor_index.cpu().numpy(), neighbor=net.neighbor.cpu().numpy(),
                    edge_w=net.edge_weight.cpu().numpy(), node_w=net.node_weight.cpu().numpy(),
                    total_w=net.get_total_edge_weight(), self_links=net.total_edge_weight_self_links)

    def network(self, node1, node2, w, n_nodes):
        colptr, row = to_csc(node1, node2, n_nodes)
        t = self.torch
        self.top = self.modularity.matrix_to_network(self._t(colptr, t.int64), self._t(row, t.int32),
                                                     self._t(w, t.float64))
        return self._dict(self.top)

    def quality(self, net, cluster, n_clusters, resolution):
        # the optimiser evaluates the top-level network only: the device copy is still there
        assert net["n_nodes"] == self.top.n_nodes
        return self.top.calc_quality_function(cluster, resolution, n_clusters=n_clusters)

    def reduce(self, net, cluster, n_clusters):
        self.levels += 1
        if net["n_nodes"] == self.top.n_nodes and net["neighbor"].size == self.top.n_edges:
            src = self.top
        else:  # a deeper level: the arrays come back from the optimiser; its total weight is not needed here
            src = self._net(net)
        return self._dict(src.create_reduced_network(cluster, n_clusters=n_clusters))
