"""The data-parallel pieces of the community detection that reads the Jaccard graph (SURVEY
section 8f, "next" row 3), on the device next to the graph they read.

Mirror of the reference's interface for these steps (src/ModularityOptimizer.{h,cpp}):

  matrix_to_network(colptr, rows, w)        ModularityOptimizer::matrixToNetwork (.cpp:761-806) on the
                                            lower-triangle CSC that snn_lower_triangle() leaves on the
                                            device (node1 = column, node2 = row), modularity function 1
  Network                                   class Network (.h:59-125): nNodes, nEdges,
                                            firstNeighborIndex, neighbor, edgeWeight, nodeWeight,
                                            totalEdgeWeightSelfLinks
  Network.get_total_edge_weight()           Network::getTotalEdgeWeight (.cpp:268-270)
  Network.create_reduced_network(cluster)   Network::createReducedNetwork (.cpp:322-373)
  Network.calc_quality_function(cluster, r) VOSClusteringTechnique::calcQualityFunction (.cpp:462-482)
  Network.cluster_weights(cluster)          the clusterWeight vector of .cpp:474-476

What is NOT here: runLocalMovingAlgorithm (.cpp:484-583) -- one node at a time in a JavaRandom
permutation, each move changing the state the next one reads; it stays on the host
(INTEGRATION.md section 8 shows where these calls sit in runLouvainAlgorithm).

Every kernel is the library's own (network_kernels.cuh), called through the C ABI; torch is the
allocator and stream provider.  Every output is bit-identical to the reference, including the
sums it forms sequentially over the whole edge list (total edge weight, quality value, self-link
total): those are replayed exactly on the device (network_kernels.cuh).
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib
from . import device as D

FLAG_NET_WEIGHT, FLAG_NET_RANGE = 32, 64


def _check_flags(flags: torch.Tensor) -> None:
    f = int(flags[0])
    if f & FLAG_NET_RANGE:
        raise ValueError("network input out of range: an entry that is not strictly below the diagonal, "
                         "or a row / cluster id outside [0, n)")
    if f & FLAG_NET_WEIGHT:
        raise ValueError("edge weights must be > 0 (the reference's reduced-network bookkeeping, "
                         "ModularityOptimizer.cpp:342, treats a zero running weight as 'not seen yet')")


def _scratch(dev, n_nodes: int, n_items: int) -> torch.Tensor:
    nbytes = int(_lib.lib().gficf_cuda_network_scratch_bytes(int(n_nodes), int(n_items)))
    return torch.empty((nbytes,), dtype=torch.uint8, device=dev)


def _n_clusters(cluster: torch.Tensor, n_clusters: int | None) -> int:
    # Clustering(IVector cluster): nClusters = max + 1 (.cpp:90-96)
    return int(cluster.max()) + 1 if n_clusters is None else int(n_clusters)


class Network:
    """Device-resident symmetric CSR network; field names follow the reference's class."""

    def __init__(self, n_nodes, first_neighbor_index, neighbor, edge_weight, node_weight, total_edge_weight,
                 total_edge_weight_self_links=0.0):
        self.n_nodes = int(n_nodes)
        self.first_neighbor_index = first_neighbor_index  # int64 [n_nodes + 1]
        self.neighbor = neighbor                          # int32 [n_edges]
        self.edge_weight = edge_weight                    # float64 [n_edges]
        self.node_weight = node_weight                    # float64 [n_nodes]
        self._total = total_edge_weight                   # float64 [1] on the device
        self.total_edge_weight_self_links = float(total_edge_weight_self_links)

    @property
    def n_edges(self) -> int:
        """Directed edge count (the reference's nEdges member; getNEdges() is half of it)."""
        return int(self.neighbor.shape[0])

    def get_total_edge_weight(self) -> float:
        return float(self._total[0])

    def _edge_ptrs(self):
        """Pointers of neighbor / edge_weight; a network without edges (everything merged into self links)
        has empty tensors, whose data pointer is null: the C entry points want a real address."""
        if self.n_edges == 0:
            dev = self.node_weight.device
            self._dummy = (torch.zeros(1, dtype=torch.int32, device=dev), torch.zeros(1, dtype=torch.float64, device=dev))
            return self._dummy[0].data_ptr(), self._dummy[1].data_ptr()
        return self.neighbor.data_ptr(), self.edge_weight.data_ptr()

    def _cluster_arg(self, cluster) -> torch.Tensor:
        cl = torch.as_tensor(cluster, device=self.node_weight.device).to(torch.int32).contiguous()
        if cl.shape != (self.n_nodes,):
            raise ValueError("one cluster id per node is required")
        return cl

    def _quality(self, cluster, resolution: float, n_clusters: int | None):
        cl = self._cluster_arg(cluster)
        nc = _n_clusters(cl, n_clusters)
        dev = cl.device
        cw = torch.empty((nc,), dtype=torch.float64, device=dev)
        q = torch.empty((1,), dtype=torch.float64, device=dev)
        flags = D.new_flags(dev)
        scratch = _scratch(dev, self.n_nodes, self.n_edges)
        p_neighbor, p_edge_w = self._edge_ptrs()
        with torch.cuda.device(dev):
            _lib.check(_lib.lib().gficf_cuda_network_quality_dev(
                self.first_neighbor_index.data_ptr(), p_neighbor, p_edge_w,
                self.node_weight.data_ptr(), self.n_nodes, self.n_edges, cl.data_ptr(), nc, float(resolution),
                self.total_edge_weight_self_links, self._total.data_ptr(), cw.data_ptr(), q.data_ptr(),
                scratch.data_ptr(), scratch.numel(), flags.data_ptr(), D._stream_ptr()))
        _check_flags(flags)
        return q, cw

    def calc_quality_function(self, cluster, resolution: float, n_clusters: int | None = None) -> float:
        return float(self._quality(cluster, resolution, n_clusters)[0][0])

    def cluster_weights(self, cluster, n_clusters: int | None = None) -> torch.Tensor:
        return self._quality(cluster, 0.0, n_clusters)[1]

    def create_reduced_network(self, cluster, n_clusters: int | None = None) -> "Network":
        cl = self._cluster_arg(cluster)
        nc = _n_clusters(cl, n_clusters)
        dev = cl.device
        L = _lib.lib()
        cap = max(1, self.n_edges)
        r_first = torch.empty((nc + 1,), dtype=torch.int64, device=dev)
        r_neighbor = torch.empty((cap,), dtype=torch.int32, device=dev)
        r_edge_w = torch.empty((cap,), dtype=torch.float64, device=dev)
        r_node_w = torch.empty((nc,), dtype=torch.float64, device=dev)
        scalars = torch.zeros((2,), dtype=torch.float64, device=dev)  # [0] self-link total, [1] total edge weight
        flags = D.new_flags(dev)
        scratch = _scratch(dev, self.n_nodes, self.n_edges)
        n_red = C.c_int64(0)
        p_neighbor, p_edge_w = self._edge_ptrs()
        with torch.cuda.device(dev):
            _lib.check(L.gficf_cuda_network_reduce_dev(
                self.first_neighbor_index.data_ptr(), p_neighbor, p_edge_w,
                self.node_weight.data_ptr(), self.n_nodes, self.n_edges, cl.data_ptr(), nc,
                self.total_edge_weight_self_links, r_first.data_ptr(),
                r_neighbor.data_ptr(), r_edge_w.data_ptr(), cap, r_node_w.data_ptr(), scalars.data_ptr(),
                scalars[1:].data_ptr(), C.byref(n_red), scratch.data_ptr(), scratch.numel(), flags.data_ptr(),
                D._stream_ptr()))
        _check_flags(flags)
        e = n_red.value
        return Network(nc, r_first, r_neighbor[:e], r_edge_w[:e], r_node_w, scalars[1:2], float(scalars[0]))


def matrix_to_network(colptr: torch.Tensor, rows: torch.Tensor, weights: torch.Tensor) -> Network:
    """Lower-triangle CSC (colptr int64 [nv+1], rows int32 [nnz] ascending inside a column, weights
    float64 [nnz]; what snn_lower_triangle returns) -> Network with node weights = total edge
    weight per node (modularityFunction 1, the clustcells() default)."""
    D._require_cuda(colptr, torch.int64)
    D._require_cuda(rows, torch.int32)
    D._require_cuda(weights, torch.float64)
    nv, nnz = int(colptr.shape[0]) - 1, int(rows.shape[0])
    if nv < 1 or nnz < 1:
        raise ValueError("Matrix contained no network data.  Check format.")  # RModularityOptimizer.cpp:84-86
    if 2 * nnz >= 2 ** 31 - 1:
        raise ValueError("the network would have %d directed edges: beyond the reference's int indices" % (2 * nnz))
    dev = rows.device
    first = torch.empty((nv + 1,), dtype=torch.int64, device=dev)
    neighbor = torch.empty((2 * nnz,), dtype=torch.int32, device=dev)
    edge_w = torch.empty((2 * nnz,), dtype=torch.float64, device=dev)
    node_w = torch.empty((nv,), dtype=torch.float64, device=dev)
    total = torch.zeros((1,), dtype=torch.float64, device=dev)
    flags = D.new_flags(dev)
    scratch = _scratch(dev, nv, nnz)
    with torch.cuda.device(dev):
        _lib.check(_lib.lib().gficf_cuda_network_dev(
            colptr.data_ptr(), rows.data_ptr(), weights.data_ptr(), nv, nnz, first.data_ptr(), neighbor.data_ptr(),
            edge_w.data_ptr(), node_w.data_ptr(), total.data_ptr(), scratch.data_ptr(), scratch.numel(),
            flags.data_ptr(), D._stream_ptr()))
    _check_flags(flags)
    return Network(nv, first, neighbor, edge_w, node_w, total, 0.0)
