"""GPU (-m gpu): the network pieces of the community detection (gficf_b200.modularity ->
gficf_cuda_network_*_dev) against the oracle (oracle/modopt_oracle.c) and, where oracle/_ref was
built, the reference's own Network / VOSClusteringTechnique classes.  The same kernels and launch
sequences run emulated on the CPU in tests/test_network_emu.py."""
import ctypes as C

import numpy as np
import pytest
import torch

from gficf_b200 import synth
from oracle import louvain
from oracle.binding import MODOPT_BIN, NetworkOracle, NetworkReference
from tests.network_cases import assert_same_network, random_lower, to_csc

pytestmark = pytest.mark.gpu


def gpu_network(n1, n2, w, nv):
    from gficf_b200 import modularity

    colptr, row = to_csc(n1, n2, nv)
    return modularity.matrix_to_network(torch.from_numpy(colptr).cuda(), torch.from_numpy(row).cuda(),
                                        torch.from_numpy(w).cuda())


def as_dict(net):
    torch.cuda.synchronize()
    return dict(n_nodes=net.n_nodes, first=net.first_neighbor_index.cpu().numpy(), neighbor=net.neighbor.cpu().numpy(),
                edge_w=net.edge_weight.cpu().numpy(), node_w=net.node_weight.cpu().numpy(),
                total_w=net.get_total_edge_weight(), self_links=net.total_edge_weight_self_links)


def full_clustering(rng, n, nc):
    cl = rng.integers(0, nc, n).astype(np.int32)
    cl[rng.permutation(n)[:nc]] = np.arange(nc)
    return cl


@pytest.mark.parametrize("nv,m,nc,seed", [(5, 6, 2, 1), (60, 300, 7, 2), (700, 6000, 40, 3), (20_000, 300_000, 500, 4),
                                           (100_000, 1_500_000, 2000, 5), (80_000, 1_000_000, 2, 6)])
def test_network_quality_and_reduction_match_oracle(cuda, nv, m, nc, seed):
    rng = np.random.default_rng(seed)
    n1, n2, w = random_lower(rng, nv, m)
    nv = int(max(n1.max(), n2.max())) + 1
    O = NetworkOracle()
    want = O.network(n1, n2, w)
    net = gpu_network(n1, n2, w, nv)
    assert_same_network(as_dict(net), want)
    nc = min(nc, nv)
    cl = full_clustering(rng, nv, nc)
    res = 0.8 / (2 * want["total_w"])
    q_want, cw_want = O.quality(want, cl, res)
    assert net.calc_quality_function(cl, res) == q_want  # the reference's double, bit for bit
    assert np.array_equal(net.cluster_weights(cl).cpu().numpy(), cw_want)  # bit-exact
    red_want = O.reduce(want, cl)
    red = net.create_reduced_network(cl)
    assert_same_network(as_dict(red), red_want)
    if nc >= 4:  # the next level: weights are sums now, self links are non-zero
        nc2 = max(2, nc // 5)
        cl2 = full_clustering(rng, nc, nc2)
        q2_want, cw2_want = O.quality(red_want, cl2, res)
        assert red.calc_quality_function(cl2, res) == q2_want
        assert np.array_equal(red.cluster_weights(cl2).cpu().numpy(), cw2_want)
        assert_same_network(as_dict(red.create_reduced_network(cl2)), O.reduce(red_want, cl2))


def test_edge_cases(cuda):
    from gficf_b200 import GficfCudaError, _lib, device as D, modularity

    rng = np.random.default_rng(9)
    n1, n2, w = random_lower(rng, 3000, 40_000)
    nv = int(max(n1.max(), n2.max())) + 1
    O = NetworkOracle()
    want = O.network(n1, n2, w)
    net = gpu_network(n1, n2, w, nv)
    one = np.zeros(nv, np.int32)  # one cluster: everything becomes self links
    red = net.create_reduced_network(one)
    assert red.n_edges == 0
    assert_same_network(as_dict(red), O.reduce(want, one))
    q_want, _ = O.quality(want, one, 0.01)
    assert net.calc_quality_function(one, 0.01) == q_want
    single = np.arange(nv, dtype=np.int32)  # singletons: the reduced network is the network
    red = as_dict(net.create_reduced_network(single))
    assert_same_network(red, O.reduce(want, single))
    assert np.array_equal(red["neighbor"], want["neighbor"])
    # output capacity too small: GFICF_E_LIMIT and the number of entries needed
    cl = full_clustering(rng, nv, 9)
    d_cl = torch.from_numpy(cl).cuda()
    L = _lib.lib()
    scratch = torch.empty((int(L.gficf_cuda_network_scratch_bytes(nv, net.n_edges)),), dtype=torch.uint8, device="cuda")
    r_first = torch.empty(10, dtype=torch.int64, device="cuda")
    r_nb = torch.empty(3, dtype=torch.int32, device="cuda")
    r_w = torch.empty(3, dtype=torch.float64, device="cuda")
    r_nw = torch.empty(9, dtype=torch.float64, device="cuda")
    sc = torch.zeros(2, dtype=torch.float64, device="cuda")
    flags = D.new_flags("cuda")
    needed = C.c_int64(0)
    rc = L.gficf_cuda_network_reduce_dev(net.first_neighbor_index.data_ptr(), net.neighbor.data_ptr(),
                                         net.edge_weight.data_ptr(), net.node_weight.data_ptr(), nv, net.n_edges,
                                         d_cl.data_ptr(), 9, 0.0, r_first.data_ptr(), r_nb.data_ptr(), r_w.data_ptr(), 3,
                                         r_nw.data_ptr(), sc.data_ptr(), sc[1:].data_ptr(), C.byref(needed),
                                         scratch.data_ptr(), scratch.numel(), flags.data_ptr(), 0)
    assert rc == 5 and needed.value == O.reduce(want, cl)["neighbor"].size
    # a scratch buffer that is too small is refused before anything is launched
    rc = L.gficf_cuda_network_reduce_dev(net.first_neighbor_index.data_ptr(), net.neighbor.data_ptr(),
                                         net.edge_weight.data_ptr(), net.node_weight.data_ptr(), nv, net.n_edges,
                                         d_cl.data_ptr(), 9, 0.0, r_first.data_ptr(), r_nb.data_ptr(), r_w.data_ptr(), 3,
                                         r_nw.data_ptr(), sc.data_ptr(), sc[1:].data_ptr(), C.byref(needed),
                                         scratch.data_ptr(), 1024, flags.data_ptr(), 0)
    assert rc == 1
    # bad inputs are flagged and surface as errors in the Python mirror
    bad_row = n2.copy()
    bad_row[5] = n1[5]
    with pytest.raises(ValueError, match="diagonal"):
        gpu_network(n1, bad_row, w, nv)
    wz = w.copy()
    wz[7] = 0.0
    with pytest.raises(ValueError, match="> 0"):
        gpu_network(n1, n2, wz, nv)
    bad_cl = cl.copy()
    bad_cl[3] = 9
    with pytest.raises(ValueError, match="cluster id"):
        net.calc_quality_function(bad_cl, 0.01, n_clusters=9)
    with pytest.raises(ValueError, match="no network data"):
        modularity.matrix_to_network(torch.zeros(3, dtype=torch.int64, device="cuda"),
                                     torch.zeros(0, dtype=torch.int32, device="cuda"),
                                     torch.zeros(0, dtype=torch.float64, device="cuda"))
    assert GficfCudaError is not None


def test_on_the_device_graph_of_a_knn_matrix_with_reference_labels(cuda, oracle):
    """kNN matrix -> (device) Jaccard counts -> SNN lower triangle -> network, all resident on the
    GPU, against the reference's steps on the host; then the quality and the reduced network of the
    clustering the reference's own Louvain run returns for this graph."""
    import os

    from gficf_b200 import device as D, modularity, snn

    n, k = 30_000, 30
    idx0 = synth.knn_index(n, k, family="planted", scramble=True)
    rel = oracle.parallel(synth.to_r_matrix(idx0))
    names, cols, rows_ref, data_ref = louvain.lower_triangle_edges(rel)
    O = NetworkOracle()
    want = O.network(cols, rows_ref, data_ref)
    padded, _ = D.pad_rows(idx0.cuda())
    colptr, rows, w, flags = snn.snn_lower_triangle(padded, n, k)
    assert int(flags[0]) & ~16 == 0
    net = modularity.matrix_to_network(colptr, rows, w)
    assert_same_network(as_dict(net), want)
    if os.path.exists(MODOPT_BIN):
        _, labels = louvain.louvain_labels(rel, n_start=2, n_iter=3)
    else:
        labels = (np.arange(names.size) * 7919 % 23).astype(np.int64)
    cl = labels.astype(np.int32)
    res = 0.8 / (2 * want["total_w"])  # resolution2 of RModularityOptimizer.cpp:101
    q_want, cw_want = O.quality(want, cl, res)
    assert net.calc_quality_function(cl, res) == q_want
    assert np.array_equal(net.cluster_weights(cl).cpu().numpy(), cw_want)
    red_want = O.reduce(want, cl)
    red = net.create_reduced_network(cl)
    assert_same_network(as_dict(red), red_want)
    if NetworkReference.available():
        R = NetworkReference()
        ref_net = R.network(cols, rows_ref, data_ref)
        assert net.calc_quality_function(cl, res) == R.quality(ref_net, cl, res)
        ref_red = R.reduce(ref_net, cl)
        assert_same_network(as_dict(red), ref_red)
        R.free(ref_red)
        R.free(ref_net)


def test_network_golden_vectors(cuda):
    """tests/golden/net_*.npz: outputs of the reference's own classes (make_golden.py)."""
    from tests.test_network_oracle import NET_GOLDEN, golden_network

    assert len(NET_GOLDEN) >= 3
    for path in NET_GOLDEN:
        g = np.load(path)
        want = golden_network(g, "net")
        net = gpu_network(g["node1"], g["node2"], g["w"], want["n_nodes"])
        assert_same_network(as_dict(net), want)
        res = float(g["resolution"])
        assert net.calc_quality_function(g["cluster"], res) == float(g["quality"])
        red = net.create_reduced_network(g["cluster"])
        assert_same_network(as_dict(red), golden_network(g, "red"))
        assert red.calc_quality_function(g["cluster2"], res) == float(g["quality2"])
        assert_same_network(as_dict(red.create_reduced_network(g["cluster2"])), golden_network(g, "red2"))
