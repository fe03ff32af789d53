"""The network kernels (gficf_b200/csrc/network_kernels.cuh) and their launch sequences
(network_plan.h) run on the CPU through tests/cuda_emu -- a fibre-per-thread emulation of CTAs,
barriers and warp collectives -- against the oracle and the reference's own classes.  This is the
no-GPU check of the code that tests/test_gpu_network.py runs on the B200."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from tests.network_cases import assert_same_network, random_lower, seq_sum, to_csc
from oracle.binding import NetworkOracle, NetworkReference

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_ll, _ip, _dp, _up, _ullp = (C.POINTER(C.c_longlong), C.POINTER(C.c_int), C.POINTER(C.c_double),
                             C.POINTER(C.c_uint), C.POINTER(C.c_ulonglong))


def _p(a, t):
    return a.ctypes.data_as(t)


def compile_emu(so, asan=False):
    cmd = ["g++", "-O1", "-std=c++17", "-ffp-contract=off", "-DGFICF_CUDA_EMU", "-I" + os.path.join(ROOT, "tests", "cuda_emu"),
           "-I" + os.path.join(ROOT, "gficf_b200", "csrc"), "-I" + os.path.join(ROOT, "include"), "-shared", "-fPIC",
           "-Wall", "-Werror", os.path.join(ROOT, "tests", "cuda_emu", "network_emu.cpp"), "-o", so]
    if asan:
        cmd[1:1] = ["-g", "-fsanitize=address", "-fno-omit-frame-pointer", "-DGFICF_NET_ASAN"]
    out = subprocess.run(cmd, capture_output=True, text=True)
    assert out.returncode == 0, out.stderr


def load_emu(so):
    L = C.CDLL(so)
    L.emu_scan.argtypes = [_ip, C.c_longlong, _ll]
    L.emu_radix_sort.argtypes = [_ullp, _up, C.c_longlong, C.c_int, C.c_int]
    L.emu_seq_sum.argtypes = [_dp, C.c_longlong, C.c_double, C.c_double, _up, C.c_int]
    L.emu_seq_sum.restype = C.c_double
    L.emu_net_build.argtypes = [_ll, _ip, _dp, C.c_longlong, C.c_longlong, _ll, _ip, _dp, _dp, _dp, C.c_int]
    L.emu_net_build.restype = C.c_uint
    L.emu_net_quality.argtypes = [_ll, _ip, _dp, _dp, C.c_longlong, _ip, C.c_int, C.c_double, C.c_double,
                                  C.c_double, _dp, _dp, C.c_int]
    L.emu_net_quality.restype = C.c_uint
    L.emu_net_reduce.argtypes = [_ll, _ip, _dp, _dp, C.c_longlong, _ip, C.c_int, C.c_double, _ll, _ip, _dp,
                                 C.c_longlong, _dp, _dp, _dp, _ll, _up, C.c_int]
    L.emu_net_reduce.restype = C.c_longlong
    return L


@pytest.fixture(scope="module")
def emu(tmp_path_factory):
    so = str(tmp_path_factory.mktemp("emu") / "libnetwork_emu.so")
    compile_emu(so)
    return load_emu(so)


def emu_build(L, node1, node2, w, nv, ctas=3):
    colptr, row = to_csc(node1, node2, nv)
    nnz = row.size
    first = np.zeros(nv + 1, np.int64)
    neighbor = np.zeros(2 * nnz, np.int32)
    edge_w = np.zeros(2 * nnz, np.float64)
    node_w = np.zeros(nv, np.float64)
    total = np.zeros(1, np.float64)
    flags = L.emu_net_build(_p(colptr, _ll), _p(row, _ip), _p(w, _dp), nv, nnz, _p(first, _ll), _p(neighbor, _ip),
                            _p(edge_w, _dp), _p(node_w, _dp), _p(total, _dp), ctas)
    return dict(n_nodes=nv, first=first, neighbor=neighbor, edge_w=edge_w, node_w=node_w, total_w=float(total[0]),
                self_links=0.0), flags


def emu_quality(L, net, cluster, nc, resolution, ctas=3):
    cw = np.zeros(nc, np.float64)
    q = np.zeros(1, np.float64)
    cl = np.ascontiguousarray(cluster, np.int32)
    flags = L.emu_net_quality(_p(net["first"], _ll), _p(net["neighbor"], _ip), _p(net["edge_w"], _dp),
                              _p(net["node_w"], _dp), net["n_nodes"], _p(cl, _ip), nc, resolution,
                              net["self_links"], net["total_w"], _p(cw, _dp), _p(q, _dp), ctas)
    return float(q[0]), cw, flags


def emu_reduce(L, net, cluster, nc, ctas=3, r_cap=None):
    cl = np.ascontiguousarray(cluster, np.int32)
    cap = max(1, net["neighbor"].size) if r_cap is None else r_cap
    r_first = np.zeros(nc + 1, np.int64)
    r_neighbor = np.zeros(max(cap, 1), np.int32)
    r_edge_w = np.zeros(max(cap, 1), np.float64)
    r_node_w = np.zeros(nc, np.float64)
    self_links = np.full(1, -1.0, np.float64)
    total = np.full(1, -1.0, np.float64)
    needed = np.zeros(1, np.int64)
    flags = np.zeros(1, np.uint32)
    n = L.emu_net_reduce(_p(net["first"], _ll), _p(net["neighbor"], _ip), _p(net["edge_w"], _dp),
                         _p(net["node_w"], _dp), net["n_nodes"], _p(cl, _ip), nc, net["self_links"],
                         _p(r_first, _ll), _p(r_neighbor, _ip), _p(r_edge_w, _dp), cap, _p(r_node_w, _dp),
                         _p(self_links, _dp),
                         _p(total, _dp), _p(needed, _ll), _p(flags, _up), ctas)
    if n < 0:
        return None, int(needed[0])
    ew = r_edge_w[:n].copy()
    return dict(n_nodes=nc, first=r_first, neighbor=r_neighbor[:n].copy(), edge_w=ew, node_w=r_node_w,
                total_w=float(total[0]), self_links=float(self_links[0])), int(flags[0])


def test_emulated_scan_sort_and_sum(emu):
    rng = np.random.default_rng(0)
    for m in [1, 31, 1024, 1025, 5000]:
        cnt = rng.integers(0, 9, m).astype(np.int32)
        out = np.zeros(m + 1, np.int64)
        emu.emu_scan(_p(cnt, _ip), m, _p(out, _ll))
        assert np.array_equal(out, np.concatenate([[0], np.cumsum(cnt)]))
    for n, nbits, ctas in [(2, 3, 1), (100, 5, 3), (2048, 9, 3), (2049, 17, 1), (7000, 20, 2), (5000, 40, 3)]:
        keys = rng.integers(0, 1 << nbits, n).astype(np.uint64)
        vals = np.arange(n, dtype=np.uint32)
        k2, v2 = keys.copy(), vals.copy()
        emu.emu_radix_sort(_p(k2, _ullp), _p(v2, _up), n, nbits, ctas)
        order = np.argsort(keys, kind="stable")
        assert np.array_equal(k2, keys[order]) and np.array_equal(v2, vals[order]), (n, nbits)


def _emu_seq(emu, x, s0=0.0, scale=1.0, ctas=3):
    x = np.ascontiguousarray(x, np.float64)
    flags = np.zeros(1, np.uint32)
    return emu.emu_seq_sum(_p(x, _dp), x.size, s0, scale, _p(flags, _up), ctas), int(flags[0])


def test_emulated_sequential_sum_is_the_sequential_sum(emu):
    """The exact replay of s <- RN(s + x[t]): every case must give the bits of the left-to-right loop."""
    rng = np.random.default_rng(3)
    k = 30
    cases = {
        "empty": np.zeros(0), "one": np.array([0.3]), "zeros": np.zeros(100),
        "jaccard weights": rng.integers(1, k + 1, 70_000) / (2 * k - rng.integers(1, k + 1, 70_000)),
        "few distinct values (correlated rounding)": np.repeat([1 / 3, 0.1, 1 / 59, 2 / 58], 30_000),
        "ties: powers of two and halves": rng.choice([0.5, 0.25, 1.0, 3.0, 1.5, 2.0 ** -20, 2.0 ** -30], 50_000),
        "uniform": rng.random(40_000),
        "wide range": np.exp(rng.normal(0, 12, 30_000)),
        "large then tiny (absorbed)": np.concatenate([[1e18], rng.random(9000), [3e18], rng.random(5000) * 1e3]),
        "with skipped elements": rng.random(50_000) * (rng.random(50_000) < 0.3),
        "subnormal start": np.concatenate([[5e-324, 1e-310, 2.5e-308], rng.random(3000) * 1e-300]),
        "one past a block": rng.random(4097), "exactly a block": rng.random(4096),
    }
    for name, x in cases.items():
        for s0 in (0.0, 7.25, 1e-3):
            got, flags = _emu_seq(emu, x, s0)
            assert flags == 0, name
            assert got == seq_sum(x, s0), (name, s0, got, seq_sum(x, s0))
    x = cases["jaccard weights"]
    assert _emu_seq(emu, x, 0.0, 0.5, ctas=1)[0] == seq_sum(x) / 2.0
    for _ in range(40):  # random draws from the same families (a 4-minute soak of 2202 such cases passed, profiles/r02_network.md)
        n = int(rng.choice([3, 17, 100, 4096, 5000, 20_000]))
        kind = int(rng.integers(0, 5))
        x = [rng.random(n), np.exp(rng.normal(0, rng.uniform(1, 30), n)),
             rng.choice(2.0 ** rng.integers(-40, 40, 8), n) * rng.choice([1, 1.5, 3, 0.75], n),
             rng.random(n) * (rng.random(n) < rng.uniform(0.01, 0.9)), np.abs(rng.standard_cauchy(n))][kind]
        s0 = float(rng.choice([0.0, rng.random() * 100, 2.0 ** float(rng.integers(-60, 60))]))
        assert _emu_seq(emu, x, s0, ctas=int(rng.integers(1, 5)))[0] == seq_sum(x, s0), (kind, n, s0)
    # the domain: non-negative finite values
    for bad in (-1.0, np.nan, np.inf):
        y = rng.random(5000)
        y[1234] = bad
        assert _emu_seq(emu, y)[1] & 32


@pytest.mark.parametrize("nv,m,nc,seed", [(5, 6, 2, 1), (60, 300, 7, 2), (700, 6000, 40, 3), (600, 3000, 100, 4),
                                           (300, 9000, 3, 5)])
def test_emulated_network_pipeline_matches_oracle(emu, nv, m, nc, seed):
    rng = np.random.default_rng(seed)
    n1, n2, w = random_lower(rng, nv, m)
    nv = int(max(n1.max(), n2.max())) + 1
    O = NetworkOracle()
    want = O.network(n1, n2, w)
    got, flags = emu_build(emu, n1, n2, w, nv)
    assert flags == 0
    assert_same_network(got, want)
    # level 0: quality and reduced network for a random clustering that uses every cluster id
    nc = min(nc, nv)
    cl = rng.integers(0, nc, nv).astype(np.int32)
    cl[rng.permutation(nv)[:nc]] = np.arange(nc)
    res = 0.8 / (2 * want["total_w"])
    q_want, cw_want = O.quality(want, cl, res)
    q, cw, flags = emu_quality(emu, got, cl, nc, res)
    assert flags == 0
    assert np.array_equal(cw, cw_want)
    assert q == q_want
    red_want = O.reduce(want, cl)
    red, flags = emu_reduce(emu, got, cl, nc)
    assert flags == 0
    assert_same_network(red, red_want)
    # level 1 on the reduced network (self links now non-zero, weights are sums)
    if nc >= 4:
        nc2 = max(2, nc // 5)
        cl2 = rng.integers(0, nc2, nc).astype(np.int32)
        cl2[rng.permutation(nc)[:nc2]] = np.arange(nc2)
        q2_want, cw2_want = O.quality(red_want, cl2, res)
        q2, cw2, _ = emu_quality(emu, red, cl2, nc2, res)
        assert np.array_equal(cw2, cw2_want)
        assert q2 == q2_want
        red2_want = O.reduce(red_want, cl2)
        red2, _ = emu_reduce(emu, red, cl2, nc2)
        assert_same_network(red2, red2_want)


def test_emulated_edge_cases(emu):
    rng = np.random.default_rng(9)
    n1, n2, w = random_lower(rng, 200, 1500)
    nv = int(max(n1.max(), n2.max())) + 1
    O = NetworkOracle()
    want = O.network(n1, n2, w)
    got, _ = emu_build(emu, n1, n2, w, nv, ctas=1)
    # one cluster: no cross edge at all, everything becomes self links
    one = np.zeros(nv, np.int32)
    red, flags = emu_reduce(emu, got, one, 1)
    red_want = O.reduce(want, one)
    assert flags == 0 and red["neighbor"].size == 0
    assert_same_network(red, red_want)
    q, cw, _ = emu_quality(emu, got, one, 1, 0.01)
    q_want, cw_want = O.quality(want, one, 0.01)
    assert np.array_equal(cw, cw_want) and q == q_want
    # singletons: the reduced network is the network itself
    single = np.arange(nv, dtype=np.int32)
    red, _ = emu_reduce(emu, got, single, nv)
    assert_same_network(red, O.reduce(want, single))
    assert np.array_equal(red["neighbor"], want["neighbor"])
    # output capacity too small: the needed number of entries is reported, nothing is written past it
    cl = rng.integers(0, 9, nv).astype(np.int32)
    cl[:9] = np.arange(9)
    none, needed = emu_reduce(emu, got, cl, 9, r_cap=3)
    assert none is None and needed == O.reduce(want, cl)["neighbor"].size
    # an entry that is not strictly lower, a row out of range, a non-positive weight: flagged
    bad_row = n2.copy()
    bad_row[5] = n1[5]
    _, flags = emu_build(emu, n1, bad_row, w, nv)
    assert flags & 64
    wz = w.copy()
    wz[7] = 0.0
    _, flags = emu_build(emu, n1, n2, wz, nv)
    assert flags & 32
    bad_cl = cl.copy()
    bad_cl[3] = 9
    _, _, flags = emu_quality(emu, got, bad_cl, 9, 0.01)
    assert flags & 64


@pytest.mark.skipif(not NetworkReference.available(), reason="oracle/_ref not built")
def test_emulated_pipeline_matches_reference_classes(emu):
    """Same check against the reference's own Network / createReducedNetwork / calcQualityFunction."""
    rng = np.random.default_rng(5)
    n1, n2, w = random_lower(rng, 400, 3000)
    nv = int(max(n1.max(), n2.max())) + 1
    R = NetworkReference()
    want = R.network(n1, n2, w)
    got, _ = emu_build(emu, n1, n2, w, nv)
    assert_same_network(got, want)
    cl = rng.integers(0, 25, nv).astype(np.int32)
    cl[:25] = np.arange(25)
    res = 0.8 / (2 * want["total_w"])
    q, _, _ = emu_quality(emu, got, cl, 25, res)
    assert q == R.quality(want, cl, res)
    red_want = R.reduce(want, cl)
    red, _ = emu_reduce(emu, got, cl, 25)
    assert_same_network(red, red_want)
    R.free(red_want)
    R.free(want)


def test_python_mirror_over_the_emulated_abi(emu, monkeypatch):
    """gficf_b200.modularity (the Python mirror of the reference's Network interface) driven end to
    end on CPU tensors: the emulated library exports the same gficf_cuda_network_* entry points, so
    every argument the mirror passes -- order, width, buffer sizes -- is checked without a GPU."""
    import contextlib

    import torch

    from gficf_b200 import _lib, device as D, modularity

    for name in ("gficf_cuda_network_scratch_bytes", "gficf_cuda_network_dev", "gficf_cuda_network_quality_dev",
                 "gficf_cuda_network_reduce_dev"):
        res, args = _lib.PROTOTYPES[name]
        getattr(emu, name).restype = res
        getattr(emu, name).argtypes = args
    monkeypatch.setattr(_lib, "lib", lambda: emu)
    monkeypatch.setattr(D, "_require_cuda", lambda t, dtype: None)
    monkeypatch.setattr(D, "_stream_ptr", lambda: 0)
    monkeypatch.setattr(torch.cuda, "device", lambda dev: contextlib.nullcontext())

    rng = np.random.default_rng(11)
    n1, n2, w = random_lower(rng, 500, 4000)
    nv = int(max(n1.max(), n2.max())) + 1
    colptr, row = to_csc(n1, n2, nv)
    O = NetworkOracle()
    want = O.network(n1, n2, w)
    net = modularity.matrix_to_network(torch.from_numpy(colptr), torch.from_numpy(row), torch.from_numpy(w))

    def as_dict(x):
        return dict(n_nodes=x.n_nodes, first=x.first_neighbor_index.numpy(), neighbor=x.neighbor.numpy(),
                    edge_w=x.edge_weight.numpy(), node_w=x.node_weight.numpy(), total_w=x.get_total_edge_weight(),
                    self_links=x.total_edge_weight_self_links)

    assert_same_network(as_dict(net), want)
    cl = rng.integers(0, 12, nv).astype(np.int32)
    cl[:12] = np.arange(12)
    res = 0.8 / (2 * want["total_w"])
    q_want, cw_want = O.quality(want, cl, res)
    assert net.calc_quality_function(cl, res) == q_want
    assert np.array_equal(net.cluster_weights(cl).numpy(), cw_want)
    red_want = O.reduce(want, cl)
    red = net.create_reduced_network(cl)
    assert_same_network(as_dict(red), red_want)
    cl2 = np.array([0, 1, 2, 0, 1, 2, 0, 1, 2, 0, 1, 2], np.int32)
    q2_want, _ = O.quality(red_want, cl2, res)
    assert red.calc_quality_function(cl2, res) == q2_want
    assert_same_network(as_dict(red.create_reduced_network(cl2)), O.reduce(red_want, cl2))
    # flags surface as errors
    bad = cl.copy()
    bad[0] = 12
    with pytest.raises(ValueError, match="cluster id"):
        net.calc_quality_function(bad, res, n_clusters=12)
    wz = w.copy()
    wz[3] = -1.0
    with pytest.raises(ValueError, match="> 0"):
        modularity.matrix_to_network(torch.from_numpy(colptr), torch.from_numpy(row), torch.from_numpy(wz))
    with pytest.raises(ValueError, match="no network data"):
        modularity.matrix_to_network(torch.zeros(3, dtype=torch.int64), torch.zeros(0, dtype=torch.int32),
                                     torch.zeros(0, dtype=torch.float64))


def test_emulated_on_a_jaccard_graph_with_the_reference_louvain_labels(emu, oracle):
    """The real input: the SNN graph of a planted kNN matrix (weights u/(2k-u), mutual pairs doubled)
    and the clustering the reference's own Louvain run returns for it."""
    from gficf_b200 import synth
    from oracle import louvain
    from oracle.binding import MODOPT_BIN

    if not os.path.exists(MODOPT_BIN):
        pytest.skip("oracle/_ref/modopt not built")
    n, k = 1500, 10
    rel = oracle.parallel(synth.to_r_matrix(synth.knn_index(n, k, family="planted", scramble=True)))
    names, cols, rows, data = louvain.lower_triangle_edges(rel)
    _, labels = louvain.louvain_labels(rel, n_start=2, n_iter=3)
    O = NetworkOracle()
    want = O.network(cols, rows, data)
    got, flags = emu_build(emu, cols.astype(np.int32), rows.astype(np.int32), data, names.size)
    assert flags == 0
    assert_same_network(got, want)
    cl = labels.astype(np.int32)
    nc = int(cl.max()) + 1
    res = 0.8 / (2 * want["total_w"])
    q_want, cw_want = O.quality(want, cl, res)
    q, cw, _ = emu_quality(emu, got, cl, nc, res)
    assert np.array_equal(cw, cw_want) and q == q_want
    red, _ = emu_reduce(emu, got, cl, nc)
    assert_same_network(red, O.reduce(want, cl))


def test_emulated_pipeline_matches_the_golden_vectors(emu):
    """tests/golden/net_*.npz: outputs of the reference's own classes (make_golden.py)."""
    from tests.test_network_oracle import NET_GOLDEN, golden_network

    for path in NET_GOLDEN:
        g = np.load(path)
        want = golden_network(g, "net")
        got, flags = emu_build(emu, g["node1"], g["node2"], g["w"], want["n_nodes"])
        assert flags == 0
        assert_same_network(got, want)
        res, cl, cl2 = float(g["resolution"]), g["cluster"], g["cluster2"]
        nc, nc2 = int(cl.max()) + 1, int(cl2.max()) + 1
        assert emu_quality(emu, got, cl, nc, res)[0] == float(g["quality"])
        red, _ = emu_reduce(emu, got, cl, nc)
        assert_same_network(red, golden_network(g, "red"))
        assert emu_quality(emu, red, cl2, nc2, res)[0] == float(g["quality2"])
        red2, _ = emu_reduce(emu, red, cl2, nc2)
        assert_same_network(red2, golden_network(g, "red2"))


def test_emulated_kernels_under_address_sanitizer(tmp_path):
    """The emulation doubles as a memory checker: the same kernels and launch sequences compiled with
    -fsanitize=address, every scratch piece fenced by a poisoned red zone, outputs in exactly-sized
    heap buffers -- an index that runs past its array is reported with the kernel's source line.
    Runs in a child process (the sanitizer runtime has to be loaded first)."""
    import shutil
    import sys

    libasan = subprocess.run(["g++", "-print-file-name=libasan.so"], capture_output=True, text=True).stdout.strip()
    if not shutil.which("g++") or not os.path.isabs(libasan) or not os.path.exists(libasan):
        pytest.skip("no AddressSanitizer runtime next to g++")
    so = str(tmp_path / "libnetwork_emu_asan.so")
    compile_emu(so, asan=True)
    env = dict(os.environ, LD_PRELOAD=libasan, ASAN_OPTIONS="detect_leaks=0:abort_on_error=0:exitcode=23",
               PYTHONPATH=ROOT)
    out = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "cuda_emu", "asan_run.py"), so],
                         capture_output=True, text=True, env=env, timeout=900)
    assert "ERROR: AddressSanitizer" not in out.stderr, out.stderr[-4000:]
    assert out.returncode == 0, (out.stdout[-2000:], out.stderr[-2000:])
    assert "asan run ok" in out.stdout
    # and the checker does see an overflow when there is one (an output array half the needed size)
    out = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "cuda_emu", "asan_run.py"), so, "overflow"],
                         capture_output=True, text=True, env=env, timeout=900)
    assert out.returncode != 0 and "heap-buffer-overflow" in out.stderr and "net_fill_kernel" in out.stderr


@pytest.mark.skipif(not NetworkReference.available(), reason="oracle/_ref not built")
def test_louvain_labels_with_the_emulated_kernels_in_the_loop(emu, oracle, monkeypatch):
    """End to end for row 3: the reference's own Louvain run (its runLocalMovingAlgorithm, JavaRandom,
    mergeClusters -- oracle/ref_modopt_entry.cpp) with network construction, every reduced network and
    every quality value supplied by this repository's kernels through gficf_b200.modularity: the labels
    and the maximum modularity are those of the unmodified reference.  Also exercises a network
    without edges (everything merged) through the Python mirror."""
    import contextlib

    import torch

    from gficf_b200 import _lib, device as D, modularity, synth
    from oracle import louvain
    from tests.network_cases import MirrorHooks

    for name in ("gficf_cuda_network_scratch_bytes", "gficf_cuda_network_dev", "gficf_cuda_network_quality_dev",
                 "gficf_cuda_network_reduce_dev"):
        res, args = _lib.PROTOTYPES[name]
        getattr(emu, name).restype = res
        getattr(emu, name).argtypes = args
    monkeypatch.setattr(_lib, "lib", lambda: emu)
    monkeypatch.setattr(D, "_require_cuda", lambda t, dtype: None)
    monkeypatch.setattr(D, "_stream_ptr", lambda: 0)
    monkeypatch.setattr(torch.cuda, "device", lambda dev: contextlib.nullcontext())

    rel = oracle.parallel(synth.to_r_matrix(synth.knn_index(700, 8, family="planted", scramble=True)))
    names, cols, rows, data = louvain.lower_triangle_edges(rel)
    R = NetworkReference()
    for algorithm in (1,):  # algorithm 2 (multilevel refinement) runs in tests/test_gpu_zz_louvain_loop.py
        want, q_want, _ = R.louvain(cols, rows, data, algorithm=algorithm, n_start=2, n_iter=3)
        hooks = MirrorHooks("cpu")
        got, q, calls = R.louvain(cols, rows, data, algorithm=algorithm, n_start=2, n_iter=3, hooks=hooks)
        assert calls[0] == 1 and calls[1] >= 2 and calls[2] >= 2
        assert np.array_equal(got, want) and q == q_want
    if os.path.exists(louvain.MODOPT_BIN):  # and the driver restated around the hooks is the reference's own main loop
        _, labels = louvain.louvain_labels(rel, algorithm=1, n_start=2, n_iter=3)
        assert np.array_equal(R.louvain(cols, rows, data, algorithm=1, n_start=2, n_iter=3)[0], labels)
    # a network whose edges all became self links still answers (null data pointers are not passed down)
    top = hooks.top
    merged = top.create_reduced_network(np.zeros(top.n_nodes, np.int32))
    assert merged.n_edges == 0 and merged.n_nodes == 1
    assert merged.calc_quality_function(np.zeros(1, np.int32), 0.01) == \
        NetworkOracle().quality(dict(n_nodes=1, first=np.zeros(2, np.int32), neighbor=np.zeros(0, np.int32),
                                     edge_w=np.zeros(0), node_w=merged.node_weight.numpy(), total_w=0.0,
                                     self_links=merged.total_edge_weight_self_links), np.zeros(1, np.int32), 0.01)[0]
    again = merged.create_reduced_network(np.zeros(1, np.int32))
    assert again.n_edges == 0 and again.total_edge_weight_self_links == merged.total_edge_weight_self_links
