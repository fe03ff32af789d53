"""CPU: the network oracle (oracle/modopt_oracle.c) against the golden vectors produced by the
reference's own Network / VOSClusteringTechnique classes (tests/golden/net_*.npz, generator
tests/golden/make_golden.py) and, where oracle/_ref was built, against those classes directly."""
import glob
import os

import numpy as np
import pytest

from oracle.binding import NetworkOracle, NetworkReference
from tests.network_cases import assert_same_network, random_lower

NET_GOLDEN = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "net_*.npz")))
KEYS = ("first", "neighbor", "edge_w", "node_w", "total_w", "self_links")


def golden_network(g, tag):
    d = {k: (g[tag + "_" + k] if g[tag + "_" + k].ndim else float(g[tag + "_" + k])) for k in KEYS}
    d["n_nodes"] = d["node_w"].size
    return d


def test_network_golden_present():
    assert len(NET_GOLDEN) >= 3


@pytest.mark.parametrize("path", NET_GOLDEN, ids=[os.path.basename(p)[:-4] for p in NET_GOLDEN])
def test_oracle_matches_network_golden(path):
    g = np.load(path)
    O = NetworkOracle()
    net = O.network(g["node1"], g["node2"], g["w"])
    assert_same_network(net, golden_network(g, "net"))
    res = float(g["resolution"])
    assert O.quality(net, g["cluster"], res)[0] == float(g["quality"])
    red = O.reduce(net, g["cluster"])
    assert_same_network(red, golden_network(g, "red"))
    assert O.quality(red, g["cluster2"], res)[0] == float(g["quality2"])
    assert_same_network(O.reduce(red, g["cluster2"]), golden_network(g, "red2"))


@pytest.mark.skipif(not NetworkReference.available(), reason="oracle/_ref not built")
@pytest.mark.parametrize("nv,m,nc,seed", [(30, 100, 4, 1), (300, 2500, 17, 2), (5000, 60_000, 400, 3), (2000, 50_000, 3, 4)])
def test_oracle_matches_reference_classes(nv, m, nc, seed):
    rng = np.random.default_rng(seed)
    n1, n2, w = random_lower(rng, nv, m)
    nv = int(max(n1.max(), n2.max())) + 1
    O, R = NetworkOracle(), NetworkReference()
    no, nr = O.network(n1, n2, w), R.network(n1, n2, w)
    assert_same_network(no, nr)
    cl = rng.integers(0, nc, nv).astype(np.int32)
    cl[rng.permutation(nv)[:nc]] = np.arange(nc)
    res = 0.8 / (2 * no["total_w"])
    assert O.quality(no, cl, res)[0] == R.quality(nr, cl, res)
    ro, rr = O.reduce(no, cl), R.reduce(nr, cl)
    assert_same_network(ro, rr)
    cl2 = rng.integers(0, 2, nc).astype(np.int32)
    cl2[:2] = [0, 1]
    assert O.quality(ro, cl2, res)[0] == R.quality(rr, cl2, res)
    r2 = R.reduce(rr, cl2)
    assert_same_network(O.reduce(ro, cl2), r2)
    for x in (r2, rr, nr):
        R.free(x)
