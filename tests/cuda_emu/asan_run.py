"""Child process of tests/test_network_emu.py::test_emulated_kernels_under_address_sanitizer: runs
the emulated network pipeline from a library built with -fsanitize=address (LD_PRELOAD=libasan).
Usage: asan_run.py <libnetwork_emu_asan.so> [overflow]"""
import sys

import numpy as np

from oracle.binding import NetworkOracle
from tests.network_cases import assert_same_network, random_lower, seq_sum
from tests.test_network_emu import _emu_seq, _ip, _ll, _dp, _p, emu_build, emu_quality, emu_reduce, load_emu, to_csc


def main():
    L = load_emu(sys.argv[1])
    rng = np.random.default_rng(21)
    if len(sys.argv) > 2 and sys.argv[2] == "overflow":
        n1, n2, w = random_lower(rng, 200, 1500)
        nv = int(max(n1.max(), n2.max())) + 1
        colptr, row = to_csc(n1, n2, nv)
        nnz = row.size
        first, node_w, total = np.zeros(nv + 1, np.int64), np.zeros(nv), np.zeros(1)
        neighbor, edge_w = np.zeros(nnz, np.int32), np.zeros(2 * nnz)  # neighbor needs 2 * nnz entries
        L.emu_net_build(_p(colptr, _ll), _p(row, _ip), _p(w, _dp), nv, nnz, _p(first, _ll), _p(neighbor, _ip),
                        _p(edge_w, _dp), _p(node_w, _dp), _p(total, _dp), 2)
        print("the overflow went unnoticed")
        return
    O = NetworkOracle()
    for nv, m, nc, ctas in ((40, 150, 5, 1), (250, 2000, 12, 2)):
        n1, n2, w = random_lower(rng, nv, m)
        nv = int(max(n1.max(), n2.max())) + 1
        want = O.network(n1, n2, w)
        got, flags = emu_build(L, n1, n2, w, nv, ctas=ctas)
        assert flags == 0
        assert_same_network(got, want)
        nc = min(nc, nv)
        cl = rng.integers(0, nc, nv).astype(np.int32)
        cl[rng.permutation(nv)[:nc]] = np.arange(nc)
        q, cw, _ = emu_quality(L, got, cl, nc, 1e-3, ctas=ctas)
        q_want, cw_want = O.quality(want, cl, 1e-3)
        assert q == q_want and np.array_equal(cw, cw_want)
        red, _ = emu_reduce(L, got, cl, nc, ctas=ctas)
        assert_same_network(red, O.reduce(want, cl))
    for n in (0, 1, 4097):
        x = rng.random(n)
        assert _emu_seq(L, x, 0.5)[0] == seq_sum(x, 0.5)
    print("asan run ok")


if __name__ == "__main__":
    main()
