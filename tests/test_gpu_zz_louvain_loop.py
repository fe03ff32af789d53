"""GPU (-m gpu): the reference's own Louvain run with its three bulk steps -- network construction,
every reduced network, every quality value -- served by the device kernels through
gficf_b200.modularity; labels and maximum modularity must be those of the unmodified reference.
The CPU suite runs the same loop on the emulated kernels
(tests/test_network_emu.py::test_louvain_labels_with_the_emulated_kernels_in_the_loop).
(Named to sort last: it was added after the round's GPU budget was spent and has run on the
emulation only.)"""
import numpy as np
import pytest

from gficf_b200 import synth
from oracle import louvain
from oracle.binding import NetworkReference
from tests.network_cases import MirrorHooks

pytestmark = pytest.mark.gpu


@pytest.mark.skipif(not NetworkReference.available(), reason="oracle/_ref not built")
@pytest.mark.parametrize("n,k,algorithm", [(3_000, 10, 1), (20_000, 15, 1), (20_000, 15, 2)])
def test_louvain_labels_with_the_device_kernels_in_the_loop(cuda, oracle, n, k, algorithm):
    rel = oracle.parallel(synth.to_r_matrix(synth.knn_index(n, k, family="planted", scramble=True)))
    names, cols, rows, data = louvain.lower_triangle_edges(rel)
    R = NetworkReference()
    want, q_want, _ = R.louvain(cols, rows, data, algorithm=algorithm, n_start=3, n_iter=4)
    hooks = MirrorHooks("cuda")
    got, q, calls = R.louvain(cols, rows, data, algorithm=algorithm, n_start=3, n_iter=4, hooks=hooks)
    assert calls[0] == 1 and calls[1] >= 3 and calls[2] >= 3
    assert np.array_equal(got, want)
    assert q == q_want
