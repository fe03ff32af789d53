"""CPU: the synthetic kNN generators give valid neighbour lists and are deterministic."""
import torch

from gficf_b200 import synth


def _valid(a, n, k):
    assert a.shape == (n, k) and a.dtype == torch.int32
    assert int(a.min()) >= 0 and int(a.max()) < n
    s = torch.sort(a, dim=1).values
    assert not bool((s[:, 1:] == s[:, :-1]).any())
    assert not bool((a == torch.arange(n, dtype=torch.int32)[:, None]).any())


def test_planted_and_uniform_valid():
    _valid(synth.knn_index(5000, 15), 5000, 15)
    _valid(synth.knn_index(3000, 30, scramble=True), 3000, 30)
    _valid(synth.knn_index(2000, 100, family="uniform"), 2000, 100)
    _valid(synth.knn_index(40, 30), 40, 30)  # cluster larger than n, rows need topping up


def test_deterministic_and_chunk_independent():
    a = synth.knn_index(4000, 30, seed=7)
    b = synth.knn_index(4000, 30, seed=7, chunk=1000)
    assert torch.equal(a, b)
    assert not torch.equal(a, synth.knn_index(4000, 30, seed=8))


def test_scramble_is_a_relabelling(oracle):
    import numpy as np

    a = synth.knn_index(600, 15)
    b = synth.scramble_ids(a)
    ra = oracle.parallel(synth.to_r_matrix(a))
    rb = oracle.parallel(synth.to_r_matrix(b))
    # the multiset of weights is invariant under relabelling
    assert np.array_equal(np.sort(ra[:, 2]), np.sort(rb[:, 2]))


def test_planted_community_follows_the_scramble():
    """The label of a cell is the block of its ORIGINAL id: ~95 % of the neighbours share it, with or
    without the id scramble."""
    n, k = 20_000, 15
    for scramble in (False, True):
        idx = synth.knn_index(n, k, family="planted", scramble=scramble).numpy()
        com = synth.planted_community(n, k, scramble=scramble).numpy()
        assert com.min() == 0 and com.max() == (n - 1) // 256
        same = (com[idx] == com[:, None]).mean()
        assert 0.9 < same < 0.99, same
