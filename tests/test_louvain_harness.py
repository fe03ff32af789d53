"""CPU: the label-parity harness (oracle/louvain.py + the reference's ModularityOptimizer built
into oracle/_ref/modopt) runs and recovers planted clusters; used by the GPU label test."""
import os

import numpy as np
import pytest

from oracle import louvain
from oracle.binding import MODOPT_BIN


@pytest.mark.skipif(not os.path.exists(MODOPT_BIN), reason="oracle/_ref/modopt not built")
def test_louvain_on_oracle_edges(oracle):
    rng = np.random.default_rng(3)
    n, k, c = 600, 15, 100
    rows = []
    for i in range(n):
        base = (i // c) * c
        cand = np.setdiff1d(np.arange(base, base + c), [i])
        rows.append(rng.choice(cand, k, replace=False))
    idx = np.asfortranarray(np.stack(rows).astype(np.float64) + 1.0)
    rel = oracle.parallel(idx)
    cells, labels = louvain.louvain_labels(rel)
    assert len(cells) == n == len(labels)
    by_cell = np.empty(n, dtype=np.int64)
    by_cell[cells - 1] = labels
    # disconnected planted clusters must come back as the communities
    assert len(np.unique(by_cell)) == n // c
    for b in range(0, n, c):
        assert len(np.unique(by_cell[b:b + c])) == 1
    cells2, labels2 = louvain.louvain_labels(rel.copy())
    assert np.array_equal(labels, labels2)
