"""GPU (-m gpu): downstream Louvain labels from the CUDA edge matrix equal those from the
reference-algorithm edge matrix (configs[1]: 100k cells, k=30, planted clusters), through the
reference's own ModularityOptimizer (oracle/_ref/modopt) with clustcells()'s "louvian 2"
parameters (R/clustCells.R:46,81)."""
import os

import numpy as np
import pytest

from gficf_b200 import synth
from oracle import louvain
from oracle.binding import MODOPT_BIN

pytestmark = pytest.mark.gpu


@pytest.mark.skipif(not os.path.exists(MODOPT_BIN), reason="oracle/_ref/modopt not shipped")
def test_louvain_labels_identical_100k(cuda, oracle):
    n, k = 100_000, 30
    r = synth.to_r_matrix(synth.knn_index(n, k))
    gpu_edges = cuda.rcpp_parallel_jaccard_coef(r)
    cpu_edges = oracle.parallel(r)
    assert np.array_equal(gpu_edges, cpu_edges)
    # one random start keeps the test short; the edges are bit-identical, so any setting agrees
    cells_g, lab_g = louvain.louvain_labels(gpu_edges, n_start=1, n_iter=10)
    cells_c, lab_c = louvain.louvain_labels(cpu_edges, n_start=1, n_iter=10)
    assert np.array_equal(cells_g, cells_c) and np.array_equal(lab_g, lab_c)
    assert 100 < len(np.unique(lab_g)) < 2000  # ~390 planted clusters of 256 cells
