"""GPU (-m gpu): the SNN graph built on the device (gficf_b200.snn) equals what the reference's
R / igraph / C++ steps produce from the edge matrix (oracle/louvain.py restates them), and gives
the same Louvain labels through the reference's ModularityOptimizer."""
import os
import subprocess
import tempfile

import numpy as np
import pytest
import torch

from gficf_b200 import synth
from oracle import louvain
from oracle.binding import MODOPT_BIN

pytestmark = pytest.mark.gpu


def _gpu_triangle(n, k, idx0):
    from gficf_b200 import device as D, snn

    padded, _ = D.pad_rows(idx0.cuda())
    colptr, rows, w, flags = snn.snn_lower_triangle(padded, n, k)
    torch.cuda.synchronize()
    return colptr.cpu().numpy(), rows.cpu().numpy(), w.cpu().numpy(), int(flags[0])


@pytest.mark.parametrize("n,k,family", [(20_000, 15, "planted"), (30_000, 30, "planted"), (5_000, 100, "planted"),
                                         (3_000, 30, "uniform"), (300, 7, "planted")])
def test_lower_triangle_matches_reference_steps(cuda, oracle, n, k, family):
    idx0 = synth.knn_index(n, k, family=family, scramble=True)
    rel = oracle.parallel(synth.to_r_matrix(idx0))
    names, cols, rows_ref, data_ref = louvain.lower_triangle_edges(rel)
    colptr, rows, w, flags = _gpu_triangle(n, k, idx0)
    if not np.array_equal(names, np.arange(1, n + 1)):
        assert flags & 16  # a cell without edges: the device path must say so
        return
    assert flags == 0
    assert np.array_equal(np.repeat(np.arange(n), np.diff(colptr)), cols)
    assert np.array_equal(rows, rows_ref)
    assert np.array_equal(w, data_ref)  # bit-exact sums


def test_isolated_cell_is_flagged(cuda):
    from gficf_b200 import device as D, snn

    n, k = 2000, 5
    idx0 = synth.knn_index(n, k, family="uniform")  # u is almost always 0 -> many cells without edges
    padded, _ = D.pad_rows(idx0.cuda())
    _, _, _, flags = snn.snn_lower_triangle(padded, n, k)
    assert int(flags[0]) & 16


@pytest.mark.skipif(not os.path.exists(MODOPT_BIN), reason="oracle/_ref/modopt not shipped")
def test_labels_from_device_graph(cuda, oracle):
    n, k = 50_000, 30
    idx0 = synth.knn_index(n, k)
    colptr, rows, w, flags = _gpu_triangle(n, k, idx0)
    assert flags == 0
    cols = np.repeat(np.arange(n), np.diff(colptr))
    with tempfile.TemporaryDirectory() as d:
        fin, fout = os.path.join(d, "e.tsv"), os.path.join(d, "l.txt")
        with open(fin, "w") as f:
            for x, y, z in zip(cols.tolist(), rows.tolist(), w.tolist()):
                f.write("%d\t%d\t%.17g\n" % (x, y, z))
        p = louvain.DEFAULTS
        subprocess.run([MODOPT_BIN, fin, fout, str(p["modularity"]), repr(p["resolution"]), str(p["algorithm"]),
                        "1", "10", str(p["seed"]), "0"], check=True, capture_output=True)
        lab_gpu = np.loadtxt(fout, dtype=np.int64)
    rel = oracle.parallel(synth.to_r_matrix(idx0))
    cells, lab_ref = louvain.louvain_labels(rel, n_start=1, n_iter=10)
    assert np.array_equal(cells, np.arange(1, n + 1))
    assert np.array_equal(lab_gpu, lab_ref)


def test_snn_timing_record(cuda):
    """Not a pass/fail gate: prints the device time of the whole step at 1M x 30."""
    from gficf_b200 import device as D, snn

    n, k = 1_000_000, 30
    idx0 = synth.knn_index(n, k, scramble=True, device="cuda")
    padded, _ = D.pad_rows(idx0)
    snn.snn_lower_triangle(padded, n, k)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    colptr, rows, w, flags = snn.snn_lower_triangle(padded, n, k)
    e1.record()
    torch.cuda.synchronize()
    print("SNN 1M x 30: %.2f ms, nnz=%d" % (e0.elapsed_time(e1), rows.numel()))
    assert int(flags[0]) == 0
