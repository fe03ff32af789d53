"""Edge matrix -> Louvain labels through the reference's own ModularityOptimizer
(TEST INFRASTRUCTURE, NOT PRODUCT CODE).

Mirrors what clustcells(community.algo="louvian 2") does with the Jaccard result:

  R/clustCells.R:66      relations <- relations[relations[,3] > 0, ]
  R/clustCells.R:67-69   igraph::graph.data.frame(relations, directed=FALSE)
                         (vertices numbered by first appearance in c(from, to))
  R/clustCells.R:81      as_adjacency_matrix(g, attr="weight", sparse=T): symmetric, the
                         weights of parallel edges (i->j and j->i) SUMMED
  src/RModularityOptimizer.cpp:67-83   strictly-lower-triangle scan in column order
                         -> (node1 = col, node2 = row, weight)
  src/ModularityOptimizer.cpp:851-1012 STANDALONE main = same loop as
                         RModularityOptimizer.cpp:101-172 (resolution scaling, JavaRandom
                         seed, random starts, orderClustersByNNodes)

The optimiser binary is oracle/_ref/modopt, built unmodified from the reference by
oracle/Makefile.  igraph itself is third-party and absent; its two steps above
are restated here with numpy/scipy, and BOTH sides of a label-parity test pass
through this same code, so bit-identical edges must give identical labels.
"""
from __future__ import annotations

import os
import subprocess
import tempfile

import numpy as np
import scipy.sparse as sp

from .binding import MODOPT_BIN

# clustcells() defaults for "louvian 2" (R/clustCells.R:46,81)
DEFAULTS = dict(modularity=1, resolution=0.8, algorithm=1, n_start=10, n_iter=10, seed=180582)


def lower_triangle_edges(relations: np.ndarray):
    rel = np.asarray(relations, dtype=np.float64)
    rel = rel[rel[:, 2] > 0]
    names = np.concatenate([rel[:, 0], rel[:, 1]])
    uniq, first = np.unique(names, return_index=True)
    order = np.argsort(first, kind="stable")
    vertex_names = uniq[order]  # igraph vertex id -> cell id (1-based)
    rank = np.empty_like(order)
    rank[order] = np.arange(order.size)
    vid = rank[np.searchsorted(uniq, names)]
    m = rel.shape[0]
    a, b = vid[:m], vid[m:]
    nv = vertex_names.size
    adj = sp.coo_matrix((np.concatenate([rel[:, 2], rel[:, 2]]),
                         (np.concatenate([a, b]), np.concatenate([b, a]))), shape=(nv, nv)).tocsc()
    adj.sum_duplicates()
    adj.sort_indices()
    low = sp.tril(adj, k=-1, format="csc")
    low.sort_indices()
    cols = np.repeat(np.arange(nv), np.diff(low.indptr))
    return vertex_names, cols, low.indices, low.data


def louvain_labels(relations: np.ndarray, **kw) -> np.ndarray:
    """Labels per igraph vertex (0-based cluster ids), plus the vertex->cell map applied."""
    if not os.path.exists(MODOPT_BIN):
        raise FileNotFoundError(MODOPT_BIN + " (build it in the container: make -C oracle ref)")
    p = dict(DEFAULTS)
    p.update(kw)
    vertex_names, n1, n2, w = lower_triangle_edges(relations)
    with tempfile.TemporaryDirectory() as d:
        fin, fout = os.path.join(d, "edges.tsv"), os.path.join(d, "labels.txt")
        with open(fin, "w") as f:
            for x, y, z in zip(n1.tolist(), n2.tolist(), w.tolist()):
                f.write("%d\t%d\t%.17g\n" % (x, y, z))
        subprocess.run([MODOPT_BIN, fin, fout, str(p["modularity"]), repr(p["resolution"]),
                        str(p["algorithm"]), str(p["n_start"]), str(p["n_iter"]), str(p["seed"]), "0"],
                       check=True, capture_output=True)
        labels = np.loadtxt(fout, dtype=np.int64, ndmin=1)
    return vertex_names.astype(np.int64), labels
