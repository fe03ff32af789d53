"""The step after the Jaccard path, on the GPU ("next" row 1): from the index
matrix straight to the graph the community detection reads -- no 24-bytes-per-edge matrix, no R
filter, no igraph.

Reference steps replaced (R/clustCells.R:66-69,81; src/RModularityOptimizer.cpp:67-83):
``relations[relations[,3]>0,]`` -> ``graph.data.frame(directed=FALSE)`` ->
``as_adjacency_matrix(attr="weight")`` (parallel edges i->j / j->i summed) -> strictly-lower-
triangle scan in column order.  The result is that triangle in CSC form.
"""
from __future__ import annotations

import torch

from . import _lib
from . import device as D

FLAG_ISOLATED = 16


def jaccard_counts_mutual(idx_i32: torch.Tensor, n: int, k: int, flags: torch.Tensor | None = None):
    """uint8 per edge: intersection count in bits 0-6, bit 7 set when the edge is mutual."""
    D._require_cuda(idx_i32, torch.int32)
    dev = idx_i32.device
    out = torch.empty((n * k,), dtype=torch.uint8, device=dev)
    if flags is None:
        flags = D.new_flags(dev)
    with torch.cuda.device(dev):
        _lib.check(_lib.lib().gficf_cuda_jaccard_counts_mutual_dev(idx_i32.data_ptr(), n, k, 0, n, out.data_ptr(),
                                                                   flags.data_ptr(), D._stream_ptr()))
    return out, flags


def snn_lower_triangle(idx_i32: torch.Tensor, n: int, k: int):
    """(colptr int64 [n+1], rows int32 [nnz], weights float64 [nnz], flags).  Column c lists the
    vertices r > c joined to c, ascending, with the summed Jaccard weight: node1 = c, node2 = r of
    the reference's edge list.  Check flags for DUP_ID / HASH_FAIL / ISOLATED before use."""
    um, flags = jaccard_counts_mutual(idx_i32, n, k)
    dev = idx_i32.device
    cap = n * k
    L = _lib.lib()
    colptr = torch.empty((n + 1,), dtype=torch.int64, device=dev)
    rows = torch.empty((cap,), dtype=torch.int32, device=dev)
    w = torch.empty((cap,), dtype=torch.float64, device=dev)
    scratch = torch.empty((int(L.gficf_cuda_snn_scratch_bytes(n, cap)),), dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        _lib.check(L.gficf_cuda_snn_lower_dev(idx_i32.data_ptr(), n, k, um.data_ptr(), colptr.data_ptr(),
                                              rows.data_ptr(), w.data_ptr(), cap, scratch.data_ptr(),
                                              flags.data_ptr(), D._stream_ptr()))
    nnz = int(colptr[n])
    return colptr, rows[:nnz], w[:nnz], flags
