"""The step after the Jaccard path, on the GPU ("next" row 1): from the index
matrix straight to the graph the community detection reads -- no 24-bytes-per-edge matrix, no R
filter, no igraph.

Reference steps replaced (R/clustCells.R:66-69,81; src/RModularityOptimizer.cpp:67-83):
``relations[relations[,3]>0,]`` -> ``graph.data.frame(directed=FALSE)`` (vertices numbered by
first appearance in c(from, to)) -> ``as_adjacency_matrix(attr="weight")`` (parallel edges
i->j / j->i summed) -> strictly-lower-triangle scan in column order.  The result is that triangle
in CSC form over igraph's vertex ids, plus the vertex -> cell map.

Three entry levels, like the Jaccard path itself:
  snn_graph(mat)                      host buffers in / out (the C ABI gficf_cuda_snn_lower)
  snn_lower_triangle(idx_i32, n, k)   device-resident index
  snn_lower_triangle_sharded(...)     one process per GPU: the count kernel (with the mutual bit) is
                                      row-sharded, the bytes are all-gathered, the host rank builds
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib
from . import device as D

FLAG_ISOLATED = 16


def jaccard_counts_mutual(idx_i32: torch.Tensor, n: int, k: int, flags: torch.Tensor | None = None,
                          row_lo: int = 0, row_hi: int | None = None, out: torch.Tensor | None = None):
    """uint8 per edge of rows [row_lo,row_hi): intersection count in bits 0-6, bit 7 set when the
    edge is mutual (k <= 127)."""
    D._require_cuda(idx_i32, torch.int32)
    dev = idx_i32.device
    hi = n if row_hi is None else row_hi
    if out is None:
        out = torch.empty(((hi - row_lo) * k,), dtype=torch.uint8, device=dev)
    if flags is None:
        flags = D.new_flags(dev)
    with torch.cuda.device(dev):
        _lib.check(_lib.lib().gficf_cuda_jaccard_counts_mutual_dev(idx_i32.data_ptr(), n, k, row_lo, hi, out.data_ptr(),
                                                                   flags.data_ptr(), D._stream_ptr()))
    return out, flags


def build_from_counts(idx_i32: torch.Tensor, n: int, k: int, um: torch.Tensor, flags: torch.Tensor):
    """Graph kernels on the count+mutual bytes of ALL n rows.  Returns
    (colptr int64 [nv+1], rows int32 [nnz], weights float64 [nnz], vertex_cell int32 [nv], flags)."""
    dev = idx_i32.device
    cap = n * k
    L = _lib.lib()
    colptr = torch.empty((n + 1,), dtype=torch.int64, device=dev)
    rows = torch.empty((cap,), dtype=torch.int32, device=dev)
    w = torch.empty((cap,), dtype=torch.float64, device=dev)
    vcell = torch.empty((n,), dtype=torch.int32, device=dev)
    nv = torch.zeros((1,), dtype=torch.int64, device=dev)
    scratch = torch.empty((int(L.gficf_cuda_snn_scratch_bytes(n, cap)),), dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        _lib.check(L.gficf_cuda_snn_lower_dev(idx_i32.data_ptr(), n, k, um.data_ptr(), colptr.data_ptr(),
                                              rows.data_ptr(), w.data_ptr(), cap, vcell.data_ptr(), nv.data_ptr(),
                                              scratch.data_ptr(), flags.data_ptr(), D._stream_ptr()))
    nvert = int(nv[0])
    nnz = int(colptr[nvert])
    return colptr[:nvert + 1], rows[:nnz], w[:nnz], vcell[:nvert], flags


def snn_lower_triangle(idx_i32: torch.Tensor, n: int, k: int, with_vertex_map: bool = False):
    """(colptr, rows, weights, flags) -- or (colptr, rows, weights, vertex_cell, flags).  Column c
    lists the vertices r > c joined to c, ascending, with the summed Jaccard weight: node1 = c,
    node2 = r of the reference's edge list.  Check flags for DUP_ID / HASH_FAIL before use;
    ISOLATED is informational (some cell keeps no edge of its own: vertex ids != cell ids)."""
    um, flags = jaccard_counts_mutual(idx_i32, n, k)
    colptr, rows, w, vcell, flags = build_from_counts(idx_i32, n, k, um, flags)
    if with_vertex_map:
        return colptr, rows, w, vcell, flags
    return colptr, rows, w, flags


def snn_lower_triangle_sharded(idx_full: torch.Tensor, n: int, k: int, group=None, host_rank: int = 0):
    """One process per GPU: every rank counts (with the mutual bit) the rows of its slab, the
    bytes are all-gathered over NCCL (1 byte per edge slot), the host rank runs the graph kernels.
    Returns (colptr, rows, weights, vertex_cell, flags) on the host rank, None elsewhere."""
    import torch.distributed as dist

    from . import sharding

    world, rank = dist.get_world_size(group), dist.get_rank(group)
    lo, hi = sharding.slab_bounds(n, world, rank)
    flags = D.new_flags(idx_full.device)
    local, _ = jaccard_counts_mutual(idx_full, n, k, flags, lo, hi)
    um = sharding.allgather_counts(local, n, k, group)
    bits = torch.stack([(flags[0] >> b) & 1 for b in range(3)])
    dist.all_reduce(bits, op=dist.ReduceOp.MAX, group=group)  # a repeated id anywhere invalidates counts everywhere
    flags[0] = bits[0] | (bits[1] << 1) | (bits[2] << 2)
    if rank != host_rank:
        return None
    return build_from_counts(idx_full, n, k, um, flags)


def snn_graph(mat, out=None):
    """Host buffers in, host buffers out (C ABI gficf_cuda_snn_lower): the n x k kNN matrix (1-based
    ids, float64 or int32 like rcpp_parallel_jaccard_coef takes it) -> dict with colptr int64
    [nv+1], row int32 [nnz], weight float64 [nnz], vertex_cell int32 [nv] (1-based cell of every
    vertex).  `out` may pass preallocated (e.g. page-locked) arrays under the same keys sized
    n+1 / n*k / n*k / n."""
    from .api import _as_numeric_matrix

    a = _as_numeric_matrix(mat)
    n, k = a.shape
    cap = n * k
    out = out or {}
    colptr = out.get("colptr") if out.get("colptr") is not None else np.empty(n + 1, dtype=np.int64)
    row = out.get("row") if out.get("row") is not None else np.empty(cap, dtype=np.int32)
    w = out.get("weight") if out.get("weight") is not None else np.empty(cap, dtype=np.float64)
    vc = out.get("vertex_cell") if out.get("vertex_cell") is not None else np.empty(n, dtype=np.int32)
    nv, nnz = C.c_int64(0), C.c_int64(0)
    err = C.create_string_buffer(512)
    rc = _lib.lib().gficf_cuda_snn_lower(a.ctypes.data, 4 if a.dtype == np.int32 else 8, n, k, colptr.ctypes.data,
                                         row.ctypes.data, w.ctypes.data, row.shape[0], vc.ctypes.data,
                                         C.byref(nv), C.byref(nnz), err, 512)
    _lib.check(rc, err)
    return {"colptr": colptr[:nv.value + 1], "row": row[:nnz.value], "weight": w[:nnz.value],
            "vertex_cell": vc[:nv.value], "n_vertices": nv.value, "nnz": nnz.value}
