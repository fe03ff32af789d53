"""Device-resident entry points on torch CUDA tensors (torch is only the allocator / stream
provider here; every kernel is the library's own, called through the C ABI of
include/gficf_cuda.h).

Index layout on the device: int32, row-major, 0-based, ``row_stride(k)`` ints per row, pad = -2.
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib

FLAG_BAD_ID, FLAG_DUP_ID, FLAG_HASH_FAIL = 1, 2, 4


def row_stride(k: int) -> int:
    return int(_lib.lib().gficf_cuda_row_stride(int(k)))


def _stream_ptr() -> int:
    return int(torch.cuda.current_stream().cuda_stream)


def _require_cuda(t: torch.Tensor, dtype) -> None:
    if not t.is_cuda:
        raise ValueError("a CUDA tensor is required (this package has no CPU implementation)")
    if t.dtype != dtype or not t.is_contiguous():
        raise ValueError("expected a contiguous %s tensor" % dtype)


def new_flags(device) -> torch.Tensor:
    return torch.zeros(2, dtype=torch.int32, device=device)


def layout_from_r_matrix(d_idx_f64_colmajor: torch.Tensor, n: int, k: int, out: torch.Tensor | None = None,
                         flags: torch.Tensor | None = None, row_lo: int = 0, row_hi: int | None = None):
    """f64 column-major 1-based (a device copy of the R matrix, passed as a 1-D or (k, n)
    contiguous tensor) -> padded int32 rows.  Returns (idx_i32 [n, stride], flags)."""
    _require_cuda(d_idx_f64_colmajor, torch.float64)
    stride = row_stride(k)
    dev = d_idx_f64_colmajor.device
    if out is None:
        out = torch.empty((n, stride), dtype=torch.int32, device=dev)
    if flags is None:
        flags = new_flags(dev)
    hi = n if row_hi is None else row_hi
    with torch.cuda.device(dev):
        _lib.check(_lib.lib().gficf_cuda_layout_dev(d_idx_f64_colmajor.data_ptr(), n, 0, n, k, row_lo, hi,
                                                    out.data_ptr(), flags.data_ptr(), _stream_ptr()))
    return out, flags


def pad_rows(d_idx_dense_i32: torch.Tensor, out: torch.Tensor | None = None,
             flags: torch.Tensor | None = None):
    """int32 [n, k] row-major 0-based (e.g. from a GPU kNN) -> padded int32 rows."""
    _require_cuda(d_idx_dense_i32, torch.int32)
    n, k = d_idx_dense_i32.shape
    stride = row_stride(k)
    dev = d_idx_dense_i32.device
    if out is None:
        out = torch.empty((n, stride), dtype=torch.int32, device=dev)
    if flags is None:
        flags = new_flags(dev)
    with torch.cuda.device(dev):
        _lib.check(_lib.lib().gficf_cuda_pad_dev(d_idx_dense_i32.data_ptr(), n, k, 0, n, out.data_ptr(),
                                                 flags.data_ptr(), _stream_ptr()))
    return out, flags


def jaccard_edges(idx_i32: torch.Tensor, n: int, k: int, row_lo: int = 0, row_hi: int | None = None,
                  out: torch.Tensor | None = None, flags: torch.Tensor | None = None):
    """Fused fast kernel: rows [row_lo,row_hi) -> out[3, slab_edges] float64 (from, to, w)."""
    _require_cuda(idx_i32, torch.int32)
    hi = n if row_hi is None else row_hi
    e = (hi - row_lo) * k
    dev = idx_i32.device
    if out is None:
        out = torch.empty((3, e), dtype=torch.float64, device=dev)
    if flags is None:
        flags = new_flags(dev)
    with torch.cuda.device(dev):
        _lib.check(_lib.lib().gficf_cuda_jaccard_dev(idx_i32.data_ptr(), n, k, row_lo, hi, out[0].data_ptr(),
                                                     out[1].data_ptr(), out[2].data_ptr(), flags.data_ptr(),
                                                     _stream_ptr()))
    return out, flags


def jaccard_counts(idx_i32: torch.Tensor, n: int, k: int, row_lo: int = 0, row_hi: int | None = None,
                   out: torch.Tensor | None = None, flags: torch.Tensor | None = None):
    """Fast kernel, intersection counts only: uint8 [slab_edges] for k <= 255, int16-typed uint16
    for 255 < k <= 1024."""
    _require_cuda(idx_i32, torch.int32)
    hi = n if row_hi is None else row_hi
    e = (hi - row_lo) * k
    dev = idx_i32.device
    if out is None:
        out = torch.empty((e,), dtype=torch.uint8 if k <= 255 else torch.int16, device=dev)
    if flags is None:
        flags = new_flags(dev)
    with torch.cuda.device(dev):
        _lib.check(_lib.lib().gficf_cuda_jaccard_counts_dev(idx_i32.data_ptr(), n, k, row_lo, hi,
                                                            out.data_ptr(), flags.data_ptr(), _stream_ptr()))
    return out, flags


def jaccard_counts_exact(idx_i32: torch.Tensor, n: int, k: int, set_semantics: bool, row_lo: int = 0,
                         row_hi: int | None = None):
    """Exact kernel (rows may repeat ids, any k): uint8 counts (k<=255) or int16-typed uint16."""
    _require_cuda(idx_i32, torch.int32)
    hi = n if row_hi is None else row_hi
    e = (hi - row_lo) * k
    dev = idx_i32.device
    out = torch.empty((e,), dtype=torch.uint8 if k <= 255 else torch.int16, device=dev)
    with torch.cuda.device(dev):
        _lib.check(_lib.lib().gficf_cuda_jaccard_exact_dev(idx_i32.data_ptr(), n, k, row_lo, hi,
                                                           int(bool(set_semantics)), out.data_ptr(),
                                                           _stream_ptr()))
    return out


def expand(idx_i32: torch.Tensor, k: int, counts: torch.Tensor, mode: int = 0, row_lo: int = 0,
           row_hi: int | None = None, out: torch.Tensor | None = None):
    """counts -> (from, to, w) rows.  mode 0 fixed slots; mode 1 compacted (returns n_written too)."""
    _require_cuda(idx_i32, torch.int32)
    hi = idx_i32.shape[0] if row_hi is None else row_hi
    e = (hi - row_lo) * k
    dev = idx_i32.device
    if out is None:
        out = torch.empty((3, e), dtype=torch.float64, device=dev)
    L = _lib.lib()
    scratch_ptr, nw, nw_ptr = 0, None, 0
    if mode == 1:
        scratch = torch.empty((max(int(L.gficf_cuda_expand_scratch_bytes(e)), 8),), dtype=torch.uint8, device=dev)
        nw = torch.zeros(1, dtype=torch.int64, device=dev)
        scratch_ptr, nw_ptr = scratch.data_ptr(), nw.data_ptr()
    with torch.cuda.device(dev):
        _lib.check(L.gficf_cuda_expand_dev(idx_i32.data_ptr(), k, row_lo, hi, counts.data_ptr(), mode,
                                           out[0].data_ptr(), out[1].data_ptr(), out[2].data_ptr(),
                                           scratch_ptr, nw_ptr, _stream_ptr()))
    return out, nw


# ---- raw-pointer variants for the peer-memory gather (buffers allocated / mapped by the library)
def ipc_alloc(nbytes: int):
    """cudaMalloc'ed, zeroed, exportable buffer on the current device -> (ptr, 64-byte handle)."""
    p = C.c_void_p()
    h = C.create_string_buffer(64)
    _lib.check(_lib.lib().gficf_cuda_ipc_alloc(nbytes, C.byref(p), h))
    return int(p.value), h.raw


def ipc_open(handle: bytes) -> int:
    p = C.c_void_p()
    _lib.check(_lib.lib().gficf_cuda_ipc_open(C.create_string_buffer(handle, 64), C.byref(p)))
    return int(p.value)


def ipc_close(ptr: int) -> None:
    _lib.lib().gficf_cuda_ipc_close(ptr)


def ipc_free(ptr: int) -> None:
    _lib.lib().gficf_cuda_ipc_free(ptr)


def jaccard_counts_to(idx_i32: torch.Tensor, n: int, k: int, row_lo: int, row_hi: int, out_ptr: int,
                      flags: torch.Tensor) -> None:
    """Count kernel writing its u8 results at the raw device address out_ptr (may be peer memory)."""
    _require_cuda(idx_i32, torch.int32)
    with torch.cuda.device(idx_i32.device):
        _lib.check(_lib.lib().gficf_cuda_jaccard_counts_dev(idx_i32.data_ptr(), n, k, row_lo, row_hi, out_ptr,
                                                            flags.data_ptr(), _stream_ptr()))


def signal(flag_ptr: int, value: int) -> None:
    _lib.check(_lib.lib().gficf_cuda_signal_dev(flag_ptr, value & 0xFFFFFFFF, _stream_ptr()))


def wait_flag(flag_ptr: int, expected: int, flags: torch.Tensor) -> None:
    _lib.check(_lib.lib().gficf_cuda_wait_dev(flag_ptr, expected & 0xFFFFFFFF, flags.data_ptr(), _stream_ptr()))


def jaccard_counts_tagged_to(idx_i32: torch.Tensor, n: int, k: int, row_lo: int, row_hi: int, out_ptr: int,
                             tag: int, flags: torch.Tensor) -> None:
    """Count kernel storing u | tag (tag = 0x00 / 0x80: the step parity of the streaming peer gather,
    k <= 127) at the raw device address out_ptr (slab-relative; typically peer memory)."""
    _require_cuda(idx_i32, torch.int32)
    with torch.cuda.device(idx_i32.device):
        _lib.check(_lib.lib().gficf_cuda_jaccard_counts_tagged_dev(idx_i32.data_ptr(), n, k, row_lo, row_hi, out_ptr,
                                                                   tag, flags.data_ptr(), _stream_ptr()))


def expand_stream(idx_i32: torch.Tensor, k: int, segments, counts_ptr: int, out3: torch.Tensor, tag: int,
                  flags: torch.Tensor, timeout_ms: int = 0) -> None:
    """Streaming expand on the host rank: the rows of `segments` [(lo, hi), ...] (absolute rows) are
    expanded into out3[3, n*k] while peers are still storing the tagged counts at counts_ptr
    (absolute edge order); every byte is polled until its parity bit equals `tag`."""
    segs = [(int(a), int(b)) for a, b in segments if b > a]
    if not segs:
        return
    lo = (C.c_int64 * len(segs))(*[a for a, _ in segs])
    hi = (C.c_int64 * len(segs))(*[b for _, b in segs])
    with torch.cuda.device(idx_i32.device):
        _lib.check(_lib.lib().gficf_cuda_expand_stream_dev(idx_i32.data_ptr(), k, lo, hi, len(segs), counts_ptr,
                                                           out3[0].data_ptr(), out3[1].data_ptr(), out3[2].data_ptr(),
                                                           tag, int(timeout_ms), flags.data_ptr(), _stream_ptr()))


def jaccard_edges_to(idx_i32: torch.Tensor, n: int, k: int, row_lo: int, row_hi: int, from_ptr: int, to_ptr: int,
                     w_ptr: int, flags: torch.Tensor) -> None:
    """Fused kernel for rows [row_lo,row_hi) writing (from, to, w) at raw device addresses (the slab's
    first element each; may be a peer GPU's memory: the host rank's mapped output)."""
    _require_cuda(idx_i32, torch.int32)
    with torch.cuda.device(idx_i32.device):
        _lib.check(_lib.lib().gficf_cuda_jaccard_dev(idx_i32.data_ptr(), n, k, row_lo, row_hi, from_ptr, to_ptr, w_ptr,
                                                     flags.data_ptr(), _stream_ptr()))


def set_launch_cap(ctas_per_sm: int) -> None:
    """Cap the resident CTAs per SM of the persistent kernels launched from this thread (0 = none)."""
    _lib.check(_lib.lib().gficf_cuda_set_launch_ctas_per_sm(int(ctas_per_sm)))


class _DeviceBuffer:
    """A raw device allocation presented through __cuda_array_interface__ (zero-copy torch view)."""

    def __init__(self, ptr: int, shape, typestr: str):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (int(ptr), False),
                                         "version": 3, "strides": None}


def tensor_from_ptr(ptr: int, shape, dtype=torch.float64) -> torch.Tensor:
    """torch view of device memory the library allocated (e.g. gficf_cuda_ipc_alloc), current device."""
    typestr = {torch.float64: "<f8", torch.uint8: "|u1", torch.int32: "<i4"}[dtype]
    return torch.as_tensor(_DeviceBuffer(ptr, shape, typestr), device=torch.device("cuda", torch.cuda.current_device()))
