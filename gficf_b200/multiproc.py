"""One process per GPU through the library's own communicator (include/gficf_cuda.h,
"One process per GPU"): every rank handles its own row slab of the SAME host matrices, so the
H2D / D2H legs of the ranks run in parallel over their own PCIe links.

    comm_init_from_torch()                       # once per process, after init_process_group
    r   = SharedHostMatrix("idx", (n, k), create=(rank == 0))   # /dev/shm mapping, page-locked
    out = SharedHostMatrix("out", (n * k, 3), create=(rank == 0))
    rcpp_parallel_jaccard_coef_rank(r.array, out.array)          # collective

torch.distributed is used only to hand the 128-byte NCCL id from rank 0 to the others.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import _lib


def comm_init(unique_id: bytes, nranks: int, rank: int, device: int) -> None:
    err = C.create_string_buffer(512)
    buf = C.create_string_buffer(unique_id, 128)
    _lib.check(_lib.lib().gficf_cuda_comm_init_rank(buf, nranks, rank, device, err, 512), err)


def comm_unique_id() -> bytes:
    buf = C.create_string_buffer(128)
    _lib.check(_lib.lib().gficf_cuda_comm_unique_id(buf))
    return buf.raw


def comm_init_from_torch(group=None) -> None:
    import torch
    import torch.distributed as dist

    rank, world = dist.get_rank(group), dist.get_world_size(group)
    box = [comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(box, src=0, group=group)
    comm_init(box[0], world, rank, torch.cuda.current_device())


def comm_destroy() -> None:
    _lib.lib().gficf_cuda_comm_destroy()


def bind_to_gpu_numa_node(device_index: int) -> str:
    """Pin this process to the CPUs of the NUMA node its GPU hangs off (so that the pages it
    first-touches and its staging threads are local to the GPU's PCIe root).  Best effort."""
    try:
        import pynvml

        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(device_index)
        bus = pynvml.nvmlDeviceGetPciInfo(h).busId
        bus = (bus.decode() if isinstance(bus, bytes) else bus).lower()
        if len(bus.split(":")[0]) == 8:  # nvml pads the domain to 8 hex digits, sysfs uses 4
            bus = bus[4:]
        node = int(open("/sys/bus/pci/devices/%s/numa_node" % bus).read())
        if node < 0:
            return "numa node unknown"
        cpus = []
        for part in open("/sys/devices/system/node/node%d/cpulist" % node).read().strip().split(","):
            a, _, b = part.partition("-")
            cpus += list(range(int(a), int(b or a) + 1))
        os.sched_setaffinity(0, cpus)
        return "node %d (%d cpus)" % (node, len(cpus))
    except Exception as e:  # containers without sysfs / nvml: run unbound
        return "unbound (%s)" % type(e).__name__


class SharedHostMatrix:
    """A Fortran-ordered float64 matrix in /dev/shm, mapped by every rank and page-locked.

    Two-step use when NUMA placement matters: construct with pin=False on every rank, let each rank
    `first_touch_rows()` the rows it will move (the pages then live on that rank's NUMA node),
    barrier, then `pin()`."""

    def __init__(self, name: str, shape, create: bool, pin: bool = True):
        self.path = os.path.join("/dev/shm", name)
        self.shape = tuple(shape)
        self.created = create
        n = int(np.prod(self.shape))
        self.mm = np.memmap(self.path, dtype=np.float64, mode="w+" if create else "r+", shape=(max(n, 1),))
        self.array = self.mm[:n].reshape(self.shape, order="F")
        self.pinned = False
        if pin:
            self.pin()

    def pin(self) -> bool:
        n = int(np.prod(self.shape))
        if n and not self.pinned:
            self.pinned = _lib.lib().gficf_cuda_host_register(self.mm.ctypes.data, n * 8) == 0
        return self.pinned

    def first_touch_rows(self, lo: int, hi: int) -> None:
        """Write zeros to rows [lo,hi) of every column (allocates those pages on the caller's node)."""
        if hi > lo:
            self.array[lo:hi, :] = 0.0

    def close(self):
        if self.pinned:
            _lib.lib().gficf_cuda_host_unregister(self.mm.ctypes.data)
            self.pinned = False
        self.array = None
        del self.mm
        if self.created and os.path.exists(self.path):
            os.unlink(self.path)


def rcpp_parallel_jaccard_coef_rank(mat: np.ndarray, out: np.ndarray) -> np.ndarray:
    """Collective: this rank's share of rcpp_parallel_jaccard_coef on shared host matrices
    (mat: n x k float64 Fortran order, 1-based; out: (n*k) x 3 float64 Fortran order)."""
    if mat.dtype != np.float64 or not mat.flags.f_contiguous or out.dtype != np.float64 or not out.flags.f_contiguous:
        raise ValueError("Fortran-ordered float64 matrices are required")
    n, k = mat.shape
    if out.shape != (n * k, 3):
        raise ValueError("out must be (n*k, 3)")
    err = C.create_string_buffer(512)
    _lib.check(_lib.lib().gficf_cuda_jaccard_rank(mat.ctypes.data, n, k, out.ctypes.data, err, 512), err)
    return out
