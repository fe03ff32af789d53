"""Row sharding of the Jaccard path over one process per GPU (torch.distributed).

The path shards by cell-row ranges (every edge (i,j) depends only on rows i and idx[i,j]):
each rank needs the whole int32 index (replicated read-only input) and produces the edge slab
of its own rows.  Collectives, as BASELINE.json's north_star names them:

  1. broadcast of the padded int32 index from the host rank (rank 0)       [NCCL broadcast]
  2. per-rank kernel on rows [lo, hi)
  3. the slabs travel back to the host rank.  What crosses NVLink is the 1-byte intersection
     count per edge (from = i+1, to = idx+1 and w = u/(2k-u) are reconstructed on the host
     rank by the library's expand kernel), 24x fewer link bytes than the three f64 columns,
     and bit-identical by construction.                                    [NCCL all-gather]

The functions take the process group explicitly so that the same code runs under ``gloo`` on
CPU (tests: partition / assembly logic, with a stand-in compute function) and ``nccl`` on GPUs.
"""
from __future__ import annotations

import os
from typing import Callable

import torch
import torch.distributed as dist


def slab_rows(n: int, world: int) -> int:
    """Rows per rank: ceil(n / world); the last ranks may own fewer (or zero) rows."""
    return (n + world - 1) // world


def slab_bounds(n: int, world: int, rank: int) -> tuple[int, int]:
    per = slab_rows(n, world)
    lo = min(n, rank * per)
    return lo, min(n, lo + per)


def broadcast_index(idx: torch.Tensor | None, shape: tuple[int, int], device, src: int = 0, group=None):
    """Rank `src` holds the padded int32 index; every rank returns a full copy."""
    if dist.get_rank(group) == src:
        buf = idx.to(device).contiguous()
        assert tuple(buf.shape) == tuple(shape)
    else:
        buf = torch.empty(shape, dtype=torch.int32, device=device)
    dist.broadcast(buf, src=src, group=group)
    return buf


def allgather_counts(local_counts: torch.Tensor, n: int, k: int, group=None) -> torch.Tensor:
    """Equal-size all-gather of the per-rank count slabs (padded to slab_rows*k) -> counts of all
    n*k edges in row order on every rank."""
    world = dist.get_world_size(group)
    per = slab_rows(n, world) * k
    send = local_counts
    if send.numel() != per:
        send = torch.zeros(per, dtype=local_counts.dtype, device=local_counts.device)
        send[: local_counts.numel()] = local_counts
    recv = torch.empty(per * world, dtype=local_counts.dtype, device=local_counts.device)
    dist.all_gather_into_tensor(recv, send, group=group)
    return recv[: n * k]


def sharded_counts(idx_full: torch.Tensor, n: int, k: int,
                   compute_counts: Callable[[torch.Tensor, int, int, int, int], torch.Tensor],
                   group=None) -> torch.Tensor:
    """Steps 2+3: this rank's slab through `compute_counts(idx, n, k, lo, hi)`, then the all-gather."""
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    lo, hi = slab_bounds(n, world, rank)
    local = compute_counts(idx_full, n, k, lo, hi)
    return allgather_counts(local, n, k, group)


def jaccard_sharded_gpu(idx_full: torch.Tensor, n: int, k: int, group=None, expand_on: int | None = 0,
                        out: torch.Tensor | None = None):
    """The GPU path: library count kernel per slab, NCCL all-gather, expand kernel on the host
    rank (expand_on=None: on every rank).  Returns (out[3, n*k] or None, flags)."""
    from . import device as D

    rank = dist.get_rank(group)
    flags = D.new_flags(idx_full.device)

    def compute(idx, n_, k_, lo, hi):
        c, _ = D.jaccard_counts(idx, n_, k_, lo, hi, flags=flags)
        return c

    counts = sharded_counts(idx_full, n, k, compute, group)
    # the caller-visible flags: OR of the flag bits across ranks (NCCL has no bitwise reduction: MAX over unpacked bits)
    bits = torch.stack([(flags[0] >> b) & 1 for b in range(3)])
    dist.all_reduce(bits, op=dist.ReduceOp.MAX, group=group)
    flags[0] = bits[0] | (bits[1] << 1) | (bits[2] << 2)
    if expand_on is None or rank == expand_on:
        out, _ = D.expand(idx_full, k, counts, mode=0, row_lo=0, row_hi=n, out=out)
    else:
        out = None
    return out, flags


_CACHE: dict = {}


def rcpp_parallel_jaccard_coef_sharded(mat, n: int, k: int, out=None, group=None, host_rank: int = 0):
    """Host buffers in, host buffers out, one process per GPU (NCCL).  The host rank passes the
    R matrix `mat` (n x k float64, column-major, 1-based; ideally gficf_b200.pinned_empty) and
    receives the (n*k) x 3 column-major result in `out`; the other ranks pass None.

      host rank: H2D -> layout pre-pass -> NCCL broadcast of the int32 index
      all ranks: count kernel on the own row slab -> NCCL all-gather of the u8 counts
      host rank: expand kernel -> D2H
    """
    import numpy as np

    from . import device as D

    rank = dist.get_rank(group)
    dev = torch.device("cuda", torch.cuda.current_device())
    stride = D.row_stride(k)
    key = (n, k, dev.index)
    ws = _CACHE.get(key)
    if ws is None:
        _CACHE.clear()
        ws = {"idx": torch.empty((n, stride), dtype=torch.int32, device=dev), "flags": D.new_flags(dev)}
        if rank == host_rank:
            ws["r"] = torch.empty((k, n), dtype=torch.float64, device=dev)
            ws["out"] = torch.empty((3, n * k), dtype=torch.float64, device=dev)
        _CACHE[key] = ws
    ws["flags"].zero_()
    if rank == host_rank:
        a = np.asfortranarray(mat, dtype=np.float64)
        if a.shape != (n, k):
            raise ValueError("mat must be n x k")
        ws["r"].copy_(torch.from_numpy(a.T), non_blocking=True)  # (k, n) C-order == column-major n x k
        D.layout_from_r_matrix(ws["r"], n, k, out=ws["idx"], flags=ws["flags"])
    dist.broadcast(ws["idx"], src=host_rank, group=group)
    res, flags = jaccard_sharded_gpu(ws["idx"], n, k, group=group, expand_on=host_rank, out=ws.get("out"))
    bits = int(flags[0]) | int(ws["flags"][0])
    if bits & D.FLAG_BAD_ID:
        from ._lib import GficfCudaError

        raise GficfCudaError(2, "neighbour ids must be integers in [1, nrow]")
    if bits & (D.FLAG_DUP_ID | D.FLAG_HASH_FAIL):
        # rows with repeated ids: exact multiset kernel per slab, same gather
        def compute(idx, n_, k_, lo, hi):
            return D.jaccard_counts_exact(idx, n_, k_, False, lo, hi)

        counts = sharded_counts(ws["idx"], n, k, compute, group)
        if rank == host_rank:
            res, _ = D.expand(ws["idx"], k, counts, mode=0, row_lo=0, row_hi=n)
    if rank != host_rank:
        return None
    if out is None:
        out = np.empty((n * k, 3), dtype=np.float64, order="F")
    torch.from_numpy(out.T).copy_(res, non_blocking=True)
    torch.cuda.current_stream().synchronize()
    return out


# ---------------------------------------------------------------------------------------------
# Pipelined gather to the host rank.
#
# With the final (E x 3) f64 matrix required in ONE GPU's HBM, the host rank must write 24 B per
# edge of the WHOLE matrix (the expand kernel) while the other ranks only count.  So the rows are
# split unevenly -- the host rank counts fewer rows (none from ~6 ranks on) -- and each rank's slab
# is cut into chunks that are sent (1 byte per edge, NCCL send/recv over NVLink) as soon as they
# are counted, so that the host rank's expansion of chunk c overlaps everybody's counting of
# chunk c+1.
# ---------------------------------------------------------------------------------------------
def share_bounds(n: int, world: int, host_share: float, host_rank: int = 0, align: int = 1):
    """Contiguous row ranges in rank order: the host rank takes round(host_share * n) rows, the
    other ranks equal parts of the rest (the last of them takes the remainder).  Every boundary
    except n itself is a multiple of `align` (the streaming peer gather sends groups of up to 16
    consecutive rows with 16-byte vector stores: its ranges start on multiples of 16 rows)."""
    if world == 1:
        return [(0, n)]
    rows0 = int(round(min(1.0, max(0.0, host_share)) * n))
    rows0 = min(n, (rows0 + align - 1) // align * align)
    rest = n - rows0
    per = (rest + world - 2) // (world - 1)
    per = (per + align - 1) // align * align
    sizes, left = [], rest
    for r in range(world):
        if r == host_rank:
            sizes.append(rows0)
        else:
            take = min(per, left)
            sizes.append(take)
            left -= take
    assert left == 0
    out, lo = [], 0
    for take in sizes:
        out.append((lo, lo + take))
        lo += take
    assert lo == n
    return out


def weighted_bounds(n: int, world: int, rho: float, host_rank: int = 0):
    """Row ranges per rank.  rho = (time to expand a row) / (time to count a row) on one GPU.
    Balancing  x*Tc + Te = (1-x)*Tc/(world-1)  gives the host rank's share x = (1 - rho*(world-1))/world."""
    if world == 1:
        return [(0, n)]
    return share_bounds(n, world, max(0.0, (1.0 - rho * (world - 1)) / world), host_rank)


def balanced_host_share(world: int, t_fused: float, t_count: float, t_expand: float) -> float:
    """Host rank's share x of the rows for the streaming peer gather.  Per row: t_fused = fused
    kernel (the host rank's own rows), t_count = count kernel (the peers' rows), t_expand = expand
    kernel (the host rank expands every peer row).  Host: x*t_fused + (1-x)*t_expand; a peer:
    (1-x)*t_count/(world-1); equal when x = (p - t_expand) / (t_fused + p - t_expand), p = t_count/(world-1)."""
    if world <= 1:
        return 1.0
    p = t_count / (world - 1)
    return min(1.0, max(0.0, (p - t_expand) / (t_fused + p - t_expand)))


def chunk_bounds(lo: int, hi: int, chunks: int):
    per = (hi - lo + chunks - 1) // chunks if hi > lo else 0
    return [(min(hi, lo + c * per), min(hi, lo + (c + 1) * per)) for c in range(chunks)]


class PipelinedGather:
    """Counts on every rank, chunked sends to the host rank, expansion there.

    compute_counts(idx, n, k, lo, hi, out_u8_view) and expand(idx, k, counts_view, lo, hi, out3)
    default to the library kernels; tests inject stand-ins to run the same schedule over gloo."""

    def __init__(self, n: int, k: int, group=None, rho: float = 0.2, chunks: int = 4, host_rank: int = 0,
                 compute_counts=None, expand=None):
        self.n, self.k, self.group, self.host = n, k, group, host_rank
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        self.bounds = weighted_bounds(n, self.world, rho, host_rank)
        self.chunks = [chunk_bounds(lo, hi, chunks) for lo, hi in self.bounds]
        self.compute_counts = compute_counts or self._lib_counts
        self.expand = expand or self._lib_expand
        self.flags = None
        self.launches = 0  # kernels this rank launched (count + expand), for the bench record

    def _lib_counts(self, idx, n, k, lo, hi, out):
        from . import device as D

        if self.flags is None:
            self.flags = D.new_flags(idx.device)
        D.jaccard_counts(idx, n, k, lo, hi, out=out, flags=self.flags)
        self.launches += 1

    def _lib_expand(self, idx, k, counts, lo, hi, out3):
        from . import device as D

        # out3 is [3, E]; the kernel writes slab-relative, so hand it the slab's columns
        D.expand(idx, k, counts, mode=0, row_lo=lo, row_hi=hi, out=out3[:, lo * k:hi * k])
        self.launches += 1

    def step(self, idx_full, counts_all, out3):
        """counts_all: uint8 [n*k] on every rank (only the own rows are used off the host rank);
        out3: float64 [3, n*k] on the host rank (None elsewhere)."""
        k, n = self.k, self.n
        if self.flags is None and idx_full.is_cuda:
            from . import device as D

            self.flags = D.new_flags(idx_full.device)
        if self.rank != self.host:
            works = []
            for lo, hi in self.chunks[self.rank]:
                if hi > lo:
                    view = counts_all[lo * k:hi * k]
                    self.compute_counts(idx_full, n, k, lo, hi, view)
                    works += dist.batch_isend_irecv([dist.P2POp(dist.isend, view, self.host, self.group)])
            for w in works:
                w.wait()
            return
        recvs = {}
        nchunks = len(self.chunks[0])
        for c in range(nchunks):  # post every receive first; they complete as the chunks arrive
            for r in range(self.world):
                lo, hi = self.chunks[r][c]
                if r != self.host and hi > lo:
                    recvs[(r, c)] = dist.batch_isend_irecv(
                        [dist.P2POp(dist.irecv, counts_all[lo * k:hi * k], r, self.group)])[0]
        for c in range(nchunks):
            lo, hi = self.chunks[self.host][c]
            if hi > lo:
                self.compute_counts(idx_full, n, k, lo, hi, counts_all[lo * k:hi * k])
                self.expand(idx_full, k, counts_all[lo * k:hi * k], lo, hi, out3)
            for r in range(self.world):
                if (r, c) in recvs:
                    recvs[(r, c)].wait()
                    lo, hi = self.chunks[r][c]
                    self.expand(idx_full, k, counts_all[lo * k:hi * k], lo, hi, out3)


class PeerGather:
    """Counts on every rank, gather FUSED into the count kernel, streaming expansion on the host rank.

    The host rank exports its count buffer (CUDA IPC); the other ranks map it and their count
    kernels store the 1-byte results straight into the host rank's HBM over NVLink -- ONE persistent
    count launch per rank and step.  Every count byte carries the step's parity in bit 7 (k <= 127),
    so a byte is its own ready flag: the host rank's expand kernel (one launch, one sub-grid per
    contributing rank, each walking its rank's rows linearly like the count kernel that produces
    them) polls the bytes it is about to expand and otherwise never synchronises -- no flags, no
    fences, no per-chunk launches, no collective kernel competing for SMs.  The host rank computes
    its own (smaller) row share with the fused kernel straight into the output before it expands.
    The only other synchronisation is one ack flag per step: a peer may overwrite the buffer for
    step s+1 only after the host rank's expand of step s has finished.

    direct_share > 0 (the peer-store split): the host rank writes 24 B per edge of the WHOLE matrix,
    which bounds this mode from ~6 ranks on.  A peer then counts only the first (1 - direct_share) of
    its rows (3 CTAs per SM) while a second, 1-CTA-per-SM launch of the FUSED kernel on a side stream
    stores the finished doubles of its last rows straight into the host rank's output over NVLink
    (peer-mapped output, one completion flag per peer and step); the host rank expands less.  The
    output then lives in a buffer this object allocates and exports (`out3`).

    torch.distributed only carries the IPC handles at construction and the flag bits in finish()."""

    def __init__(self, n: int, k: int, group=None, host_share: float | None = None, host_rank: int = 0,
                 timeout_ms: int = 0, rho: float | None = None, chunks: int | None = None,
                 direct_share: float | None = None):
        from . import device as D

        if k > 127:
            raise ValueError("the streaming peer gather carries the step parity in bit 7 of the count byte: k <= 127")
        self.D = D
        self.n, self.k, self.group, self.host = n, k, group, host_rank
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        if host_share is None:
            # rho (legacy argument): expand / count cost ratio -> the same balance with t_fused = t_count
            host_share = balanced_host_share(self.world, 1.0, 1.0 if rho is not None else 0.85,
                                             rho if rho is not None else 0.2)
        self.host_share = host_share
        self.bounds = share_bounds(n, self.world, host_share, host_rank, align=16)
        if direct_share is None:
            direct_share = float(os.environ.get("GFICF_CUDA_PEER_DIRECT", "0") or 0.0)
        self.direct_share = min(0.9, max(0.0, direct_share)) if k <= 32 else 0.0  # the k <= 32 kernels only
        # [lo, mid): counted, expanded by the host rank; [mid, hi): finished doubles stored by the peer itself
        self.mids = []
        for r, (lo, hi) in enumerate(self.bounds):
            d = 0 if r == host_rank else int(round(self.direct_share * (hi - lo))) // 16 * 16
            self.mids.append(hi - d)
        self.timeout_ms = timeout_ms
        self.epoch = 0
        self.launches = 0
        # "stream": the host rank's expand polls the bytes while the peers count (default);
        # "wait": it first waits for a completion flag per peer (A/B switch for measurements)
        self.mode = os.environ.get("GFICF_PEER_MODE", "stream")
        # how a peer's count kernel stores into the host rank's memory: groups of 8 rows as 16-byte
        # vectors (needed from ~6 ranks on: row-by-row stores of 7 peers saturate the host rank's NVLink
        # packet rate, 0.68 instead of 0.47 ms) or row by row (the faster kernel; measured better at 2 and 4)
        st = os.environ.get("GFICF_CUDA_PEER_STORE", "auto")
        self.row_stores = st == "bytes" or (st == "auto" and self.world < 6)
        self.trace = None  # set to [] to collect (start, mid, end) CUDA events of every host-rank step
        # host rank, "side": its own rows are computed by a capped fused launch on a side stream NEXT TO the
        # streaming expand (which mostly waits for the peers at 3-5 ranks) instead of before it.
        # GFICF_CUDA_HOST_OWN = "first" (default) | "side:<fused CTAs per SM>:<expand CTAs per SM>"
        own = os.environ.get("GFICF_CUDA_HOST_OWN", "first").split(":")
        self.host_side = own[0] == "side"
        self.host_side_caps = (int(own[1]) if len(own) > 1 else 1, int(own[2]) if len(own) > 2 else 4)
        self.side_host = torch.cuda.Stream() if self.host_side and self.rank == host_rank else None
        e = n * k
        self.flag_off = (e + 255) // 256 * 256          # the ack flag (+ one done flag per rank) lives behind the counts
        nbytes = self.flag_off + 256
        self.out3 = None
        self.out_base = 0
        self.side = torch.cuda.Stream() if self.direct_share > 0 and self.rank != host_rank else None
        box = [None]
        ok, self.base = 1, 0
        if self.rank == host_rank:
            try:
                self.base, handle = D.ipc_alloc(nbytes)   # zero-filled: parity 0, the first step uses 0x80
                h_out = None
                if self.direct_share > 0:
                    self.out_base, h_out = D.ipc_alloc(3 * e * 8)
                    self.out3 = D.tensor_from_ptr(self.out_base, (3, e), torch.float64)
                box = [(handle, h_out)]
            except Exception:
                ok = 0
        dist.broadcast_object_list(box, src=host_rank, group=group)
        if self.rank != host_rank:
            try:
                if box[0] is None:
                    raise RuntimeError("no handle")
                self.base = D.ipc_open(box[0][0])
                if self.direct_share > 0:
                    self.out_base = D.ipc_open(box[0][1])
            except Exception:
                ok = 0
        self.flags = D.new_flags(torch.device("cuda", torch.cuda.current_device()))
        # every rank learns whether EVERY rank could map the buffers (no peer access between some
        # GPUs, IPC disabled in a container ...): all raise together, the caller falls back to NCCL
        agree = torch.tensor([ok], dtype=torch.int32, device=self.flags.device)
        dist.all_reduce(agree, op=dist.ReduceOp.MIN, group=group)
        if int(agree[0]) == 0:
            self._release()
            raise RuntimeError("peer-memory gather unavailable: the count buffer could not be mapped on every rank")

    def _release(self):
        D = self.D
        self.out3 = None
        for ptr in (self.base, self.out_base):
            if ptr:
                (D.ipc_free if self.rank == self.host else D.ipc_close)(ptr)
        self.base = self.out_base = 0

    def rows_of(self, rank: int) -> int:
        return self.bounds[rank][1] - self.bounds[rank][0]

    def close(self):
        torch.cuda.synchronize()
        dist.barrier(group=self.group)
        self._release()

    def step(self, idx_full, out3=None):
        """One pass: asynchronous on the current stream.  out3: float64 [3, n*k] on the host rank
        (ignored with direct_share > 0: the output is self.out3)."""
        D, k, n = self.D, self.k, self.n
        self.epoch += 1
        tag = (self.epoch & 1) << 7
        ack = self.base + self.flag_off
        lo, hi = self.bounds[self.rank]
        mid = self.mids[self.rank]
        if self.rank != self.host:
            # the host rank must have expanded the previous step's counts before they are overwritten
            D.wait_flag(ack, self.epoch - 1, self.flags)
            main = torch.cuda.current_stream()
            if hi > mid:  # the peer-store share: fused kernel on a side stream, 1 CTA per SM, next to the count kernel
                self.side.wait_stream(main)
                D.set_launch_cap(3)
            if mid > lo:
                D.jaccard_counts_tagged_to(idx_full, n, k, lo, mid, self.base + lo * k,
                                           tag | (0x100 if self.row_stores else 0), self.flags)
                self.launches += 1
            if hi > mid:
                e = n * k
                with torch.cuda.stream(self.side):
                    D.set_launch_cap(1)
                    D.jaccard_edges_to(idx_full, n, k, mid, hi, self.out_base + 8 * (mid * k),
                                       self.out_base + 8 * (e + mid * k), self.out_base + 8 * (2 * e + mid * k), self.flags)
                    D.signal(ack + 4 * (1 + self.rank), self.epoch)
                D.set_launch_cap(0)
                main.wait_stream(self.side)
                self.launches += 1
            elif self.mode == "wait":
                D.signal(ack + 4 * (1 + self.rank), self.epoch)
            return
        if self.direct_share > 0:
            out3 = self.out3
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)] if self.trace is not None else None
        if ev:
            ev[0].record()
        side = self.side_host if (self.host_side and hi > lo) else None
        if side is not None:  # own rows on a side stream, capped, next to the expand
            main = torch.cuda.current_stream()
            side.wait_stream(main)
            with torch.cuda.stream(side):
                D.set_launch_cap(self.host_side_caps[0])
                D.jaccard_edges(idx_full, n, k, lo, hi, out=out3[:, lo * k:hi * k], flags=self.flags)
            D.set_launch_cap(self.host_side_caps[1])
            self.launches += 1
        elif hi > lo:  # own rows: fused kernel straight into the output, while the peers count
            D.jaccard_edges(idx_full, n, k, lo, hi, out=out3[:, lo * k:hi * k], flags=self.flags)
            self.launches += 1
        if ev:
            ev[1].record()
        segs = [(b[0], self.mids[r]) for r, b in enumerate(self.bounds) if r != self.host and self.mids[r] > b[0]]
        if self.mode == "wait":
            for r in range(self.world):
                if r != self.host:
                    D.wait_flag(ack + 4 * (1 + r), self.epoch, self.flags)
        if segs:
            D.expand_stream(idx_full, k, segs, self.base, out3, tag, self.flags, self.timeout_ms)
            self.launches += 1
        if side is not None:
            D.set_launch_cap(0)
            torch.cuda.current_stream().wait_stream(side)
        if self.direct_share > 0 and self.mode != "wait":  # the peers' own stores must have landed
            for r in range(self.world):
                if r != self.host and self.bounds[r][1] > self.mids[r]:
                    D.wait_flag(ack + 4 * (1 + r), self.epoch, self.flags)
        D.signal(ack, self.epoch)
        if ev:
            ev[2].record()
            self.trace.append(ev)

    def finish(self, idx_full, out3=None) -> int:
        """Collective, after the last step(): agrees on the flag bits of all ranks.  A peer timeout
        raises on every rank (the output is invalid); rows with repeated ids (DUP_ID / HASH_FAIL on
        any rank invalidate fast counts everywhere) are recomputed by the host rank with the exact
        multiset kernel + expand.  Returns the OR of the flags."""
        D = self.D
        if self.direct_share > 0:
            out3 = self.out3
        torch.cuda.synchronize()
        bits = torch.stack([(self.flags[0] >> b) & 1 for b in range(4)])
        dist.all_reduce(bits, op=dist.ReduceOp.MAX, group=self.group)
        allf = int(bits[0]) | (int(bits[1]) << 1) | (int(bits[2]) << 2) | (int(bits[3]) << 3)
        if allf & 8:
            raise RuntimeError("peer-memory gather timed out (GFICF_FLAG_PEER_TIMEOUT): a rank did not deliver its "
                               "counts / its ack within the spin bound; the output is invalid")
        if allf & (D.FLAG_DUP_ID | D.FLAG_HASH_FAIL) and self.rank == self.host:
            counts = D.jaccard_counts_exact(idx_full, self.n, self.k, False)
            D.expand(idx_full, self.k, counts, mode=0, row_lo=0, row_hi=self.n, out=out3)
            torch.cuda.current_stream().synchronize()
        self.flags.zero_()
        return allf
