"""CPU, world_size 2 over gloo: the row partition + collectives of gficf_b200.sharding
(broadcast of the index, per-rank slab, all-gather of the counts) reassemble exactly the
whole-matrix result.  The per-slab compute is a stand-in (the oracle) because the product has
no CPU implementation; on the GPU box the same functions run the library kernels over NCCL
(tests/test_gpu_multi.py)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from gficf_b200 import sharding, synth


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n, k, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle.binding import Oracle

        orc = Oracle()
        idx0 = synth.knn_index(n, k, seed=11) if rank == 0 else None
        full = sharding.broadcast_index(idx0, (n, k), "cpu", src=0)

        def compute(idx, n_, k_, lo, hi):
            slab = orc.parallel_rows(synth.to_r_matrix(idx), lo, hi, nthreads=1)
            w = slab[:, 2]
            lut = np.array([u / (2.0 * k_ - u) for u in range(k_ + 1)])
            u = np.searchsorted(lut, w)  # the LUT is strictly increasing
            assert np.array_equal(lut[u], w)
            return torch.from_numpy(u.astype(np.uint8))

        counts = sharding.sharded_counts(full, n, k, compute)
        q.put((rank, sharding.slab_bounds(n, world, rank), counts.numpy().copy(), full.numpy().copy()))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n,k", [(1001, 15), (64, 30)])
def test_two_rank_partition_and_allgather(oracle, n, k):
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n, k, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = [q.get(timeout=180) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    idx = synth.knn_index(n, k, seed=11)
    whole = oracle.parallel(synth.to_r_matrix(idx))
    lut = np.array([u / (2.0 * k - u) for u in range(k + 1)])
    bounds = sorted(b for _, b, _, _ in got)
    assert bounds[0][0] == 0 and bounds[-1][1] == n and bounds[0][1] == bounds[1][0]
    for rank, _, counts, full in got:
        assert np.array_equal(full, idx.numpy())  # broadcast delivered the index
        assert counts.shape == (n * k,)
        assert np.array_equal(lut[counts], whole[:, 2])  # every rank holds all edges, in row order


def test_slab_bounds_cover_rows_exactly():
    for n in (1, 7, 64, 1000, 4_000_000):
        for world in (1, 2, 3, 4, 8):
            spans = [sharding.slab_bounds(n, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            assert all(hi - lo <= sharding.slab_rows(n, world) for lo, hi in spans)


def _pipe_worker(rank, world, port, n, k, rho, host, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle.binding import Oracle

        orc = Oracle()
        idx = synth.knn_index(n, k, seed=3)
        r = synth.to_r_matrix(idx)
        lut = np.array([u / (2.0 * k - u) for u in range(k + 1)])

        def compute(idx_, n_, k_, lo, hi, out):
            w = orc.parallel_rows(r, lo, hi, nthreads=1)[:, 2]
            out.copy_(torch.from_numpy(np.searchsorted(lut, w).astype(np.uint8)))

        def expand(idx_, k_, counts, lo, hi, out3):
            u = counts.numpy().astype(np.int64)
            rows = np.repeat(np.arange(lo, hi), k_) + 1.0
            nz = u > 0
            out3[0, lo * k_:hi * k_] = torch.from_numpy(np.where(nz, rows, 0.0))
            out3[1, lo * k_:hi * k_] = torch.from_numpy(np.where(nz, idx_[lo:hi].numpy().reshape(-1) + 1.0, 0.0))
            out3[2, lo * k_:hi * k_] = torch.from_numpy(lut[u])

        pg = sharding.PipelinedGather(n, k, rho=rho, chunks=3, host_rank=host, compute_counts=compute, expand=expand)
        counts = torch.zeros(n * k, dtype=torch.uint8)
        out3 = torch.full((3, n * k), -1.0, dtype=torch.float64) if rank == host else None
        pg.step(idx, counts, out3)
        pg.step(idx, counts, out3)  # the schedule is re-entrant
        q.put((rank, pg.bounds, out3.numpy().copy() if rank == host else None))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,rho,host", [(2, 0.2, 0), (3, 0.6, 0), (2, 0.0, 0), (3, 0.3, 2), (2, 0.2, 1)])
def test_pipelined_gather_schedule(oracle, world, rho, host):
    """Uneven row split + chunked send/recv + host-rank expansion reproduce the whole matrix."""
    n, k = 999, 15
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_pipe_worker, args=(r, world, port, n, k, rho, host, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = [q.get(timeout=180) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    want = oracle.parallel(synth.to_r_matrix(synth.knn_index(n, k, seed=3)))
    for rank, bounds, out in got:
        assert bounds[0][0] == 0 and bounds[-1][1] == n and all(a[1] == b[0] for a, b in zip(bounds, bounds[1:]))
        if rank == host:
            assert np.array_equal(out.T, want)
        else:
            assert out is None
    if rho >= 0.5 and world == 3:
        assert got[0][1][host] == (0, 0) or got[0][1][host][0] == got[0][1][host][1]  # the host rank only expands


def test_weighted_bounds():
    for n in (10, 1000, 4_000_000):
        for world in (1, 2, 4, 8):
            for rho in (0.0, 0.19, 0.5, 2.0):
                b = sharding.weighted_bounds(n, world, rho)
                assert len(b) == world and b[0][0] == 0 and b[-1][1] == n
                assert all(x[1] == y[0] for x, y in zip(b, b[1:])) and all(hi >= lo for lo, hi in b)
    # rho = 0: an even split
    assert [hi - lo for lo, hi in sharding.weighted_bounds(1000, 4, 0.0)] == [250, 250, 250, 250]
