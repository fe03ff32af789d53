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
