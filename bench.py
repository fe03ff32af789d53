#!/usr/bin/env python
"""bench.py -- Jaccard edges/s of the Phenograph graph-weighting path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload (BASELINE.json `metric`: "Jaccard edges/s at 4M cells k=30"): a synthetic 4M x 30
planted-cluster kNN index with scrambled cell ids (gficf_b200.synth), E = 1.2e8 edge slots.
A step is one pass of the hot path over the whole matrix.

  value     whole-job edges/s with the padded int32 index resident in HBM on every GPU;
            N=1: one launch of the fused kernel; N>1: rows sharded over the ranks (strong
            scaling: the 4M cells are fixed), per-rank count kernel, NCCL all-gather of the
            1-byte counts, expand kernel on the host rank
  e2e       the same metric through the reference-facing call with HOST buffers (pinned):
            H2D of the f64 R matrix, layout pre-pass, kernel(s), D2H of the (E x 3) doubles
  roofline  the fused/count kernel against MEASURED_PEAKS.json hbm_gbs, algorithmic bytes
            (4k+28) per edge
  cpu_baseline  the reference's own sources (oracle/_ref, unmodified, R runtime stubbed) on a
            bounded row sample with all host threads

`--impl reference` times only that CPU implementation (rank 0).
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "jaccard_edges_per_s"
UNIT = "edges/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--cells", type=int, default=4_000_000)
    ap.add_argument("--k", type=int, default=30)
    ap.add_argument("--family", default="planted", choices=["planted", "uniform"])
    ap.add_argument("--no-scramble", action="store_true")
    ap.add_argument("--seed", type=int, default=180582)
    ap.add_argument("--e2e-steps", type=int, default=0, help="0 = min(steps, 10)")
    ap.add_argument("--cpu-seconds", type=float, default=15.0, help="CPU baseline sample budget")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--chunks", type=int, default=4, help="N>1: row chunks of the pipelined gather")
    ap.add_argument("--gather", default="peer", choices=["peer", "nccl"],
                    help="N>1: how the counts reach the host rank (fused peer stores / NCCL send-recv)")
    return ap.parse_args()


def workload_name(a):
    return "%dM cells k=%d %s%s kNN index (E=%.3g edge slots)" % (
        a.cells // 1_000_000, a.k, a.family, "" if a.no_scramble else " scrambled-ids", a.cells * a.k) \
        if a.cells % 1_000_000 == 0 else "%d cells k=%d %s%s" % (a.cells, a.k, a.family,
                                                                 "" if a.no_scramble else " scrambled-ids")


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """SM clock + throttle reasons of one GPU sampled with NVML during the timed region."""

    def __init__(self, index: int):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._t = None
        try:
            import pynvml

            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = int(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
        except Exception:
            self.nv = None

    def sample_once(self):
        """One synchronous sample (NVML calls can take tens of ms: the thread alone may see few)."""
        if self.nv is None:
            return
        try:
            self.samples.append(int(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM)))
        except Exception:
            pass

    def _loop(self):
        nv = self.nv
        names = {
            getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4): "sw_power_cap",
            getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8): "hw_slowdown",
            getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
            getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
            getattr(nv, "nvmlClocksThrottleReasonHwPowerBrakeSlowdown", 0x80): "hw_power_brake_slowdown",
        }
        while not self._stop.is_set():
            try:
                self.samples.append(int(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
                r = int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
                for bit, nm in names.items():
                    if r & bit:
                        self.reasons.add(nm)
            except Exception:
                pass
            time.sleep(0.01)

    def start(self):
        if self.nv is not None:
            self._t = threading.Thread(target=self._loop, daemon=True)
            self._t.start()

    def stop(self):
        self.sample_once()  # right behind the last timed kernel
        if self._t is not None:
            self._stop.set()
            self._t.join()
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2] if s else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(s)}


# --------------------------------------------------------------------------- CPU reference
def reference_sampler(r_matrix, k, budget_s, nthreads=0):
    """Times the reference's unmodified JCoefficient worker (oracle/_ref) on rows [0,m) gathering
    from the full matrix; m calibrated so one sample costs about budget_s seconds."""
    from oracle.binding import Oracle, Reference

    if Reference.available():
        ref, kind = Reference(), "reference"
        cores = ref.hw_threads() if nthreads == 0 else nthreads

        def run(lo, hi):
            ref.parallel_rows(r_matrix, lo, hi, nthreads=nthreads)
            return ref.last_seconds
    else:
        orc, kind = Oracle(), "port"
        cores = os.cpu_count() or 1

        def run(lo, hi):
            t0 = time.perf_counter()
            orc.parallel_rows(r_matrix, lo, hi, nthreads=cores)
            return time.perf_counter() - t0

    n = r_matrix.shape[0]
    probe = min(n, 4096 * max(1, cores // 4))
    t = run(0, probe)
    rate = probe / max(t, 1e-6)
    m = int(max(probe, min(n, rate * budget_s)))
    return run, m, kind, cores


def run_reference_arm(a):
    import numpy as np
    import torch

    from gficf_b200 import synth

    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    idx0 = synth.knn_index(a.cells, a.k, family=a.family, seed=a.seed, scramble=not a.no_scramble,
                           device="cuda" if torch.cuda.is_available() else "cpu")
    r = synth.to_r_matrix(idx0)
    del idx0
    per_step = max(0.5, min(10.0, 150.0 / max(1, a.steps + a.warmup)))
    run, m, kind, cores = reference_sampler(r, a.k, per_step)
    for _ in range(a.warmup):
        run(0, m)
    times = [run(0, m) for _ in range(a.steps)]
    tot = float(np.sum(times))
    val = m * a.k * a.steps / tot
    sample = "rows [0,%d) of %d (%.3g edges/step) gathering from the full matrix" % (m, a.cells, m * a.k)
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": a.gpus, "steps": a.steps,
        "warmup": a.warmup, "ms_per_step": 1e3 * tot / a.steps, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(a), "sample": sample},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))


# --------------------------------------------------------------------------- GPU arm
def run_ours(a):
    import numpy as np
    import torch
    import torch.distributed as dist

    import gficf_b200
    from gficf_b200 import device as D
    from gficf_b200 import sharding, synth

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the product path has no CPU fallback)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # the chunk sends must get SMs while the persistent count kernels own the GPU
        os.environ.setdefault("TORCH_NCCL_HIGH_PRIORITY", "1")
        dist.init_process_group("nccl", device_id=dev)
    gficf_b200.lib()  # fail loudly now if the extension is missing

    n, k = a.cells, a.k
    E = n * k
    bytes_per_edge = 4 * k + 28
    hbm_peak, peak_src = peaks()

    # ---- inputs: every rank generates the same matrix (counter-based generator)
    idx0 = synth.knn_index(n, k, family=a.family, seed=a.seed, scramble=not a.no_scramble, device=dev)
    padded, flags = D.pad_rows(idx0)
    torch.cuda.synchronize()
    assert int(flags[0]) == 0
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2

    lo, hi = sharding.slab_bounds(n, world, rank)
    if world == 1:
        out = torch.empty((3, E), dtype=torch.float64, device=dev)

        def step():
            D.jaccard_edges(padded, n, k, out=out, flags=flags)
        launches_per_step = 1
    else:
        # calibrate the uneven row split: rho = expand time per row / count time per row (this GPU)
        m = min(n, 1_000_000)
        cal_c = torch.empty(m * k, dtype=torch.uint8, device=dev)
        cal_o = torch.empty((3, m * k), dtype=torch.float64, device=dev)
        t = []
        for fn in (lambda: D.jaccard_counts(padded, n, k, 0, m, out=cal_c, flags=flags),
                   lambda: D.expand(padded, k, cal_c, mode=0, row_lo=0, row_hi=m, out=cal_o)):
            for _ in range(2):
                fn()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(5):
                fn()
            e1.record()
            torch.cuda.synchronize()
            t.append(e0.elapsed_time(e1) / 5)
        del cal_c, cal_o
        rho_t = torch.tensor([t[1] / t[0]], dtype=torch.float64, device=dev)
        dist.broadcast(rho_t, src=0)  # every rank must use the same split
        rho = float(rho_t[0])
        out = torch.empty((3, E), dtype=torch.float64, device=dev) if rank == 0 else None
        counts_all = torch.empty(E, dtype=torch.uint8, device=dev)
        if a.gather == "peer":
            try:
                pg = sharding.PeerGather(n, k, rho=rho, chunks=a.chunks)
            except RuntimeError as ex:  # raised on every rank together
                sys.stderr.write("bench.py: %s; falling back to NCCL send/recv\n" % ex)
                a.gather = "nccl"
        if a.gather == "peer":

            def step():
                pg.step(padded, out)
        else:
            pg = sharding.PipelinedGather(n, k, rho=rho, chunks=a.chunks)

            def step():
                pg.step(padded, counts_all, out)
        pg.flags = flags
        if a.gather == "peer":  # a contiguous range of this rank's row count, for the kernel-only timing
            lo = min(pg.plan[0][rank][0], n - pg.rows_of(rank))
            hi = lo + pg.rows_of(rank)
            host_share = pg.rows_of(0) / n
        else:
            lo, hi = pg.bounds[rank]
            host_share = (pg.bounds[0][1] - pg.bounds[0][0]) / n
        launches_per_step = None  # counted by pg

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    torch.cuda.profiler.start()  # ncu --profile-from-start off: skip the input generator's launches
    for _ in range(max(3, a.warmup)):
        step()
    barrier()
    sampler = ClockSampler(local)
    ev0 = [torch.cuda.Event(enable_timing=True) for _ in range(a.steps)]
    ev1 = [torch.cuda.Event(enable_timing=True) for _ in range(a.steps)]
    kev0 = [torch.cuda.Event(enable_timing=True) for _ in range(a.steps)]
    kev1 = [torch.cuda.Event(enable_timing=True) for _ in range(a.steps)]
    sampler.start()
    t_wall0 = time.perf_counter()
    for i in range(a.steps):
        flush.fill_(i & 0xFF)  # evict the index and the previous outputs from L2 (not timed)
        if world > 1:
            dist.barrier()
        ev0[i].record()
        step()
        ev1[i].record()
    barrier()
    t_wall = time.perf_counter() - t_wall0
    clocks = sampler.stop()
    step_ms = torch.tensor([e0.elapsed_time(e1) for e0, e1 in zip(ev0, ev1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(step_ms, op=dist.ReduceOp.MAX)  # max over ranks, per step
        # the dominant kernel alone (count kernel over this rank's slab), for the roofline record
        nl = pg.launches
        kt = []
        if hi > lo:
            for _ in range(3):
                D.jaccard_counts(padded, n, k, lo, hi, out=counts_all[lo * k:hi * k], flags=flags)
            for i in range(5):
                flush.fill_(i)
                kev0[0].record()
                D.jaccard_counts(padded, n, k, lo, hi, out=counts_all[lo * k:hi * k], flags=flags)
                kev1[0].record()
                torch.cuda.synchronize()
                kt.append(kev0[0].elapsed_time(kev1[0]))
        kern_t = torch.tensor([sum(kt) / len(kt) if kt else 0.0, float(hi - lo), float(nl)], dtype=torch.float64, device=dev)
        allk = [torch.zeros_like(kern_t) for _ in range(world)]
        dist.all_gather(allk, kern_t)
        slowest = max(allk, key=lambda v: float(v[0]))
        kern_ms = slowest[:1]
        kern_rows = int(slowest[1])
        total_launches = int(sum(float(v[2]) for v in allk))
    else:
        kern_ms = step_ms
    assert int(flags[0]) == 0, "fast kernel flagged the synthetic input"
    total_ms = float(step_ms.sum())
    ms_per_step = total_ms / a.steps
    value = E / (ms_per_step * 1e-3)
    kern_avg_ms = float(kern_ms.mean())
    edges_per_launch = E if world == 1 else kern_rows * k
    bpe = bytes_per_edge if world == 1 else (4 * k + 4 + 1)  # count kernel writes 1 B/edge
    achieved = edges_per_launch * bpe / (kern_avg_ms * 1e-3) / 1e9

    # ---- end to end through the reference-facing call, host buffers (pinned)
    e2e = None
    if not a.no_e2e:
        e2e_steps = a.e2e_steps or min(a.steps, 10)
        if world == 1:
            r_host = gficf_b200.pinned_empty((n, k))
            r_host[...] = synth.to_r_matrix(idx0)
            out_host = gficf_b200.pinned_empty((E, 3))
            for _ in range(2):
                gficf_b200.rcpp_parallel_jaccard_coef(r_host, False, 1, out=out_host)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(e2e_steps):
                gficf_b200.rcpp_parallel_jaccard_coef(r_host, False, 1, out=out_host)
            dt = (time.perf_counter() - t0) / e2e_steps
            tm = gficf_b200.last_timings()
            e2e = {"value": E / dt, "unit": UNIT, "h2d_bytes_per_step": 8 * E, "d2h_bytes_per_step": 24 * E,
                   "ms_per_step": dt * 1e3, "call": "gficf_b200.rcpp_parallel_jaccard_coef (pinned host buffers)",
                   "breakdown_ms": {kk: round(v, 3) for kk, v in tm.items() if kk != "reserved"}}
            # the same call on ordinary pageable memory (what R hands over): staged through pinned slots
            r_page = np.asfortranarray(np.array(r_host))
            out_page = np.zeros((E, 3), dtype=np.float64, order="F")
            gficf_b200.rcpp_parallel_jaccard_coef(r_page, False, 1, out=out_page)
            t0 = time.perf_counter()
            for _ in range(3):
                gficf_b200.rcpp_parallel_jaccard_coef(r_page, False, 1, out=out_page)
            dtp = (time.perf_counter() - t0) / 3
            e2e["pageable"] = {"value": E / dtp, "ms_per_step": dtp * 1e3,
                               "matches_pinned": bool(np.array_equal(out_page[: 10**6], np.asarray(out_host)[: 10**6])
                                                      and np.array_equal(out_page[-10**6:], np.asarray(out_host)[-10**6:])),
                               "breakdown_ms": {kk: round(v, 3) for kk, v in gficf_b200.last_timings().items()
                                                if kk != "reserved"}}
            del r_page, out_page
            # uwot's integer matrix taken as it is (gficf_cuda_jaccard_i32): half the H2D bytes
            r_i32 = gficf_b200.pinned_empty((n, k), dtype=np.int32)
            r_i32[...] = np.asarray(r_host).astype(np.int32)
            gficf_b200.rcpp_parallel_jaccard_coef(r_i32, False, 1, out=out_host)
            t0 = time.perf_counter()
            for _ in range(3):
                gficf_b200.rcpp_parallel_jaccard_coef(r_i32, False, 1, out=out_host)
            dti = (time.perf_counter() - t0) / 3
            e2e["int32_input"] = {"value": E / dti, "ms_per_step": dti * 1e3, "h2d_bytes_per_step": 4 * E}
            del r_i32
        else:
            # one process per GPU on SHARED host matrices: every rank moves its own rows over its
            # own PCIe link (gficf_cuda_jaccard_rank); rank 0 owns / fills / checks the matrices
            from gficf_b200 import multiproc

            multiproc.comm_init_from_torch()
            tag = [("gficf_bench_%d_%d" % (os.getpid(), int(time.time()))) if rank == 0 else None]
            dist.broadcast_object_list(tag, src=0)
            # NUMA: every rank runs on the CPUs next to its GPU and first-touches the rows it will
            # move, so the D2H streams of 8 GPUs do not all land on one socket's memory
            numa = multiproc.bind_to_gpu_numa_node(local)
            numa_aware = numa.startswith("node")  # only then is first-touch placement meaningful
            r_sh = out_sh = None
            if rank == 0:
                r_sh = multiproc.SharedHostMatrix(tag[0] + "_idx", (n, k), create=True, pin=not numa_aware)
                out_sh = multiproc.SharedHostMatrix(tag[0] + "_out", (E, 3), create=True, pin=not numa_aware)
                if not numa_aware:
                    r_sh.array[...] = synth.to_r_matrix(idx0)
            dist.barrier()
            if rank != 0:
                r_sh = multiproc.SharedHostMatrix(tag[0] + "_idx", (n, k), create=False, pin=not numa_aware)
                out_sh = multiproc.SharedHostMatrix(tag[0] + "_out", (E, 3), create=False, pin=not numa_aware)
            if numa_aware:
                s_lo, s_hi = sharding.slab_bounds(n, world, rank)
                r_sh.first_touch_rows(s_lo, s_hi)
                out_sh.first_touch_rows(s_lo * k, s_hi * k)
                dist.barrier()
                if rank == 0:
                    r_sh.array[...] = synth.to_r_matrix(idx0)
                dist.barrier()
                r_sh.pin()
                out_sh.pin()
            dist.barrier()
            for _ in range(2):
                multiproc.rcpp_parallel_jaccard_coef_rank(r_sh.array, out_sh.array)
            barrier()
            t0 = time.perf_counter()
            for _ in range(e2e_steps):
                multiproc.rcpp_parallel_jaccard_coef_rank(r_sh.array, out_sh.array)
            barrier()
            dt = torch.tensor([(time.perf_counter() - t0) / e2e_steps], dtype=torch.float64, device=dev)
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
            e2e = {"value": E / float(dt[0]), "unit": UNIT, "h2d_bytes_per_step": 8 * E,
                   "d2h_bytes_per_step": 24 * E, "ms_per_step": float(dt[0]) * 1e3,
                   "call": "gficf_b200.multiproc.rcpp_parallel_jaccard_coef_rank (one rank per GPU, shared "
                           "page-locked host matrices; each rank moves its own row slab)",
                   "pinned": bool(r_sh.pinned and out_sh.pinned), "numa_binding_rank0": numa,
                   "breakdown_ms_rank0": {kk: round(v, 3) for kk, v in gficf_b200.last_timings().items()
                                          if kk != "reserved"}}
            if rank == 0:
                # the shared result must be the single-GPU result
                chk = out[:, : 3000 * k].cpu().numpy().T
                e2e["matches_device_path_on_sample"] = bool(np.array_equal(out_sh.array[: 3000 * k], chk)) and \
                    bool(np.array_equal(out_sh.array[E - 3000 * k:], out[:, E - 3000 * k:].cpu().numpy().T))
            barrier()
            r_sh.close()
            out_sh.close()
            multiproc.comm_destroy()

    # ---- the step after the path, on the device: counts+mutual -> SNN lower triangle
    snn_rec = None
    if world == 1 and k <= 127:
        try:
            from gficf_b200 import snn

            snn.snn_lower_triangle(padded, n, k)
            torch.cuda.synchronize()
            ts = []
            for i in range(3):
                flush.fill_(i)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                colptr, srows, sw, sflags = snn.snn_lower_triangle(padded, n, k)
                e1.record()
                torch.cuda.synchronize()
                ts.append(e0.elapsed_time(e1))
            snn_rec = {"ms": sum(ts) / len(ts), "nnz": int(srows.numel()), "flags": int(sflags[0]),
                       "what": "count kernel with mutual bit + CSC of the strictly lower triangle of the summed "
                               "adjacency (replaces the R filter, igraph and the triangle scan), device resident"}
            del colptr, srows, sw
        except Exception as ex:  # an extra record, never a reason to lose the bench line
            snn_rec = {"error": str(ex)[:200]}

    # ---- CPU baseline beside it (rank 0, N=1 only)
    cpu = None
    if world == 1 and not a.no_cpu_baseline:
        r = synth.to_r_matrix(idx0)
        run, m, kind, cores = reference_sampler(r, k, a.cpu_seconds)
        t = run(0, m)
        cpu = {"value": m * k / t, "unit": UNIT, "cores": cores, "kind": kind,
               "sample": "rows [0,%d) of %d (%.3g edges) gathering from the full matrix, %.1f s" % (m, n, m * k, t)}
        # and the GPU result of those rows must be what the reference computed
        ref_rows = min(m, 2000)
        from oracle.binding import Reference, Oracle
        chk = (Reference() if Reference.available() else Oracle()).parallel_rows(r, 0, ref_rows)
        got = out[:, : ref_rows * k].cpu().numpy().T
        cpu["gpu_matches_on_sample"] = bool(np.array_equal(got, chk))

    traffic = None
    try:  # ncu-measured DRAM bytes per launch of this exact kernel/workload, if a capture was committed
        tj = json.load(open(os.path.join(ROOT, "profiles", "r01_traffic.json")))
        key = ("jaccard_small_k_kernel<32,false>" if 16 < k <= 32 else "jaccard_wide_k_kernel<false>") + \
            " cells=%d k=%d" % (n, k)
        if world == 1 and key in tj:
            traffic = tj[key]["traffic_bytes"]
    except Exception:
        pass
    if rank == 0:
        g, b, s, v = (C_int32() for _ in range(4))
        gficf_b200.lib().gficf_cuda_last_launch(g, b, s, v)
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": max(3, a.warmup),
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "int32 ids / f64 weights", "data": "synthetic",
            "config": {"workload": workload_name(a), "cells": n, "k": k, "edges": E,
                       "l2": "explicit 256 MiB flush write between timed steps; index 4*n*32 B = %.0f MB > 126 MB L2"
                             % (n * 32 * 4 / 1e6),
                       "sharding": "none" if world == 1 else
                       "rows over %d ranks, host rank takes %.1f%% (expand/count cost ratio %.3f measured); resident "
                       "replicated int32 index; %d chunks per rank; u8 counts reach rank 0 %s; expand kernel on rank 0 "
                       "overlapping the next chunk" % (
                           world, 100.0 * host_share, rho, a.chunks,
                           "by peer stores from the count kernel's epilogue (CUDA IPC mapping over NVLink, flag per "
                           "chunk)" if a.gather == "peer" else "over NCCL send/recv as counted"),
                       "launch": {"grid": g.value, "block": b.value, "smem": s.value}},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s",
                         "frac": achieved / hbm_peak, "traffic": traffic,
                         "kernel": ("jaccard_small_k_kernel<%d,%s>" % (D.row_stride(k), "false" if world == 1 else "true")) if k <= 32
                         else ("jaccard_wide_k_kernel<%s>" % ("false" if world == 1 else "true")),
                         "bytes_per_edge": bpe, "edges_per_launch": edges_per_launch,
                         "kernel_ms": kern_avg_ms, "peak_source": peak_src},
            "kernel_only": None if world == 1 else {
                "value": E / (kern_avg_ms * 1e-3), "unit": UNIT, "ms": kern_avg_ms,
                "what": "all ranks counting their rows concurrently, slowest rank's kernel; no gather / expand"},
            "clocks": clocks,
            "gpu_launches": launches_per_step * a.steps if world == 1 else int(total_launches * a.steps / (a.steps + max(3, a.warmup))),
            "loop_wall_ms": t_wall * 1e3,
        }
        if snn_rec:
            line["snn_next_row"] = snn_rec
        if e2e:
            line["e2e"] = e2e
        if cpu:
            line["cpu_baseline"] = cpu
        print(json.dumps(line))
    if world > 1:
        if a.gather == "peer":
            pg.close()
        dist.destroy_process_group()


def C_int32():
    import ctypes

    return ctypes.c_int32(0)


def main():
    a = parse()
    # exactly ONE JSON line on stdout: libraries (NCCL's version banner ...) write to fd 1 too, so
    # everything else goes to stderr and only the record is written to the real stdout
    real = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    sys.stdout = real
    if a.impl == "reference":
        run_reference_arm(a)
    else:
        run_ours(a)
    real.flush()


if __name__ == "__main__":
    main()
