#!/bin/bash
# Multi-GPU session (gpurun --gpus N): multi-device tests, the torchrun bench at N, plus 1-GPU k=100 bench.
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo.txt 2>&1
timeout 1200 python -m pytest tests/test_gpu_multi.py -x -q -m gpu > gpurun_out/pytest_multi.log 2>&1; echo "pytest multi rc=$?" | tee -a gpurun_out/summary_multi.txt
tail -15 gpurun_out/pytest_multi.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 30 --warmup 5 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; echo "bench N=$N rc=$?" | tee -a gpurun_out/summary_multi.txt
cat gpurun_out/bench_n$N.json; tail -5 gpurun_out/bench_n$N.err
# in-process sharding e2e timing through the C ABI (pinned host buffers)
timeout 900 python - > gpurun_out/inproc_n$N.log 2>&1 <<PY
import time, numpy as np, torch, gficf_b200
from gficf_b200 import synth
n,k=4_000_000,30
idx0=synth.knn_index(n,k,scramble=True,device="cuda")
r=gficf_b200.pinned_empty((n,k)); r[...]=synth.to_r_matrix(idx0)
out=gficf_b200.pinned_empty((n*k,3))
for d in (1,$N):
    for _ in range(2): gficf_b200.rcpp_parallel_jaccard_coef(r,False,d,out=out)
    t0=time.perf_counter()
    for _ in range(5): gficf_b200.rcpp_parallel_jaccard_coef(r,False,d,out=out)
    dt=(time.perf_counter()-t0)/5
    print("n_devices=%d e2e %.2f ms  %.3f Gedges/s"%(d,dt*1e3,n*k/dt/1e9), gficf_b200.last_timings())
PY
cat gpurun_out/inproc_n$N.log
# k=100 (configs[4] shape) on one GPU, 2M cells
timeout 900 python bench.py --cells 2000000 --k 100 --steps 10 --warmup 3 --no-e2e --cpu-seconds 10 > gpurun_out/bench_k100.json 2> gpurun_out/bench_k100.err; echo "bench k100 rc=$?" | tee -a gpurun_out/summary_multi.txt
cat gpurun_out/bench_k100.json; tail -3 gpurun_out/bench_k100.err
