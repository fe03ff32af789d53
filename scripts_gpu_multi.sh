#!/bin/bash
# Multi-GPU session (gpurun --gpus N): multi-device tests, the torchrun bench at N.
N=${1:-2}
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_multi.py tests/test_rpkg_sources.py -x -q -m gpu > gpurun_out/pytest_multi.log 2>&1; echo "pytest multi rc=$?" | tee -a gpurun_out/summary_multi.txt
tail -15 gpurun_out/pytest_multi.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 30 --warmup 5 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; echo "bench N=$N rc=$?" | tee -a gpurun_out/summary_multi.txt
cat gpurun_out/bench_n$N.json; tail -5 gpurun_out/bench_n$N.err
