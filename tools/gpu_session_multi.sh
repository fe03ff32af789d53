#!/bin/bash
# Multi-GPU session (gpurun --gpus N): multi-device tests + the driver's bench command at N.
N=${1:-2}
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_multi.py -x -q -m gpu > gpurun_out/pytest_multi.log 2>&1; echo "pytest multi rc=$?" | tee -a gpurun_out/summary_multi.txt
tail -4 gpurun_out/pytest_multi.log
for ch in 4; do
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 30 --warmup 5 --no-e2e --chunks $ch > gpurun_out/bench_n${N}_c$ch.json 2> gpurun_out/bench_n${N}.err; echo "bench N=$N chunks=$ch rc=$?" | tee -a gpurun_out/summary_multi.txt
python -c "
import json; l=json.load(open('gpurun_out/bench_n${N}_c$ch.json')); print('chunks $ch:', l['value']/1e9, l['ms_per_step'], l['kernel_only']['value']/1e9, l['gpu_launches'])"; grep -v "Warning\|^\*\|OMP" gpurun_out/bench_n${N}.err | tail -3
done
