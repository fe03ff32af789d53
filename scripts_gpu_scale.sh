#!/bin/bash
# Scaling session: bench at N (and optionally cfg5 10M x 100 at N).
N=${1:-4}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_multi.py -x -q -m gpu > gpurun_out/pytest_multi_n$N.log 2>&1; echo "pytest multi rc=$?"; tail -3 gpurun_out/pytest_multi_n$N.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 30 --warmup 5 > gpurun_out/bench_n${N}.json 2> gpurun_out/bench_n${N}.err; echo "bench N=$N rc=$?"
python -c "
import json; l=json.load(open('gpurun_out/bench_n${N}.json')); print('4Mx30 N=$N', l['value']/1e9, l['kernel_only'], l['ms_per_step'], l['roofline']['kernel_ms'], l['config']['sharding'][:90], 'e2e', l.get('e2e',{}).get('ms_per_step'), l.get('e2e',{}).get('breakdown_ms_rank0'))"; grep -v "Warning\|^\*\|OMP" gpurun_out/bench_n${N}.err | tail -5
