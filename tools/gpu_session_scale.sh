#!/bin/bash
# Scaling session: the driver's bench command at N.
N=${1:-4}
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo_n$N.txt 2>&1; lscpu | grep -E "NUMA|Socket|Model name|^CPU\(s\)" >> gpurun_out/topo_n$N.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 30 --warmup 5 > gpurun_out/bench_n${N}.json 2> gpurun_out/bench_n${N}.err; echo "bench N=$N rc=$?"
python -c "
import json; l=json.load(open('gpurun_out/bench_n${N}.json')); print('4Mx30 N=$N', l['value']/1e9, l['kernel_only'], l['ms_per_step'], l['roofline']['kernel_ms'], l['config']['sharding'][:90], 'e2e', l.get('e2e',{}).get('ms_per_step'), l.get('e2e',{}).get('breakdown_ms_rank0'), l.get('e2e',{}).get('numa_binding_rank0'))"; grep -v "Warning\|^\*\|OMP" gpurun_out/bench_n${N}.err | tail -5
