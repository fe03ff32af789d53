#!/bin/bash
N=${1:-2}
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -x -q -m gpu --durations=8 > gpurun_out/pytest_multi.log 2>&1; echo "pytest all gpu rc=$?" | tee -a gpurun_out/summary_multi.txt
tail -16 gpurun_out/pytest_multi.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 30 --warmup 5 > gpurun_out/bench_n${N}.json 2> gpurun_out/bench_n${N}.err; echo "bench N=$N rc=$?" | tee -a gpurun_out/summary_multi.txt
python -c "
import json; l=json.load(open('gpurun_out/bench_n${N}.json')); print(l['value']/1e9, l['ms_per_step'], l['roofline']['kernel_ms'], l['gpu_launches'], l['config']['sharding'][:80], l.get('e2e',{}).get('ms_per_step'))"; grep -v "Warning\|^\*\|OMP" gpurun_out/bench_n${N}.err | tail -5
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:jaccard_wide_k -s 3 -c 1 -o gpurun_out/prof_wide_k3 -f python bench.py --cells 2000000 --k 100 --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_wide.log 2>&1; echo "ncu wide rc=$?" | tee -a gpurun_out/summary_multi.txt
