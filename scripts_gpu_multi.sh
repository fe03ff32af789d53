#!/bin/bash
# Multi-GPU session (gpurun --gpus N): multi-device tests, the torchrun bench at N, plus 1-GPU k=100 bench.
N=${1:-2}
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_multi.py tests/test_gpu_parity.py -x -q -m gpu > gpurun_out/pytest_multi.log 2>&1; echo "pytest multi rc=$?" | tee -a gpurun_out/summary_multi.txt
tail -15 gpurun_out/pytest_multi.log
# k=100 (configs[4] shape) on one GPU, 2M cells
timeout 900 python bench.py --cells 2000000 --k 100 --steps 10 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/bench_k100.json 2> gpurun_out/bench_k100.err; echo "bench k100 rc=$?" | tee -a gpurun_out/summary_multi.txt
cat gpurun_out/bench_k100.json; tail -3 gpurun_out/bench_k100.err
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:jaccard_wide_k -s 3 -c 1 -o gpurun_out/prof_wide_k -f python bench.py --cells 2000000 --k 100 --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_wide.log 2>&1; echo "ncu wide rc=$?" | tee -a gpurun_out/summary_multi.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1; echo "ncu list rc=$?" | tee -a gpurun_out/summary_multi.txt
