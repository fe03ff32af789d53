#!/bin/bash
# r02 session A (1 GPU): smoke, the new GPU tests, e2e output-mode sweep, host write microbenchmark, bench.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
nproc >> gpurun_out/gpu.txt; lscpu | grep -E 'Model name|Thread|Core|Socket|NUMA|L3' >> gpurun_out/gpu.txt; free -g >> gpurun_out/gpu.txt
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" | tee gpurun_out/summary.txt
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu > gpurun_out/pytest_parity.log 2>&1; echo "pytest parity rc=$?" | tee -a gpurun_out/summary.txt
tail -5 gpurun_out/pytest_parity.log
g++ -O3 -std=c++17 -pthread -I gficf_b200/csrc tools/host_expand_bench.cpp gficf_b200/csrc/host_expand.cpp -o /tmp/hxb && /tmp/hxb 4000000 30 1 4 8 12 14 16 > gpurun_out/host_expand_bench.txt 2>&1
for c in 0 1 2; do echo "col $c (GB/s x3 overcounted)" >> gpurun_out/host_expand_bench.txt; HXB_COL=$c /tmp/hxb 4000000 30 8 14 16 2>&1 | grep expand >> gpurun_out/host_expand_bench.txt; done
cat gpurun_out/host_expand_bench.txt
timeout 1500 python tools/e2e_sweep.py > gpurun_out/e2e_sweep.md 2> gpurun_out/e2e_sweep.err; echo "e2e sweep rc=$?" | tee -a gpurun_out/summary.txt
cat gpurun_out/e2e_sweep.md
timeout 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?" | tee -a gpurun_out/summary.txt
cat gpurun_out/bench.json; tail -5 gpurun_out/bench.err
timeout 1500 python -m pytest tests -x -q -m gpu --deselect tests/test_gpu_parity.py > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rest rc=$?" | tee -a gpurun_out/summary.txt
tail -5 gpurun_out/pytest_gpu.log
