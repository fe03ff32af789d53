#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_parity.py -x -q -m gpu --durations=5 > gpurun_out/pytest_ab.log 2>&1; echo "pytest rc=$?"; tail -12 gpurun_out/pytest_ab.log
for kk in 200 500; do
timeout 600 python bench.py --no-e2e --no-cpu-baseline --cells 500000 --k $kk --steps 5 --warmup 3 2>gpurun_out/bench_k$kk.err | python -c "
import sys,json
l=json.loads(sys.stdin.read()); print('k=$kk', '%.3f Gedges/s  %.3f ms  frac %.3f'%(l['value']/1e9, l['ms_per_step'], l['roofline']['frac']))"
done
