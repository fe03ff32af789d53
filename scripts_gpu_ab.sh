#!/bin/bash
# A/B of library build variants (gficf_b200/variants/*.so) on the headline workload + k=100 check.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu > gpurun_out/pytest_ab.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_ab.log
: > gpurun_out/ab.txt
for so in gficf_b200/variants/*.so; do
  for rep in 1 2; do
    GFICF_CUDA_LIB=$PWD/$so timeout 600 python bench.py --steps 30 --warmup 5 --no-e2e --no-cpu-baseline $AB_ARGS 2>/dev/null | python -c "
import sys,json
l=json.loads(sys.stdin.read()); print('$so', 'rep$rep', '%.3f Gedges/s  %.3f ms  frac %.3f'%(l['value']/1e9, l['ms_per_step'], l['roofline']['frac']))" | tee -a gpurun_out/ab.txt
  done
done
timeout 900 python bench.py --cells 2000000 --k 100 --steps 10 --warmup 3 --no-e2e --no-cpu-baseline 2>/dev/null | python -c "
import sys,json
l=json.loads(sys.stdin.read()); print('k100 default', '%.3f Gedges/s  %.3f ms  frac %.3f'%(l['value']/1e9, l['ms_per_step'], l['roofline']['frac']))" | tee -a gpurun_out/ab.txt
