#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu > gpurun_out/pytest_ab.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/pytest_ab.log
: > gpurun_out/ab.txt
for rep in 1 2; do
for so in gficf_b200/variants/*.so; do
    GFICF_CUDA_LIB=$PWD/$so timeout 600 python bench.py --no-e2e --no-cpu-baseline --steps 30 --warmup 5 2>/dev/null | python -c "
import sys,json
l=json.loads(sys.stdin.read()); print('$so', 'rep$rep', '%.3f Gedges/s  %.3f ms  frac %.3f'%(l['value']/1e9, l['ms_per_step'], l['roofline']['frac']))" | tee -a gpurun_out/ab.txt
done
done
GFICF_CUDA_LIB=$PWD/gficf_b200/variants/lib_match1.so timeout 600 python bench.py --no-e2e --no-cpu-baseline --cells 4000000 --k 15 --steps 30 --warmup 5 2>/dev/null | python -c "
import sys,json
l=json.loads(sys.stdin.read()); print('k15 match1', '%.3f Gedges/s  %.3f ms  frac %.3f'%(l['value']/1e9, l['ms_per_step'], l['roofline']['frac']))"
GFICF_CUDA_LIB=$PWD/gficf_b200/variants/lib_match0.so timeout 600 python bench.py --no-e2e --no-cpu-baseline --cells 4000000 --k 15 --steps 30 --warmup 5 2>/dev/null | python -c "
import sys,json
l=json.loads(sys.stdin.read()); print('k15 match0', '%.3f Gedges/s  %.3f ms  frac %.3f'%(l['value']/1e9, l['ms_per_step'], l['roofline']['frac']))"
