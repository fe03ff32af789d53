#!/bin/bash
# quick check of kernel variants: parity subset + a few shapes
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_snn.py -x -q -m gpu > gpurun_out/pytest_ab.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/pytest_ab.log
for cfg in "4000000 30" "4000000 25" "4000000 28" "4000000 12" "2000000 100"; do set -- $cfg
timeout 600 python bench.py --no-e2e --no-cpu-baseline --cells $1 --k $2 --steps 10 --warmup 3 2>/dev/null | python -c "
import sys,json
l=json.loads(sys.stdin.read()); print('n=$1 k=$2', '%.3f Gedges/s  %.3f ms  frac %.3f'%(l['value']/1e9, l['ms_per_step'], l['roofline']['frac']), (l.get('snn_next_row') or {}).get('ms'))"
done
