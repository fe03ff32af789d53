#!/bin/bash
# r02 final session, N GPUs: multi-GPU tests, the full bench record at N (and cfg5 when asked).
N=${1:-2}
export GFICF_CUDA_PEER_TIMEOUT_MS=5000
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo_n$N.txt 2>&1; nproc >> gpurun_out/topo_n$N.txt; free -g >> gpurun_out/topo_n$N.txt
if [ "$2" != "notests" ]; then
timeout 900 python -m pytest tests/test_gpu_multi.py -x -q -m gpu > gpurun_out/pytest_multi_n$N.log 2>&1; echo "pytest multi rc=$?" | tee gpurun_out/summary_f_n$N.txt
tail -3 gpurun_out/pytest_multi_n$N.log
fi
run() { timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29577 bench.py --gpus $N "$@"; }
show() { python -c "
import json,sys
try:
    d=json.loads(sys.stdin.read())
    print('$1', 'ms/step %.3f' % d['ms_per_step'], 'Gedges/s %.1f' % (d['value']/1e9), 'kernel_only %.3f' % d['kernel_only']['ms'], d['per_rank'], 'e2e', d.get('e2e',{}).get('ms_per_step'), 'parity', d.get('parity'), 'clocks', d['clocks'])
except Exception as ex:
    print('$1 failed', ex)
"; }
run > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; echo "bench n=$N rc=$?" | tee -a gpurun_out/summary_f_n$N.txt
show cfg4 < gpurun_out/bench_n$N.json | tee gpurun_out/scale_n$N.txt; tail -2 gpurun_out/bench_n$N.err
if [ "$3" = "cfg5" ]; then
run --config cfg5 --no-e2e --steps 10 > gpurun_out/bench_cfg5_n$N.json 2> gpurun_out/bench_cfg5_n$N.err; echo "bench cfg5 n=$N rc=$?" | tee -a gpurun_out/summary_f_n$N.txt
show cfg5 < gpurun_out/bench_cfg5_n$N.json | tee -a gpurun_out/scale_n$N.txt; tail -2 gpurun_out/bench_cfg5_n$N.err
fi
