#!/bin/bash
# r02 session C (N GPUs): multi-GPU tests, the full bench record at N, host-share variants, optionally cfg5.
N=${1:-4}
export GFICF_CUDA_PEER_TIMEOUT_MS=5000
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo_n$N.txt 2>&1; nproc >> gpurun_out/topo_n$N.txt; free -g >> gpurun_out/topo_n$N.txt
timeout 900 python -m pytest tests/test_gpu_multi.py -x -q -m gpu > gpurun_out/pytest_multi_n$N.log 2>&1; echo "pytest multi rc=$?" | tee gpurun_out/summary_c_n$N.txt
tail -5 gpurun_out/pytest_multi_n$N.log
run() { timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29577 bench.py --gpus $N "$@"; }
show() { python -c "
import json,sys
try:
    d=json.loads(sys.stdin.read())
    print('$1', 'ms/step %.3f' % d['ms_per_step'], 'Gedges/s %.1f' % (d['value']/1e9), 'kernel_only %.3f' % d['kernel_only']['ms'], d['per_rank'], 'e2e', d.get('e2e',{}).get('ms_per_step'), 'parity', d.get('parity'))
except Exception as ex:
    print('$1 failed', ex)
"; }
run > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; echo "bench n=$N rc=$?" | tee -a gpurun_out/summary_c_n$N.txt
show full < gpurun_out/bench_n$N.json | tee gpurun_out/scale_n$N.txt; tail -2 gpurun_out/bench_n$N.err
for s in $2; do
run --no-e2e --no-parity --steps 20 --host-share $s 2>/dev/null | show share-$s | tee -a gpurun_out/scale_n$N.txt
done
if [ "$N" != "8" ]; then GFICF_PEER_MODE=wait run --no-e2e --no-parity --steps 20 2>/dev/null | show wait-mode | tee -a gpurun_out/scale_n$N.txt; fi
if [ "$3" = "cfg5" ]; then
run --config cfg5 --no-e2e --steps 10 > gpurun_out/bench_cfg5_n$N.json 2> gpurun_out/bench_cfg5_n$N.err; echo "bench cfg5 n=$N rc=$?" | tee -a gpurun_out/summary_c_n$N.txt
show cfg5 < gpurun_out/bench_cfg5_n$N.json | tee -a gpurun_out/scale_n$N.txt; tail -2 gpurun_out/bench_cfg5_n$N.err
fi
