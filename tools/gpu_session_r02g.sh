#!/bin/bash
# r02 session G (N GPUs): the peer-store split of the streaming gather, value path only.
N=${1:-2}
export GFICF_CUDA_PEER_TIMEOUT_MS=5000
mkdir -p gpurun_out
if [ "$2" = "tests" ]; then
timeout 900 python -m pytest tests/test_gpu_multi.py -x -q -m gpu > gpurun_out/pytest_multi_n$N.log 2>&1; echo "pytest multi rc=$?"
tail -5 gpurun_out/pytest_multi_n$N.log
fi
run() { timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29577 bench.py --gpus $N --no-e2e --steps 20 "$@" 2>gpurun_out/direct_n$N.err; }
show() { python -c "
import json,sys
try:
    d=json.loads(sys.stdin.read())
    print('$1', 'ms/step %.3f' % d['ms_per_step'], 'Gedges/s %.1f' % (d['value']/1e9), d['per_rank'], 'parity', (d.get('parity') or {}).get('full_matrix_equal'), (d.get('parity') or {}).get('value_path_equals_oracle'))
except Exception as ex:
    print('$1 failed', ex)
"; }
: > gpurun_out/direct_n$N.txt
for ds in $3; do
run --direct-share $ds | show direct-$ds | tee -a gpurun_out/direct_n$N.txt
tail -2 gpurun_out/direct_n$N.err | cut -c1-300
done
