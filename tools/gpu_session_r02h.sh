#!/bin/bash
# r02 session H (N GPUs): host rank's own rows next to (instead of before) its streaming expand.
N=${1:-4}
export GFICF_CUDA_PEER_TIMEOUT_MS=5000
mkdir -p gpurun_out
run() { timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29577 bench.py --gpus $N --no-e2e --steps 20 "$@" 2>gpurun_out/hostside_n$N.err; }
show() { python -c "
import json,sys
try:
    d=json.loads(sys.stdin.read())
    print('$1', 'ms/step %.3f' % d['ms_per_step'], 'Gedges/s %.1f' % (d['value']/1e9), d['per_rank'], 'parity', (d.get('parity') or {}).get('full_matrix_equal'), (d.get('parity') or {}).get('value_path_equals_oracle'))
except Exception as ex:
    print('$1 failed', ex)
"; }
: > gpurun_out/hostside_n$N.txt
GFICF_CUDA_HOST_OWN=side:1:4 run --host-share 0.08 | show side-1-4-share0.08 | tee -a gpurun_out/hostside_n$N.txt; tail -1 gpurun_out/hostside_n$N.err | cut -c1-200
GFICF_CUDA_HOST_OWN=side:2:3 run --host-share 0.12 | show side-2-3-share0.12 | tee -a gpurun_out/hostside_n$N.txt; tail -1 gpurun_out/hostside_n$N.err | cut -c1-200
GFICF_CUDA_HOST_OWN=side:1:5 run --host-share 0.05 | show side-1-5-share0.05 | tee -a gpurun_out/hostside_n$N.txt; tail -1 gpurun_out/hostside_n$N.err | cut -c1-200
