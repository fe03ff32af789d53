#!/bin/bash
# r02 session D (N GPUs): grouped 16-byte remote stores against row-by-row byte stores, value path only.
N=${1:-8}
export GFICF_CUDA_PEER_TIMEOUT_MS=5000
mkdir -p gpurun_out
run() { timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29577 bench.py --gpus $N --no-e2e --no-parity --steps 20 "$@" 2>/dev/null; }
show() { python -c "
import json,sys
try:
    d=json.loads(sys.stdin.read())
    print('$1', 'ms/step %.3f' % d['ms_per_step'], 'Gedges/s %.1f' % (d['value']/1e9), 'kernel_only %.3f' % d['kernel_only']['ms'], d['per_rank'])
except Exception as ex:
    print('$1 failed', ex)
"; }
GFICF_CUDA_PEER_STORE=grouped run | show grouped | tee gpurun_out/peer_store_n$N.txt
GFICF_CUDA_PEER_STORE=bytes run | show bytes | tee -a gpurun_out/peer_store_n$N.txt
GFICF_PEER_MODE=wait run | show grouped-wait | tee -a gpurun_out/peer_store_n$N.txt
