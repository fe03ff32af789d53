#!/bin/bash
# r02 session B3 (N GPUs): where does the streaming gather's step time go?  A/B of the polling expand
# against a wait-then-expand schedule, host shares around the calibrated one.
N=${1:-2}
export GFICF_CUDA_PEER_TIMEOUT_MS=5000
mkdir -p gpurun_out
run() { timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29577 bench.py --gpus $N --no-e2e --no-parity --steps 20 "$@" 2>/dev/null; }
show() { python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('$1', 'ms/step %.3f' % d['ms_per_step'], 'kernel_only %.3f' % d['kernel_only']['ms'], d['per_rank'], d['config']['sharding'][:60])
"; }
run | show stream-calibrated | tee gpurun_out/peer_ab_n$N.txt
GFICF_PEER_MODE=wait run | show wait-calibrated | tee -a gpurun_out/peer_ab_n$N.txt
for s in 0.0 0.2 0.3 0.45; do
run --host-share $s | show stream-share-$s | tee -a gpurun_out/peer_ab_n$N.txt
done
GFICF_PEER_MODE=wait run --host-share 0.3 | show wait-share-0.3 | tee -a gpurun_out/peer_ab_n$N.txt
run --gather nccl | show nccl-pipelined | tee -a gpurun_out/peer_ab_n$N.txt
