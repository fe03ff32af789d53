#!/bin/bash
# r02 session B2 (N GPUs, default 2): multi-GPU tests, bench at N with the calibrated and two other host shares.
N=${1:-2}
export GFICF_CUDA_PEER_TIMEOUT_MS=5000
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo_n$N.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_multi.py -x -q -m gpu > gpurun_out/pytest_multi_n$N.log 2>&1; echo "pytest multi rc=$?" | tee gpurun_out/summary_b2_n$N.txt
tail -15 gpurun_out/pytest_multi_n$N.log
run() { timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29577 bench.py --gpus $N "$@"; }
run > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; echo "bench n=$N rc=$?" | tee -a gpurun_out/summary_b2_n$N.txt
cat gpurun_out/bench_n$N.json; tail -3 gpurun_out/bench_n$N.err
python - "$N" <<'PY'
import json, sys
n = sys.argv[1]
try:
    d = json.load(open("gpurun_out/bench_n%s.json" % n))
    print("N=%s value %.1f Gedges/s %.3f ms/step  kernel_only %.3f ms  e2e %.2f ms  parity %s" % (
        n, d["value"] / 1e9, d["ms_per_step"], d["kernel_only"]["ms"], d["e2e"]["ms_per_step"], d.get("parity")))
    print(d["config"]["sharding"])
except Exception as ex:
    print("failed", ex)
PY
