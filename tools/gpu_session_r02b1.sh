#!/bin/bash
# r02 session B1 (1 GPU): the whole GPU suite, default-mode e2e table, bench.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee gpurun_out/summary_b1.txt
tail -15 gpurun_out/pytest_gpu.log
g++ -O3 -std=c++17 -pthread -I gficf_b200/csrc tools/host_expand_bench.cpp gficf_b200/csrc/host_expand.cpp -o /tmp/hxb && /tmp/hxb 4000000 30 1 8 14 16 > gpurun_out/host_expand_bench.txt 2>&1
for c in 0 1 2; do echo "col $c (GB/s x3 overcounted)" >> gpurun_out/host_expand_bench.txt; HXB_COL=$c /tmp/hxb 4000000 30 8 14 2>&1 | grep expand >> gpurun_out/host_expand_bench.txt; done
cat gpurun_out/host_expand_bench.txt
timeout 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?" | tee -a gpurun_out/summary_b1.txt
cat gpurun_out/bench.json; tail -5 gpurun_out/bench.err
GFICF_CUDA_H2D_NARROW=0 timeout 600 python bench.py --steps 5 --no-cpu-baseline --no-traffic-probe > gpurun_out/bench_nonarrow.json 2> gpurun_out/bench_nonarrow.err
python - <<'PY'
import json
for f in ("bench", "bench_nonarrow"):
    try:
        e = json.load(open("gpurun_out/%s.json" % f))["e2e"]
        print(f, "e2e ms", round(e["ms_per_step"], 2), e["breakdown_ms"], "pageable", round(e["pageable"]["ms_per_step"], 2), "int32", round(e["int32_input"]["ms_per_step"], 2))
    except Exception as ex:
        print(f, "failed", ex)
PY
