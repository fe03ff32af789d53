#!/bin/bash
# One GPU-box session: smoke, GPU tests, bench (+reference arm), ncu launch list + full captures.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
nproc >> gpurun_out/gpu.txt; lscpu | grep 'Model name' >> gpurun_out/gpu.txt
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" | tee gpurun_out/summary.txt
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/summary.txt
tail -3 gpurun_out/pytest_gpu.log
timeout 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?" | tee -a gpurun_out/summary.txt
cat gpurun_out/bench.json
if [ "$1" = "full" ]; then
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "bench ref rc=$?" | tee -a gpurun_out/summary.txt
cat gpurun_out/bench_ref.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1; echo "ncu list rc=$?" | tee -a gpurun_out/summary.txt
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:jaccard_small_k -s 3 -c 1 -o gpurun_out/prof_small_k_final -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_full.log 2>&1; echo "ncu full rc=$?" | tee -a gpurun_out/summary.txt
timeout 1500 python tools/all_configs.py > gpurun_out/all_configs.md 2> gpurun_out/all_configs.err; cat gpurun_out/all_configs.md
fi
if [ "$1" = "full" ]; then
# race / memory checks of the hand-written kernels on small inputs (shared-memory hash tables, warp-synchronous code)
cat > /tmp/sanitize_case.py <<'PY'
import numpy as np, sys
sys.path.insert(0, '.')
import gficf_b200
from oracle.binding import Oracle
from tests.conftest import random_knn
orc = Oracle(); rng = np.random.default_rng(0)
for n, k, distinct in ((3000, 30, True), (2000, 15, True), (600, 100, True), (500, 64, True), (300, 30, False), (200, 7, True)):
    idx = random_knn(rng, n, k, distinct=distinct)
    assert np.array_equal(gficf_b200.rcpp_parallel_jaccard_coef(idx), orc.parallel(idx)), (n, k)
    assert np.array_equal(gficf_b200.jaccard_coeff(idx), orc.serial(idx)), (n, k)
print("SANITIZE_CASE_OK")
PY
for tool in memcheck racecheck; do
timeout 900 compute-sanitizer --tool $tool --print-limit 20 python /tmp/sanitize_case.py > gpurun_out/sanitizer_$tool.log 2>&1; echo "sanitizer $tool rc=$?" | tee -a gpurun_out/summary.txt
grep -E "ERROR SUMMARY|RACECHECK SUMMARY|SANITIZE_CASE_OK|Race reported|Invalid" gpurun_out/sanitizer_$tool.log | head -5
done
fi
