#!/bin/bash
# r02 final session, 1 GPU: smoke, GPU suite, bench (+reference arm), ncu launch list + full captures,
# L2-side counters for cfg2/cfg3, all five configs, sanitizer runs of the new kernels.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
nproc >> gpurun_out/gpu.txt; lscpu | grep -E 'Model name|Core|Socket|L3' >> gpurun_out/gpu.txt; free -g >> gpurun_out/gpu.txt
S=gpurun_out/summary_f1.txt
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" | tee $S
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $S
tail -4 gpurun_out/pytest_gpu.log
timeout 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?" | tee -a $S
cat gpurun_out/bench.json | cut -c1-1500
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "bench ref rc=$?" | tee -a $S
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-traffic-probe > gpurun_out/bench_under_ncu.log 2>&1; echo "ncu list rc=$?" | tee -a $S
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:jaccard_small_k -s 3 -c 1 -o gpurun_out/prof_small_k_r02 -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-traffic-probe --no-parity > gpurun_out/ncu_full.log 2>&1; echo "ncu full rc=$?" | tee -a $S
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"expand_stream|jaccard_small_k" -s 7 -c 9 -o gpurun_out/prof_stream_r02 -f python tools/stream_kernels_probe.py > gpurun_out/ncu_stream.log 2>&1; echo "ncu stream rc=$?" | tee -a $S
for cfg in cfg2 cfg3; do
timeout 600 ncu --metrics gpu__time_duration.sum,lts__t_bytes.sum,lts__t_sector_hit_rate.pct,dram__bytes_read.sum,dram__bytes_write.sum,lts__throughput.avg.pct_of_peak_sustained_elapsed --clock-control none --profile-from-start off -k regex:jaccard_small_k -c 3 --csv --log-file gpurun_out/l2_$cfg.csv python bench.py --traffic-probe --config $cfg > gpurun_out/l2_$cfg.log 2>&1; echo "l2 $cfg rc=$?" | tee -a $S
done
timeout 1800 python tools/all_configs.py > gpurun_out/all_configs.md 2> gpurun_out/all_configs.err; echo "all_configs rc=$?" | tee -a $S; cat gpurun_out/all_configs.md
timeout 600 python tools/bench_expand.py > gpurun_out/bench_expand.txt 2>&1; cat gpurun_out/bench_expand.txt
cat > /tmp/sanitize_case.py <<'PY'
import numpy as np, sys, torch
sys.path.insert(0, '.')
import gficf_b200
from gficf_b200 import device as D, snn, synth
from oracle.binding import Oracle
from oracle import louvain
from tests.conftest import random_knn
orc = Oracle(); rng = np.random.default_rng(0)
for n, k, distinct in ((3000, 30, True), (2000, 15, True), (600, 100, True), (300, 30, False), (200, 7, True)):
    idx = random_knn(rng, n, k, distinct=distinct)
    assert np.array_equal(gficf_b200.rcpp_parallel_jaccard_coef(idx), orc.parallel(idx)), (n, k)
    assert np.array_equal(gficf_b200.jaccard_coeff(idx), orc.serial(idx)), (n, k)
for n, k, fam in ((4000, 30, "planted"), (3000, 6, "uniform"), (900, 100, "planted")):
    idx0 = synth.knn_index(n, k, family=fam, scramble=True)
    names, cols, rows_ref, data_ref = louvain.lower_triangle_edges(orc.parallel(synth.to_r_matrix(idx0)))
    g = snn.snn_graph(synth.to_r_matrix(idx0))
    assert np.array_equal(g["row"], rows_ref) and np.array_equal(g["weight"], data_ref) and np.array_equal(g["vertex_cell"], names)
    padded, fl = D.pad_rows(idx0.cuda())
    want, _ = D.jaccard_edges(padded, n, k)
    buf = torch.zeros(n * k + 64, dtype=torch.uint8, device="cuda")
    segs = [(0, 16 * (n // 40)), (16 * (n // 40), n)]
    for lo, hi in segs:
        D.jaccard_counts_tagged_to(padded, n, k, lo, hi, buf.data_ptr() + lo * k, 0x80, fl)
    out = torch.empty((3, n * k), dtype=torch.float64, device="cuda")
    D.expand_stream(padded, k, segs, buf.data_ptr(), out, 0x80, fl, timeout_ms=2000)
    torch.cuda.synchronize()
    assert torch.equal(out, want) and int(fl[0]) == 0
print("SANITIZE_CASE_OK")
PY
for tool in memcheck racecheck; do
timeout 900 compute-sanitizer --tool $tool --print-limit 20 python /tmp/sanitize_case.py > gpurun_out/sanitizer_$tool.log 2>&1; echo "sanitizer $tool rc=$?" | tee -a $S
grep -E "ERROR SUMMARY|RACECHECK SUMMARY|SANITIZE_CASE_OK|Race reported|Invalid" gpurun_out/sanitizer_$tool.log | head -5
done
