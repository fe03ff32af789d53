"""Mann-Whitney U per gene (next row 4): the GPU call through the C ABI against the reference's own
sources (oracle/_ref/libgficf_ref_wmu.so) on the host cores, same matrices, results compared.
    python tools/wmu_bench.py [genes] [cluster_cells] [other_cells]
"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gficf_b200
from oracle.binding import WmuOracle, WmuReference

genes = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
n1 = int(sys.argv[2]) if len(sys.argv) > 2 else 5000
n2 = int(sys.argv[3]) if len(sys.argv) > 3 else 95000
rng = np.random.default_rng(1)
N = n1 + n2
# CPM-like: ~10 % non-zero, library-size scaled (distinct values), one giant zero tie group per gene
lib = rng.uniform(0.5, 2.0, size=N)
m = np.asfortranarray((rng.poisson(0.15, size=(genes, N)) * (1e6 / 5000.0) / lib[None, :]).astype(np.float64))
x, y = np.asfortranarray(m[:, :n1]), np.asfortranarray(m[:, n1:])
del m
for _ in range(2):
    got = gficf_b200.rcpp_parallel_WMU_test(x, y)
t0 = time.perf_counter()
reps = 3
for _ in range(reps):
    got = gficf_b200.rcpp_parallel_WMU_test(x, y)
dt = (time.perf_counter() - t0) / reps
tm = gficf_b200.last_timings()
ref = WmuReference() if WmuReference.available() else None
sub = min(genes, 256)
if ref is not None:
    want = ref.wmu(x[:sub], y[:sub]); cpu_s = ref.last_seconds; kind = "reference sources"
else:
    t0 = time.perf_counter(); want = WmuOracle().wmu(x[:sub], y[:sub]); cpu_s = time.perf_counter() - t0; kind = "oracle port"
ok = np.array_equal(got[:sub], want, equal_nan=True)
vals = genes * N
print("| genes x (n1 + n2) | values | GPU call ms (host buffers) | H2D ms | kernels ms | Gvalues/s (call) | CPU %s, %d threads, %d genes: s | CPU Mvalues/s | speed-up (call) | GPU == CPU on those genes |" % (kind, os.cpu_count(), sub))
print("|---|---|---|---|---|---|---|---|---|---|")
print("| %d x (%d + %d) | %.3g | %.1f | %.1f | %.1f | %.2f | %.2f | %.1f | %.0fx | %s |" % (
    genes, n1, n2, vals, dt * 1e3, tm["h2d_ms"], tm["jaccard_ms"], vals / dt / 1e9, cpu_s, sub * N / cpu_s / 1e6,
    (cpu_s / sub) / (dt / genes), ok))
