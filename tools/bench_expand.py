"""Micro-benchmark of the expand / count kernels (GB/s), 4M x 30."""
import sys, torch
sys.path.insert(0, '.')
from gficf_b200 import device as D, synth
n, k = 4_000_000, 30
idx0 = synth.knn_index(n, k, scramble=True, device="cuda")
padded, flags = D.pad_rows(idx0)
cnt, _ = D.jaccard_counts(padded, n, k)
out = torch.empty((3, n * k), dtype=torch.float64, device="cuda")
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
def timeit(fn, reps=10):
    for _ in range(3): fn()
    ts = []
    for i in range(reps):
        flush.fill_(i)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return sum(ts) / len(ts)
t = timeit(lambda: D.expand(padded, k, cnt, mode=0, out=out))
print("expand fixed   %.3f ms  %.0f GB/s (29 B/edge)" % (t, n * k * 29 / t / 1e6))
t = timeit(lambda: D.jaccard_counts(padded, n, k, out=cnt, flags=flags))
print("count kernel   %.3f ms  %.2f Gedges/s" % (t, n * k / t / 1e6))
t = timeit(lambda: D.jaccard_edges(padded, n, k, out=out, flags=flags))
print("fused kernel   %.3f ms  %.2f Gedges/s" % (t, n * k / t / 1e6))
t = timeit(lambda: D.expand(padded, k, cnt, mode=1))
print("expand compact %.3f ms" % t)
t = timeit(lambda: out.fill_(1.0))
print("torch fill 2.88 GB  %.3f ms  %.0f GB/s" % (t, out.numel() * 8 / t / 1e6))
o2 = torch.empty_like(out)
t = timeit(lambda: o2.copy_(out))
print("torch copy 2.88 GB  %.3f ms  %.0f GB/s (r+w)" % (t, 2 * out.numel() * 8 / t / 1e6))
t = timeit(lambda: out[0].fill_(1.0))
print("torch fill 0.96 GB  %.3f ms  %.0f GB/s" % (t, out[0].numel() * 8 / t / 1e6))
