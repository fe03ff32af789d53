"""Standalone speed of the streaming expand (all count bytes already tagged: no polling), 4M x 30,
for 1 / 3 / 7 segments -- the host rank's floor in the peer gather at N = 2 / 4 / 8."""
import sys, torch
sys.path.insert(0, '.')
from gficf_b200 import device as D, synth, sharding
n, k = 4_000_000, 30
idx0 = synth.knn_index(n, k, scramble=True, device="cuda")
padded, flags = D.pad_rows(idx0)
cnt, _ = D.jaccard_counts(padded, n, k)
cnt |= 0x80
out = torch.empty((3, n * k), dtype=torch.float64, device="cuda")
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
def timeit(fn, reps=8):
    for _ in range(2): fn()
    ts = []
    for i in range(reps):
        flush.fill_(i)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return sum(ts) / len(ts)
res = []
for nseg in (1, 3, 7):
    segs = [b for b in sharding.share_bounds(n, nseg + 1, 0.0, 0, align=16)[1:]]
    t = timeit(lambda: D.expand_stream(padded, k, segs, cnt.data_ptr(), out, 0x80, flags))
    res.append("%d seg %.3f ms" % (nseg, t))
cnt &= 0x7F
t = timeit(lambda: D.expand(padded, k, cnt, mode=0, out=out))
print("  ".join(res), "  expand_fixed %.3f ms" % t, " flags", int(flags[0]))
