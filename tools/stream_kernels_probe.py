"""The two kernels of the multi-GPU peer gather run on ONE device (for ncu / sanitizer captures):
tagged count kernel (grouped 16-byte stores) into a local buffer, then the streaming expand."""
import sys, torch
sys.path.insert(0, '.')
from gficf_b200 import device as D, synth, sharding
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4_000_000
k = int(sys.argv[2]) if len(sys.argv) > 2 else 30
idx0 = synth.knn_index(n, k, scramble=True, device="cuda")
padded, flags = D.pad_rows(idx0)
buf = torch.zeros(n * k + 256, dtype=torch.uint8, device="cuda")
out = torch.empty((3, n * k), dtype=torch.float64, device="cuda")
segs = sharding.share_bounds(n, 8, 0.0, 0, align=16)[1:]
torch.cuda.synchronize()
torch.cuda.profiler.start()
for epoch in (1, 2):
    tag = (epoch & 1) << 7
    for lo, hi in segs:
        D.jaccard_counts_tagged_to(padded, n, k, lo, hi, buf.data_ptr() + lo * k, tag, flags)
    D.expand_stream(padded, k, segs, buf.data_ptr(), out, tag, flags, timeout_ms=2000)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
def ms(fn, reps=5):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
lo, hi = segs[3]
scr = torch.empty((hi - lo) * k, dtype=torch.uint8, device="cuda")
print("count kernel over 1/7 of the rows: tagged+grouped <3> %.3f ms, plain <1> %.3f ms" % (
    ms(lambda: D.jaccard_counts_tagged_to(padded, n, k, lo, hi, buf.data_ptr() + lo * k, 0x80, flags)),
    ms(lambda: D.jaccard_counts(padded, n, k, lo, hi, out=scr, flags=flags))))
want, _ = D.jaccard_edges(padded, n, k)
assert torch.equal(out, want) and int(flags[0]) == 0
print("STREAM_PROBE_OK")
