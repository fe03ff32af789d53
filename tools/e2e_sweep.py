#!/usr/bin/env python
"""tools/e2e_sweep.py -- the host-buffer call (gficf_b200.rcpp_parallel_jaccard_coef) at 4M x 30 under
each output mode and a range of host thread counts; one markdown table.  One GPU.
    python tools/e2e_sweep.py [cells] [k]
"""
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

CHILD = r"""
import os, sys, time, json, numpy as np
sys.path.insert(0, %(root)r)
import torch, gficf_b200
from gficf_b200 import synth
n, k = %(n)d, %(k)d
E = n * k
idx0 = synth.knn_index(n, k, scramble=True, device="cuda")
r = synth.to_r_matrix(idx0)
del idx0
res = []
for pinned in (True, False):
    for as_int in (False, True):
        src = r.astype(np.int32) if as_int else r
        if pinned:
            buf = gficf_b200.pinned_empty((n, k), dtype=src.dtype); buf[...] = src; src = buf
            out = gficf_b200.pinned_empty((E, 3))
        else:
            src = np.asfortranarray(src); out = np.zeros((E, 3), order="F")
        for _ in range(2):
            gficf_b200.rcpp_parallel_jaccard_coef(src, False, 1, out=out)
        t0 = time.perf_counter()
        reps = 5
        for _ in range(reps):
            gficf_b200.rcpp_parallel_jaccard_coef(src, False, 1, out=out)
        dt = (time.perf_counter() - t0) / reps
        tm = gficf_b200.last_timings(); om = gficf_b200.last_output()
        res.append({"pinned": pinned, "int32": as_int, "ms": dt * 1e3, "h2d": tm["h2d_ms"], "kern": tm["jaccard_ms"],
                    "out": tm["d2h_ms"], "mode": om["mode"], "host_share": om["host_share"], "d2h_gb": om["d2h_bytes"] / 1e9})
        del src, out
print("RESULT " + json.dumps(res))
"""


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 4_000_000
    k = int(sys.argv[2]) if len(sys.argv) > 2 else 30
    print("| out mode | host threads | buffers | input | ms/call | H2D ms | kernels ms | output phase ms | resolved | host share | D2H GB |")
    print("|---|---|---|---|---|---|---|---|---|---|---|")
    hw = os.cpu_count() or 8
    combos = [("dma", 0)] + [(m, t) for m in ("hybrid", "host") for t in sorted({4, 8, 12, hw - 2, hw})]
    for mode, t in combos:
        env = dict(os.environ, GFICF_CUDA_OUT_MODE=mode)
        if t:
            env["GFICF_CUDA_EXPAND_THREADS"] = str(t)
        p = subprocess.run([sys.executable, "-c", CHILD % {"root": ROOT, "n": n, "k": k}], env=env, capture_output=True,
                           text=True, timeout=600)
        line = [ln for ln in p.stdout.splitlines() if ln.startswith("RESULT ")]
        if not line:
            print("| %s | %s | FAILED | %s |" % (mode, t, p.stderr[-300:].replace("\n", " ")))
            continue
        import json
        for r in json.loads(line[0][7:]):
            print("| %s | %s | %s | %s | %.1f | %.1f | %.1f | %.1f | %s | %.2f | %.2f |" % (
                mode, t or "-", "pinned" if r["pinned"] else "pageable", "int32" if r["int32"] else "f64", r["ms"], r["h2d"],
                r["kern"], r["out"], r["mode"], r["host_share"], r["d2h_gb"]))
        sys.stdout.flush()


if __name__ == "__main__":
    main()
