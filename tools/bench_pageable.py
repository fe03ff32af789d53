"""e2e of the host-buffer call on pageable memory vs the number of staging threads."""
import os, sys, time, subprocess
if len(sys.argv) > 1:
    import numpy as np
    sys.path.insert(0, '.')
    import gficf_b200
    from gficf_b200 import synth
    n, k = 4_000_000, 30
    r = synth.to_r_matrix(synth.knn_index(n, k, scramble=True, device="cuda"))
    out = np.zeros((n * k, 3), order="F")
    gficf_b200.rcpp_parallel_jaccard_coef(r, False, 1, out=out)
    t0 = time.perf_counter()
    for _ in range(3):
        gficf_b200.rcpp_parallel_jaccard_coef(r, False, 1, out=out)
    dt = (time.perf_counter() - t0) / 3
    print("threads", sys.argv[1], "%.1f ms" % (dt * 1e3), {k_: round(v, 1) for k_, v in gficf_b200.last_timings().items()})
else:
    for t in (2, 4, 6, 8, 12, 16):
        env = dict(os.environ, GFICF_CUDA_COPY_THREADS=str(t))
        subprocess.run([sys.executable, __file__, str(t)], env=env)
