import numpy as np, sys
sys.path.insert(0, '.')
import gficf_b200
from gficf_b200 import synth
from oracle.binding import Oracle
orc=Oracle()
for (n,k) in [(100_000,30),(10_001,15),(3000,100),(250_000,30)]:
    r=synth.to_r_matrix(synth.knn_index(n,k,scramble=True))
    want=orc.parallel(r)
    nbad=0
    for trial in range(20):
        got=gficf_b200.rcpp_parallel_jaccard_coef(r,False,2)
        bad=np.nonzero((got!=want).any(axis=1))[0]
        if bad.size:
            nbad+=1
            rows=np.unique(bad//k)
            print(n,k,'trial',trial,'mismatch edge rows',bad.size,'cells',rows.size, rows[:4], rows[-4:], 'cols', (got!=want).any(axis=0))
            b=bad[0]; print(' got',got[b], 'want', want[b])
    print(n,k,'bad trials',nbad)
