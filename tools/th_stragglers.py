"""Development aid: where the expensive columns of a TH step come from -- for every step of the benchmark's TH batch, the columns that need more
than 60 residual evaluations and what the SAME columns needed in the step before (the evidence behind the launch order of
vsfm_kernels.cuh:order_bucket).  Runs on a GPU:  python tools/th_stragglers.py [ncol]"""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, bench, mpp_b200
from mpp_b200 import problems as PB
ncol = int(sys.argv[1]) if len(sys.argv) > 1 else 2097152
d = bench.shard_inputs_th(0, ncol)
p, ids = PB.build_elm_th(mpp_b200.TH, d)
p.set_column_ordering(0)
prev = None
for s in range(13):
    conv, reason, out = PB.elm_th_step(p, ids, d, 1800.0, s + 1)
    nf = p.stats()["nfuncs"].copy()
    big = np.where(nf > 60)[0]
    msg = "step %d ms %.2f nf>60: %d" % (s + 1, p.last_step_ms(), big.size)
    if prev is not None and big.size:
        msg += " prev nf of those: " + str(sorted(prev[big].tolist())[:40]) + " now: " + str(sorted(nf[big].tolist())[-8:])
    print(msg, "hist", np.bincount(np.minimum(nf, 12)).tolist(), flush=True)
    prev = nf
