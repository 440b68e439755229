"""Quick device-resident timing of the VSFM step kernel (development aid; not the headline bench)."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "tests"))
import problems as PB, bench
import mpp_b200
ncol = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
d = bench.shard_inputs(0, ncol)
p, ids = PB.build_elm_vsfm(mpp_b200.VSFM, d)
bench.set_forcing_host(p, ids, d)
ms = []
for s in range(8):
    p.pre_step_dt(); p.step_dt(1800.0, s + 1); p.post_step_dt()
    ms.append(p.last_step_ms())
st = p.stats()
print(os.environ.get("MPPGPU_LIB_PATH", "default"), "ncol", ncol, "ms/step", ["%.2f" % m for m in ms], "col-steps/s %.3e" % (ncol / (np.mean(ms[3:]) * 1e-3)),
      "its mean %.2f nf mean %.2f" % (st["newton_its"].mean(), st["nfuncs"].mean()))
