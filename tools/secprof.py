"""Development aid: per-section cycle counters of vsfm_step2_kernel (needs the -DVSFM2_PROFILE build in build/variants)."""
import ctypes as C, os, sys, numpy as np
os.environ["MPPGPU_LIB_PATH"] = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "build", "variants", "libmppgpu_prof.so")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import bench, mpp_b200
from mpp_b200 import problems as PB
from mpp_b200._lib import lib
ncol = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
d = bench.shard_inputs(0, ncol)
p, ids = PB.build_elm_vsfm(mpp_b200.VSFM, d); bench.set_forcing_host(p, ids, d)
L = lib(); out = (C.c_longlong * 12)()
for s in range(8):
    p.pre_step_dt(); p.step_dt(1800.0, s + 1); p.post_step_dt()
    L.mppgpu_dbg_profile(out)
    o = list(out)
    if s >= 4:
        nw = o[6]
        print("step %d ms %.2f | per warp: total %.0f cyc | newton %.1f x %.0f cyc (assemble %.0f, eliminate %.0f, pcr %.0f, rest %.0f) | eval %.1f x %.0f cyc (curves %.0f, flux+derivs %.0f, norms %.0f) | logic %.0f cyc | setup+teardown %.0f" % (
            s + 1, p.last_step_ms(), o[5] / nw, o[1] / nw, o[0] / max(o[1], 1), o[9] / max(o[1], 1), o[10] / max(o[1], 1), o[11] / max(o[1], 1), (o[0] - o[9] - o[10] - o[11]) / max(o[1], 1), o[3] / nw, o[2] / max(o[3], 1), o[7] / max(o[3], 1), (o[2] - o[7] - o[8]) / max(o[3], 1), o[8] / max(o[3], 1), o[4] / nw, (o[5] - o[0] - o[2] - o[4]) / nw))
