"""Development aid: GPU-vs-oracle deviations of candidate parity tests at several SNES tolerances (prints maxima; asserts nothing)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import mpp_b200
from mpp_b200 import problems as PB, constants as K
from oracle import oracle as O


def relmax(a, b):
    return float(np.max(np.abs(a - b) / np.maximum(np.abs(b), 1e-300))) if a.size else 0.0


def relmax_p(a, b):
    return float(np.max(np.abs(a - b) / np.maximum(np.abs(b), np.abs(b - K.PRESSURE_REF)))) if a.size else 0.0


def th_case(satfunc, dens, iee, rtol, stol, ncol=500, nsteps=3):
    d = PB.elm_th_inputs(ncol, 15, satfunc=satfunc, density_type=dens, iee_type=iee)
    p, ids = PB.build_elm_th(mpp_b200.TH, d)
    o, oids = PB.build_elm_th(O.OracleTH, d, per_column=True, nthreads=8)
    for s in (p, o):
        s.set_tolerances(1e-50, rtol, stol, 50, 10000)
    for step in range(nsteps):
        conv, reason, out = PB.elm_th_step(p, ids, d, 1800.0, step + 1)
        convo, reasono, outo = PB.elm_th_step(o, oids, d, 1800.0, step + 1)
        sg, so = p.stats(), o.stats()
        same = sg["newton_its"] == so["newton_its"]
        devs = {k: (relmax_p if k == "pressure" else relmax)(out[k], outo[k]) for k in ("pressure", "temperature", "sat", "mass")}
        print("TH", satfunc, dens, "rtol %.0e" % rtol, "step", step, conv, convo, "cuts equal", bool(np.array_equal(sg["dt_cuts"], so["dt_cuts"])), "its differ %d/%d" % (int((~same).sum()), ncol),
              " ".join("%s %.2e" % kv for kv in devs.items()), "| reasons differ", int((sg["reasons"] != so["reasons"]).sum()), flush=True)


def vsfm_case(rtol, stol, max_it, scale, ncol=1000, nsteps=3, zwt_min=1.0):
    d = PB.elm_vsfm_inputs(ncol, 15, zwt_min=zwt_min)
    p, ids = PB.build_elm_vsfm(mpp_b200.VSFM, d)
    o, oids = PB.build_elm_vsfm(O.OracleVSFM, d, per_column=True, nthreads=8)
    for s in (p, o):
        s.set_tolerances(1e-50, rtol, stol, max_it, 10000)
    for step in range(nsteps):
        conv, reason, out = PB.elm_vsfm_step(p, ids, d, 1800.0, step + 1, scale=scale)
        convo, reasono, outo = PB.elm_vsfm_step(o, oids, d, 1800.0, step + 1, scale=scale)
        sg, so = p.stats(), o.stats()
        ok = so["reasons"] > 0
        P, Po = out["pressure"].reshape(ncol, 15), outo["pressure"].reshape(ncol, 15)
        S, So = out["sat"].reshape(ncol, 15), outo["sat"].reshape(ncol, 15)
        line = "VSFM rtol %.0e max_it %d scale %g step %d conv %s/%s cuts equal %s (max %d, %d cols cut) its differ %d reasons differ %d failed %d |" % (
            rtol, max_it, scale, step, conv, convo, bool(np.array_equal(sg["dt_cuts"], so["dt_cuts"])), int(so["dt_cuts"].max()), int((so["dt_cuts"] > 0).sum()),
            int((sg["newton_its"] != so["newton_its"]).sum()), int((sg["reasons"] != so["reasons"]).sum()), int((~ok).sum()))
        for name, m in (("all converged", ok), ("no cut", ok & (so["dt_cuts"] == 0)), ("1-2 cuts", ok & (so["dt_cuts"] > 0) & (so["dt_cuts"] <= 2)), (">2 cuts", ok & (so["dt_cuts"] > 2))):
            line += " %s[%d]: P %.1e sat %.1e;" % (name, int(m.sum()), relmax_p(P[m], Po[m]), relmax(S[m], So[m]))
        print(line, flush=True)


if __name__ == "__main__":
    what = sys.argv[1] if len(sys.argv) > 1 else "all"
    TC = (K.DENSITY_TGDPB01, K.INT_ENERGY_ENTHALPY_CONSTANT)
    IF = (K.DENSITY_IFC67, K.INT_ENERGY_ENTHALPY_IFC67)
    if what in ("all", "th"):
        for sf, (dn, ie) in (("van_genuchten", TC), ("smooth_brooks_corey_bz3", TC), ("van_genuchten", IF)):
            for rtol, stol in ((1e-8, 1e-10), (1e-10, 1e-12), (1e-11, 1e-13)):
                if dn == K.DENSITY_IFC67 and rtol < 1e-8:
                    continue
                th_case(sf, dn, ie, rtol, stol)
    if what in ("all", "vsfm"):
        vsfm_case(1e-8, 1e-10, 50, 1.0)
        vsfm_case(1e-10, 1e-12, 50, 1.0)
        vsfm_case(1e-8, 1e-10, 2, 5.0, ncol=200, nsteps=1)
        vsfm_case(1e-10, 1e-12, 4, 5.0, ncol=1000, nsteps=3)
