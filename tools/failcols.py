"""Development aid: run the benchmark batch on the GPU and list the columns whose StepDT cut dt or failed, so the
same columns can be replayed through the oracle on the CPU (tools_failcols.py replay)."""
import json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import problems as PB, bench
from mpp_b200 import constants as K


def column_inputs(d, cols):
    cols = np.asarray(cols)
    nlev = d["nlev"]
    out = {"ncol": len(cols), "nlev": nlev, "satfunc": d["satfunc"]}
    for key in ("dz", "watsat", "hksat", "bsw", "sucsat", "residual_sat"):
        out[key] = d[key][cols]
    for key in ("area", "infil", "dew", "snow", "sublim"):
        out[key] = d[key][cols]
    for key in ("press_ic", "et", "drain", "frac_liq"):
        out[key] = d[key].reshape(-1, nlev)[cols].reshape(-1)
    return out


if __name__ == "__main__":
    mode = sys.argv[1]
    if mode == "gpu":
        import mpp_b200
        ncol = int(sys.argv[2]); nsteps = int(sys.argv[3])
        d = bench.shard_inputs(0, ncol)
        p, ids = PB.build_elm_vsfm(mpp_b200.VSFM, d)
        bench.set_forcing_host(p, ids, d)
        rec = []
        for s in range(nsteps):
            p.pre_step_dt(); conv, reason = p.step_dt(1800.0, s + 1); p.post_step_dt()
            st = p.stats()
            bad = np.nonzero((st["reasons"] < 0) | (st["dt_cuts"] > 0))[0]
            sums, maxs = p.mass_balance()
            print("step", s + 1, "conv", conv, "reason", reason, "nbad", len(bad), "its max", st["newton_its"].max(), "max err", maxs[0], flush=True)
            for c in bad[:200]:
                rec.append({"step": s + 1, "col": int(c), "reason": int(st["reasons"][c]), "cuts": int(st["dt_cuts"][c]),
                            "its": int(st["newton_its"][c]), "nf": int(st["nfuncs"][c])})
        P = p.get_data(K.AUXVAR_INTERNAL, K.VAR_PRESSURE, 1).reshape(ncol, -1)
        cols = sorted({r["col"] for r in rec})[:64]
        json.dump({"ncol": ncol, "nsteps": nsteps, "records": rec, "cols": cols, "P_final": {str(c): P[c].tolist() for c in cols}},
                  open(os.path.join(ROOT, "gpurun_out", "failcols.json"), "w"))
    else:
        from oracle import oracle as O
        r = json.load(open(os.path.join(ROOT, "gpurun_out", "failcols.json")))
        cols = r["cols"]
        print("replaying", len(cols), "columns", cols[:10])
        d = bench.shard_inputs(0, r["ncol"])
        dc = column_inputs(d, cols)
        o, ids = PB.build_elm_vsfm(O.OracleVSFM, dc, per_column=True, nthreads=8)
        for s in range(r["nsteps"]):
            conv, reason, out = PB.elm_vsfm_step(o, ids, dc, 1800.0, s + 1)
            st = o.stats()
            print("step", s + 1, conv, reason, "its", st["newton_its"].tolist()[:16], "cuts", st["dt_cuts"].tolist()[:16], "reasons", st["reasons"].tolist()[:16])
        gpu = [x for x in r["records"]]
        print("gpu records:", gpu[:40])
        Pf = out["pressure"].reshape(len(cols), -1)
        for i, c in enumerate(cols[:8]):
            g = np.array(r["P_final"][str(c)])
            print(c, "max rel dev P", np.max(np.abs(g - Pf[i]) / np.abs(Pf[i])))
