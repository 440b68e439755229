"""Development aid: long runs of the benchmark batches (many coupling steps on the same handle) -- convergence, mass balance and step time
must stay put.  The opt-in step budget is ON here (mppgpu_set_step_budget): after ~45 steps of the SAME synthetic forcing the columns whose draw
has strong transpiration and no infiltration have dried their root zone, and a fixed-rate sink on a dry cell is not solvable (DESIGN.md
section 2) -- the run reports how many give up.
Runs on a GPU:  python tools/soak.py [ncol] [vsfm_steps] [th_steps] [budget] [elm_solves]   (BTRAN=1, SATFUNC=... in the environment)"""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
import mpp_b200
from mpp_b200 import problems as PB

ncol = int(sys.argv[1]) if len(sys.argv) > 1 else 1048576
nv = int(sys.argv[2]) if len(sys.argv) > 2 else 300
nt = int(sys.argv[3]) if len(sys.argv) > 3 else 100
budget = int(sys.argv[4]) if len(sys.argv) > 4 else 2000
btran = bool(int(os.environ.get("BTRAN", "0")))   # BTRAN=1: transpiration falls with the soil water, as a host model's does (ELM's btran)
from mpp_b200 import constants as K

d = bench.shard_inputs(0, ncol)
d["satfunc"] = os.environ.get("SATFUNC", d["satfunc"])          # e.g. SATFUNC=smooth_brooks_corey_bz3: ELM's default curve
p, ids = PB.build_elm_vsfm(mpp_b200.VSFM, d)
bench.set_forcing_host(p, ids, d)
p.set_step_budget(budget)
failed_cols = 0
ms, bad, worst_err, worst_nf, cuts = [], 0, 0.0, 0, 0
for s in range(nv):
    if btran and s:
        # what the host model does between two steps (ELM's plant wilting factor, problems.plant_wilting_factor)
        p.set_data(K.AUXVAR_SS, K.VAR_BC_SS_CONDITION, ids["et"], d["et"] * PB.plant_wilting_factor(p.get_data(K.AUXVAR_INTERNAL, K.VAR_PRESSURE, 1)))
    p.pre_step_dt(); conv, reason = p.step_dt(1800.0, s + 1); p.post_step_dt()
    ms.append(p.last_step_ms()); bad += (not conv)
    sums, maxs = p.mass_balance()
    worst_err = max(worst_err, maxs[0]); worst_nf = max(worst_nf, int(p.stats()["nfuncs"].max())) if s % 25 == 0 else worst_nf
    cuts = max(cuts, int(maxs[3]))
    if not conv:
        failed_cols = max(failed_cols, int((p.stats()["reasons"] < 0).sum()))
    if s % 50 == 49:
        print("  vsfm step %d: %.2f ms, steps with a failed column so far %d (most failed columns in one step %d)" % (s + 1, ms[-1], bad, failed_cols), flush=True)
if nv:
    print(d["satfunc"], "btran" if btran else "fixed-rate ET", "vsfm %d columns x %d steps: not converged %d, worst |mass error| %.2e kg, max dt cuts %d, max evaluations (sampled) %d, ms/step first 5 %s last 5 %s" % (
        ncol, nv, bad, worst_err, cuts, worst_nf, ["%.2f" % x for x in ms[1:6]], ["%.2f" % x for x in ms[-5:]]), flush=True)
p.close()

d = bench.shard_inputs_th(0, ncol)
p, ids = PB.build_elm_th(mpp_b200.TH, d)
p.set_step_budget(budget)
ms, bad, worst_nf = [], 0, 0
for s in range(nt):
    conv, reason, out = PB.elm_th_step(p, ids, d, 1800.0, s + 1)
    ms.append(p.last_step_ms()); bad += (not conv)
    if s % 10 == 0:
        worst_nf = max(worst_nf, int(p.stats()["nfuncs"].max()))
        assert np.isfinite(out["pressure"]).all() and np.isfinite(out["temperature"]).all()
if nt:
    print("th %d columns x %d steps: not converged %d, max evaluations (sampled) %d, T range %.2f..%.2f K, ms/step first 5 %s last 5 %s" % (
        ncol, nt, bad, worst_nf, out["temperature"].min(), out["temperature"].max(), ["%.2f" % x for x in ms[1:6]], ["%.2f" % x for x in ms[-5:]]), flush=True)

# ---- the ELM driver call (MPPVSFMALM_Solve: packing, StepDT, the per-column retry loop, unpacking) step after step on ELM's own arrays ----
ne = int(sys.argv[5]) if len(sys.argv) > 5 else 0
if ne:
    d = bench.shard_inputs(0, ncol); d["satfunc"] = "smooth_brooks_corey_bz3"          # ELM's default curve
    p, ids = PB.build_elm_vsfm(mpp_b200.VSFM, d)
    st = PB.page_aligned_state(PB.elm_vsfm_raw_state(p, d, patches=False))
    p.elm_set_geometry(st["zi"], st["dz"], st["nlevsoi"], ids)
    p.set_step_budget(budget)
    out = PB.page_aligned_state(p.elm_solve(1800.0, st, 1))
    for v in list(st.values()) + list(out.values()):
        if isinstance(v, np.ndarray) and v.nbytes:
            mpp_b200.host_register(v)
    qtran0, rootr = st["qflx_tran_veg_col"].copy(), st["rootr_col"].reshape(ncol, -1)
    nfail, natt, worst, wall = 0, 0, 0.0, []
    import time
    for s in range(1, ne):
        # what the host model does between two solves: transpiration demand times the root-weighted wilting factor (ELM's btran) of the
        # matric potentials the last solve returned
        wilt = PB.plant_wilting_factor(K.PRESSURE_REF + out["smp_l"].reshape(ncol, -1) * 1.0e-3 * (998.2 * 9.80665))    # smp_l [mm] -> [Pa]
        st["qflx_tran_veg_col"][...] = qtran0 * (rootr * wilt).sum(axis=1)
        t0 = time.perf_counter(); r = p.elm_solve(1800.0, st, s + 1, out=out); wall.append((time.perf_counter() - t0) * 1e3)
        nfail = max(nfail, r["nfailed"]); natt = max(natt, r["nattempts"]); worst = max(worst, float(out["abs_mass_error"][out["status"] == 1].max()))
        if s % 25 == 0:
            print("  elm_solve step %d: %.1f ms, most failed columns in one solve so far %d, most StepDT calls in one solve %d" % (s, wall[-1], nfail, natt), flush=True)
    print("elm_solve (smooth_brooks_corey_bz3, btran) %d columns x %d solves: most failed columns in one solve %d, most StepDT calls in one solve %d, "
          "worst |mass error| of an accepted column %.2e kg, ms/solve median %.1f" % (ncol, ne, nfail, natt, worst, float(np.median(wall))), flush=True)
