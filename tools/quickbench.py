"""Quick device-resident timing of the step kernels (development aid; not the headline bench)."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import bench
import mpp_b200
from mpp_b200 import problems as PB
from mpp_b200 import constants as K
mode = sys.argv[1] if len(sys.argv) > 1 else "vsfm"
ncol = int(sys.argv[2]) if len(sys.argv) > 2 else 1 << 20
if mode == "vsfm":
    zmin = float(os.environ.get("ZWT_MIN", "2.0"))
    d = bench.shard_inputs(0, ncol, zwt_min=zmin)
    for ordering in ([int(os.environ["ORDERING"])] if "ORDERING" in os.environ else [0, 1]):
        p, ids = PB.build_elm_vsfm(mpp_b200.VSFM, d)
        p.set_column_ordering(ordering)
        bench.set_forcing_host(p, ids, d)
        if os.environ.get("STEP_BUDGET"):
            p.set_step_budget(int(os.environ["STEP_BUDGET"]))
        if os.environ.get("WITH_BC"):           # a Dirichlet head at the bottom of every column: the HAS_BC instance of the step kernel
            bot = p.add_condition(1, K.COND_BC, K.COND_DIRICHLET, K.SOIL_BOTTOM_CELLS)
            # hydrostatic continuation of the initial profile to the bottom face: no flux through it at the start
            p.set_data(K.AUXVAR_BC, K.VAR_BC_SS_CONDITION, bot, d["press_ic"].reshape(ncol, -1)[:, -1] + 998.2 * 9.80665 * 0.5 * d["dz"].reshape(ncol, -1)[:, -1])
        ms, bad = [], 0
        for s in range(12):
            p.pre_step_dt(); conv, reason = p.step_dt(1800.0, s + 1); p.post_step_dt()
            ms.append(p.last_step_ms()); bad += (not conv)
        st = p.stats(); sums, maxs = p.mass_balance()
        print(os.environ.get("MPPGPU_LIB_PATH", "default"), "vsfm ncol", ncol, "zwt_min", zmin, "ordering", ordering, "ms/step", ["%.2f" % m for m in ms],
              "col-steps/s %.3e" % (ncol / (np.mean(ms[4:]) * 1e-3)), "its mean %.2f nf mean %.2f nf max %d" % (st["newton_its"].mean(), st["nfuncs"].mean(), st["nfuncs"].max()),
              "steps not converged", bad, "cuts", int((st["dt_cuts"] > 0).sum()), "max mass err %.2e" % maxs[0], "warp waste (batch order) %.3f" % bench.warp_waste(st["nfuncs"]), flush=True)
        p.close()
elif mode == "th":
    # the benchmark's TH batch (ELM's default curve, water tables from 2 m); TH_SATFUNC=van_genuchten for the survey's original draw
    d = bench.shard_inputs_th(0, ncol, satfunc=os.environ.get("TH_SATFUNC", "smooth_brooks_corey_bz3"))
    p, ids = PB.build_elm_th(mpp_b200.TH, d)
    if "ORDERING" in os.environ:
        p.set_column_ordering(int(os.environ["ORDERING"]))
    ms = []
    for s in range(int(os.environ.get("NSTEPS", "8"))):
        conv, reason, out = PB.elm_th_step(p, ids, d, 1800.0, s + 1)
        ms.append(p.last_step_ms())
    st = p.stats()
    m = float(np.mean(ms[3:]))
    print("th ncol", ncol, d["satfunc"], "ms/step", ["%.2f" % x for x in ms], "col-steps/s %.3e" % (ncol / (m * 1e-3)),
          "alg GB/s (1824 B/col) %.0f" % (1824 * ncol / (m * 1e-3) / 1e9), "its mean %.2f nf mean %.2f nf max %d conv %s cuts %d" % (
              st["newton_its"].mean(), st["nfuncs"].mean(), st["nfuncs"].max(), conv, int((st["dt_cuts"] > 0).sum())))
elif mode == "elm":
    d = bench.shard_inputs(0, ncol); d["satfunc"] = os.environ.get("ELM_SATFUNC", "smooth_brooks_corey_bz3")
    p, ids = PB.build_elm_vsfm(mpp_b200.VSFM, d)
    st = PB.elm_vsfm_raw_state(p, d, patches=True)
    p.elm_set_geometry(st["zi"], st["dz"], st["nlevsoi"], ids)
    if len(sys.argv) > 3:
        p.set_step_budget(int(sys.argv[3]))
    for s in range(5):
        t0 = time.time(); out = p.elm_solve(1800.0, st, s + 1); wall = time.time() - t0
        print("elm_solve ncol", ncol, "device ms %.2f wall ms %.1f attempts %d nfailed %d max its %d" % (p.last_step_ms(), wall * 1e3, out["nattempts"], out["nfailed"], out["iter_count"].max()),
              "col-steps/s device %.3e e2e %.3e" % (ncol / (p.last_step_ms() * 1e-3), ncol / wall))
elif mode == "elmhost":
    import torch
    d = bench.shard_inputs(0, ncol)
    p, ids = PB.build_elm_vsfm(mpp_b200.VSFM, d)
    st = PB.elm_vsfm_raw_state(p, d, patches=True)
    p.elm_set_geometry(st["zi"], st["dz"], st["nlevsoi"], ids)
    p.set_step_budget(2000)
    o = p.elm_solve(1800.0, st, 1)
    nbytes = sum(v.nbytes for v in list(st.values()) + list(o.values()) if isinstance(v, np.ndarray))
    print("bytes of all state + output arrays %.1f MB, npft %d" % (nbytes / 1e6, st["pft_wtcol"].size))
    x = torch.empty(ncol * 15, dtype=torch.float64).pin_memory(); g = torch.empty_like(x, device="cuda")
    for name, f in (("h2d", lambda: g.copy_(x, non_blocking=True)), ("d2h", lambda: x.copy_(g, non_blocking=True))):
        f(); torch.cuda.synchronize(); t0 = time.time()
        for _ in range(5):
            f()
        torch.cuda.synchronize(); print("raw pinned %s %.1f GB/s" % (name, 5 * x.nbytes / (time.time() - t0) / 1e9))
    def run(tag, S, O):
        for s in range(3):
            t0 = time.time(); p.elm_solve(1800.0, S, s + 2, out=O); w = time.time() - t0
            print(tag, "wall ms %.1f device ms %.2f" % (w * 1e3, p.last_step_ms()))
    run("pageable", st, o)
    sp, op = PB.page_aligned_state(st), PB.page_aligned_state(o)
    lk = [v for v in list(sp.values()) + list(op.values()) if isinstance(v, np.ndarray) and v.nbytes]
    for v in lk:
        mpp_b200.host_register(v)
    run("host_register", sp, op)
    for v in lk:
        mpp_b200.host_unregister(v)
    pin = lambda S: {k: (torch.from_numpy(v).pin_memory().numpy() if isinstance(v, np.ndarray) and v.nbytes else v) for k, v in S.items()}
    keep = (pin(st), pin(o))
    run("torch pin_memory", *keep)
elif mode == "elmpipe":
    # the two ELM solve entry points with page-locked host arrays, over the number of pipeline chunks
    d = bench.shard_inputs(0, ncol); d["satfunc"] = "smooth_brooks_corey_bz3"
    p, ids = PB.build_elm_vsfm(mpp_b200.VSFM, d)
    st = PB.elm_vsfm_raw_state(p, d, patches=True)
    p.elm_set_geometry(st["zi"], st["dz"], st["nlevsoi"], ids)
    sp = PB.page_aligned_state(st)
    op = PB.page_aligned_state(p.elm_solve(1800.0, sp, 1))
    lk = [v for v in list(sp.values()) + list(op.values()) if isinstance(v, np.ndarray) and v.nbytes]
    for v in lk:
        mpp_b200.host_register(v)
    for nch in (1, 2, 4, 8, 16, 32):
        p.elm_set_pipeline(nch)
        w = []
        for s in range(4):
            t0 = time.time(); r = p.elm_solve(1800.0, sp, s + 2, out=op); w.append((time.time() - t0) * 1e3)
        print("vsfm elm_solve chunks %2d wall ms %s span ms %.2f attempts %d" % (nch, ["%.1f" % x for x in w], p.last_step_ms(), r["nattempts"]), flush=True)
    for v in lk:
        mpp_b200.host_unregister(v)
    p.close()
    base = 4096
    d0 = PB.elm_snow_thermal_inputs(base, 15, 5)
    d, o = PB.tile_snow_thermal(d0, PB.pack_elm_snow_thermal(d0), max(1, ncol // base))
    p = PB.build_elm_snow_thermal(mpp_b200.ThermalSnow, d)
    e0 = PB.elm_thermal_raw_arrays(d0); reps = ncol // base
    e = PB.page_aligned_state({k: (np.tile(v, (1, reps)) if v.ndim == 2 else np.tile(v, reps)) for k, v in e0.items()})
    for v in e.values():
        mpp_b200.host_register(v)
    for static in (False, True):
        for nch in (1, 4, 8, 16):
            p.elm_set_pipeline(nch, static_soil_geometry=static)
            w = []
            for s in range(4):
                t0 = time.time(); p.elm_solve(1800.0, e, s + 2); w.append((time.time() - t0) * 1e3)
            print("thermal elm_solve static_soil %d chunks %2d wall ms %s span ms %.2f" % (static, nch, ["%.1f" % x for x in w], p.last_step_ms()), flush=True)
    for v in e.values():
        mpp_b200.host_unregister(v)
elif mode == "snow":
    base = 4096
    d0 = PB.elm_snow_thermal_inputs(base, 15, 5)
    d, o = PB.tile_snow_thermal(d0, PB.pack_elm_snow_thermal(d0), max(1, ncol // base))
    ncol = d["ncol"]
    p = PB.build_elm_snow_thermal(mpp_b200.ThermalSnow, d)
    conv, T = PB.elm_snow_thermal_step(p, o, 1800.0, 1)
    ms = []
    for s in range(10):
        p.step_dt(1800.0, s + 2)
        ms.append(p.last_step_ms())
    m = float(np.mean(ms[3:]))
    print("thermal snow+ssw+soil ncol", ncol, "ms/step", ["%.3f" % x for x in ms], "col-steps/s %.3e" % (ncol / (m * 1e-3)),
          "alg GB/s (2356 B/col) %.0f" % (2356 * ncol / (m * 1e-3) / 1e9))
else:
    d = PB.elm_thermal_inputs(ncol, 15)
    for mode in (0, 1):
        p, ids = PB.build_elm_thermal(mpp_b200.Thermal, d)
        p.set_bulk_copy(mode)
        T = d["T0"]
        conv, T = PB.elm_thermal_step(p, ids, d, T, 1800.0, 1)
        ms = []
        for s in range(12):
            p.step_dt(1800.0, s + 2)           # device-resident chain: soln -> soln_prev
            ms.append(p.last_step_ms())
        m = float(np.mean(ms[3:]))
        print("thermal ncol", ncol, "bulk_copy", mode, "ms/step", ["%.3f" % x for x in ms], "col-steps/s %.3e" % (ncol / (m * 1e-3)),
              "alg GB/s (1224 B/col) %.0f" % (1224 * ncol / (m * 1e-3) / 1e9), flush=True)
        p.close()
