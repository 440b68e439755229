"""Pin the CPU oracle to the reference's own golden vectors."""
import os
import numpy as np
import pytest

import problems as PB
from mpp_b200 import constants as K


def test_eos_density_known_answers(oracle, golden):
    # src/tests/test_eos_{constant,tgdp01,ifc67}_density.F90:16-25
    g = golden["eos_density"]
    for name, itype in (("constant", K.DENSITY_CONSTANT), ("tgdpb01", K.DENSITY_TGDPB01), ("ifc67", K.DENSITY_IFC67)):
        den, ddp, ddt = oracle.density(g["p"], g["t_K"], itype)
        assert abs(den - g[name]["den"]) < g["tol"]["den"]
        assert abs(ddp - g[name]["dden_dp"]) < g["tol"]["dden_dp"]
        assert abs(ddt - g[name]["dden_dT"]) < g["tol"]["dden_dT"]


def _assert_matches_printed(ours, ref_val, tol_abs=None):
    """The baseline holds 13 printed significant digits; require agreement to the last printed digit
    (and to the reference's own absolute tolerance when that is looser)."""
    quantum = 10.0 ** (np.floor(np.log10(abs(ref_val))) - 12) if ref_val != 0 else 1e-13
    tol = max(0.51 * quantum, tol_abs or 0.0)
    assert abs(ours - ref_val) <= tol, (ours, ref_val, tol)


@pytest.mark.parametrize("per_column", [False, True])
def test_celia1990_reproduces_reference_baseline(oracle, golden, per_column):
    # regression_tests/vsfm/vsfm_celia1990.regression.baseline (pressure 1e-10 abs: vsfm.cfg:4-5)
    p, top, bot = PB.build_celia(oracle.OracleVSFM, per_column=per_column)
    P, S, its = PB.run_celia(p, top, bot)
    assert sum(its) == 226 and its[0] == 34
    for name, data, tol in (("liquid_pressure", P, 1e-10), ("liquid_saturation", S, 1e-16)):
        ref = golden["vsfm_celia1990"][name]
        _assert_matches_printed(data.min(), ref["min"], tol)
        _assert_matches_printed(data.max(), ref["max"], tol)
        _assert_matches_printed(data.sum() / data.size, ref["mean"], tol)
        for key, val in ref.items():
            if key.startswith("cell"):
                _assert_matches_printed(data[int(key.split()[1]) - 1], val, tol)


def test_celia_regression_file_format(oracle, golden):
    # src/driver/standalone/util/regression.F90:76-124 -- the writer must reproduce the baseline text
    p, top, bot = PB.build_celia(oracle.OracleVSFM, per_column=False)
    P, S, _ = PB.run_celia(p, top, bot)
    lines = PB.regression_block("liquid_pressure", "pressure", P, 5)
    assert lines[0] == "[liquid_pressure]" and lines[1] == "category = pressure"
    assert lines[2] == "min =   0.3535500000000E+04"
    assert lines[5].startswith("cell    1 = ") and lines[9].startswith("cell   81 = ")
    assert lines[3] == "max =   0.9398362808479E+05"


def test_saturation_curves_are_consistent(oracle):
    # no reference unit test exists for SaturationFunction.F90; check analytic derivatives against central differences
    rng = np.random.default_rng(3)
    for name in ("van_genuchten", "brooks_corey", "smooth_brooks_corey_bz2", "smooth_brooks_corey_bz3"):
        for _ in range(20):
            alpha = 1.0 / (rng.uniform(50, 600) * K.GRAV)
            lam = 1.0 / rng.uniform(3, 12)
            sp = oracle.satparams(name, 0.05, alpha, lam)
            pc = -rng.uniform(1.2, 50.0) / alpha
            press = K.PRESSURE_REF + pc
            h = 1e-4 * abs(pc)
            s, ds = oracle.press_to_sat(sp, press)
            k, dk = oracle.press_to_relperm(sp, press)
            sp_, _ = oracle.press_to_sat(sp, press + h); sm_, _ = oracle.press_to_sat(sp, press - h)
            kp_, _ = oracle.press_to_relperm(sp, press + h); km_, _ = oracle.press_to_relperm(sp, press - h)
            assert 0 < s < 1 and 0 < k < 1
            assert abs((sp_ - sm_) / (2 * h) - ds) <= 1e-6 * abs(ds) + 1e-18
            assert abs((kp_ - km_) / (2 * h) - dk) <= 1e-5 * abs(dk) + 1e-22
        # saturated branch
        s, ds = oracle.press_to_sat(sp, K.PRESSURE_REF + 10.0)
        assert s == 1.0 and ds == 0.0


def test_sbc_smoothing_is_continuous(oracle):
    # SatFunc_Set_SBC_bz2/bz3 choose pu so that the cubic joins Brooks-Corey with matching value at pu (:296-372)
    alpha, lam = 1.0 / (200.0 * K.GRAV), 0.2
    for name in ("smooth_brooks_corey_bz2", "smooth_brooks_corey_bz3"):
        sp = oracle.satparams(name, 0.0, alpha, lam)
        assert sp.sbc_pu < sp.sbc_ps < 0
        eps = 1e-9 * abs(sp.sbc_pu)
        s0, _ = oracle.press_to_sat(sp, K.PRESSURE_REF + sp.sbc_pu - eps)
        s1, _ = oracle.press_to_sat(sp, K.PRESSURE_REF + sp.sbc_pu + eps)
        assert abs(s0 - s1) < 1e-8
        s2, _ = oracle.press_to_sat(sp, K.PRESSURE_REF + sp.sbc_ps - eps)
        assert abs(s2 - 1.0) < 1e-8


def test_elm_like_batch_global_vs_per_column(oracle):
    """The reference runs ONE SNES over all columns of a rank (global norms, one lambda, one dt cut); the GPU runs
    one per column.  Both stop at rtol = 1e-8 of their own ||F0||, so at the default tolerances the two answers differ
    at the O(rtol) level (the reference's own answer depends on how columns are dealt to MPI ranks at that level);
    with the tolerances tightened the fixed points agree to 1e-10 relative (SURVEY.md section 7 hard part (a))."""
    d = PB.elm_vsfm_inputs(24, 15)

    def run(per_column, tight):
        p, ids = PB.build_elm_vsfm(oracle.OracleVSFM, d, per_column=per_column)
        if tight:
            p.set_tolerances(1e-50, 1e-10, 1e-16, 50, 10000)
        for step in range(3):
            conv, reason, out = PB.elm_vsfm_step(p, ids, d, 1800.0, step + 1)
            assert conv and reason > 0
        return out

    def relmax(a, b):
        return np.max(np.abs(a - b) / np.maximum(np.abs(a), 1e-300))

    loose = {pc: run(pc, False) for pc in (False, True)}
    tight = {pc: run(pc, True) for pc in (False, True)}
    for k in ("pressure", "sat", "mass"):
        assert relmax(loose[False][k], loose[True][k]) < 1e-6, k
        assert relmax(tight[False][k], tight[True][k]) < 1e-10, k
        assert relmax(tight[True][k], loose[True][k]) < 1e-6, k


def test_elm_like_mass_balance(oracle):
    # MPPVSFMALM_Driver.F90:860-863: |m_beg - m_end + sum(q) dt| < 1e-5 kg per column
    d = PB.elm_vsfm_inputs(16, 15)
    p, ids = PB.build_elm_vsfm(oracle.OracleVSFM, d, per_column=True)
    conv, reason, out0 = PB.elm_vsfm_step(p, ids, d, 1800.0, 1)
    m0 = out0["mass"].reshape(16, 15).sum(1)
    conv, reason, out1 = PB.elm_vsfm_step(p, ids, d, 1800.0, 2)
    m1 = out1["mass"].reshape(16, 15).sum(1)
    q = d["infil"] + d["et"].reshape(16, 15).sum(1)
    err = np.abs(m0 - m1 + q * 1800.0)
    assert conv and err.max() < 1e-5


def test_thermal_mms_reproduces_reference_baseline(oracle, golden):
    # regression_tests/thermal/thermal_mms.regression.baseline (category general => 1e-16 abs: every printed digit)
    T = PB.run_thermal_mms(PB.build_thermal_mms(oracle.OracleThermal))
    ref = golden["thermal_mms"]["temperature"]
    _assert_matches_printed(T.min(), ref["min"])
    _assert_matches_printed(T.max(), ref["max"])
    _assert_matches_printed(T.sum() / T.size, ref["mean"])
    for key, val in ref.items():
        if key.startswith("cell"):
            _assert_matches_printed(T[int(key.split()[1]) - 1], val)
    lines = PB.regression_block("temperature", "general", T, 5)
    assert lines[2] == "min =   0.2707677262973E+03" and lines[5] == "cell    1 =   0.2707677262973E+03"
    assert lines[9] == "cell   17 =   0.2752526149775E+03"


def test_thermal_energy_balance_oracle(oracle):
    d = PB.elm_thermal_inputs(32, 15)
    d["dhsdT"] = np.zeros(32)
    o, ids = PB.build_elm_thermal(oracle.OracleThermal, d)
    conv, T1 = PB.elm_thermal_step(o, ids, d, d["T0"], 1800.0, 1)
    hc = d["csol"] * (1 - d["watsat"]) * d["dz"] + d["ice"].reshape(32, 15) * 2.11727e3 + d["liq"].reshape(32, 15) * 4.188e3
    dE = (hc * (T1.reshape(32, 15) - d["T0"].reshape(32, 15))).sum(1)
    assert np.max(np.abs(dE - d["hs"] * 1800.0)) < 1e-6 * np.max(np.abs(d["hs"] * 1800.0))


def test_mass_and_heat_reproduces_reference_baseline(oracle, golden):
    """regression_tests/th/mass_and_heat.regression.baseline (100 cells along x, IFC-67, one 3600 s step).
    The reference solves its Newton systems inexactly (GMRES + ILU(0) on segregated unknowns, KSP rtol 1e-5) and only
    asks for 1e-8 K / 1e-12 Pa(abs) of itself (regression_tests/th/th.cfg); an exact-Newton restatement lands within
    ~1e-12 relative of the printed baseline.  Bar used here: 1e-11 relative."""
    p, b0, b1 = PB.build_mass_and_heat(oracle.OracleTH)
    conv, reason, P, T = PB.run_mass_and_heat(p, b0, b1)
    assert conv and reason == 3
    for name, data in (("liquid_pressure", P), ("temperature", T)):
        ref = golden["mass_and_heat"][name]
        for key, val in ref.items():
            if key == "category":
                continue
            ours = {"min": data.min(), "max": data.max(), "mean": data.sum() / data.size}.get(key)
            if ours is None:
                ours = data[int(key.split()[1]) - 1]
            assert abs(ours - val) <= 1e-11 * abs(val), (name, key, ours, val)


# th_mms temperature bar.  Two things keep the reference's 1e-8 K (regression_tests/th/th.cfg:8-9) out of reach of anything but the
# reference binary itself.  (i) Its driver differentiates the enthalpy and the Kersten number NUMERICALLY with a step of 1e-6 m
# (th_mms_problem.F90:1203, 1396-1440): the ~1e-13 relative round-off of the IFC-67 polynomials is divided by 2e-6, so the heat source
# depends on the EOS implementation's last ulp to ~1e-5 relative (measured between this repo's two EOS implementations: 7e-6), and the
# temperature to ~1e-5 K.  (ii) The reference solves every Newton system inexactly (GMRES + ILU(0) on the segregated [P | T] ordering,
# KSP rtol 1e-5), so its last iterate is only as converged as ||F|| <= 1e-8 ||F0|| = 3.3e-6 W demands.  With constant density the mass
# equation sees neither effect -- no numerical derivative survives, its Newton systems are tridiagonal (ILU(0) exact) -- and there the
# restatement reproduces EVERY printed digit of the baseline's pressures.  Measured temperature deviation: 2.1e-5 K at most.
TH_MMS_T_ABS = 5.0e-5


def test_th_mms_reproduces_reference_baseline(oracle, golden):
    """regression_tests/th/th_mms.regression.baseline (src/driver/standalone/thermal-e/th_mms_problem.F90: manufactured steady state,
    20 cells along x, constant density, IFC-67 enthalpy, per-cell permeability in BOTH governing equations, Dirichlet P and T at both ends,
    per-cell mass and heat sources built from the driver's own finite-difference formulas)."""
    p, ids, d = PB.build_th_mms(oracle.OracleTH, oracle.OraclePhysics())
    conv, reason, P, T = PB.run_th_mms(p, ids, d)
    assert conv and reason == 3
    for name, data, check in (("liquid_pressure", P, lambda a, b: "%.13E" % a == "%.13E" % b or abs(a - b) <= 1e-8),
                              ("temperature", T, lambda a, b: abs(a - b) <= TH_MMS_T_ABS)):
        for key, val in golden["th_mms"][name].items():
            if key == "category":
                continue
            ours = {"min": data.min(), "max": data.max(), "mean": data.sum() / data.size}.get(key)
            if ours is None:
                ours = data[int(key.split()[1]) - 1]
            assert check(ours, val), (name, key, ours, val)
    # and the discrete solution sits where a second-order scheme on 20 cells should: within 1 % of the manufactured fields
    assert np.max(np.abs(P - d["P_exact"])) < 1e-2 * 15000.0 * 7 and np.max(np.abs(T - d["T_exact"])) < 0.3


@pytest.mark.parametrize("dens,iee", [(K.DENSITY_TGDPB01, K.INT_ENERGY_ENTHALPY_CONSTANT), (K.DENSITY_IFC67, K.INT_ENERGY_ENTHALPY_IFC67)])
def test_th_analytic_jacobian_blocks_vs_finite_differences(oracle, dens, iee):
    """The 2x2 block-tridiagonal Jacobian restated from GoveqnRichards...:1941-2200, 2333-2613 and
    GoveqnThermalEnthalpySoilType.F90:1223-1295, 1501-1716, 1847-2377 against central differences of the residual."""
    ncol, nlev = 3, 15
    d = PB.elm_th_inputs(ncol, nlev, density_type=dens, iee_type=iee)
    o, ids = PB.build_elm_th(oracle.OracleTH, d)
    PB.elm_th_step(o, ids, d, 1800.0, 1)                       # loads the conditions; gives a non-trivial state
    n = ncol * nlev
    x = np.empty(2 * n); x[0::2] = o.get_data(K.AUXVAR_INTERNAL, K.VAR_PRESSURE, 1); x[1::2] = o.get_data(K.AUXVAR_INTERNAL, K.VAR_TEMPERATURE, 1, ieqn=2)
    rng = np.random.default_rng(5)
    x[0::2] += rng.uniform(-50.0, 50.0, n); x[1::2] += rng.uniform(-0.5, 0.5, n)
    f, ja, jb, jc = o.eval(1800.0, x, x)
    J = np.zeros((2 * n, 2 * n))
    for c in range(n):
        for r in range(2):
            for cc in range(2):
                J[2 * c + r, 2 * c + cc] = jb[4 * c + 2 * r + cc]
                if c % nlev > 0:
                    J[2 * c + r, 2 * (c - 1) + cc] = ja[4 * c + 2 * r + cc]
                if c % nlev < nlev - 1:
                    J[2 * c + r, 2 * (c + 1) + cc] = jc[4 * c + 2 * r + cc]
    Jfd = np.zeros_like(J)
    for k in range(2 * n):
        h = 1e-2 if k % 2 == 0 else 1e-5
        xp, xm = x.copy(), x.copy(); xp[k] += h; xm[k] -= h
        fp = o.eval(1800.0, x, xp)[0]; fm = o.eval(1800.0, x, xm)[0]
        Jfd[:, k] = (fp - fm) / (2 * h)
    scale = np.abs(J).max(axis=1, keepdims=True)
    # the reference's dFT/dT uses the true temperature derivative of the mass flux but its dFP/dP keeps upwinded kr
    # frozen at the switch, so compare away from upwind switches: relative to the row scale
    assert np.max(np.abs(J - Jfd) / scale) < 5e-5


def test_snow_ssw_soil_oracle_reduces_to_pinned_soil_oracle(oracle):
    """The snow / standing-water extension has no reference baseline (parity unpinned); with both inactive it must reproduce the
    soil-only restatement, which is pinned to thermal_mms."""
    import problems as PB
    ncol, nlev, nsno = 60, 15, 5
    d = PB.elm_snow_thermal_inputs(ncol, nlev, nsno, snow="none", water="none")
    o = PB.pack_elm_snow_thermal(d)
    p = PB.build_elm_snow_thermal(oracle.OracleThermalSnow, d)
    conv, T = PB.elm_snow_thermal_step(p, o)
    off = ncol * (nsno + 1)
    q, ids = PB.build_elm_thermal(oracle.OracleThermal, d)
    d2 = dict(d); d2["tuning"] = o["tuning"][off:]; d2["frac"] = o["frac_soil"]; d2["sabg"] = o["sabg_soil"]
    conv2, T2 = PB.elm_thermal_step(q, ids, d2, d["T0"])
    assert np.max(np.abs(T[off:] - T2) / T2) < 1e-14
    assert np.all(T[:off] == 0.0)


def test_snow_ssw_soil_oracle_energy_budget(oracle):
    """Independent check of the restated coupled system: with cnfac = 0 (fully implicit) each active row is
    C (T_new - T_old) = sum of the implicit fluxes + sources; summing the rows weighted as the reference weights them is not
    conservative (the snow->soil and soil->snow links differ by design), so check the rows of an all-snow, no-water column
    set through a hand-built dense system instead."""
    import problems as PB
    ncol, nlev, nsno = 6, 15, 5
    d = PB.elm_snow_thermal_inputs(ncol, nlev, nsno, snow="all", water="none")
    o = PB.pack_elm_snow_thermal(d)
    p = PB.build_elm_snow_thermal(oracle.OracleThermalSnow, d)
    conv, T = PB.elm_snow_thermal_step(p, o)
    # interior snow rows (neither top nor bottom active layer): C (Tn - To) = cnfac*(F_up - F_dn)(old) + (1-cnfac)*(F_up - F_dn)(new) + sabg
    tkair, tkice, cpliq, cpice = 0.023, 2.29, 4.188e3, 2.11727e3
    checked = 0
    for c in range(ncol):
        ns = -d["snl"][c]
        for k in range(nsno - ns + 1, nsno - 1):
            i = c * nsno + k
            def cond(ii):
                bw = (o["ice"][ii] + o["liq"][ii]) / (o["frac"][ii] * o["dz"][ii])
                return tkair + (7.75e-5 * bw + 1.105e-6 * bw * bw) * (tkice - tkair)
            def flux(Tv, up, dn):
                du, dd = o["dist_up"][up], o["dist_dn"][dn]
                ku, kd = cond(up), cond(dn)
                kk = ku * kd * (du + dd) / (ku * dd + kd * du)
                return -kk * (Tv[up] - Tv[dn]) / (du + dd)
            C_ = max(1e-6, (cpliq * o["liq"][i] + cpice * o["ice"][i]) / o["frac"][i]) / o["dz"][i] * o["dz"][i] / (1800.0 * o["tuning"][i])
            lhs = C_ * (T[i] - o["T"][i])
            rhs = 0.5 * (flux(o["T"], i, i + 1) - flux(o["T"], i - 1, i)) + 0.5 * (flux(T, i, i + 1) - flux(T, i - 1, i)) + o["sabg_snow"][i]
            assert abs(lhs - rhs) <= 1e-9 * max(abs(lhs), abs(rhs), 1.0), (c, k, lhs, rhs)
            checked += 1
    assert checked > 0


def test_elm_driver_restatement_agrees_with_an_independent_numpy_packing(oracle):
    """orc_vsfm_elm_solve (the C restatement of MPPVSFMALM_Solve) against the plain SetData / StepDT / GetData sequence fed with
    sources packed here in numpy from the same raw ELM arrays (MPPVSFMALM_Driver.F90:325-372, 404, 435-450)."""
    import problems as PB
    ncol, nlev, nlevsoi, dt = 80, 15, 10, 1800.0
    d = PB.elm_vsfm_inputs(ncol)
    a, aids = PB.build_elm_vsfm(oracle.OracleVSFM, d, per_column=True)
    b, bids = PB.build_elm_vsfm(oracle.OracleVSFM, d, per_column=True)
    st = PB.elm_vsfm_raw_state(a, d, patches=False)
    raw = PB.copy_state(st)
    a.elm_set_geometry(st["zi"], st["dz"], nlevsoi, aids)
    out = a.elm_solve(dt, st)
    assert out["nfailed"] == 0 and out["iter_count"].max() == 1
    # --- numpy packing ---
    conv = 1.0 * 1000.0 * 1.0e-3
    et = np.zeros((ncol, nlev)); et[:, :nlevsoi] = -raw["qflx_tran_veg_col"][:, None] * raw["rootr_col"][:, :nlevsoi] * conv
    nosnow = raw["snl"] >= 0
    dew = np.where(nosnow, (raw["qflx_dew_snow"] + raw["qflx_dew_grnd"]) * (1.0 - raw["frac_h2osfc"]) * conv, 0.0)
    sub = np.where(nosnow, -raw["qflx_sub_snow"] * (1.0 - raw["frac_h2osfc"]) * conv, 0.0)
    drain = np.zeros((ncol, nlev)); qd_new = raw["qflx_drain"].copy()
    zi, dz = raw["zi"], raw["dz"]
    for c in range(ncol):
        if raw["qflx_drain"][c] > 0.0:
            hits = np.nonzero(raw["zwt"][c] <= zi[c, 1:])[0]
            jwt = max((hits[0] + 1) - 1 if hits.size else nlev, 1)
            js = np.arange(jwt, nlev + 1)
            ql = raw["qflx_drain"][c] * dz[c, js - 1] / dz[c, js - 1].sum()
            ql = np.minimum(ql, (raw["h2osoi_liq"][c, js - 1] - 0.01) / dt)
            drain[c, js - 1] = -ql * conv
            qd_new[c] = ql.sum()
    drain += raw["mflx_drain_perched"]
    snow = raw["mflx_snowlyr_col"] + raw["mflx_neg_snow_col"]
    fliq = 1.0 - raw["h2osoi_ice"] / (raw["h2osoi_liq"] + raw["h2osoi_ice"])
    for name, val in (("infil", raw["qflx_infl"] * conv), ("et", et.reshape(-1)), ("dew", dew), ("drain", drain.reshape(-1)), ("snow", snow), ("sublim", sub)):
        b.set_data(K.AUXVAR_SS, K.VAR_BC_SS_CONDITION, bids[name], val)
    b.set_data(K.AUXVAR_INTERNAL, K.VAR_FRAC_LIQ_SAT, 1, fliq.reshape(-1))
    b.pre_step_dt(); conv_b, reason = b.step_dt(dt, 1); b.post_step_dt()
    assert conv_b
    P = b.get_data(K.AUXVAR_INTERNAL, K.VAR_PRESSURE, 1)
    mass = b.get_data(K.AUXVAR_INTERNAL, K.VAR_MASS, 1).reshape(ncol, nlev)
    smp = b.get_data(K.AUXVAR_INTERNAL, K.VAR_SOIL_MATRIX_POT, 1)
    # summation order of the drained thickness differs (numpy pairwise vs the driver's running sum): 1e-13, not bitwise
    assert np.max(np.abs(out["soilp_col"] - P) / np.maximum(np.abs(P), 1e4)) < 1e-12
    assert np.max(np.abs(out["smp_l"] - smp * 1000.0) / np.maximum(np.abs(smp * 1000.0), 1e3)) < 1e-11
    assert np.max(np.abs(st["h2osoi_liq"] + st["h2osoi_ice"] - mass) / mass) < 1e-12
    assert np.max(np.abs(st["qflx_drain"] - qd_new)) < 1e-18 + 1e-12 * np.max(qd_new)


@pytest.mark.parametrize("problem", ["drying", "wetting"])
def test_sy1991_layered_column_relaxes_to_the_new_steady_state(oracle, problem):
    """src/driver/standalone/vsfm/vsfm_sy1991_problem.F90 (Srivastava & Yeh 1991; no regression baseline in the reference): the driver
    restated with its own initial-pressure tables (tests/golden/sy1991_ic.json, extracted by make_golden.py).  Each problem starts from
    the steady state of the OTHER recharge rate, so each table doubles as the answer of the other run: the head at the top must move
    monotonically towards the other table's value, the bottom cell stays pinned by the Dirichlet head, every step converges, and with
    the run extended the column arrives at the other table (which the reference's authors computed independently)."""
    import json
    ic = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "sy1991_ic.json")))
    start = np.array(ic["press_ic_%s" % problem]); other = np.array(ic["press_ic_%s" % ("wetting" if problem == "drying" else "drying")])
    p, top, bot = PB.build_sy1991(oracle.OracleVSFM, start, per_column=False)
    P24, sat, its = PB.run_sy1991(p, top, bot, start, problem, nstep=24)
    assert max(its) <= 12 and np.all((sat > 0.15) & (sat <= 1.0))
    d0, d24 = abs(start[-1] - other[-1]), abs(P24[-1] - other[-1])
    assert d24 < d0                                                      # the top of the column moves towards the new steady state
    assert abs(P24[0] - start[0]) < 60.0                                 # first cell: 5 mm above the Dirichlet face (rho g 0.005 m = 49 Pa)
    P_long, _, _ = PB.run_sy1991(p, top, bot, start, problem, nstep=24 * 40, first_step=25)
    print(problem, "24 h: top %.2f -> %.2f (target %.2f); long run max |P - table| = %.3f Pa" % (start[-1], P24[-1], other[-1], np.max(np.abs(P_long - other))))
    # The two drivers hold the bottom FACE at their own table's first value (101320.2 vs 101281.1 Pa), so the two steady states differ by
    # up to that much in the low-permeability half; in the upper (ten times more permeable) half the column must land on the other table
    # to its printed precision plus what 997.16 kg/m3 vs the Tanaka density does over 2 m.
    top_half = slice(100, 200)
    print(problem, "upper half max |P - table| = %.3f Pa" % np.max(np.abs(P_long - other)[top_half]))
    assert np.max(np.abs(P_long - other)[top_half]) < 2.0, np.max(np.abs(P_long - other)[top_half])
    assert np.max(np.abs(P_long - other)) < 50.0, np.max(np.abs(P_long - other))
