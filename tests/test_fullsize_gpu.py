"""Parity at BASELINE.json's full sizes: size-independent properties over the whole batch (per-column mass / energy balance,
determinism, sortedness of nothing -- columns are independent) plus the oracle on a random sample of the batch's own columns."""
import numpy as np
import pytest

import problems as PB
from mpp_b200 import constants as K

pytestmark = pytest.mark.gpu
RTOL = 1e-10


@pytest.fixture(scope="module")
def mpp():
    import mpp_b200
    from mpp_b200._lib import lib
    assert lib().mppgpu_device_count() > 0
    return mpp_b200


def relmax(a, b):
    return float(np.max(np.abs(a - b) / np.maximum(np.abs(b), 1e-300)))


def relmax_p(a, b):
    return float(np.max(np.abs(a - b) / np.maximum(np.abs(b), np.abs(b - K.PRESSURE_REF))))


def _take_columns(d, cols, per_col, per_cell, nlev):
    out = {k: d[k][cols] for k in per_col}
    out.update({k: d[k].reshape(-1, nlev)[cols].reshape(-1) for k in per_cell})
    return out


def test_vsfm_4Mi_columns_mass_balance_determinism_and_sampled_parity(mpp, oracle):
    """configs[3]: 4 Mi columns x 15 layers.  (i) every converged column closes its mass balance to 1e-5 kg
    (MPPVSFMALM_Driver.F90:860-863); (ii) two runs are bitwise identical; (iii) 1024 columns drawn from the batch, run
    through the oracle on their own, agree to 1e-10 after two steps."""
    import bench
    ncol, nlev = 4 * 1024 * 1024, 15
    d = bench.shard_inputs(0, ncol)
    runs = []
    for rep in range(2):
        p, ids = PB.build_elm_vsfm(mpp.VSFM, d)
        bench.set_forcing_host(p, ids, d)
        for step in range(2):
            p.pre_step_dt(); conv, reason = p.step_dt(1800.0, step + 1); p.post_step_dt()
            if step == 0:
                m0 = p.get_data(K.AUXVAR_INTERNAL, K.VAR_MASS, 1).reshape(ncol, nlev).sum(1)
        P = p.get_data(K.AUXVAR_INTERNAL, K.VAR_PRESSURE, 1)
        m1 = p.get_data(K.AUXVAR_INTERNAL, K.VAR_MASS, 1).reshape(ncol, nlev).sum(1)
        st = p.stats()
        runs.append((P, m0, m1, st))
        p.close()
    P, m0, m1, st = runs[0]
    assert np.array_equal(P, runs[1][0]) and np.array_equal(st["newton_its"], runs[1][3]["newton_its"])
    assert conv and (st["reasons"] > 0).all() and (st["dt_cuts"] == 0).all()    # the benchmark batch converges column by column, no dt cut
    q = d["infil"] + d["et"].reshape(ncol, nlev).sum(1)
    err = np.abs(m0 - m1 + q * 1800.0)
    assert err.max() < 1e-5                                              # the reference's gate, MPPVSFMALM_Driver.F90:140
    assert np.isfinite(P).all()
    rng = np.random.default_rng(7)
    cols = np.sort(rng.choice(ncol, 1024, replace=False))
    ds = _take_columns(d, cols, ("dz", "watsat", "hksat", "bsw", "sucsat", "residual_sat", "area", "infil", "dew", "snow", "sublim"),
                       ("press_ic", "et", "drain", "frac_liq"), nlev)
    ds.update(ncol=len(cols), nlev=nlev, satfunc="van_genuchten")
    o, oids = PB.build_elm_vsfm(oracle.OracleVSFM, ds, per_column=True, nthreads=16)
    for step in range(2):
        convo, reasono, outo = PB.elm_vsfm_step(o, oids, ds, 1800.0, step + 1)
    so = o.stats()
    assert convo and (so["dt_cuts"] == 0).all()
    Pg = P.reshape(ncol, nlev)[cols]
    assert relmax_p(Pg, outo["pressure"].reshape(-1, nlev)) < RTOL       # every sampled column
    assert np.mean(st["newton_its"][cols] != so["newton_its"]) < 0.005   # (a convergence test may sit on a rounding edge)


def test_thermal_1Mi_columns_energy_balance_and_sampled_parity(mpp, oracle):
    """configs[1]: 1 Mi columns x 15 layers.  With dhsdT = 0 the column's heat content changes by exactly the surface flux
    times dt (no other source); 512 sampled columns agree with the oracle to 1e-10."""
    ncol, nlev = 1 << 20, 15
    d = PB.elm_thermal_inputs(ncol, nlev)
    d["dhsdT"] = np.zeros(ncol)
    p, ids = PB.build_elm_thermal(mpp.Thermal, d)
    conv, T1 = PB.elm_thermal_step(p, ids, d, d["T0"], 1800.0, 1)
    hc = d["csol"] * (1 - d["watsat"]) * d["dz"] + d["ice"].reshape(ncol, nlev) * 2.11727e3 + d["liq"].reshape(ncol, nlev) * 4.188e3
    dE = (hc * (T1.reshape(ncol, nlev) - d["T0"].reshape(ncol, nlev))).sum(1)
    assert conv and np.max(np.abs(dE - d["hs"] * 1800.0)) < 1e-6 * np.max(np.abs(d["hs"] * 1800.0))
    rng = np.random.default_rng(8)
    cols = np.sort(rng.choice(ncol, 512, replace=False))
    ds = _take_columns(d, cols, ("dz", "area", "dist_up", "dist_dn", "watsat", "csol", "tkmg", "tkdry", "lun_type", "hs", "dhsdT", "frac"),
                       ("ice", "liq", "T0", "snow_water", "nsnow", "tuning", "sabg"), nlev)
    ds.update(ncol=len(cols), nlev=nlev, nlevsoi=d["nlevsoi"])
    o, oids = PB.build_elm_thermal(oracle.OracleThermal, ds, nthreads=8)
    convo, To = PB.elm_thermal_step(o, oids, ds, ds["T0"], 1800.0, 1)
    assert relmax(T1.reshape(ncol, nlev)[cols].reshape(-1), To) < RTOL


def test_th_256Ki_columns_sampled_parity(mpp, oracle):
    """configs[4] per-GPU share (2 Mi columns over 8 GPUs): 256 Ki columns x 15 layers of the benchmark's TH batch (ELM's default curve
    smooth_brooks_corey_bz3), two steps; every column converges, and 512 sampled columns agree with the oracle to 1e-10 -- all of them,
    including any that needed a dt cut."""
    import bench
    ncol, nlev = 1 << 18, 15
    d = bench.shard_inputs_th(0, ncol)
    p, ids = PB.build_elm_th(mpp.TH, d)
    for step in range(2):
        conv, reason, out = PB.elm_th_step(p, ids, d, 1800.0, step + 1)
        assert conv
    st = p.stats()
    assert (st["reasons"] > 0).all() and np.isfinite(out["pressure"]).all() and np.isfinite(out["temperature"]).all()
    rng = np.random.default_rng(9)
    cols = np.sort(np.unique(np.concatenate([rng.choice(ncol, 512, replace=False), np.nonzero(st["dt_cuts"] > 0)[0][:16]])))
    ds = _take_columns(d, cols, ("dz", "watsat", "hksat", "bsw", "sucsat", "residual_sat", "area", "infil", "dew", "snow", "sublim", "csol", "tkdry", "T_top", "P_top_bc"),
                       ("press_ic", "et", "drain", "frac_liq", "temp_ic", "heat"), nlev)
    ds.update(ncol=len(cols), nlev=nlev, satfunc=d["satfunc"], density_type=d["density_type"], iee_type=d["iee_type"])
    o, oids = PB.build_elm_th(oracle.OracleTH, ds, per_column=True, nthreads=8)
    for step in range(2):
        convo, reasono, outo = PB.elm_th_step(o, oids, ds, 1800.0, step + 1)
        assert convo
    so = o.stats()
    assert np.array_equal(so["dt_cuts"], st["dt_cuts"][cols])
    for k in ("pressure", "temperature", "sat"):
        a, b = out[k].reshape(ncol, nlev)[cols], outo[k].reshape(-1, nlev)
        rm = (relmax_p if k == "pressure" else relmax)(a, b)
        assert rm < RTOL, (k, rm)


def test_snow_thermal_1Mi_columns_uniform_state_is_a_fixed_point_and_sampled_parity(mpp, oracle):
    """Snow + standing water + soil at bench size.  (i) With every heat flux, its derivative and the absorbed radiation set to zero, a
    column at one uniform temperature must stay there whatever its snow / water configuration: all conduction terms vanish and each
    row reduces to C T_new = C T_old.  (ii) The oracle on a random sample of the forced batch's own columns."""
    base, reps, nlev, nsno = 4096, 256, 15, 5
    d0 = PB.elm_snow_thermal_inputs(base, nlev, nsno)
    o0 = PB.pack_elm_snow_thermal(d0)
    d, o = PB.tile_snow_thermal(d0, o0, reps)
    ncol = d["ncol"]
    g = PB.build_elm_snow_thermal(mpp.ThermalSnow, d)
    # (ii) forced step, sample = the first `base` columns (the tile itself: every column of the batch is one of them)
    conv, T = PB.elm_snow_thermal_step(g, o)
    r = PB.build_elm_snow_thermal(oracle.OracleThermalSnow, d0, nthreads=8)
    convo, To = PB.elm_snow_thermal_step(r, o0)
    a0, a1, a2 = ncol * nsno, ncol * (nsno + 1), base * nsno
    act0 = o0["active"] == 1
    Tt = np.concatenate([T[:a2], T[a0:a0 + base], T[a1:a1 + base * nlev]])
    assert relmax(Tt[act0], To[act0]) < RTOL
    # every replica of the tile gives the same answer: determinism across the batch
    assert np.array_equal(T[a1:].reshape(reps, base * nlev)[0], T[a1:].reshape(reps, base * nlev)[reps - 1])
    # (i) fixed point
    q = dict(o)
    act = o["active"] == 1
    q["T"] = np.where(act, 268.0, 273.15)
    for k in ("hs_snow", "hs_sh2o", "hs_soil", "dhsdT_snow", "dhsdT_sh2o", "dhsdT_soil", "sabg_snow", "sabg_soil"):
        q[k] = np.zeros_like(o[k])
    conv, T = PB.elm_snow_thermal_step(g, q)
    assert np.max(np.abs(T[act] - 268.0)) < 1e-9
