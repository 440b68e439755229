"""Soil thermal (KSP path) parity: CUDA path through the C ABI vs the oracle and the reference's thermal_mms baseline."""
import numpy as np
import pytest

import problems as PB
from mpp_b200 import constants as K

pytestmark = pytest.mark.gpu
RTOL = 1e-10


def relmax(a, b):
    return float(np.max(np.abs(a - b) / np.maximum(np.abs(b), 1e-300)))


@pytest.fixture(scope="module")
def mpp():
    import mpp_b200
    from mpp_b200._lib import lib
    assert lib().mppgpu_device_count() > 0
    return mpp_b200


def test_thermal_mms_vs_reference_baseline(mpp, golden, oracle):
    # regression_tests/thermal/thermal_mms.regression.baseline; nx = 20 -> thermal_step_kernel<32>
    T = PB.run_thermal_mms(PB.build_thermal_mms(mpp.Thermal))
    To = PB.run_thermal_mms(PB.build_thermal_mms(oracle.OracleThermal))
    assert relmax(T, To) < RTOL
    ref = golden["thermal_mms"]["temperature"]

    def ok(v, r):
        return abs(v - r) <= 0.51 * 10.0 ** (np.floor(np.log10(abs(r))) - 12)
    assert ok(T.min(), ref["min"]) and ok(T.max(), ref["max"]) and ok(T.mean(), ref["mean"])
    for key, val in ref.items():
        if key.startswith("cell"):
            assert ok(T[int(key.split()[1]) - 1], val), key


@pytest.mark.parametrize("nx", [5, 16, 40, 100])
def test_thermal_mms_other_lengths(mpp, oracle, nx):
    # 40, 100 > 32 layers -> thermal_step_generic_kernel
    T = PB.run_thermal_mms(PB.build_thermal_mms(mpp.Thermal, nx=nx))
    To = PB.run_thermal_mms(PB.build_thermal_mms(oracle.OracleThermal, nx=nx))
    assert relmax(T, To) < RTOL
    x = (np.arange(nx) + 0.5) / nx
    assert np.max(np.abs(T - (10 * np.sin(np.pi * x) + 270.0))) < 0.5 * (20.0 / nx) ** 2 + 1e-9     # converges to the manufactured solution


@pytest.mark.parametrize("ncol,nlev", [(5000, 15), (4737, 10), (6016, 16)])
def test_bulk_async_thermal_kernel_is_bit_identical_and_matches_oracle(mpp, oracle, ncol, nlev):
    """Batches that fill the GPU run on the persistent kernel whose inputs arrive by bulk-async copies (cp.async.bulk into shared-memory
    stages, thermal_step2_tma_kernel); the arithmetic is the same device function as the register-load kernel, so the two are
    bit-identical -- including the last, partially filled tile (ncol not a multiple of 16) -- and both agree with the oracle."""
    d = PB.elm_thermal_inputs(ncol, nlev, nlevsoi=max(1, min(10, nlev - 2)))
    res = {}
    for mode in (1, 0):
        p, ids = PB.build_elm_thermal(mpp.Thermal, d)
        p.set_bulk_copy(mode)
        T = d["T0"].copy()
        hist = []
        for step in range(3):
            conv, T = PB.elm_thermal_step(p, ids, d, T, 1800.0, step + 1)
            assert conv
            hist.append(T.copy())
        p.step_dt(1800.0, 4)                      # device-resident chain (soln -> soln_prev), no new mailbox data
        hist.append(p.get_soln())
        res[mode] = hist
    for a, b in zip(res[1], res[0]):
        assert np.array_equal(a, b)
    o, oids = PB.build_elm_thermal(oracle.OracleThermal, d, nthreads=8)
    To = d["T0"].copy()
    for step in range(3):
        convo, To = PB.elm_thermal_step(o, oids, d, To, 1800.0, step + 1)
        assert relmax(res[1][step], To) < RTOL, step


@pytest.mark.parametrize("ncol,nlev,varying", [(1, 15, False), (127, 15, False), (128, 15, True), (1000, 15, False), (37, 10, True),
                                               (9, 24, False), (33, 24, True), (5, 16, True), (3, 2, False)])
def test_elm_like_thermal_batch_matches_oracle(mpp, oracle, ncol, nlev, varying):
    # nlev <= 16: two cells per lane; 17..32: one cell per lane.  `varying`: connection distances differ from column to column
    # (per-cell arrays in HBM); otherwise they are uniform and the kernels read them per layer from their constant bank
    d = PB.elm_thermal_inputs(ncol, nlev, nlevsoi=max(1, min(10, nlev - 2)))
    if varying:
        rng = np.random.default_rng(ncol)
        f = rng.uniform(0.8, 1.2, (ncol, 1))
        d["dist_up"] = d["dist_up"] * f; d["dist_dn"] = d["dist_dn"] * f
    p, ids = PB.build_elm_thermal(mpp.Thermal, d)
    o, oids = PB.build_elm_thermal(oracle.OracleThermal, d, nthreads=4)
    T, To = d["T0"].copy(), d["T0"].copy()
    for step in range(4):
        conv, T = PB.elm_thermal_step(p, ids, d, T, 1800.0, step + 1)
        convo, To = PB.elm_thermal_step(o, oids, d, To, 1800.0, step + 1)
        assert conv and convo
        assert relmax(T, To) < RTOL, (ncol, nlev, step)
    assert np.all(np.isfinite(T)) and T.min() > 200.0 and T.max() < 350.0


def test_thermal_energy_conservation(mpp):
    """Crank-Nicolson step with a pure heat-flux BC: the change of stored heat equals the boundary heat input."""
    ncol, nlev = 512, 15
    d = PB.elm_thermal_inputs(ncol, nlev)
    d["dhsdT"] = np.zeros(ncol)                     # flux independent of T so the balance closes exactly
    p, ids = PB.build_elm_thermal(mpp.Thermal, d)
    p.get_data  # noqa
    T0 = d["T0"].copy()
    conv, T1 = PB.elm_thermal_step(p, ids, d, T0, 1800.0, 1)
    # heat capacity per unit area as the reference computes it (ThermalKSPTemperatureSoilAuxType.F90:120-134)
    por, csol = d["watsat"], d["csol"]
    liq, ice = d["liq"].reshape(ncol, nlev), d["ice"].reshape(ncol, nlev)
    hc = csol * (1 - por) * d["dz"] + ice * 2.11727e3 + liq * 4.188e3
    dE = (hc * (T1.reshape(ncol, nlev) - T0.reshape(ncol, nlev))).sum(1)
    assert np.max(np.abs(dE - d["hs"] * 1800.0)) < 1e-6 * np.max(np.abs(d["hs"] * 1800.0))


def test_thermal_pre_step_rollback_and_chain(mpp):
    d = PB.elm_thermal_inputs(64, 15)
    p, ids = PB.build_elm_thermal(mpp.Thermal, d)
    conv, T1 = PB.elm_thermal_step(p, ids, d, d["T0"], 1800.0, 1)
    conv, T1b = PB.elm_thermal_step(p, ids, d, d["T0"], 1800.0, 1)     # SetSolnPrevCLM + PreStepDT again: same answer
    assert np.array_equal(T1, T1b)
    p.step_dt(1800.0, 2)                                                # without PreStepDT: continues from soln
    T2 = p.get_soln()
    conv, T2b = PB.elm_thermal_step(p, ids, d, T1, 1800.0, 2)
    assert np.array_equal(T2, T2b)


def test_thermal_error_behaviour(mpp):
    p = mpp.Thermal(4, 15)
    with pytest.raises(mpp.MPPError):
        p.step_dt(1800.0, 1)
    d = PB.elm_thermal_inputs(4, 15)
    p.set_mesh(K.MESH_ALONG_GRAVITY, d["dz"], d["area"])
    with pytest.raises(mpp.MPPError):
        p.add_condition(1, K.COND_BC, K.COND_MASS_RATE, K.SOIL_TOP_CELLS)
    with pytest.raises(mpp.MPPError):
        p.set_data(K.AUXVAR_INTERNAL, K.VAR_PRESSURE, 1, np.zeros(60))


# ---- snow + standing surface water + soil (MPPThermalTBasedALM_{Initialize,Driver}.F90) ----------------------
def _snow_advance(d, o, T, rng):
    """Feed the solution back as ELM would (t_soisno / t_h2osfc) and perturb the forcing for the next step."""
    ncol, nlev, nsno = d["ncol"], d["nlev"], d["nlevsno"]
    act = o["active"] == 1
    Tn = np.where(act, T, 273.15)
    d["t_snow"] = np.where(d["snow_dz"] > 0, Tn[:ncol * nsno].reshape(ncol, nsno), 0.0)
    d["t_h2osfc"] = np.where(d["frac_h2osfc"] > 0, Tn[ncol * nsno:ncol * (nsno + 1)], d["t_h2osfc"])
    d["t_soil"] = Tn[ncol * (nsno + 1):].reshape(ncol, nlev)
    d["hs_top_snow"] = d["hs_top_snow"] + rng.uniform(-5.0, 5.0, ncol)
    d["hs_soil"] = d["hs_soil"] + rng.uniform(-5.0, 5.0, ncol)


@pytest.mark.parametrize("ncol,nlev,nsno,snow,water", [(300, 15, 5, "mixed", "mixed"), (64, 15, 5, "all", "all"), (33, 10, 3, "mixed", "none"),
                                                       (5, 26, 5, "mixed", "mixed"), (129, 15, 0, "none", "mixed")])
def test_snow_ssw_soil_thermal_matches_oracle(mpp, oracle, ncol, nlev, nsno, snow, water):
    d = PB.elm_snow_thermal_inputs(ncol, nlev, nsno, snow=snow, water=water, nlevsoi=min(10, nlev))
    g = PB.build_elm_snow_thermal(mpp.ThermalSnow, d)
    r = PB.build_elm_snow_thermal(oracle.OracleThermalSnow, d, nthreads=4)
    rng = np.random.default_rng(3)
    for step in range(3):
        o = PB.pack_elm_snow_thermal(d)
        conv, T = PB.elm_snow_thermal_step(g, o, 1800.0, step + 1)
        convo, To = PB.elm_snow_thermal_step(r, o, 1800.0, step + 1)
        assert conv and convo
        act = o["active"] == 1
        assert relmax(T[act], To[act]) < RTOL, (step, relmax(T[act], To[act]))
        assert np.all(T[~act] == 0.0) and np.all(To[~act] == 0.0)          # identity rows with a zero right-hand side
        assert np.abs(T[act] - o["T"][act]).max() > 1e-3                    # the step did something
        _snow_advance(d, o, To, rng)


def test_snow_mode_without_snow_or_water_equals_soil_only_kernel(mpp):
    ncol, nlev, nsno = 200, 15, 5
    d = PB.elm_snow_thermal_inputs(ncol, nlev, nsno, snow="none", water="none")
    o = PB.pack_elm_snow_thermal(d)
    g = PB.build_elm_snow_thermal(mpp.ThermalSnow, d)
    conv, T = PB.elm_snow_thermal_step(g, o)
    off = ncol * (nsno + 1)
    s, ids = PB.build_elm_thermal(mpp.Thermal, d)
    d2 = dict(d); d2["tuning"] = o["tuning"][off:]; d2["frac"] = o["frac_soil"]; d2["sabg"] = o["sabg_soil"]
    conv2, T2 = PB.elm_thermal_step(s, ids, d2, d["T0"])
    assert relmax(T[off:], T2) < 1e-12
    assert np.all(T[:off] == 0.0)


def test_snow_layer_count_changes_between_steps(mpp, oracle):
    """The heat-flux condition follows the top ACTIVE snow layer (ThermKSPTempSnowUpdateBoundaryConn): grow, shrink and lose the pack."""
    ncol = 40
    d = PB.elm_snow_thermal_inputs(ncol, 15, 5, snow="all", water="none")
    g = PB.build_elm_snow_thermal(mpp.ThermalSnow, d)
    r = PB.build_elm_snow_thermal(oracle.OracleThermalSnow, d)
    rng = np.random.default_rng(11)
    for step, mode in enumerate(("all", "mixed", "none", "all")):
        d2 = PB.elm_snow_thermal_inputs(ncol, 15, 5, seed=PB.SEED + step, snow=mode, water="none")
        for k in ("snl", "snow_dz", "snow_z", "snow_zi", "frac_sno_eff", "snow_liq", "snow_ice", "h2osno", "t_snow"):
            d[k] = d2[k]
        o = PB.pack_elm_snow_thermal(d)
        conv, T = PB.elm_snow_thermal_step(g, o, 1800.0, step + 1)
        convo, To = PB.elm_snow_thermal_step(r, o, 1800.0, step + 1)
        act = o["active"] == 1
        assert relmax(T[act], To[act]) < RTOL, (step, mode)
        _snow_advance(d, o, To, rng)


def test_snow_mode_error_behaviour(mpp):
    d = PB.elm_snow_thermal_inputs(4, 15, 5)
    g = PB.build_elm_snow_thermal(mpp.ThermalSnow, d)
    with pytest.raises(mpp.MPPError):
        g.add_condition(1, K.COND_BC, K.COND_HEAT_FLUX, K.SOIL_TOP_CELLS)          # the configuration brings its own conditions
    with pytest.raises(mpp.MPPError):
        g.set_data(K.AUXVAR_INTERNAL, K.VAR_LIQ_AREAL_DEN, 1, np.zeros(g.ncells + 1))
    with pytest.raises(mpp.MPPError):
        g.set_data(K.AUXVAR_BC, K.VAR_BC_SS_CONDITION, 4, np.zeros(4))
    with pytest.raises(mpp.MPPError):
        g.restart(np.zeros(4 * 15))                                                  # soil-only length
    big = mpp.ThermalSnow(2, 30, 5)
    with pytest.raises(mpp.MPPError):                                                # 36 rows per column
        big.set_mesh(np.ones((2, 30)), np.ones(2), np.ones((2, 29)), np.ones((2, 29)), np.ones(2))


@pytest.mark.parametrize("snow", [False, True])
def test_thermal_landunit_types(mpp, oracle, snow):
    """ThermKSPTempSoilAuxVarCompute branches on the landunit type (ThermalKSPTemperatureSoilAuxType.F90:71-171): soil / crop (Johansen
    conductivity, bedrock below nlevsoi), wetland (ice / water conductivity, no mineral heat capacity), land ice and ice_mec."""
    ncol, nlev, nsno = 120, 15, 5
    types = np.array([K.ISTSOIL, K.ISTCROP, K.ISTICE, K.ISTICE_MEC, K.ISTWET], dtype=np.int32)
    if snow:
        d = PB.elm_snow_thermal_inputs(ncol, nlev, nsno)
        d["lun_type"] = types[np.arange(ncol) % 5]
        g = PB.build_elm_snow_thermal(mpp.ThermalSnow, d)
        r = PB.build_elm_snow_thermal(oracle.OracleThermalSnow, d)
        o = PB.pack_elm_snow_thermal(d)
        conv, T = PB.elm_snow_thermal_step(g, o)
        convo, To = PB.elm_snow_thermal_step(r, o)
        act = o["active"] == 1
        assert relmax(T[act], To[act]) < RTOL
    else:
        d = PB.elm_thermal_inputs(ncol, nlev)
        d["lun_type"] = types[np.arange(ncol) % 5]
        d["snow_water"] = np.where(np.arange(ncol * nlev) % nlev == 0, 3.0, 0.0)       # h2osno enters the top layer's heat capacity when snl = 0
        g, ids = PB.build_elm_thermal(mpp.Thermal, d)
        r, rids = PB.build_elm_thermal(oracle.OracleThermal, d)
        T, To = d["T0"].copy(), d["T0"].copy()
        for step in range(2):
            conv, T = PB.elm_thermal_step(g, ids, d, T, 1800.0, step + 1)
            convo, To = PB.elm_thermal_step(r, rids, d, To, 1800.0, step + 1)
            assert relmax(T, To) < RTOL, step
    # the types really take different branches: same forcing, different temperatures
    Tt = (T[-ncol * nlev:] if snow else T).reshape(ncol, nlev)
    assert np.abs(Tt[0::5].mean() - Tt[2::5].mean()) > 1e-3


@pytest.mark.parametrize("ncol,snow,water", [(257, "mixed", "mixed"), (64, "none", "none"), (100, "all", "all")])
def test_thermal_elm_solve_raw_arrays(mpp, oracle, ncol, snow, water):
    """mppgpu_thermal_elm_solve: ELM's (c, j) arrays in, tvector out, against the driver's packing restated in Python
    (pack_elm_snow_thermal / unpack_elm_snow_thermal = MPPThermalTBasedALM_Driver.F90:204-330, 460-505) around the oracle's StepDT."""
    nlev, nsno = 15, 5
    d = PB.elm_snow_thermal_inputs(ncol, nlev, nsno, snow=snow, water=water)
    g = PB.build_elm_snow_thermal(mpp.ThermalSnow, d)
    r = PB.build_elm_snow_thermal(oracle.OracleThermalSnow, d, nthreads=4)
    rng = np.random.default_rng(17)
    for step in range(2):
        e = PB.elm_thermal_raw_arrays(d)
        tv = g.elm_solve(1800.0, e, step + 1)
        o = PB.pack_elm_snow_thermal(d)
        conv, To = PB.elm_snow_thermal_step(r, o, 1800.0, step + 1)
        tvo = PB.unpack_elm_snow_thermal(d, o, To, np.full((nsno + 1 + nlev, ncol), -999.0))
        untouched = tvo == -999.0
        assert np.array_equal(tv == -999.0, untouched)                       # only the entries the driver assigns change
        assert relmax(tv[~untouched], tvo[~untouched]) < RTOL, step
        # the packed mailbox itself is bit-identical to the driver's
        for var, key in ((K.VAR_TUNING_FACTOR, "tuning"), (K.VAR_DZ, "dz"), (K.VAR_DIST_UP, "dist_up"), (K.VAR_FRAC, "frac"), (K.VAR_LIQ_AREAL_DEN, "liq")):
            assert np.array_equal(g.get_data(K.AUXVAR_INTERNAL, var, 1), o[key]), key
        assert np.array_equal(g.get_data(K.AUXVAR_BC, K.VAR_FRAC, 3, n=ncol), o["frac_soil"])
        assert np.array_equal(g.get_data(K.AUXVAR_SS, K.VAR_BC_SS_CONDITION, 1, n=ncol * nsno), o["sabg_snow"])
        _snow_advance(d, o, To, rng)


def test_thermal_elm_solve_page_locked_arrays(mpp):
    """The caller's arrays page-locked in place with mppgpu_host_register: same tvector bit for bit."""
    d = PB.elm_snow_thermal_inputs(257, 15, 5)
    a = PB.build_elm_snow_thermal(mpp.ThermalSnow, d)
    b = PB.build_elm_snow_thermal(mpp.ThermalSnow, d)
    e = PB.elm_thermal_raw_arrays(d)
    el = PB.page_aligned_state(e)
    for v in el.values():
        mpp.host_register(v)
    for step in range(2):
        ta = a.elm_solve(1800.0, e, step + 1).copy()
        tb = b.elm_solve(1800.0, el, step + 1)
        assert np.array_equal(ta, tb)
        e["t_soisno"][...] = ta[1:]; el["t_soisno"][...] = tb[1:]
        e["t_soisno"][ta[1:] == -999.0] = 270.0; el["t_soisno"][tb[1:] == -999.0] = 270.0
    for v in el.values():
        mpp.host_unregister(v)


def test_thermal_elm_solve_pipeline_is_bit_identical(mpp):
    """mppgpu_elm_set_pipeline on the thermal SoE: column chunks on three streams (ragged last chunk) and the soil rows of z / dz / zi
    uploaded by the first solve only -- same tvector and the same packed mailbox bit for bit as the unpipelined solve that uploads
    every row every time, while the snow pack (snow rows of z / dz / zi, snl) changes from step to step."""
    ncol, nlev, nsno = 2500, 15, 5
    d = PB.elm_snow_thermal_inputs(ncol, nlev, nsno)
    a = PB.build_elm_snow_thermal(mpp.ThermalSnow, d)
    b = PB.build_elm_snow_thermal(mpp.ThermalSnow, d)
    a.elm_set_pipeline(1)
    b.elm_set_pipeline(3, static_soil_geometry=True)
    rng = np.random.default_rng(23)
    e = PB.elm_thermal_raw_arrays(d)
    el = PB.page_aligned_state(e)
    for v in el.values():
        mpp.host_register(v)
    for step in range(3):
        ta = a.elm_solve(1800.0, e, step + 1).copy()
        if step > 0:
            # what a stale upload would pick up: the soil rows of the caller's z / dz / zi are NOT read again on this handle
            for k, r0 in (("z", nsno), ("dz", nsno), ("zi", nsno + 1)):
                el[k][r0:] = np.nan
        tb = b.elm_solve(1800.0, el, step + 1)
        assert np.array_equal(ta, tb), step
        for var in (K.VAR_TUNING_FACTOR, K.VAR_DZ, K.VAR_DIST_UP, K.VAR_DIST_DN, K.VAR_FRAC, K.VAR_LIQ_AREAL_DEN, K.VAR_TEMPERATURE):
            assert np.array_equal(a.get_data(K.AUXVAR_INTERNAL, var, 1), b.get_data(K.AUXVAR_INTERNAL, var, 1)), var
        # next step: temperatures carried on, the snow pack settles (thinner snow layers: the snow rows of z / dz / zi move)
        for s_ in (e, el):
            s_["t_soisno"][...] = np.where(ta[1:] == -999.0, 270.0, ta[1:])
            shrink = 1.0 - 0.05 * (step + 1)
            s_["dz"][:nsno] *= shrink; s_["z"][:nsno] *= shrink; s_["zi"][:nsno] *= shrink
            s_["hs_soil"] += 3.0
    for v in el.values():
        mpp.host_unregister(v)
    # asking again re-arms the one-time upload
    b.elm_set_pipeline(3, static_soil_geometry=True)
    assert np.isnan(b.elm_solve(1800.0, el, 4)).any()                 # the NaN soil rows are read this time


def test_thermal_elm_solve_error_behaviour(mpp):
    d = PB.elm_thermal_inputs(4, 15)
    p, ids = PB.build_elm_thermal(mpp.Thermal, d)
    d2 = PB.elm_snow_thermal_inputs(4, 15, 5)
    e = PB.elm_thermal_raw_arrays(d2)
    from mpp_b200._lib import ElmThermalColumns
    import ctypes as C
    cols = ElmThermalColumns()
    assert p.L.mppgpu_thermal_elm_solve(p.h, 1800.0, 1, C.byref(cols), 0.34) != 0      # soil-only handle: snow / ssw equations not added
    g = PB.build_elm_snow_thermal(mpp.ThermalSnow, d2)
    with pytest.raises(ValueError):
        bad = dict(e); bad["zi"] = e["z"]
        g.elm_solve(1800.0, bad)
    with pytest.raises(mpp.MPPError):
        g.elm_solve(0.0, e)


@pytest.mark.parametrize("interior", [False, True])
def test_thermal_sparse_mailbox_arrays(mpp, oracle, interior):
    """tuning_factor / snow_water / num_snow_layer: ELM sets them in the top layer only, but the mailbox takes any values; both shapes,
    alternating between steps.  (A kernel variant that read only the top layer of such arrays -- 24 of 104 B per cell less -- was
    slower, 0.436 vs 0.405 ms per 1 Mi columns, and was dropped: the kernel sits at the latency knee, not at the byte count.)"""
    ncol, nlev = 257, 15
    d = PB.elm_thermal_inputs(ncol, nlev)
    rng = np.random.default_rng(9)
    g, ids = PB.build_elm_thermal(mpp.Thermal, d)
    r, rids = PB.build_elm_thermal(oracle.OracleThermal, d)
    T, To = d["T0"].copy(), d["T0"].copy()
    for step in range(3):
        dense = interior and step == 1                       # step 0 sparse, step 1 dense, step 2 sparse again
        tun = np.ones((ncol, nlev)); tun[:, 0] = rng.uniform(1.2, 2.5, ncol)
        sw = np.zeros((ncol, nlev)); sw[:, 0] = rng.uniform(0.0, 8.0, ncol)
        ns = np.zeros((ncol, nlev), dtype=np.int32); ns[:, 0] = rng.integers(0, 3, ncol)
        if dense:
            tun = rng.uniform(0.8, 2.0, (ncol, nlev)); sw = rng.uniform(0.0, 3.0, (ncol, nlev)); ns = rng.integers(0, 2, (ncol, nlev)).astype(np.int32)
        d["tuning"], d["snow_water"], d["nsnow"] = tun.reshape(-1), sw.reshape(-1), ns.reshape(-1)
        conv, T = PB.elm_thermal_step(g, ids, d, T, 1800.0, step + 1)
        convo, To = PB.elm_thermal_step(r, rids, d, To, 1800.0, step + 1)
        assert relmax(T, To) < RTOL, (step, dense)
