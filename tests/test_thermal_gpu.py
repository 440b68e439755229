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
