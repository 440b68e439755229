"""The __host__ __device__ constitutive relations of mpp_b200/csrc/physics.cuh (the code the kernels run),
compiled for the CPU (libmpp_hostcheck.so), against the oracle.  No GPU needed."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from mpp_b200 import constants as K

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
c_dp = C.POINTER(C.c_double)


@pytest.fixture(scope="module")
def hc():
    so = os.path.join(ROOT, "mpp_b200", "libmpp_hostcheck.so")
    src = os.path.join(ROOT, "mpp_b200", "csrc")
    if (not os.path.exists(so)) or os.path.getmtime(os.path.join(src, "physics.cuh")) > os.path.getmtime(so):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-x", "c++", "-o", so,
                               os.path.join(src, "hostcheck.cpp"), "-lm"])
    L = C.CDLL(so)
    L.hc_convert_soil.argtypes = [C.c_int] + [C.c_double] * 5 + [c_dp]
    L.hc_sat.argtypes = [C.c_int, c_dp, C.c_double, C.c_double, c_dp]
    L.hc_density.argtypes = [C.c_int, C.c_double, C.c_double, c_dp]
    L.hc_log_exp.argtypes = [C.c_int, c_dp, c_dp, c_dp]
    L.hc_density_fixedT.argtypes = [C.c_int, C.c_double, C.c_double, c_dp]
    L.hc_enthalpy_ifc67.argtypes = [C.c_double, C.c_double, c_dp]
    L.hc_internal_energy_enthalpy.argtypes = [C.c_int] + [C.c_double] * 5 + [c_dp]
    return L


def rel(a, b):
    return abs(a - b) / max(abs(b), 1e-300)


NAMES = ["van_genuchten", "brooks_corey", "smooth_brooks_corey_bz2", "smooth_brooks_corey_bz3"]


@pytest.mark.parametrize("name", NAMES)
def test_soil_conversion_and_curves_match_oracle(hc, oracle, name):
    rng = np.random.default_rng(11)
    satfunc_name = K.SATFUNC[name]
    kind = {0: 0, 1: 1, 2: 2, 3: 2}[satfunc_name]
    worst, worst_dry = 0.0, 0.0
    for _ in range(200):
        watsat, hksat = rng.uniform(0.35, 0.55), np.exp(rng.uniform(np.log(5e-4), np.log(5e-2)))
        bsw, sucsat, sr = rng.uniform(3, 12), rng.uniform(50, 600), rng.uniform(0, 0.2)
        out = np.zeros(10)
        bad = hc.hc_convert_soil(satfunc_name, watsat, hksat, bsw, sucsat, sr, out.ctypes.data_as(c_dp))
        assert bad == 0
        alpha, lam = 1.0 / (sucsat * K.GRAV), 1.0 / bsw
        sp = oracle.satparams(name, sr, alpha, lam)
        assert out[0] == watsat and out[1] == hksat * 0.001002 / (1000.0 * K.GRAV) * 0.001
        assert out[3] == sp.alpha and out[4] == lam
        if kind == 0:
            assert out[5] == sp.vg_n
        if kind == 2:
            for a, b in zip(out[6:10], (sp.sbc_pu, sp.sbc_ps, sp.sbc_b2, sp.sbc_b3)):
                assert rel(a, b) < 1e-13 or a == b
        params = np.array([out[2], out[3], out[4], out[5], out[6], out[7], out[8], out[9]])
        # pressures from very dry to ponded, and around the SBC smoothing window
        pcs = np.concatenate([-np.exp(rng.uniform(np.log(0.05), np.log(400.0), 30)) / alpha, [0.0, 50.0, -0.95 / alpha, -1.0 / alpha]])
        for pc in pcs:
            press = K.PRESSURE_REF + pc
            for frac in (1.0, 0.6):
                got = np.zeros(4)
                hc.hc_sat(kind, params.ctypes.data_as(c_dp), press, frac, got.ctypes.data_as(c_dp))
                s, ds = oracle.press_to_sat(sp, press)
                k, dk = oracle.press_to_relperm(sp, press, frac)
                for i, (a, b) in enumerate(zip(got, (s, ds, k, dk))):
                    if b == 0.0:
                        assert a == 0.0
                    elif i >= 2 and k < 1e-4:
                        # very dry soil: kr = sqrt(Se) (1 - AA^m)^2 with AA -> 1 cancels in the reference's own
                        # formula (SaturationFunction.F90:832-836); both sides carry ~ulp / (1 - AA^m) of noise
                        worst_dry = max(worst_dry, rel(a, b))
                    else:
                        worst = max(worst, rel(a, b))
    # log/exp restatement of the reference's pow() chains: a few ulp amplified by the exponents (<= ~20)
    assert worst < 2e-13, worst
    assert worst_dry < 1e-10, worst_dry


def test_density_matches_oracle_and_reference_known_answers(hc, oracle, golden):
    g = golden["eos_density"]
    out = np.zeros(3)
    for name, itype in (("constant", 1), ("tgdpb01", 2), ("ifc67", 3)):
        hc.hc_density(itype, g["p"], g["t_K"], out.ctypes.data_as(c_dp))
        assert abs(out[0] - g[name]["den"]) < g["tol"]["den"]
        assert abs(out[1] - g[name]["dden_dp"]) < g["tol"]["dden_dp"]
        assert abs(out[2] - g[name]["dden_dT"]) < 2 * g["tol"]["dden_dT"]
    rng = np.random.default_rng(5)
    for _ in range(300):
        p, t = rng.uniform(2e4, 5e5), rng.uniform(274.0, 320.0)
        for itype in (1, 2, 3):
            hc.hc_density(itype, p, t, out.ctypes.data_as(c_dp))
            ref = oracle.density(p, t, itype)
            for a, b in zip(out, ref):
                assert (a == b) or rel(a, b) < 2e-12
        # VSFM path: temperature pinned at 298.15 K
        o2 = np.zeros(2)
        for itype in (1, 2):
            hc.hc_density_fixedT(itype, 298.15, p, o2.ctypes.data_as(c_dp))
            ref = oracle.density(p, 298.15, itype)
            assert rel(o2[0], ref[0]) < 1e-15 and ((o2[1] == ref[1]) or rel(o2[1], ref[1]) < 1e-14)


def test_enthalpy_matches_oracle(hc, oracle):
    rng = np.random.default_rng(6)
    out = np.zeros(3)
    for _ in range(300):
        p, tc = rng.uniform(5e4, 5e5), rng.uniform(1.0, 60.0)
        hc.hc_enthalpy_ifc67(tc, p, out.ctypes.data_as(c_dp))
        ref = oracle.enthalpy_ifc67(tc, p)
        # IFC-67 sums terms of opposite sign scaled by pc1*vc1mol = 1.26e6 J/kmol: round-off of the integer powers
        # (the reference's pow(theta,18.) vs repeated squaring here) is amplified to ~1e-5 J/kmol, i.e. ~1e-10 K
        assert abs(out[0] - ref[0]) < 2e-4 and rel(out[0], ref[0]) < 1e-9
        assert rel(out[1], ref[1]) < 1e-13
        assert rel(out[2], ref[2]) < 1e-10


def test_lean_log_exp_are_accurate_to_2ulp(hc):
    """mpp_log / mpp_exp of physics.cuh (branch-free, table-driven, constant-bank coefficients) against libm over the ranges
    the soil curves produce: x = -alpha pc in [1e-12, 1e12], 1 + x^n up to 1e300, exponents n L1 and -m L2 / 2 in [-700, 700].
    The table log is log(c_i) + log1p(r): away from 1 it is held to 2.5 ulp; within 2^-6 of log = 0 the two terms cancel and
    the ABSOLUTE error (what the curves see: every log is multiplied by an exponent and fed to exp) is held to 4e-18."""
    rng = np.random.default_rng(17)
    x = np.concatenate([np.exp(rng.uniform(np.log(1e-12), np.log(1e12), 200000)), 1.0 + np.exp(rng.uniform(-40, 690, 100000)),
                        1.0 + rng.uniform(-0.3, 0.45, 100000), [1.0, 2.0, 0.5, np.sqrt(2.0), np.sqrt(0.5), 1e300, 1e-300]])
    lg, ex = np.zeros_like(x), np.zeros_like(x)
    hc.hc_log_exp(x.size, x.ctypes.data_as(c_dp), lg.ctypes.data_as(c_dp), ex.ctypes.data_as(c_dp))
    ref = np.log(x)
    far = np.abs(ref) > 2.0 ** -6
    ulp = np.abs(lg - ref)[far] / np.spacing(np.abs(ref[far]))
    assert ulp.max() <= 2.5, ulp.max()
    assert np.abs(lg - ref)[~far].max() <= 4e-18
    assert abs(lg[-7]) <= 1e-18                                       # log(1)
    # exp on its own argument set (the log output of this call is ignored)
    y = np.concatenate([rng.uniform(-700.0, 700.0, 200000), rng.uniform(-2.0, 2.0, 200000), [0.0, -707.9, 707.9, 1e-300, -1e-300]])
    dummy, ex = np.zeros_like(y), np.zeros_like(y)
    hc.hc_log_exp(y.size, y.ctypes.data_as(c_dp), dummy.ctypes.data_as(c_dp), ex.ctypes.data_as(c_dp))
    ref = np.exp(y)
    ulp = np.abs(ex - ref) / np.spacing(ref)
    assert ulp.max() <= 2.0, ulp.max()
    # clamped tails stay finite and normal
    out, big = np.zeros(2), np.array([-1e6, 1e6])
    hc.hc_log_exp(2, big.ctypes.data_as(c_dp), dummy[:2].ctypes.data_as(c_dp), out.ctypes.data_as(c_dp))
    assert 0.0 < out[0] < 1e-300 and 1e307 < out[1] < np.inf
