"""ctypes binding of the CPU oracle (oracle/libmpp_oracle.so).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs -- never by mpp_b200 (the product).
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

c_dp = C.POINTER(C.c_double)
c_ip = C.POINTER(C.c_int)


def build(force=False):
    so = os.path.join(_HERE, "libmpp_oracle.so")
    srcs = [os.path.join(_HERE, f) for f in os.listdir(_HERE) if f.endswith((".c", ".h"))]
    if force or not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        subprocess.check_call(["make", "-C", _HERE, "-s"], env={**os.environ, "CC": "gcc"})
    return so


_LIB_PATH = None


def use_library(path):
    """Bind to another build of the same sources (bench.py's timing-only -O3 -march=native build); must precede the first lib()."""
    global _LIB, _LIB_PATH
    _LIB, _LIB_PATH = None, path


def lib():
    global _LIB
    if _LIB is None:
        _LIB = C.CDLL(_LIB_PATH or build())
        L = _LIB
        L.orc_vsfm_create.restype = C.c_void_p
        for name in ("orc_thermal_create", "orc_th_create", "orc_thermal3_create"):
            if hasattr(L, name):
                getattr(L, name).restype = C.c_void_p
        L.orc_findgu_sbc_zerocoeff.restype = C.c_double
        L.orc_findgu_sbc_zerocoeff.argtypes = [C.c_double, C.c_int, C.c_double]
    return _LIB


def dp(a):
    return a.ctypes.data_as(c_dp)


def ip(a):
    return a.ctypes.data_as(c_ip)


def f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def table(a, ncol, nlev):
    """(ncol, nlev) array -> Fortran column-major flat buffer t[j*ncol + c] (the reference's (c,j) tables)."""
    a = np.asarray(a, dtype=np.float64)
    if a.shape != (ncol, nlev):
        raise ValueError("table must have shape (ncol, nlev)")
    return np.ascontiguousarray(a.T).reshape(-1)


# ---- scalar physics helpers (for unit tests) ---------------------------------
def density(p, t_K, itype):
    L = lib()
    den, dp_, dt_ = C.c_double(), C.c_double(), C.c_double()
    L.orc_density(C.c_double(p), C.c_double(t_K), C.c_int(itype), C.byref(den), C.byref(dp_), C.byref(dt_))
    return den.value, dp_.value, dt_.value


def enthalpy_ifc67(t_C, p):
    L = lib()
    h, hp, ht = C.c_double(), C.c_double(), C.c_double()
    L.orc_enthalpy_ifc67(C.c_double(t_C), C.c_double(p), C.c_int(1), C.byref(h), C.byref(hp), C.byref(ht))
    return h.value, hp.value, ht.value


class SatParams(C.Structure):
    _fields_ = [("sat_func_type", C.c_int), ("relperm_func_type", C.c_int), ("sat_res", C.c_double),
                ("alpha", C.c_double), ("vg_m", C.c_double), ("vg_n", C.c_double), ("bc_lambda", C.c_double),
                ("sbc_pu", C.c_double), ("sbc_ps", C.c_double), ("sbc_b2", C.c_double), ("sbc_b3", C.c_double)]


def satparams(name, sat_res, alpha, lam):
    L = lib()
    sp = SatParams()
    if name == "van_genuchten":
        rc = L.orc_satfunc_set_vg(C.byref(sp), C.c_double(sat_res), C.c_double(alpha), C.c_double(lam))
    elif name == "brooks_corey":
        rc = L.orc_satfunc_set_bc(C.byref(sp), C.c_double(sat_res), C.c_double(alpha), C.c_double(lam))
    elif name == "smooth_brooks_corey_bz2":
        rc = L.orc_satfunc_set_sbc_bz2(C.byref(sp), C.c_double(sat_res), C.c_double(alpha), C.c_double(lam), C.c_double(-0.9 / alpha))
    elif name == "smooth_brooks_corey_bz3":
        rc = L.orc_satfunc_set_sbc_bz3(C.byref(sp), C.c_double(sat_res), C.c_double(alpha), C.c_double(lam), C.c_double(-0.9 / alpha))
    else:
        raise ValueError(name)
    if rc:
        raise ValueError("bad saturation parameters rc=%d" % rc)
    return sp


def press_to_sat(sp, press):
    s, ds = C.c_double(), C.c_double()
    lib().orc_press_to_sat(C.byref(sp), C.c_double(press), C.byref(s), C.byref(ds))
    return s.value, ds.value


def press_to_relperm(sp, press, frac_liq=1.0):
    k, dk = C.c_double(), C.c_double()
    lib().orc_press_to_relperm(C.byref(sp), C.c_double(press), C.c_double(frac_liq), C.byref(k), C.byref(dk))
    return k.value, dk.value



class OraclePhysics:
    """The reference's scalar EOS / saturation-function calls (EOSWaterMod.F90, SaturationFunction.F90 as restated in eos_water.c /
    satfunc.c) for problem set-up code (manufactured sources of th_mms).  Same five calls as mpp_b200.hostphysics.HostPhysics."""

    def density(self, P, T, itype):
        return density(P, T, itype)

    def viscosity(self, P, T):
        v, a, b = C.c_double(), C.c_double(), C.c_double()
        lib().orc_viscosity(C.c_double(P), C.c_double(T), C.byref(v), C.byref(a), C.byref(b))
        return v.value

    def internal_energy_enthalpy(self, P, T, itype, rho, drho_dT, drho_dP):
        o = [C.c_double() for _ in range(6)]
        lib().orc_internal_energy_enthalpy(C.c_double(P), C.c_double(T), C.c_int(itype), C.c_double(rho), C.c_double(drho_dT), C.c_double(drho_dP),
                                           *[C.byref(x) for x in o])
        return tuple(x.value for x in o)            # U, H, dU_dT, dH_dT, dU_dP, dH_dP

    def vg_sat(self, P, sat_res, alpha, m):
        return press_to_sat(satparams("van_genuchten", sat_res, alpha, m), P)

    def vg_relperm(self, P, sat_res, alpha, m):
        return press_to_relperm(satparams("van_genuchten", sat_res, alpha, m), P, 1.0)

SATFUNC_NAMES = {"van_genuchten": 0, "brooks_corey": 1, "smooth_brooks_corey_bz2": 2, "smooth_brooks_corey_bz3": 3}


class OracleVSFM:
    """Same call surface as mpp_b200.soe.VSFM (the product wrapper) so tests drive both identically."""

    def __init__(self, ncol, nlev, per_column=True, nthreads=1):
        self.L = lib()
        self.ncol, self.nlev, self.ncells = ncol, nlev, ncol * nlev
        self.h = C.c_void_p(self.L.orc_vsfm_create(ncol, nlev))
        self.L.orc_vsfm_set_mode(self.h, int(per_column), int(nthreads))

    def __del__(self):
        try:
            self.L.orc_vsfm_destroy(self.h)
        except Exception:
            pass

    def set_mesh(self, orientation, dz, area, col_active=None):
        dz, area = table(dz, self.ncol, self.nlev), f64(area)
        assert area.size == self.ncol
        ca = i32(col_active) if col_active is not None else None
        return self.L.orc_vsfm_set_mesh(self.h, int(orientation), dp(dz), dp(area), ip(ca) if ca is not None else None)

    def add_condition(self, ieqn, ss_or_bc, cond_type, region):
        return self.L.orc_vsfm_add_condition(self.h, int(ss_or_bc), int(cond_type), int(region))

    def set_soils(self, watsat, hksat, bsw, sucsat, residual_sat, satfunc_type="van_genuchten", density_type=2):
        a = [table(x, self.ncol, self.nlev) for x in (watsat, hksat, bsw, sucsat, residual_sat)]
        rc = self.L.orc_vsfm_set_soils(self.h, *[dp(x) for x in a], SATFUNC_NAMES[satfunc_type], int(density_type))
        if rc:
            raise ValueError("set_soils rc=%d" % rc)
        self._soils_set = True

    def set_tolerances(self, atol, rtol, stol, max_it, max_funcs):
        self.L.orc_vsfm_set_tolerances(self.h, C.c_double(atol), C.c_double(rtol), C.c_double(stol), int(max_it), int(max_funcs))

    def restart(self, press):
        press = f64(press)
        assert press.size == self.ncells
        rc = self.L.orc_vsfm_restart(self.h, dp(press))
        if getattr(self, "_soils_set", False):
            self.L.orc_vsfm_fill_mailbox(self.h)          # as mppgpu_restart does (ELM's initialisation leaves the mailbox consistent)
        return rc

    ELM_COND_ORDER = ("infil", "et", "dew", "drain", "snow", "sublim")

    def elm_set_geometry(self, zi, dz, nlevsoi, ids, watmin=0.01):
        self._zi = np.ascontiguousarray(np.asarray(zi, dtype=np.float64).reshape(self.ncol, self.nlev + 1))
        self._dz = np.ascontiguousarray(np.asarray(dz, dtype=np.float64).reshape(self.ncol, self.nlev))
        self._nlevsoi, self._watmin = int(nlevsoi), float(watmin)
        self._cids = i32([ids[k] for k in self.ELM_COND_ORDER] if isinstance(ids, dict) else list(ids))

    def elm_solve(self, dt, st, nstep=1):
        ncol, n = self.ncol, self.ncells
        out = {"smp_l": np.zeros(n), "soilp_col": np.zeros(n), "qcharge": np.zeros(ncol), "abs_mass_error": np.zeros(ncol),
               "iter_count": np.zeros(ncol, dtype=np.int32), "status": np.zeros(ncol, dtype=np.int32)}
        pat = st.get("col_pfti") is not None
        nul_i, nul_d = C.POINTER(C.c_int)(), C.POINTER(C.c_double)()
        self.L.orc_vsfm_elm_solve.restype = C.c_int
        rc = self.L.orc_vsfm_elm_solve(
            self.h, C.c_double(dt), C.c_int(self._nlevsoi), C.c_double(self._watmin), ip(self._cids),
            C.c_int(int(st["max_patch_per_col"]) if pat else 0),
            ip(st["col_pfti"]) if pat else nul_i, ip(st["col_npfts"]) if pat else nul_i, ip(st["pft_active"]) if pat else nul_i,
            dp(st["pft_wtcol"]) if pat else nul_d, dp(st["rootr_pft"]) if pat else nul_d, dp(st["qflx_tran_veg_pft"]) if pat else nul_d,
            dp(st["rootr_col"]), dp(st["qflx_tran_veg_col"]), dp(st["qflx_infl"]), dp(st["qflx_dew_snow"]), dp(st["qflx_dew_grnd"]),
            dp(st["qflx_sub_snow"]), dp(st["frac_h2osfc"]), ip(st["snl"]), dp(st["qflx_drain"]), dp(st["zwt"]), dp(self._zi), dp(self._dz),
            dp(st["h2osoi_liq"]), dp(st["h2osoi_ice"]), dp(st["mflx_snowlyr_col"]), dp(st["mflx_neg_snow_col"]), dp(st["mflx_drain_perched"]),
            dp(out["smp_l"]), dp(out["soilp_col"]), dp(out["qcharge"]), dp(out["abs_mass_error"]), ip(out["iter_count"]), ip(out["status"]))
        if rc > 0:
            raise ValueError("orc_vsfm_elm_solve rc=%d" % rc)
        out["nfailed"] = -rc
        out["nattempts"] = int(out["iter_count"].max())
        return out

    def set_data(self, auxvar_type, var_type, cond_id, data, ieqn=1):
        data = f64(data)
        rc = self.L.orc_vsfm_set_data(self.h, int(auxvar_type), int(var_type), int(cond_id), dp(data), int(data.size))
        if rc:
            raise ValueError("set_data rc=%d" % rc)

    def get_data(self, auxvar_type, var_type, cond_id, n=None, ieqn=1):
        n = self.ncells if n is None else n
        out = np.empty(n, dtype=np.float64)
        rc = self.L.orc_vsfm_get_data(self.h, int(auxvar_type), int(var_type), int(cond_id), dp(out), int(n))
        if rc:
            raise ValueError("get_data rc=%d" % rc)
        return out

    def pre_step_dt(self):
        self.L.orc_vsfm_pre_step_dt(self.h)

    def post_step_dt(self):
        self.L.orc_vsfm_post_step_dt(self.h)

    def step_dt(self, dt, nstep=1):
        conv, reason = C.c_int(), C.c_int()
        self.L.orc_vsfm_step_dt(self.h, C.c_double(dt), int(nstep), C.byref(conv), C.byref(reason))
        return bool(conv.value), reason.value

    def stats(self):
        its, rs, cuts, nf = (np.zeros(self.ncol, dtype=np.int32) for _ in range(4))
        self.L.orc_vsfm_get_stats(self.h, ip(its), ip(rs), ip(cuts), ip(nf))
        return {"newton_its": its, "reasons": rs, "dt_cuts": cuts, "nfuncs": nf}

    def eval(self, dt, x_prev, x):
        x_prev, x = f64(x_prev), f64(x)
        f, ja, jb, jc = (np.zeros(self.ncells) for _ in range(4))
        self.L.orc_vsfm_eval(self.h, C.c_double(dt), dp(x_prev), dp(x), dp(f), dp(ja), dp(jb), dp(jc))
        return f, ja, jb, jc


class OracleThermal:
    """Same call surface as mpp_b200.soe.Thermal."""

    def __init__(self, ncol, nlev, nthreads=1, **kw):
        self.L = lib()
        self.ncol, self.nlev, self.ncells = ncol, nlev, ncol * nlev
        self.h = C.c_void_p(self.L.orc_thermal_create(ncol, nlev))
        self.L.orc_thermal_set_threads(self.h, int(nthreads))

    def __del__(self):
        try:
            self.L.orc_thermal_destroy(self.h)
        except Exception:
            pass

    def set_mesh(self, orientation, dz, area, col_active=None):
        dz, area = table(dz, self.ncol, self.nlev), f64(area)
        return self.L.orc_thermal_set_mesh(self.h, int(orientation), dp(dz), dp(area), None)

    def set_connection_distances(self, dist_up, dist_dn):
        du, dd = table(dist_up, self.ncol, self.nlev - 1), table(dist_dn, self.ncol, self.nlev - 1)
        return self.L.orc_thermal_set_conn_dist(self.h, dp(du), dp(dd))

    def add_condition(self, ieqn, ss_or_bc, cond_type, region):
        return self.L.orc_thermal_add_condition(self.h, int(ss_or_bc), int(cond_type), int(region))

    def set_soils(self, watsat, csol, tkmg, tkdry, lun_type, nlevsoi, istsoil=1):
        a = [table(x, self.ncol, self.nlev) for x in (watsat, csol, tkmg, tkdry)]
        lt = i32(lun_type)
        return self.L.orc_thermal_set_soils(self.h, *[dp(x) for x in a], ip(lt), int(nlevsoi), int(istsoil))

    def set_cnfac(self, cnfac):
        self.L.orc_thermal_set_cnfac(self.h, C.c_double(cnfac))

    def set_soln_prev(self, T):
        T = f64(T)
        assert T.size == self.ncells
        self.L.orc_thermal_set_soln_prev(self.h, dp(T))

    restart = set_soln_prev

    def set_data(self, auxvar_type, var_type, cond_id, data, ieqn=1):
        data = f64(data)
        rc = self.L.orc_thermal_set_rdata(self.h, int(auxvar_type), int(var_type), int(cond_id), dp(data), int(data.size))
        if rc:
            raise ValueError("set_rdata rc=%d" % rc)

    set_rdata = set_data

    def set_idata(self, auxvar_type, var_type, cond_id, data):
        data = i32(data)
        rc = self.L.orc_thermal_set_idata(self.h, int(auxvar_type), int(var_type), int(cond_id), ip(data), int(data.size))
        if rc:
            raise ValueError("set_idata rc=%d" % rc)

    def pre_step_dt(self):
        self.L.orc_thermal_pre_step_dt(self.h)

    def post_step_dt(self):
        pass

    def step_dt(self, dt, nstep=1):
        conv = C.c_int()
        self.L.orc_thermal_step_dt(self.h, C.c_double(dt), int(nstep), C.byref(conv))
        return bool(conv.value), 0

    def get_soln(self):
        out = np.empty(self.ncells)
        self.L.orc_thermal_get_soln(self.h, dp(out))
        return out

    def get_data(self, auxvar_type, var_type, cond_id, n=None, ieqn=1):
        if var_type == 605:
            return self.get_soln()
        out = np.empty(self.ncells)
        rc = self.L.orc_thermal_get_aux(self.h, int(var_type), dp(out))
        if rc:
            raise ValueError("get_aux rc=%d" % rc)
        return out


class OracleThermalSnow:
    """Snow + standing-surface-water + soil thermal SoE (same call surface as mpp_b200.soe.ThermalSnow).  Arrays of the
    internal aux vars have ncol*(nlevsno+1+nlev) entries in the reference's SoE order [snow | ssw | soil]."""

    def __init__(self, ncol, nlev, nlevsno, nthreads=1, **kw):
        self.L = lib()
        self.ncol, self.nlev, self.nlevsno = ncol, nlev, nlevsno
        self.ncells = ncol * (nlevsno + 1 + nlev)
        self.h = C.c_void_p(self.L.orc_thermal3_create(ncol, nlev, nlevsno))
        self.L.orc_thermal3_set_threads(self.h, int(nthreads))

    def __del__(self):
        try:
            self.L.orc_thermal3_destroy(self.h)
        except Exception:
            pass

    def set_mesh(self, dz, area, conn_dist_up, conn_dist_dn, soil_top_dist_dn, snow_dz0=None):
        dz, area = table(dz, self.ncol, self.nlev), f64(area)
        du, dd = table(conn_dist_up, self.ncol, self.nlev - 1), table(conn_dist_dn, self.ncol, self.nlev - 1)
        st = f64(soil_top_dist_dn)
        s0 = table(snow_dz0, self.ncol, self.nlevsno) if snow_dz0 is not None else None
        return self.L.orc_thermal3_set_mesh(self.h, dp(dz), dp(area), dp(du), dp(dd), dp(st), dp(s0) if s0 is not None else None)

    def set_soils(self, watsat, csol, tkmg, tkdry, lun_type, nlevsoi, istsoil=1):
        a = [table(x, self.ncol, self.nlev) for x in (watsat, csol, tkmg, tkdry)]
        lt = i32(lun_type)
        return self.L.orc_thermal3_set_soils(self.h, *[dp(x) for x in a], ip(lt), int(nlevsoi), int(istsoil))

    def set_cnfac(self, cnfac):
        self.L.orc_thermal3_set_cnfac(self.h, C.c_double(cnfac))

    def set_soln_prev(self, T):
        T = f64(T)
        assert T.size == self.ncells
        self.L.orc_thermal3_set_soln_prev(self.h, dp(T))

    def set_data(self, auxvar_type, var_type, cond_id, data, ieqn=1):
        data = f64(data)
        rc = self.L.orc_thermal3_set_rdata(self.h, int(auxvar_type), int(var_type), int(cond_id), dp(data), int(data.size))
        if rc:
            raise ValueError("set_rdata rc=%d" % rc)

    def set_idata(self, auxvar_type, var_type, cond_id, data):
        data = i32(data)
        rc = self.L.orc_thermal3_set_idata(self.h, int(auxvar_type), int(var_type), int(cond_id), ip(data), int(data.size))
        if rc:
            raise ValueError("set_idata rc=%d" % rc)

    def pre_step_dt(self):
        self.L.orc_thermal3_pre_step_dt(self.h)

    def post_step_dt(self):
        pass

    def step_dt(self, dt, nstep=1):
        conv = C.c_int()
        self.L.orc_thermal3_step_dt(self.h, C.c_double(dt), int(nstep), C.byref(conv))
        return bool(conv.value), 0

    def get_soln(self):
        out = np.empty(self.ncells)
        self.L.orc_thermal3_get_soln(self.h, dp(out))
        return out


class OracleTH:
    """Same call surface as mpp_b200.soe.TH."""

    def __init__(self, ncol, nlev, per_column=True, nthreads=1, **kw):
        self.L = lib()
        self.ncol, self.nlev, self.ncells = ncol, nlev, ncol * nlev
        self.h = C.c_void_p(self.L.orc_th_create(ncol, nlev))
        self.L.orc_th_set_mode(self.h, int(per_column), int(nthreads))

    def __del__(self):
        try:
            self.L.orc_th_destroy(self.h)
        except Exception:
            pass

    def set_mesh(self, orientation, dz, area, col_active=None):
        dz, area = table(dz, self.ncol, self.nlev), f64(area)
        return self.L.orc_th_set_mesh(self.h, int(orientation), dp(dz), dp(area), None)

    def add_condition(self, ieqn, ss_or_bc, cond_type, region):
        return self.L.orc_th_add_condition(self.h, int(ieqn), int(ss_or_bc), int(cond_type), int(region))

    def set_soils(self, watsat, hksat, bsw, sucsat, residual_sat, csol, tkdry, satfunc_type="van_genuchten",
                  density_type=2, int_energy_enthalpy_type=1):
        a = [table(x, self.ncol, self.nlev) for x in (watsat, hksat, bsw, sucsat, residual_sat, csol, tkdry)]
        rc = self.L.orc_th_set_soils(self.h, *[dp(x) for x in a], SATFUNC_NAMES[satfunc_type], int(density_type), int(int_energy_enthalpy_type))
        if rc:
            raise ValueError("set_soils rc=%d" % rc)

    def set_energy_permeability(self, perm):
        perm = f64(perm)
        assert perm.size == self.ncells
        self.L.orc_th_set_energy_perm(self.h, dp(perm))

    def set_tolerances(self, atol=1e-50, rtol=1e-8, stol=1e-10, max_it=50, max_funcs=10000):
        self.L.orc_th_set_tolerances(self.h, C.c_double(atol), C.c_double(rtol), C.c_double(stol), int(max_it), int(max_funcs))

    def restart(self, press, temp):
        press, temp = f64(press), f64(temp)
        assert press.size == self.ncells and temp.size == self.ncells
        self.L.orc_th_restart(self.h, dp(press), dp(temp))

    def set_data(self, auxvar_type, var_type, cond_id, data, ieqn=1):
        data = f64(data)
        if int(auxvar_type) == 702 and int(var_type) == 604:      # AUXVAR_BC, VAR_PRESSURE: boundary aux-var pressure of the energy equation
            rc = self.L.orc_th_set_bc_pressure(self.h, int(ieqn), int(cond_id), dp(data), int(data.size))
        else:
            rc = self.L.orc_th_set_data(self.h, int(ieqn), int(auxvar_type), int(var_type), int(cond_id), dp(data), int(data.size))
        if rc:
            raise ValueError("th set_data rc=%d" % rc)

    def get_data(self, auxvar_type, var_type, cond_id, n=None, ieqn=1):
        out = np.empty(self.ncells)
        rc = self.L.orc_th_get_data(self.h, int(var_type), dp(out), int(out.size))
        if rc:
            raise ValueError("th get_data rc=%d" % rc)
        return out

    def pre_step_dt(self):
        pass

    def post_step_dt(self):
        pass

    def step_dt(self, dt, nstep=1):
        conv, reason = C.c_int(), C.c_int()
        self.L.orc_th_step_dt(self.h, C.c_double(dt), int(nstep), C.byref(conv), C.byref(reason))
        return bool(conv.value), reason.value

    def stats(self):
        its, rs, cuts, nf = (np.zeros(self.ncol, dtype=np.int32) for _ in range(4))
        self.L.orc_th_get_stats(self.h, ip(its), ip(rs), ip(cuts), ip(nf))
        return {"newton_its": its, "reasons": rs, "dt_cuts": cuts, "nfuncs": nf}

    def eval(self, dt, x_prev, x):
        """x, x_prev interleaved (P,T) per cell; returns f (2N) and the 2x2 block bands ja, jb, jc (4N each)."""
        x_prev, x = f64(x_prev), f64(x)
        n = self.ncells
        f = np.zeros(2 * n)
        ja, jb, jc = (np.zeros(4 * n) for _ in range(3))
        self.L.orc_th_eval(self.h, C.c_double(dt), dp(x_prev), dp(x), dp(f), dp(ja), dp(jb), dp(jc))
        return f, ja, jb, jc
