"""Host-side mirrors of the reference's system-of-equations objects, over the C ABI.

Method names follow the type-bound procedures every reference driver calls on
``class(sysofeqns_base_type)`` (src/mpp/soe/SystemOfEquationsBaseType.F90:34-93) and on the
MPP objects (src/mpp/mpp/MultiPhysicsProbVSFM.F90): ``AddConditionInGovEqn``, ``SetSoils``,
``Restart``, ``SetDataFromCLM``, ``GetDataForCLM``, ``PreStepDT``, ``StepDT``, ``PostStepDT``.
snake_case aliases are provided.  Everything here only marshals numpy arrays into
libmppgpu.so; there is no Python or CPU compute path.
"""
import ctypes as C

import numpy as np

from . import constants as K
from ._lib import MPPError, Xfer, c_dp, c_ip, check, lib


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def _dp(a):
    return a.ctypes.data_as(c_dp)


def _ip(a):
    return a.ctypes.data_as(c_ip)


def host_register(a):
    """mppgpu_host_register: page-lock a numpy array the caller keeps for the run (copies of it become direct DMA)."""
    if not a.flags["C_CONTIGUOUS"] or a.nbytes == 0:
        raise ValueError("host_register: a non-empty C-contiguous array is required")
    check(lib().mppgpu_host_register(C.c_void_p(a.ctypes.data), int(a.nbytes)))
    return a


def host_unregister(a):
    check(lib().mppgpu_host_unregister(C.c_void_p(a.ctypes.data)))


def _table(a, ncol, nlev):
    """(ncol, nlev) array -> Fortran column-major flat buffer t[j*ncol + c]."""
    a = np.asarray(a, dtype=np.float64)
    if a.shape != (ncol, nlev):
        raise ValueError("soil table must have shape (ncol, nlev) = (%d, %d), got %s" % (ncol, nlev, a.shape))
    return np.ascontiguousarray(a.T).reshape(-1)


class _SoE:
    soe_itype = None

    def __init__(self, ncol, nlev, device=0):
        self.L = lib()
        self.ncol, self.nlev, self.ncells = int(ncol), int(nlev), int(ncol) * int(nlev)
        self.h = C.c_void_p()
        check(self.L.mppgpu_create(self.soe_itype, self.ncol, self.nlev, int(device), C.byref(self.h)))
        self.device = int(device)

    def close(self):
        if getattr(self, "h", None) is not None and self.h:
            self.L.mppgpu_destroy(self.h)
            self.h = None

    def elm_set_pipeline(self, nchunks=0, static_soil_geometry=False):
        """How the ELM solve entry points (VSFM.elm_solve, Thermal.elm_solve) pipeline their host copies: number of column chunks (0 default,
        1 none); thermal only: upload the soil rows of z / dz / zi (ELM's fixed vertical grid) with the first solve only."""
        check(self.L.mppgpu_elm_set_pipeline(self.h, int(nchunks), 1 if static_soil_geometry else 0))

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- setup -------------------------------------------------------------------------------------
    def set_mesh(self, orientation, dz, area, col_active=None):
        """dz: (ncol, nlev) [m]; area: (ncol,) [m^2] (MeshType.F90:173-269, 401-428)."""
        dzt = _table(dz, self.ncol, self.nlev)
        area = _f64(area)
        if area.size != self.ncol:
            raise ValueError("area must have ncol entries")
        act = _i32(col_active) if col_active is not None else None
        check(self.L.mppgpu_set_mesh(self.h, int(orientation), _dp(dzt), _dp(area), _ip(act) if act is not None else None))

    SetMesh = set_mesh

    def set_connection_distances(self, dist_up, dist_dn):
        """mpp%CreateAndAddConnectionSet(..., dist_up, dist_dn, ...): (ncol, nlev-1) centroid-to-face distances."""
        du, dd = _table(dist_up, self.ncol, self.nlev - 1), _table(dist_dn, self.ncol, self.nlev - 1)
        check(self.L.mppgpu_set_connection_distances(self.h, _dp(du), _dp(dd)))

    def add_condition(self, ieqn, ss_or_bc, cond_type, region):
        """soe%AddConditionInGovEqn(ieqn, COND_BC|COND_SS, name, units, cond_type, region) -> 1-based condition id."""
        cid = C.c_int()
        check(self.L.mppgpu_add_condition(self.h, int(ieqn), int(ss_or_bc), int(cond_type), int(region), C.byref(cid)))
        return cid.value

    AddConditionInGovEqn = add_condition

    def set_tolerances(self, atol=1e-50, rtol=1e-8, stol=1e-10, max_it=50, max_funcs=10000):
        check(self.L.mppgpu_set_tolerances(self.h, atol, rtol, stol, int(max_it), int(max_funcs)))

    SNESSetTolerances = set_tolerances

    def set_step_budget(self, max_residual_evaluations):
        """Optional give-up budget per column per StepDT (not in the reference; 0 = unlimited): see include/mppgpu.h."""
        check(self.L.mppgpu_set_step_budget(self.h, int(max_residual_evaluations)))

    def set_column_ordering(self, mode):
        """Scheduling only (include/mppgpu.h): 1 = visit columns grouped by their previous step's cost (default), 0 = batch order."""
        check(self.L.mppgpu_set_column_ordering(self.h, int(mode)))

    def restart(self, x):
        x = _f64(x)
        check(self.L.mppgpu_restart(self.h, _dp(x), int(x.size)))

    Restart = restart

    # -- data exchange -----------------------------------------------------------------------------
    def set_data(self, auxvar_type, var_type, cond_id, data, ieqn=1):
        data = _f64(data)
        check(self.L.mppgpu_set_data(self.h, int(ieqn), int(auxvar_type), int(var_type), int(cond_id), _dp(data), int(data.size)))

    SetDataFromCLM = set_data

    def get_data(self, auxvar_type, var_type, cond_id, n=None, ieqn=1, out=None):
        if out is None:
            out = np.empty(self.ncells if n is None else int(n), dtype=np.float64)
        check(self.L.mppgpu_get_data(self.h, int(ieqn), int(auxvar_type), int(var_type), int(cond_id), _dp(out), int(out.size)))
        return out

    GetDataForCLM = get_data

    def set_data_device(self, auxvar_type, var_type, cond_id, dev_ptr, n, ieqn=1):
        check(self.L.mppgpu_set_data_device(self.h, int(ieqn), int(auxvar_type), int(var_type), int(cond_id), C.c_void_p(dev_ptr), int(n)))

    def get_data_device(self, auxvar_type, var_type, cond_id, dev_ptr, n, ieqn=1):
        check(self.L.mppgpu_get_data_device(self.h, int(ieqn), int(auxvar_type), int(var_type), int(cond_id), C.c_void_p(dev_ptr), int(n)))

    # -- time stepping -----------------------------------------------------------------------------
    def pre_step_dt(self):
        check(self.L.mppgpu_pre_step_dt(self.h))

    PreStepDT = pre_step_dt

    def post_step_dt(self):
        check(self.L.mppgpu_post_step_dt(self.h))

    PostStepDT = post_step_dt

    def step_dt(self, dt, nstep=1):
        """soe%StepDT(dt, nstep, converged, converged_reason, ierr) -> (converged, converged_reason)."""
        conv, reason = C.c_int(), C.c_int()
        check(self.L.mppgpu_step_dt(self.h, float(dt), int(nstep), C.byref(conv), C.byref(reason)))
        return bool(conv.value), reason.value

    StepDT = step_dt

    def step_dt_async(self, dt, nstep=1):
        check(self.L.mppgpu_step_dt_async(self.h, float(dt), int(nstep)))

    def step_result(self):
        conv, reason = C.c_int(), C.c_int()
        check(self.L.mppgpu_step_result(self.h, C.byref(conv), C.byref(reason)))
        return bool(conv.value), reason.value

    def set_stream(self, cuda_stream):
        check(self.L.mppgpu_set_stream(self.h, C.c_void_p(cuda_stream)))

    def synchronize(self):
        check(self.L.mppgpu_synchronize(self.h))

    # -- diagnostics -------------------------------------------------------------------------------
    def stats(self):
        its, rs, cuts, nf = (np.zeros(self.ncol, dtype=np.int32) for _ in range(4))
        check(self.L.mppgpu_get_column_stats(self.h, _ip(its), _ip(rs), _ip(cuts), _ip(nf)))
        return {"newton_its": its, "reasons": rs, "dt_cuts": cuts, "nfuncs": nf}

    def mass_balance(self, dt=0.0):
        sums, maxs = np.zeros(4), np.zeros(4)
        check(self.L.mppgpu_vsfm_mass_balance(self.h, float(dt), _dp(sums), _dp(maxs)))
        return sums, maxs

    # -- global reductions over ranks (include/mppgpu.h "global reductions") -----------------------------
    def comm_init(self, nranks=1, rank=0, unique_id=None):
        """`unique_id`: the MPPGPU_COMM_ID_BYTES bytes of mpp_b200.comm_unique_id(), identical on every rank (None for one rank)."""
        buf = C.create_string_buffer(bytes(unique_id), 128) if unique_id is not None else None
        check(self.L.mppgpu_comm_init(self.h, int(nranks), int(rank), buf))

    def global_reduce_async(self):
        check(self.L.mppgpu_global_reduce_async(self.h))

    def global_mass_balance(self):
        """Collective.  -> dict of the global sums / maxima / worst SNES reason of the last StepDT over all ranks."""
        sums, maxs, worst = np.zeros(4), np.zeros(4), C.c_int()
        check(self.L.mppgpu_global_mass_balance(self.h, _dp(sums), _dp(maxs), C.byref(worst)))
        return {"mass_begin": sums[0], "mass_end": sums[1], "source_dt": sums[2], "boundary_exchanged": sums[3],
                "max_abs_mass_error": maxs[0], "max_newton_its": int(maxs[1]), "any_diverged": bool(maxs[2]),
                "max_dt_cuts": int(maxs[3]), "worst_reason": int(worst.value)}

    def reduction_buffer_ptr(self):
        p = C.c_void_p()
        check(self.L.mppgpu_reduction_buffer_device(self.h, C.byref(p)))
        return p.value

    def launch_count(self):
        n = C.c_longlong()
        check(self.L.mppgpu_launch_count(self.h, C.byref(n)))
        return n.value

    def last_step_ms(self):
        ms = C.c_float()
        check(self.L.mppgpu_last_step_ms(self.h, C.byref(ms)))
        return ms.value


def comm_unique_id():
    """The NCCL unique id (bytes) rank 0 creates and hands to every rank's `comm_init` (mppgpu_comm_unique_id)."""
    buf = C.create_string_buffer(128)
    check(lib().mppgpu_comm_unique_id(buf))
    return buf.raw


class VSFM(_SoE):
    """sysofeqns_vsfm_type + mpp_vsfm_type for batches of independent soil columns."""
    soe_itype = K.SOE_RE_ODE

    def eval(self, dt, x_prev, x):
        """Residual and the three Jacobian bands (sub, diagonal, super; N values each) the fused step kernel assembles at x, with the
        accumulation of the start of the step taken at x_prev (VSFMSOEResidual / VSFMJacobian, SystemOfEquationsVSFMType.F90:94-403)."""
        x_prev, x = _f64(x_prev), _f64(x)
        f, ja, jb, jc = (np.zeros(self.ncells) for _ in range(4))
        check(self.L.mppgpu_eval(self.h, float(dt), _dp(x_prev), _dp(x), _dp(f), _dp(ja), _dp(jb), _dp(jc)))
        return f, ja, jb, jc

    def set_soils(self, watsat, hksat, bsw, sucsat, residual_sat, satfunc_type="van_genuchten", density_type=K.DENSITY_TGDPB01):
        """VSFMMPPSetSoils (MultiPhysicsProbVSFM.F90:211-475); tables are (ncol, nlev)."""
        if satfunc_type not in K.SATFUNC:
            raise MPPError("ERROR:: Unknown vsfm_satfunc_type = " + str(satfunc_type))
        t = [_table(x, self.ncol, self.nlev) for x in (watsat, hksat, bsw, sucsat, residual_sat)]
        check(self.L.mppgpu_vsfm_set_soils(self.h, *[_dp(x) for x in t], K.SATFUNC[satfunc_type], int(density_type)))

    VSFMMPPSetSoils = set_soils

    def coupled_step(self, dt, nstep, inputs, outputs, nchunks=0):
        """One ELM coupling step with host buffers, pipelined over column chunks (mppgpu_vsfm_coupled_step):
        SetDataFromCLM for every (auxvar_type, var_type, cond_id, array) of `inputs`, PreStepDT, StepDT, GetDataForCLM into every
        array of `outputs`.  Arrays must be C-contiguous float64 (pinned host memory makes the copies asynchronous)."""
        def pack(items):
            arr = (Xfer * max(len(items), 1))()
            for i, (at, vt, cid, a) in enumerate(items):
                if a.dtype != np.float64 or not a.flags["C_CONTIGUOUS"]:
                    raise ValueError("coupled_step needs C-contiguous float64 arrays")
                arr[i] = Xfer(1, int(at), int(vt), int(cid), _dp(a))
            return arr
        conv, reason = C.c_int(), C.c_int()
        check(self.L.mppgpu_vsfm_coupled_step(self.h, float(dt), int(nstep), len(inputs), pack(inputs), len(outputs), pack(outputs),
                                              int(nchunks), C.byref(conv), C.byref(reason)))
        return bool(conv.value), reason.value


    # -- MPPVSFMALM_Solve with ELM's raw column arrays (MPPVSFMALM_Driver.F90:204-923) --------------------------------
    ELM_COND_ORDER = ("infil", "et", "dew", "drain", "snow", "sublim")
    ELM_INOUT = ("rootr_col", "qflx_drain", "zwt", "h2osoi_liq", "h2osoi_ice", "mflx_snowlyr_col")

    def elm_set_geometry(self, zi, dz, nlevsoi, ids, watmin=0.01, fortran_order=False):
        """zi: (ncol, nlev+1) interface depths, dz: (ncol, nlev) thicknesses (ELM's col%zi(c,0:), col%dz); ids: the condition ids of
        infiltration, ET, dew, drainage, snow, sublimation (dict as returned by the set-up, or a sequence in that order).
        fortran_order: the arrays are ELM's own (c, j) arrays, i.e. numpy (nlev+1, ncol) / (nlev, ncol)."""
        cid = _i32([ids[k] for k in self.ELM_COND_ORDER] if isinstance(ids, dict) else list(ids))
        if fortran_order:
            zi = np.ascontiguousarray(np.asarray(zi, dtype=np.float64).reshape(self.nlev + 1, self.ncol))
            dz = np.ascontiguousarray(np.asarray(dz, dtype=np.float64).reshape(self.nlev, self.ncol))
            check(self.L.mppgpu_vsfm_elm_set_geometry_f(self.h, _dp(zi), _dp(dz), int(nlevsoi), float(watmin), _ip(cid)))
            return
        zi = np.ascontiguousarray(np.asarray(zi, dtype=np.float64).reshape(self.ncol, self.nlev + 1))
        dz = np.ascontiguousarray(np.asarray(dz, dtype=np.float64).reshape(self.ncol, self.nlev))
        check(self.L.mppgpu_vsfm_elm_set_geometry(self.h, _dp(zi), _dp(dz), int(nlevsoi), float(watmin), _ip(cid)))

    def elm_solve(self, dt, st, nstep=1, fortran_order=False, out=None):
        """One MPPVSFMALM_Solve.  `st`: dict of ELM's column arrays (float64 / int32, C-contiguous; cell arrays (ncol, nlev)); the in/out
        ones (ELM_INOUT) are updated in place.  Returns dict(smp_l, soilp_col, qcharge, abs_mass_error, iter_count, status, nfailed, nattempts);
        pass a previous result as `out` to write into the same (possibly page-locked) arrays again."""
        from ._lib import ElmColumns
        ncol, n = self.ncol, self.ncells
        cols = ElmColumns()
        cols.fortran_order = 1 if fortran_order else 0      # rootr_col, h2osoi_liq/ice, smp_l, soilp_col as (nlev, ncol) = ELM's (c, j)
        keep = []

        def dptr(a, size, name):
            if a.dtype != np.float64 or not a.flags["C_CONTIGUOUS"] or a.size != size:
                raise ValueError("elm_solve: %s must be a C-contiguous float64 array of %d values" % (name, size))
            keep.append(a)
            return _dp(a)

        def iptr(a, size, name):
            if a.dtype != np.int32 or not a.flags["C_CONTIGUOUS"] or a.size != size:
                raise ValueError("elm_solve: %s must be a C-contiguous int32 array of %d values" % (name, size))
            keep.append(a)
            return _ip(a)
        if st.get("col_pfti") is not None:
            npft = int(st["pft_wtcol"].size)
            cols.npft, cols.max_patch_per_col = npft, int(st["max_patch_per_col"])
            cols.col_pfti, cols.col_npfts = iptr(st["col_pfti"], ncol, "col_pfti"), iptr(st["col_npfts"], ncol, "col_npfts")
            cols.pft_active, cols.pft_wtcol = iptr(st["pft_active"], npft, "pft_active"), dptr(st["pft_wtcol"], npft, "pft_wtcol")
            cols.rootr_pft = dptr(st["rootr_pft"], npft * self.nlev, "rootr_pft")
            cols.qflx_tran_veg_pft = dptr(st["qflx_tran_veg_pft"], npft, "qflx_tran_veg_pft")
        for k in ("qflx_tran_veg_col", "qflx_infl", "qflx_dew_snow", "qflx_dew_grnd", "qflx_sub_snow", "frac_h2osfc", "qflx_drain", "zwt",
                  "mflx_snowlyr_col", "mflx_neg_snow_col"):
            setattr(cols, k, dptr(st[k], ncol, k))
        for k in ("rootr_col", "h2osoi_liq", "h2osoi_ice", "mflx_drain_perched"):
            setattr(cols, k, dptr(st[k], n, k))
        cols.snl = iptr(st["snl"], ncol, "snl")
        if out is None:
            out = {"smp_l": np.zeros(n), "soilp_col": np.zeros(n), "qcharge": np.zeros(ncol), "abs_mass_error": np.zeros(ncol),
                   "iter_count": np.zeros(ncol, dtype=np.int32), "status": np.zeros(ncol, dtype=np.int32)}
        for k, size in (("smp_l", n), ("soilp_col", n), ("qcharge", ncol), ("abs_mass_error", ncol)):
            setattr(cols, k, dptr(out[k], size, k))
        cols.iter_count, cols.status = iptr(out["iter_count"], ncol, "iter_count"), iptr(out["status"], ncol, "status")
        nf, na = C.c_int(), C.c_int()
        check(self.L.mppgpu_vsfm_elm_solve(self.h, float(dt), int(nstep), C.byref(cols), C.byref(nf), C.byref(na)))
        out["nfailed"], out["nattempts"] = nf.value, na.value
        return out


class Thermal(_SoE):
    """sysofeqns_thermal_type (soil governing equation, KSP path)."""
    soe_itype = K.SOE_THERMAL_TBASED

    def set_soils(self, watsat, csol, tkmg, tkdry, lun_type, nlevsoi, istsoil=K.ISTSOIL):
        """MPPThermalSetSoils (MultiPhysicsProbThermal.F90:76-208); tables are (ncol, nlev); lun_type is (ncol,)."""
        t = [_table(x, self.ncol, self.nlev) for x in (watsat, csol, tkmg, tkdry)]
        lt = _i32(lun_type)
        check(self.L.mppgpu_thermal_set_soils(self.h, *[_dp(x) for x in t], _ip(lt), int(nlevsoi), int(istsoil)))

    MPPThermalSetSoils = set_soils

    def set_bulk_copy(self, mode):
        """0 (default): register loads; 1: the persistent bulk-async (1-D TMA) kernel where its shape fits (measured slower on B200).
        Results are bit-identical."""
        check(self.L.mppgpu_thermal_set_bulk_copy(self.h, int(mode)))

    def set_cnfac(self, cnfac):
        check(self.L.mppgpu_thermal_set_cnfac(self.h, float(cnfac)))

    def set_soln_prev(self, T):
        self.restart(T)

    SetSolnPrevCLM = set_soln_prev

    def set_rdata(self, auxvar_type, var_type, cond_id, data):
        self.set_data(auxvar_type, var_type, cond_id, data)

    SetRDataFromCLM = set_rdata

    def set_idata(self, auxvar_type, var_type, cond_id, data):
        data = _i32(data)
        check(self.L.mppgpu_set_idata(self.h, 1, int(auxvar_type), int(var_type), int(cond_id), _ip(data), int(data.size)))

    SetIDataFromCLM = set_idata
    SetBDataFromCLM = set_idata

    def get_soln(self):
        return self.get_data(K.AUXVAR_INTERNAL, K.VAR_TEMPERATURE, -1)

    GetSoln = get_soln


class ThermalSnow(Thermal):
    """sysofeqns_thermal_type with the three governing equations ELM uses: snow (nlevsno layers), standing surface water
    (one cell) and soil, coupled into one linear system (MPPThermalTBasedALM_Initialize.F90:150-727).  Internal aux-var
    arrays hold ncol*(nlevsno+1+nlev) values in the SoE order [snow | ssw | soil]."""

    def __init__(self, ncol, nlev, nlevsno, device=0):
        super().__init__(ncol, nlev, device)
        self.nlevsno = int(nlevsno)
        self.nsoil = self.ncells
        self.ncells = self.ncol * (self.nlevsno + 1 + self.nlev)

    def set_mesh(self, dz, area, conn_dist_up, conn_dist_dn, soil_top_dist_dn, snow_dz0=None):
        """Soil mesh + internal connection distances (add_meshes, :150-500), then the snow / standing-water equations and
        the coupling conditions (add_goveqns, add_conditions_to_goveqns, allocate_auxvars)."""
        ncells, self.ncells = self.ncells, self.nsoil
        try:
            Thermal.set_mesh(self, K.MESH_ALONG_GRAVITY, dz, area)
            self.set_connection_distances(conn_dist_up, conn_dist_dn)
        finally:
            self.ncells = ncells
        st = _f64(soil_top_dist_dn)
        if st.size != self.ncol:
            raise ValueError("soil_top_dist_dn must have ncol entries")
        check(self.L.mppgpu_thermal_add_snow_ssw(self.h, self.nlevsno, _dp(st)))

    ELM_FIELDS = ("z", "dz", "zi", "t_soisno", "h2osoi_liq", "h2osoi_ice", "frac_sno_eff", "h2osno", "h2osfc", "frac_h2osfc", "t_h2osfc",
                  "sabg_lyr", "dhsdT", "hs_soil", "hs_top_snow", "hs_h2osfc")

    def elm_solve(self, dt, elm, nstep=1, capr=0.34):
        """MPPThermalTBasedALM_Solve (MPPThermalTBasedALM_Driver.F90:150-452).  `elm`: dict of ELM's column arrays in Fortran (c, j)
        order, i.e. numpy arrays of shape (nlayers, ncol), C-contiguous: z, dz, t_soisno, h2osoi_liq, h2osoi_ice over j = -nlevsno+1..nlev;
        zi and tvector over j = -nlevsno..nlev; sabg_lyr over j = -nlevsno+1..1; the rest (ncol,).  `tvector` is updated in place."""
        from ._lib import ElmThermalColumns
        cols = ElmThermalColumns()
        keep = []
        for k in self.ELM_FIELDS + ("tvector",):
            a = elm[k]
            if a.dtype != np.float64 or not a.flags["C_CONTIGUOUS"]:
                raise ValueError("elm_solve: %s must be a C-contiguous float64 array" % k)
            keep.append(a)
            setattr(cols, k, _dp(a))
        snl = elm["snl"]
        if snl.dtype != np.int32 or snl.size != self.ncol:
            raise ValueError("elm_solve: snl must be int32 with ncol entries")
        cols.snl = _ip(snl)
        nl = self.nlevsno + self.nlev
        want = {"z": nl, "dz": nl, "t_soisno": nl, "h2osoi_liq": nl, "h2osoi_ice": nl, "zi": nl + 1, "tvector": nl + 1, "sabg_lyr": self.nlevsno + 1}
        for k, n in want.items():
            if elm[k].size != n * self.ncol:
                raise ValueError("elm_solve: %s must hold %d layers x ncol values" % (k, n))
        check(self.L.mppgpu_thermal_elm_solve(self.h, float(dt), int(nstep), C.byref(cols), float(capr)))
        return elm["tvector"]


class TH(_SoE):
    """sysofeqns_th_type: Richards (ieqn 1) + enthalpy (ieqn 2) on the same columns."""
    soe_itype = K.SOE_TH

    def set_soils(self, watsat, hksat, bsw, sucsat, residual_sat, csol, tkdry, satfunc_type="van_genuchten",
                  density_type=K.DENSITY_TGDPB01, int_energy_enthalpy_type=K.INT_ENERGY_ENTHALPY_CONSTANT):
        t = [_table(x, self.ncol, self.nlev) for x in (watsat, hksat, bsw, sucsat, residual_sat, csol, tkdry)]
        check(self.L.mppgpu_th_set_soils(self.h, *[_dp(x) for x in t], K.SATFUNC[satfunc_type], int(density_type),
                                         int(int_energy_enthalpy_type)))

    MPPTHSetSoils = set_soils

    def set_energy_permeability(self, perm):
        """goveq_enthalpy%SetSoilPermeability: per-cell permeability of the energy equation's aux vars (cell order)."""
        perm = _f64(perm)
        check(self.L.mppgpu_th_set_energy_permeability(self.h, _dp(perm), int(perm.size)))

    def restart(self, press, temp=None):
        x = _f64(press) if temp is None else np.concatenate([_f64(press), _f64(temp)])
        check(self.L.mppgpu_restart(self.h, _dp(x), int(x.size)))

    def eval(self, dt, x_prev, x):
        x_prev, x = _f64(x_prev), _f64(x)
        n = self.ncells
        f = np.zeros(2 * n)
        ja, jb, jc = (np.zeros(4 * n) for _ in range(3))
        check(self.L.mppgpu_eval(self.h, float(dt), _dp(x_prev), _dp(x), _dp(f), _dp(ja), _dp(jb), _dp(jc)))
        return f, ja, jb, jc
