"""Problem recipes shared by the oracle and the CUDA library (same call surface on both).

Each builder follows a reference driver:
  celia          src/driver/standalone/vsfm/vsfm_celia1990_problem.F90
  elm_vsfm_batch src/driver/alm/MPPVSFMALM_Initialize.F90 (mesh :401-530, conditions :814-858, IC :1058-1060)
                 + SURVEY.md section 8(d) synthetic inputs (seed 20240607)
"""
import numpy as np

from mpp_b200 import constants as K

SEED = 20240607


# ---------------------------------------------------------------------------------------------------
# Celia et al. (1990) infiltration column -- config #1
# ---------------------------------------------------------------------------------------------------
def build_celia(cls, nz=100, **kw):
    p = cls(1, nz, **kw)
    dz = np.full((1, nz), 1.0 / nz)                                   # z_column = 1 m, MeshCreate :198-200
    p.set_mesh(K.MESH_AGAINST_GRAVITY, dz, np.array([1.0]))
    top = p.add_condition(1, K.COND_BC, K.COND_DIRICHLET, K.SOIL_TOP_CELLS)      # :252-254
    bot = p.add_condition(1, K.COND_BC, K.COND_DIRICHLET, K.SOIL_BOTTOM_CELLS)   # :256-258
    porosity, lam, alpha, perm = 0.368, 0.5, 3.4257e-4, 8.3913e-12    # :295-298
    vish2o, denh2o, grav = 0.001002, 1000.0, K.GRAV
    hksat = perm / vish2o * (denh2o * grav) / 0.001                   # :325
    sucsat = 1.0 / (alpha * K.GRAVITY_CONSTANT)                       # :327 (the driver itself inverts with 9.80665; VSFMMPPSetSoils converts with 9.80616)
    full = lambda v: np.full((1, nz), v)
    p.set_soils(full(porosity), full(hksat), full(1.0 / lam), full(sucsat), full(0.2772),
                "van_genuchten", K.DENSITY_TGDPB01)                   # :331-335
    p.restart(np.full(nz, 3.5355e3))                                  # :358-362
    return p, top, bot


def run_celia(p, top, bot, nstep=24, dt=3600.0):
    its = []
    for istep in range(nstep):
        p.set_data(K.AUXVAR_BC, K.VAR_BC_SS_CONDITION, top, np.array([9.3991e4]))    # :383-394
        p.set_data(K.AUXVAR_BC, K.VAR_BC_SS_CONDITION, bot, np.array([3.5355e3]))
        conv, reason = p.step_dt(dt, istep + 1)
        assert conv, "Celia step %d did not converge (reason %d)" % (istep + 1, reason)
        its.append(int(p.stats()["newton_its"][0]))
    P = p.get_data(K.AUXVAR_INTERNAL, K.VAR_PRESSURE, -1)
    S = p.get_data(K.AUXVAR_INTERNAL, K.VAR_LIQ_SAT, -1)
    return P, S, its


# ---------------------------------------------------------------------------------------------------
# vsfm_wt_dynamics -- src/driver/standalone/vsfm/vsfm_wt_dynamics_problem.F90 (no regression baseline exists for it):
#   the Celia column (:11, :58-60, :366-369) with a hydrostatic start whose water table sits at mid-height (:435-444), a mass-rate
#   source of 0.025 kg/s into the top cell and a constant-head Dirichlet condition at the bottom (:326-332, :462-476), 24 x 3600 s
# ---------------------------------------------------------------------------------------------------
def build_wt_dynamics(cls, nz=100, **kw):
    p = cls(1, nz, **kw)
    dz = 1.0 / nz
    p.set_mesh(K.MESH_AGAINST_GRAVITY, np.full((1, nz), dz), np.array([1.0]))
    top = p.add_condition(1, K.COND_SS, K.COND_MASS_RATE, K.SOIL_TOP_CELLS)
    bot = p.add_condition(1, K.COND_BC, K.COND_DIRICHLET, K.SOIL_BOTTOM_CELLS)
    porosity, lam, alpha, perm = 0.368, 0.5, 3.4257e-4, 8.3913e-12
    hksat = perm / 0.001002 * (1000.0 * K.GRAV) / 0.001
    sucsat = 1.0 / (alpha * K.GRAVITY_CONSTANT)
    full = lambda v: np.full((1, nz), v)
    p.set_soils(full(porosity), full(hksat), full(1.0 / lam), full(sucsat), full(0.2772), "van_genuchten", K.DENSITY_TGDPB01)
    z = dz / 2.0 + dz * np.arange(nz)
    p.restart(101325.0 + (0.5 - z) * 997.16 * 9.80868)
    return p, top, bot


def run_wt_dynamics(p, top, bot, nstep=24, dt=3600.0):
    its = []
    for istep in range(nstep):
        p.set_data(K.AUXVAR_SS, K.VAR_BC_SS_CONDITION, top, np.array([2.5e-5 * 1e3]))
        p.set_data(K.AUXVAR_BC, K.VAR_BC_SS_CONDITION, bot, np.array([101325.0 + 0.5 * 997.16 * 9.80868]))
        conv, reason = p.step_dt(dt, istep + 1)
        assert conv, "wt_dynamics step %d did not converge (reason %d)" % (istep + 1, reason)
        its.append(int(p.stats()["newton_its"][0]))
    return p.get_data(K.AUXVAR_INTERNAL, K.VAR_PRESSURE, -1), p.get_data(K.AUXVAR_INTERNAL, K.VAR_LIQ_SAT, -1), its


# ---------------------------------------------------------------------------------------------------
# Srivastava & Yeh (1991) layered-soil infiltration -- src/driver/standalone/vsfm/vsfm_sy1991_problem.F90
#   2 m column, 200 cells, bottom-up mesh, low-permeability lower half under a ten times more permeable upper half, constant recharge at
#   the top (mass-rate source), constant head at the bottom (Dirichlet): the transient between the steady states of two recharge rates.
#   `ic`: the driver's initial-pressure table of the chosen problem (200 values; tests/golden/sy1991_ic.json).  No regression baseline exists.
# ---------------------------------------------------------------------------------------------------
def build_sy1991(cls, ic, nz=200, **kw):
    ic = np.asarray(ic, dtype=np.float64)
    assert ic.size == nz
    p = cls(1, nz, **kw)
    p.set_mesh(K.MESH_AGAINST_GRAVITY, np.full((1, nz), 2.0 / nz), np.array([1.0]))        # MeshCreate, z_column = 2 m (:293-294)
    top = p.add_condition(1, K.COND_SS, K.COND_MASS_RATE, K.SOIL_TOP_CELLS)                  # 'Constant flux condition at top' (:348-350)
    bot = p.add_condition(1, K.COND_BC, K.COND_DIRICHLET, K.SOIL_BOTTOM_CELLS)               # 'Constant head condition at bottom' (:352-354)
    porosity, lam, alpha, perm_high, perm_low = 0.4, 0.5455, 4.0e-4, 2.5281e-12, 2.5281e-13  # :388-392
    vish2o, denh2o, grav = 0.001002, 1000.0, K.GRAV
    hksat = np.empty((1, nz))
    hksat[:, :nz // 2] = perm_low / vish2o * (denh2o * grav) / 0.001                          # :422-423
    hksat[:, nz // 2:] = perm_high / vish2o * (denh2o * grav) / 0.001
    full = lambda v: np.full((1, nz), v)
    p.set_soils(full(porosity), hksat, full(1.0 / lam), full(1.0 / (alpha * K.GRAVITY_CONSTANT)), full(0.15),
                "van_genuchten", K.DENSITY_TGDPB01)                                          # :424-433
    p.restart(ic)                                                                            # :446-463
    return p, top, bot


def run_sy1991(p, top, bot, ic, problem="drying", nstep=24, dt=3600.0, first_step=1):
    recharge = {"wetting": 2.5e-6, "drying": 2.7778e-7}[problem] * 997.16                    # :476-492
    its = []
    for step in range(first_step, first_step + nstep):
        p.set_data(K.AUXVAR_SS, K.VAR_BC_SS_CONDITION, top, np.array([recharge]))
        p.set_data(K.AUXVAR_BC, K.VAR_BC_SS_CONDITION, bot, np.array([ic[0]]))
        conv, reason = p.step_dt(dt, step)
        if not conv:
            raise RuntimeError("sy1991 step %d did not converge (reason %d)" % (step, reason))
        its.append(int(p.stats()["newton_its"][0]))
    P = p.get_data(K.AUXVAR_INTERNAL, K.VAR_PRESSURE, -1)
    sat = p.get_data(K.AUXVAR_INTERNAL, K.VAR_LIQ_SAT, -1)
    return P, sat, its


# ---------------------------------------------------------------------------------------------------
# regression file format -- src/driver/standalone/util/regression.F90:76-124
# ---------------------------------------------------------------------------------------------------
def regression_block(name, category, data, num_cells):
    def fmt(v):
        v = 0.0 if abs(v) < 1e-50 else v
        s = "%.12E" % v                       # 13 significant digits; Fortran e21.13 prints them as 0.dddddddddddddE+ee
        mant, exp = s.split("E")
        sign = "-" if mant.startswith("-") else ""
        digits = mant.replace("-", "").replace(".", "")
        e = int(exp) + 1
        if float(v) == 0.0:
            e = 0
        return ("%s0.%sE%+03d" % (sign, digits[:13], e)).rjust(21)
    lines = ["[%s]" % name, "category = %s" % category,
             "min = " + fmt(np.min(data)), "max = " + fmt(np.max(data)), "mean = " + fmt(np.sum(data) / data.size)]
    n = min(num_cells, data.size)
    inc = data.size // n
    for cell in range(1, data.size + 1, inc):
        lines.append("cell %4d = %s" % (cell, fmt(data[cell - 1])))
    lines.append("")
    return lines


def parse_regression(path):
    out, cur = {}, None
    for line in open(path):
        line = line.strip()
        if not line:
            continue
        if line.startswith("["):
            cur = line[1:-1]
            out[cur] = {}
        elif "=" in line:
            k, v = [t.strip() for t in line.split("=")]
            if k != "category":
                out[cur][k] = float(v)
    return out


# ---------------------------------------------------------------------------------------------------
# ELM-like batched columns -- configs #3/#4
# ---------------------------------------------------------------------------------------------------
def elm_layers(nlev=15):
    """ELM node depths z_j = 0.025 (exp(0.5 (j - 0.5)) - 1), interfaces at mid-points."""
    j = np.arange(1, nlev + 1)
    z = 0.025 * (np.exp(0.5 * (j - 0.5)) - 1.0)
    zi = np.zeros(nlev + 1)
    zi[1:nlev] = 0.5 * (z[:-1] + z[1:])
    zi[nlev] = z[-1] + 0.5 * (z[-1] - z[-2])
    dz = zi[1:] - zi[:-1]
    return z, zi, dz


def elm_vsfm_inputs(ncol, nlev=15, seed=SEED, satfunc="van_genuchten", zwt_min=1.0):
    """SURVEY.md 8(d) synthetic ELM-like VSFM columns.  `zwt_min`: shallowest initial water table [m]; the draw is
    U(zwt_min, 20) from the same random stream whatever zwt_min is, so every other input is unchanged.  zwt_min = 1 is the
    survey's batch (kept for the dt-cut / hard-column tests); the benchmark uses 2 m, see `shard_inputs`."""
    rng = np.random.default_rng(seed)
    z, zi, dz = elm_layers(nlev)
    d = {}
    d["dz"] = np.tile(dz, (ncol, 1))
    d["area"] = np.ones(ncol)
    d["watsat"] = rng.uniform(0.35, 0.55, (ncol, nlev))
    d["hksat"] = np.exp(rng.uniform(np.log(5e-4), np.log(5e-2), (ncol, nlev)))      # mm/s
    d["bsw"] = rng.uniform(3.0, 12.0, (ncol, nlev))
    d["sucsat"] = rng.uniform(50.0, 600.0, (ncol, nlev))                            # mm
    d["residual_sat"] = np.zeros((ncol, nlev))                                      # MPPVSFMALM_Initialize.F90:966
    zwt = rng.uniform(1.0, 20.0, ncol)
    if zwt_min != 1.0:
        zwt = zwt_min + (zwt - 1.0) * ((20.0 - zwt_min) / 19.0)
    depth = 0.5 * (zi[:-1] + zi[1:])
    # MPPVSFMALM_Initialize.F90:1058-1060
    d["press_ic"] = (K.PRESSURE_REF + 997.16 * K.GRAVITY_CONSTANT * (-zwt[:, None] - (-depth[None, :]))).reshape(-1)
    d["infil"] = rng.uniform(0.0, 2e-4, ncol)                                       # kg/s into the top cell
    et_tot = -rng.uniform(0.0, 8e-5, ncol)
    w = np.exp(-z[:10]); w /= w.sum()
    et = np.zeros((ncol, nlev)); et[:, :10] = et_tot[:, None] * w[None, :]
    d["et"] = et.reshape(-1)
    d["dew"] = np.zeros(ncol); d["snow"] = np.zeros(ncol); d["sublim"] = np.zeros(ncol)
    d["drain"] = np.zeros(ncol * nlev)
    d["frac_liq"] = np.ones(ncol * nlev)
    d["satfunc"] = satfunc
    d["ncol"], d["nlev"] = ncol, nlev
    return d


# Benchmark batch (bench.py, BASELINE.json configs[2]/[3]).  The water table starts at 2-20 m instead of the survey's 1-20 m.
# Why: with van Genuchten curves the reference algorithm (newtonls + bt) cannot always carry a cell through pc = 0 -- there
# dsat/dP jumps from 0 to O(alpha m n) within a fraction of a Pa and dkr/dSe is unbounded -- and root-zone ET (layers 1-10, down
# to 2.9 m) drags barely saturated cells through exactly that point when the table sits inside the root zone: of 4 Mi columns
# drawn with 1-20 m, 1-26 per step cut dt and ~2 per step fail all 21 attempts (oracle and GPU alike; tests/golden/
# hard_columns.json -- every one of them has its water table within 0.1 m above a cell centre at 1.0, 1.7 or 2.9 m).
# The reference's ELM driver calls endrun on such a column (MPPVSFMALM_Driver.F90:918-921), so a batch that contains them is not
# a valid reference run, and the metric is Newton-CONVERGED column-timesteps.  With 2-20 m the oracle converges every column of
# 400 000 over 30 steps without a single dt cut.  ELM itself defaults to smooth_brooks_corey_bz3 for the same reason
# (mpp_varctl.F90:17).  The 1-20 m batch remains available (`zwt_min=1.0`, bench.py --zwt-min 1) and is what the dt-cut tests use.
BENCH_ZWT_MIN = 2.0
BENCH_CHUNK = 65536                # columns per seeded chunk (shards are unions of chunks)


def shard_inputs(c0, c1, chunk=BENCH_CHUNK, nlev=15, zwt_min=BENCH_ZWT_MIN, builder=None):
    """Columns [c0, c1) of the global seeded benchmark batch: chunk k uses seed SEED + k, so a rank builds only its shard and
    the union over ranks is the same batch whatever the rank count.  `builder(ncol, nlev, seed=..)` defaults to the VSFM inputs."""
    parts = []
    k0, k1 = c0 // chunk, (c1 - 1) // chunk
    mk = builder or (lambda n, nl, seed: elm_vsfm_inputs(n, nl, seed=seed, zwt_min=zwt_min))
    for k in range(k0, k1 + 1):
        d = mk(chunk, nlev, seed=SEED + k)
        lo, hi = max(c0, k * chunk) - k * chunk, min(c1, (k + 1) * chunk) - k * chunk
        parts.append((d, lo, hi))
    d0 = parts[0][0]
    out = {k: v for k, v in d0.items() if not isinstance(v, np.ndarray)}
    out["ncol"] = c1 - c0
    for key, v in d0.items():
        if not isinstance(v, np.ndarray):
            continue
        if v.shape[0] == chunk:                                    # per column, or (ncol, nlev)
            out[key] = np.concatenate([d[key][lo:hi] for d, lo, hi in parts], axis=0)
        else:                                                      # flat per cell
            out[key] = np.concatenate([d[key].reshape(chunk, -1)[lo:hi].reshape(-1) for d, lo, hi in parts])
    return out


def build_elm_vsfm(cls, d, **kw):
    ncol, nlev = d["ncol"], d["nlev"]
    p = cls(ncol, nlev, **kw)
    p.set_mesh(K.MESH_ALONG_GRAVITY, d["dz"], d["area"])
    ids = {}
    # MPPVSFMALM_Initialize.F90:836-858, in this order
    ids["infil"] = p.add_condition(1, K.COND_SS, K.COND_MASS_RATE, K.SOIL_TOP_CELLS)
    ids["et"] = p.add_condition(1, K.COND_SS, K.COND_MASS_RATE, K.SOIL_CELLS)
    ids["dew"] = p.add_condition(1, K.COND_SS, K.COND_MASS_RATE, K.SOIL_TOP_CELLS)
    ids["drain"] = p.add_condition(1, K.COND_SS, K.COND_MASS_RATE, K.SOIL_CELLS)
    ids["snow"] = p.add_condition(1, K.COND_SS, K.COND_MASS_RATE, K.SOIL_TOP_CELLS)
    ids["sublim"] = p.add_condition(1, K.COND_SS, K.COND_MASS_RATE, K.SOIL_TOP_CELLS)
    p.set_soils(d["watsat"], d["hksat"], d["bsw"], d["sucsat"], d["residual_sat"], d["satfunc"], d.get("density_type", K.DENSITY_TGDPB01))
    p.restart(d["press_ic"])
    return p, ids


def elm_vsfm_step(p, ids, d, dt=1800.0, nstep=1, scale=1.0):
    """One ELM coupling step as MPPVSFMALM_Solve does it (MPPVSFMALM_Driver.F90:379-463, 603, 642, 674-705, 935)."""
    for name in ("infil", "et", "dew", "drain", "snow", "sublim"):
        p.set_data(K.AUXVAR_SS, K.VAR_BC_SS_CONDITION, ids[name], d[name] * scale)
    p.set_data(K.AUXVAR_INTERNAL, K.VAR_FRAC_LIQ_SAT, 1, d["frac_liq"])
    p.pre_step_dt()
    conv, reason = p.step_dt(dt, nstep)
    out = {k: p.get_data(K.AUXVAR_INTERNAL, v, 1) for k, v in
           (("sat", K.VAR_LIQ_SAT), ("mass", K.VAR_MASS), ("smp", K.VAR_SOIL_MATRIX_POT), ("pressure", K.VAR_PRESSURE))}
    p.post_step_dt()
    return conv, reason, out


def plant_wilting_factor(pressure):
    """What a host model multiplies its transpiration demand with before it hands it to the VSFM (ELM's plant wilting factor in its
    canopy fluxes): 1 above the soil water potential at which stomata are fully open (smpso = -66 m of water), 0 below the one at which
    they close (smpsc = -255 m), linear in between.  The synthetic forcing of this module is otherwise fixed-rate, which over many
    coupling steps keeps pulling water out of cells that have none left (tools/soak.py)."""
    psi = (np.asarray(pressure) - K.PRESSURE_REF) / (998.2 * 9.80665)          # [m]
    return np.clip((psi + 255.0) / (255.0 - 66.0), 0.0, 1.0)


def elm_vsfm_raw_state(p, d, seed=SEED, patches=True, nlevsoi=10, drain_frac=0.7):
    """ELM's raw column arrays for MPPVSFMALM_Solve (MPPVSFMALM_Driver.F90:110-200), consistent with the state of `p` (h2osoi_liq +
    h2osoi_ice = the SoE's liquid mass, as ELM keeps them after every solve).  Cell arrays are (ncol, nlev), C order."""
    ncol, nlev = d["ncol"], d["nlev"]
    rng = np.random.default_rng(seed + 23)
    z, zi, dz = elm_layers(nlev)
    st = {"zi": np.tile(zi, (ncol, 1)), "dz": np.tile(dz, (ncol, 1)), "nlevsoi": nlevsoi}
    mass = p.get_data(K.AUXVAR_INTERNAL, K.VAR_MASS, 1).reshape(ncol, nlev)
    fi = np.where(rng.uniform(size=(ncol, nlev)) < 0.2, rng.uniform(0.0, 0.6, (ncol, nlev)), 0.0)
    st["h2osoi_liq"] = np.ascontiguousarray((1.0 - fi) * mass); st["h2osoi_ice"] = np.ascontiguousarray(fi * mass)
    w = np.exp(-z[:nlevsoi] * rng.uniform(0.5, 3.0, (ncol, 1))); w /= w.sum(axis=1, keepdims=True)
    rootr = np.zeros((ncol, nlev)); rootr[:, :nlevsoi] = w
    st["rootr_col"] = rootr
    st["qflx_tran_veg_col"] = rng.uniform(0.0, 8e-5, ncol)
    if patches:
        npf = rng.integers(1, 4, ncol).astype(np.int32)
        st["col_npfts"] = npf; st["col_pfti"] = np.concatenate([[0], np.cumsum(npf)[:-1]]).astype(np.int32)
        npft = int(npf.sum()); st["max_patch_per_col"] = 3
        st["pft_active"] = (rng.uniform(size=npft) < 0.9).astype(np.int32)
        wt = rng.uniform(0.1, 1.0, npft)
        owner = np.repeat(np.arange(ncol), npf)
        wt /= np.bincount(owner, wt, ncol)[owner]
        st["pft_wtcol"] = wt
        rp = np.zeros((npft, nlev)); wp = np.exp(-z[:nlevsoi] * rng.uniform(0.5, 3.0, (npft, 1))); rp[:, :nlevsoi] = wp / wp.sum(axis=1, keepdims=True)
        st["rootr_pft"] = rp
        st["qflx_tran_veg_pft"] = rng.uniform(0.0, 1.2e-4, npft)
        st["qflx_tran_veg_col"] = np.bincount(owner, st["qflx_tran_veg_pft"] * wt * st["pft_active"], ncol)
    st["qflx_infl"] = rng.uniform(0.0, 2e-4, ncol)
    st["qflx_dew_snow"] = rng.uniform(0.0, 1e-6, ncol); st["qflx_dew_grnd"] = rng.uniform(0.0, 2e-6, ncol); st["qflx_sub_snow"] = rng.uniform(0.0, 1e-6, ncol)
    st["frac_h2osfc"] = np.where(rng.uniform(size=ncol) < 0.5, rng.uniform(0.0, 0.3, ncol), 0.0)
    st["snl"] = -rng.integers(0, 3, ncol).astype(np.int32)
    st["qflx_drain"] = np.where(rng.uniform(size=ncol) < drain_frac, rng.uniform(0.0, 1e-6, ncol), 0.0)     # baseflow-sized: <= 0.09 mm/day
    # water table where the initial pressure profile crosses P_ref (ELM diagnoses zwt from the VSFM solution, :866-876); a table
    # placed at random would have the drainage pull water out of dry layers, which no Newton iteration can deliver
    P0 = np.asarray(d["press_ic"]).reshape(ncol, nlev)[:, 0]
    st["zwt"] = 0.5 * (zi[0] + zi[1]) + (K.PRESSURE_REF - P0) / (997.16 * K.GRAVITY_CONSTANT)
    st["mflx_snowlyr_col"] = np.where(rng.uniform(size=ncol) < 0.1, rng.uniform(0.0, 1e-5, ncol), 0.0)
    st["mflx_neg_snow_col"] = np.where(rng.uniform(size=ncol) < 0.05, -rng.uniform(0.0, 1e-6, ncol), 0.0)
    # perched drainage only where perched water can be: cells at or near saturation (ELM diagnoses a perched table from the saturated
    # layers above a frozen one); a sink drawn at random would pull water out of dry or frozen cells, and the reference algorithm then
    # halves dt a dozen times and grinds through 10^4 - 10^6 residual evaluations on that one column
    wet = np.asarray(d["press_ic"]).reshape(ncol, nlev) > K.PRESSURE_REF - 5.0e3
    st["mflx_drain_perched"] = np.where((rng.uniform(size=(ncol, nlev)) < 0.05) & wet, -rng.uniform(0.0, 1e-6, (ncol, nlev)), 0.0)
    return st


def copy_state(st):
    return {k: (v.copy() if isinstance(v, np.ndarray) else v) for k, v in st.items()}


def page_aligned_copy(a, page=4096):
    """A copy of `a` that owns whole pages (what a host model's large allocations look like), for mppgpu_host_register."""
    nb = (a.nbytes + page - 1) // page * page
    raw = np.empty(nb + page, dtype=np.uint8)
    off = (-raw.ctypes.data) % page
    out = raw[off:off + a.nbytes].view(a.dtype).reshape(a.shape)
    out[...] = a
    return out


def page_aligned_state(st):
    return {k: (page_aligned_copy(v) if isinstance(v, np.ndarray) and v.nbytes else v) for k, v in st.items()}


# ---------------------------------------------------------------------------------------------------
# thermal_mms (1-D steady state, KSP path) -- src/driver/standalone/thermal/thermal_mms_problem.F90
#   + thermal_mms_steady_state_problem_1D.F90; baseline regression_tests/thermal/thermal_mms.regression.baseline
# ---------------------------------------------------------------------------------------------------
def build_thermal_mms(cls, nx=20, **kw):
    """20 cells along x (one chain = one 'column' with nx layers), cnfac = 0, conductivity exp(x) through tkdry,
    no water (so kappa = kappa_dry and zero heat capacity), Dirichlet at both ends, manufactured source."""
    dx = 1.0 / nx
    xc = dx / 2.0 + dx * np.arange(nx)
    p = cls(1, nx, **kw)
    p.set_mesh(K.MESH_HORIZONTAL, np.full((1, nx), dx), np.array([1.0]))           # area = dy*dz = 1
    p.set_cnfac(0.0)                                                               # :72
    b0 = p.add_condition(1, K.COND_BC, K.COND_DIRICHLET, K.SOIL_TOP_CELLS)          # ii = 1 end
    b1 = p.add_condition(1, K.COND_BC, K.COND_DIRICHLET, K.SOIL_BOTTOM_CELLS)       # ii = nx end
    ss = p.add_condition(1, K.COND_SS, K.COND_HEAT_RATE, K.SOIL_CELLS)              # ALL_CELLS
    full = lambda v: np.full((1, nx), v)
    p.set_soils(full(0.1), full(0.0), full(0.0), np.exp(xc)[None, :], np.array([K.ISTSOIL]), nx, K.ISTSOIL)   # :521-541
    p.set_soln_prev(np.full(nx, 290.0))                                            # :602, :619
    p.set_data(K.AUXVAR_INTERNAL, K.VAR_TUNING_FACTOR, 1, np.ones(nx))
    p.set_data(K.AUXVAR_INTERNAL, K.VAR_LIQ_AREAL_DEN, 1, np.zeros(nx))
    T = lambda x: 10.0 * np.sin(np.pi * x) + 270.0
    for cid, xb in ((b0, xc[0] - dx / 2.0), (b1, xc[-1] + dx / 2.0)):
        p.set_data(K.AUXVAR_BC, K.VAR_BC_SS_CONDITION, cid, np.array([T(xb)]))
        p.set_data(K.AUXVAR_BC, K.VAR_ACTIVE, cid, np.array([1.0]))                # :633-637
        p.set_data(K.AUXVAR_BC, K.VAR_FRAC, cid, np.array([1.0]))
    lam, dlam = np.exp(xc), np.exp(xc)
    dT, d2T = 10.0 * np.pi * np.cos(np.pi * xc), -10.0 * np.pi * np.pi * np.sin(np.pi * xc)
    src = (-dlam * dT - lam * d2T) * dx * 1.0 * 1.0                                # 1D.F90:152-163
    p.set_data(K.AUXVAR_SS, K.VAR_BC_SS_CONDITION, ss, src)
    return p


def run_thermal_mms(p):
    p.pre_step_dt()
    conv, _ = p.step_dt(1.0, 1)                                                    # :200
    assert conv
    return p.get_data(K.AUXVAR_INTERNAL, K.VAR_TEMPERATURE, 1)


# ---------------------------------------------------------------------------------------------------
# ELM-like batched soil thermal columns -- config #2 (MPPThermalTBasedALM_Driver.F90:290-452)
# ---------------------------------------------------------------------------------------------------
def elm_thermal_inputs(ncol, nlev=15, seed=SEED, nlevsoi=10):
    rng = np.random.default_rng(seed + 1)
    z, zi, dz = elm_layers(nlev)
    d = {"ncol": ncol, "nlev": nlev, "nlevsoi": nlevsoi}
    d["dz"] = np.tile(dz, (ncol, 1)); d["area"] = np.ones(ncol)
    # MPPThermalTBasedALM_Initialize.F90:379-381: dist_up = zi(j) - z(j), dist_dn = z(j+1) - zi(j)
    d["dist_up"] = np.tile(zi[1:nlev] - z[:nlev - 1], (ncol, 1)); d["dist_dn"] = np.tile(z[1:] - zi[1:nlev], (ncol, 1))
    d["watsat"] = rng.uniform(0.35, 0.55, (ncol, nlev))
    d["csol"] = rng.uniform(1.9e6, 2.4e6, (ncol, nlev))
    d["tkmg"] = rng.uniform(1.5, 3.5, (ncol, nlev))
    d["tkdry"] = rng.uniform(0.15, 0.3, (ncol, nlev))
    d["lun_type"] = np.full(ncol, K.ISTSOIL, dtype=np.int32)
    T0 = 270.0 + 20.0 * rng.uniform(size=(ncol, nlev))
    theta = rng.uniform(0.1, 0.9, (ncol, nlev)) * d["watsat"]
    water = theta * 1000.0 * d["dz"]                                   # kg m^-2
    frozen = T0 <= 273.15
    d["ice"] = np.where(frozen, 0.3 * water, 0.0).reshape(-1)
    d["liq"] = np.where(frozen, 0.7 * water, water).reshape(-1)
    d["T0"] = T0.reshape(-1)
    d["snow_water"] = np.zeros(ncol * nlev)
    d["nsnow"] = np.zeros(ncol * nlev, dtype=np.int32)
    d["tuning"] = np.ones(ncol * nlev)
    d["hs"] = rng.uniform(-50.0, 150.0, ncol)                          # W m^-2
    d["dhsdT"] = rng.uniform(-20.0, -5.0, ncol)
    d["frac"] = np.ones(ncol)
    d["sabg"] = np.zeros(ncol * nlev)
    return d


def build_elm_thermal(cls, d, **kw):
    ncol, nlev = d["ncol"], d["nlev"]
    p = cls(ncol, nlev, **kw)
    p.set_mesh(K.MESH_ALONG_GRAVITY, d["dz"], d["area"])
    p.set_connection_distances(d["dist_up"], d["dist_dn"])
    ids = {"hs": p.add_condition(1, K.COND_BC, K.COND_HEAT_FLUX, K.SOIL_TOP_CELLS),       # MPPThermalTBasedALM_Initialize.F90:596-601
           "sabg": p.add_condition(1, K.COND_SS, K.COND_HEAT_RATE, K.SOIL_CELLS)}
    p.set_soils(d["watsat"], d["csol"], d["tkmg"], d["tkdry"], d["lun_type"], d["nlevsoi"], K.ISTSOIL)
    return p, ids


def elm_thermal_step(p, ids, d, T, dt=1800.0, nstep=1):
    """One ELM thermal coupling step (MPPThermalTBasedALM_Driver.F90:331-452)."""
    p.set_soln_prev(T)
    for var, key in ((K.VAR_LIQ_AREAL_DEN, "liq"), (K.VAR_ICE_AREAL_DEN, "ice"), (K.VAR_SNOW_WATER, "snow_water"),
                     (K.VAR_TUNING_FACTOR, "tuning")):
        p.set_data(K.AUXVAR_INTERNAL, var, 1, d[key])
    p.set_idata(K.AUXVAR_INTERNAL, K.VAR_NUM_SNOW_LYR, 1, d["nsnow"])
    p.set_data(K.AUXVAR_BC, K.VAR_BC_SS_CONDITION, ids["hs"], d["hs"])
    p.set_data(K.AUXVAR_BC, K.VAR_DHS_DT, ids["hs"], d["dhsdT"])
    p.set_data(K.AUXVAR_BC, K.VAR_FRAC, ids["hs"], d["frac"])
    p.set_data(K.AUXVAR_SS, K.VAR_BC_SS_CONDITION, ids["sabg"], d["sabg"])
    p.pre_step_dt()
    conv, _ = p.step_dt(dt, nstep)
    return conv, p.get_data(K.AUXVAR_INTERNAL, K.VAR_TEMPERATURE, 1)


# ---------------------------------------------------------------------------------------------------
# ELM's real thermal column: snow + standing surface water + soil
#   set-up    src/driver/alm/MPPThermalTBasedALM_Initialize.F90:150-813
#   per step  src/driver/alm/MPPThermalTBasedALM_Driver.F90:150-452 (the packing below mirrors it index for index)
# ---------------------------------------------------------------------------------------------------
CAPR = 0.34                                                            # mpp_varcon.F90:30


def elm_snow_thermal_inputs(ncol, nlev=15, nlevsno=5, seed=SEED, nlevsoi=10, snow="mixed", water="mixed"):
    """ELM-like column state in ELM's own index space: soil layers 1..nlev, snow layers -nlevsno+1..0 (stored at array index
    j + nlevsno), snl = -(active snow layers)."""
    d = elm_thermal_inputs(ncol, nlev, seed, nlevsoi)
    rng = np.random.default_rng(seed + 7)
    z, zi, dz = elm_layers(nlev)
    d["nlevsno"] = nlevsno
    d["soil_top_dist_dn"] = np.full(ncol, z[0] - zi[0])              # z(c,1) - zi(c,0)  (Initialize.F90:575-577)
    if snow == "none":
        nsn = np.zeros(ncol, dtype=np.int64)
    elif snow == "all":
        nsn = rng.integers(1, nlevsno + 1, ncol)
    else:
        nsn = rng.integers(0, nlevsno + 1, ncol)
    d["snl"] = -nsn
    # snow geometry: layer thicknesses grow downward; z / zi negative upward from the soil surface (zi(c,0) = 0)
    sdz = np.zeros((ncol, nlevsno)); sz = np.zeros((ncol, nlevsno)); szi = np.zeros((ncol, nlevsno + 1))   # szi[:, k] = zi(c, k - nlevsno)
    for c in range(ncol):
        th = rng.uniform(0.02, 0.25, nlevsno) * np.linspace(0.5, 2.0, nlevsno)
        for k in range(nlevsno - 1, nlevsno - 1 - nsn[c], -1):       # from the bottom snow layer (j = 0) upwards
            sdz[c, k] = th[k]
            szi[c, k] = szi[c, k + 1] - th[k]
            sz[c, k] = 0.5 * (szi[c, k] + szi[c, k + 1])
    d["snow_dz"], d["snow_z"], d["snow_zi"] = sdz, sz, szi
    d["frac_sno_eff"] = np.where(nsn > 0, rng.uniform(0.2, 1.0, ncol), 0.0)
    rho = rng.uniform(80.0, 400.0, (ncol, nlevsno))                   # bulk density [kg m^-3]
    tot = rho * sdz * d["frac_sno_eff"][:, None]
    wet = rng.uniform(0.0, 0.1, (ncol, nlevsno))
    d["snow_liq"], d["snow_ice"] = tot * wet, tot * (1.0 - wet)
    d["h2osno"] = (d["snow_liq"] + d["snow_ice"]).sum(axis=1)
    d["t_snow"] = np.where(sdz > 0, rng.uniform(250.0, 273.15, (ncol, nlevsno)), 0.0)
    if water == "none":
        fw = np.zeros(ncol)
    elif water == "all":
        fw = rng.uniform(0.01, 0.3, ncol)
    else:
        fw = np.where(rng.uniform(size=ncol) < 0.5, rng.uniform(1e-8, 0.3, ncol), 0.0)   # includes "thin layer" cases
    fw = np.minimum(fw, 0.95 - d["frac_sno_eff"]).clip(min=0.0)
    d["frac_h2osfc"] = fw
    d["h2osfc"] = np.where(fw > 0, rng.uniform(1e-4, 30.0, ncol), 0.0)  # mm
    d["t_h2osfc"] = rng.uniform(273.2, 285.0, ncol)
    d["hs_top_snow"] = rng.uniform(-80.0, 60.0, ncol); d["hs_h2osfc"] = rng.uniform(-50.0, 150.0, ncol); d["hs_soil"] = d["hs"]
    d["sabg_lyr"] = rng.uniform(0.0, 40.0, (ncol, nlevsno + 1))       # (c, -nlevsno+1 : 1)
    d["t_soil"] = d["T0"].reshape(ncol, nlev).copy()
    d["z"], d["zi"], d["dz1"] = z, zi, dz
    return d


def pack_elm_snow_thermal(d):
    """MPPThermalTBasedALM_Driver.F90:204-330: ELM column state -> the 1-D SoE mailbox arrays [snow | ssw | soil]."""
    ncol, nlev, nsno = d["ncol"], d["nlev"], d["nlevsno"]
    N = ncol * (nlev + nsno + 1)
    o = {"T": np.full(N, 273.15), "liq": np.zeros(N), "ice": np.zeros(N), "snow_water": np.zeros(N), "dz": np.zeros(N),
         "dist_up": np.zeros(N), "dist_dn": np.zeros(N), "frac": np.ones(N), "nsnow": np.zeros(N, dtype=np.int32),
         "active": np.zeros(N, dtype=np.int32), "tuning": np.ones(N),
         "hs_snow": np.zeros(ncol), "hs_sh2o": np.zeros(ncol), "hs_soil": np.zeros(ncol), "dhsdT_snow": np.zeros(ncol),
         "dhsdT_sh2o": np.zeros(ncol), "dhsdT_soil": np.zeros(ncol), "frac_soil": np.ones(ncol),
         "sabg_snow": np.zeros(ncol * nsno), "sabg_soil": np.zeros(ncol * nlev)}
    z, zi, dz = d["z"], d["zi"], d["dz1"]
    snl = d["snl"]

    def zs(c, j):                     # z(c, j) for j <= 0 snow, j >= 1 soil
        return d["snow_z"][c, j + nsno - 1] if j <= 0 else z[j - 1]

    def zis(c, j):                    # zi(c, j): interface below layer j
        return d["snow_zi"][c, j + nsno] if j <= 0 else zi[j]
    for c in range(ncol):
        for j in range(-nsno + 1, 1):
            if j >= snl[c] + 1:
                k = j + nsno - 1
                idx = c * nsno + k
                o["T"][idx] = d["t_snow"][c, k]; o["dz"][idx] = d["snow_dz"][c, k]
                o["liq"][idx] = d["snow_liq"][c, k]; o["ice"][idx] = d["snow_ice"][c, k]
                o["nsnow"][idx] = -snl[c]; o["active"][idx] = 1
                o["dist_up"][idx] = zis(c, j) - zs(c, j); o["dist_dn"][idx] = zs(c, j) - zis(c, j - 1)
                o["frac"][idx] = d["frac_sno_eff"][c]
                if j != snl[c] + 1:
                    o["sabg_snow"][idx] = d["sabg_lyr"][c, k]
                if j == snl[c] + 1:
                    o["tuning"][idx] = d["snow_dz"][c, k] / (0.5 * (zs(c, j) - zis(c, j - 1) + CAPR * (zs(c, j + 1) - zis(c, j - 1))))
                    o["hs_snow"][c] = d["hs_top_snow"][c]; o["dhsdT_snow"][c] = d["dhsdT"][c]
                    o["frac_soil"][c] -= d["frac_sno_eff"][c]
    off = ncol * nsno
    for c in range(ncol):
        if d["frac_h2osfc"][c] > 0.0:
            idx = off + c
            o["T"][idx] = d["t_h2osfc"][c]; o["dz"][idx] = 1.0e-3 * d["h2osfc"][c]; o["active"][idx] = 1
            o["frac"][idx] = d["frac_h2osfc"][c]; o["dist_up"][idx] = o["dz"][idx] / 2.0; o["dist_dn"][idx] = o["dz"][idx] / 2.0
            o["frac_soil"][c] -= d["frac_h2osfc"][c]; o["dhsdT_sh2o"][c] = d["dhsdT"][c]; o["hs_sh2o"][c] = d["hs_h2osfc"][c]
    off = ncol * (nsno + 1)
    liq, ice = d["liq"].reshape(ncol, nlev), d["ice"].reshape(ncol, nlev)
    for c in range(ncol):
        for j in range(1, nlev + 1):
            idx = off + c * nlev + j - 1
            o["T"][idx] = d["t_soil"][c, j - 1]; o["dz"][idx] = dz[j - 1]; o["active"][idx] = 1
            o["liq"][idx] = liq[c, j - 1]; o["ice"][idx] = ice[c, j - 1]; o["frac"][idx] = 1.0
            o["dist_up"][idx] = zi[j] - z[j - 1]; o["dist_dn"][idx] = zi[j] - z[j - 1]
            if j == 1:
                o["dz"][idx] = z[0] * 2.0; o["nsnow"][idx] = -snl[c]
                if snl[c] != 0:
                    o["sabg_soil"][c * nlev] = d["frac_sno_eff"][c] * d["sabg_lyr"][c, nsno]
                    o["snow_water"][idx] = d["h2osno"][c]
                else:
                    o["tuning"][idx] = dz[0] / (0.5 * (z[0] - zi[0] + CAPR * (z[1] - zi[0])))
                o["hs_soil"][c] = d["hs_soil"][c]; o["dhsdT_soil"][c] = d["dhsdT"][c]
    return o


def tile_snow_thermal(d, o, reps):
    """Replicate a packed batch `reps` times (bench-sized inputs without the per-column Python packing loop): every segment of the
    SoE order [snow | ssw | soil] is tiled on its own."""
    ncol, nlev, nsno = d["ncol"], d["nlev"], d["nlevsno"]
    D = dict(d); D["ncol"] = ncol * reps
    for k in ("dz", "dist_up", "dist_dn", "watsat", "csol", "tkmg", "tkdry"):
        D[k] = np.tile(d[k], (reps, 1))
    for k in ("area", "lun_type", "soil_top_dist_dn"):
        D[k] = np.tile(d[k], reps)
    O = {}
    a, b = ncol * nsno, ncol * (nsno + 1)
    for k, v in o.items():
        if v.size == ncol * (nsno + 1 + nlev):
            O[k] = np.concatenate([np.tile(v[:a], reps), np.tile(v[a:b], reps), np.tile(v[b:], reps)])
        else:
            O[k] = np.tile(v, reps)
    return D, O


def elm_thermal_raw_arrays(d):
    """The same column state as ELM holds it (MPPThermalTBasedALM_Driver.F90:60-150): Fortran (c, j) arrays, here numpy (nlayers, ncol)."""
    ncol, nlev, nsno = d["ncol"], d["nlev"], d["nlevsno"]
    z, zi, dz = d["z"], d["zi"], d["dz1"]
    col = lambda a: np.ascontiguousarray(a.T)
    e = {"snl": d["snl"].astype(np.int32),
         "z": col(np.hstack([d["snow_z"], np.tile(z, (ncol, 1))])), "dz": col(np.hstack([d["snow_dz"], np.tile(dz, (ncol, 1))])),
         "zi": col(np.hstack([d["snow_zi"], np.tile(zi[1:], (ncol, 1))])),
         "t_soisno": col(np.hstack([d["t_snow"], d["t_soil"]])),
         "h2osoi_liq": col(np.hstack([d["snow_liq"], d["liq"].reshape(ncol, nlev)])), "h2osoi_ice": col(np.hstack([d["snow_ice"], d["ice"].reshape(ncol, nlev)])),
         "sabg_lyr": col(d["sabg_lyr"]), "tvector": np.full((nsno + 1 + nlev, ncol), -999.0)}
    for k in ("frac_sno_eff", "h2osno", "h2osfc", "frac_h2osfc", "t_h2osfc", "dhsdT", "hs_soil", "hs_top_snow", "hs_h2osfc"):
        e[k] = np.ascontiguousarray(d[k], dtype=np.float64)
    return e


def unpack_elm_snow_thermal(d, o, T, tvector):
    """MPPThermalTBasedALM_Driver.F90:460-505: SoE solution -> tvector(c, -nlevsno:nlev), here (nlayers, ncol)."""
    ncol, nlev, nsno = d["ncol"], d["nlev"], d["nlevsno"]
    snl = d["snl"]
    for c in range(ncol):
        for j in range(-nsno + 1, 1):
            if j >= snl[c] + 1:
                tvector[j - 1 + nsno, c] = T[c * nsno + j + nsno - 1]
        if d["frac_h2osfc"][c] > 0.0:
            tvector[nsno, c] = T[ncol * nsno + c]
        for j in range(1, nlev + 1):
            tvector[j + nsno, c] = T[ncol * (nsno + 1) + c * nlev + j - 1]
    return tvector


def build_elm_snow_thermal(cls, d, **kw):
    ncol, nlev, nsno = d["ncol"], d["nlev"], d["nlevsno"]
    p = cls(ncol, nlev, nsno, **kw)
    p.set_mesh(d["dz"], d["area"], d["dist_up"], d["dist_dn"], d["soil_top_dist_dn"])
    p.set_soils(d["watsat"], d["csol"], d["tkmg"], d["tkdry"], d["lun_type"], d["nlevsoi"], K.ISTSOIL)
    return p


def elm_snow_thermal_step(p, o, dt=1800.0, nstep=1):
    """The SetSolnPrevCLM / Set{R,I,B}DataFromCLM / PreStepDT / StepDT / GetSoln sequence of MPPThermalTBasedALM_Driver.F90:332-452."""
    p.set_soln_prev(o["T"])
    for var, key in ((K.VAR_LIQ_AREAL_DEN, "liq"), (K.VAR_ICE_AREAL_DEN, "ice"), (K.VAR_SNOW_WATER, "snow_water"), (K.VAR_DZ, "dz"),
                     (K.VAR_DIST_UP, "dist_up"), (K.VAR_DIST_DN, "dist_dn"), (K.VAR_TUNING_FACTOR, "tuning"), (K.VAR_FRAC, "frac")):
        p.set_data(K.AUXVAR_INTERNAL, var, 1, o[key])
    p.set_idata(K.AUXVAR_INTERNAL, K.VAR_NUM_SNOW_LYR, 1, o["nsnow"])
    p.set_idata(K.AUXVAR_INTERNAL, K.VAR_ACTIVE, 1, o["active"])
    for cid, a, b in ((1, "hs_snow", "dhsdT_snow"), (2, "hs_sh2o", "dhsdT_sh2o"), (3, "hs_soil", "dhsdT_soil")):
        p.set_data(K.AUXVAR_BC, K.VAR_BC_SS_CONDITION, cid, o[a])
        p.set_data(K.AUXVAR_BC, K.VAR_DHS_DT, cid, o[b])
    p.set_data(K.AUXVAR_BC, K.VAR_FRAC, 3, o["frac_soil"])
    p.set_data(K.AUXVAR_SS, K.VAR_BC_SS_CONDITION, 1, o["sabg_snow"])
    p.set_data(K.AUXVAR_SS, K.VAR_BC_SS_CONDITION, 2, o["sabg_soil"])
    p.pre_step_dt()
    conv, _ = p.step_dt(dt, nstep)
    return conv, p.get_soln()


# ---------------------------------------------------------------------------------------------------
# TH mass_and_heat -- src/driver/standalone/thermal-e/mass_and_heat_model_problem.F90
#   baseline regression_tests/th/mass_and_heat.regression.baseline
# ---------------------------------------------------------------------------------------------------
def build_mass_and_heat(cls, nx=100, **kw):
    p = cls(1, nx, **kw)
    dx = 1.0 / nx
    p.set_mesh(K.MESH_HORIZONTAL, np.full((1, nx), dx), np.array([1.0]))                  # CONN_IN_X_DIR, area = dy*dz = 1
    b0 = p.add_condition(2, K.COND_BC, K.COND_DIRICHLET, K.SOIL_TOP_CELLS)                 # cell 1, unit vector (+1,0,0)  :289-310
    b1 = p.add_condition(2, K.COND_BC, K.COND_DIRICHLET, K.SOIL_BOTTOM_CELLS)              # cell nx, unit vector (-1,0,0) :312-327
    porosity, lam, alpha, perm = 0.368, 0.5, 3.4257e-4, 8.3913e-12
    hksat = perm / 0.001002 * (1000.0 * K.GRAV) / 0.001
    sucsat = 1.0 / (alpha * K.GRAVITY_CONSTANT)                                             # mass_and_heat_model_problem.F90:458 (the driver inverts with 9.80665, MPPTHSetSoils converts with 9.80616)
    full = lambda v: np.full((1, nx), v)
    p.set_soils(full(porosity), full(hksat), full(1.0 / lam), full(sucsat), full(0.2772), full(837.0), full(0.25),
                "van_genuchten", K.DENSITY_IFC67, K.INT_ENERGY_ENTHALPY_IFC67)             # :458-472
    p.restart(np.full(nx, 91325.0), np.full(nx, 283.15))                                   # :529-531
    return p, b0, b1


def run_mass_and_heat(p, b0, b1, dt=3600.0):
    p.set_data(K.AUXVAR_BC, K.VAR_BC_SS_CONDITION, b0, np.array([303.15]), ieqn=2)         # :586-597
    p.set_data(K.AUXVAR_BC, K.VAR_BC_SS_CONDITION, b1, np.array([293.15]), ieqn=2)
    p.set_data(K.AUXVAR_BC, K.VAR_PRESSURE, b0, np.array([91325.0]), ieqn=2)               # :616-621 (pokes aux_vars_bc%pressure)
    p.set_data(K.AUXVAR_BC, K.VAR_PRESSURE, b1, np.array([91325.0]), ieqn=2)
    conv, reason = p.step_dt(dt, 1)
    P = p.get_data(K.AUXVAR_INTERNAL, K.VAR_PRESSURE, -1, ieqn=1)
    T = p.get_data(K.AUXVAR_INTERNAL, K.VAR_TEMPERATURE, -1, ieqn=2)
    return conv, reason, P, T


# ---------------------------------------------------------------------------------------------------
# TH th_mms -- src/driver/standalone/thermal-e/th_mms_problem.F90 (manufactured steady state, 20 cells along x)
#   baseline regression_tests/th/th_mms.regression.baseline; registered in regression_tests/th/th.cfg:8-9 (temperature 1e-8 K)
# `phys` supplies the scalar EOS / saturation-function calls the reference's driver makes while building its source terms
# (set_variable_for_problem :1158-1458): oracle.OraclePhysics for the oracle, mpp_b200.hostphysics.HostPhysics for the CUDA path.
# ---------------------------------------------------------------------------------------------------
def th_mms_data(phys, nx=20):
    x_min, x_max = 0.0, 10.0                                                         # set_default_problem :937-955
    xlim = x_max - x_min
    dx = (x_max - x_min) / nx                                                        # set_problem :958-988
    pi = 4.0 * np.arctan(1.0)
    dens_type, iee_type = K.DENSITY_CONSTANT, K.INT_ENERGY_ENTHALPY_IFC67            # :962-964 (the second assignment wins)
    FMW = K.FMWH2O
    xc = np.array([dx / 2.0 + dx * i + x_min for i in range(nx)])                    # mpp_mesh_utils.F90:212

    def pres(x):                                                                     # compute_pressure_or_deriv :991-1020
        a0, a1 = 15000.0, -20000.0
        return (a0 * np.sin((x - x_min) / xlim * pi) + a1 + K.PRESSURE_REF, a0 * pi / xlim * np.cos((x - x_min) / xlim * pi),
                -a0 * pi * pi / xlim / xlim * np.sin((x - x_min) / xlim * pi))

    def temp(x):                                                                     # compute_temperature_or_deriv :1023-1043 (second set)
        a0, a1 = 5.0, 290.0
        return (a0 * np.sin((x - x_min) / xlim * pi) * 1.0 + a1, a0 * pi / xlim * np.cos((x - x_min) / xlim * pi) * 1.0,
                -a0 * pi * pi / xlim / xlim * np.sin((x - x_min) / xlim * pi) * 1.0)

    def perm(x):                                                                     # compute_permeability_or_deriv :1046-1059
        p0 = 1.0e-11
        return p0 * (2.0 - np.cos((x - x_min) / xlim * pi)), p0 * pi / xlim * (+np.sin((x - x_min) / xlim * pi))

    alpha, lam, sat_res, kdry, kwet, talpha, pert = 1.0 / 4000.0, 0.5, 0.0, 0.25, 1.3, 0.45, 1.0e-6     # :1062-1155, :1203
    d = {"nx": nx, "dx": dx, "xc": xc, "alpha": alpha, "lam": lam, "sat_res": sat_res, "tkdry": kdry,
         "density_type": dens_type, "iee_type": iee_type}
    d["perm"] = np.array([perm(x)[0] for x in xc])
    Pex = np.array([pres(x)[0] for x in xc]); Tex = np.array([temp(x)[0] for x in xc])
    d["P_exact"], d["T_exact"] = Pex, Tex
    P0 = 0.0
    for v in Pex:
        P0 = P0 + 1.0 / nx * v                                                       # DATA_INITIAL_PRESSURE :1212-1219
    T0 = 0.0
    for v in Tex:
        T0 = T0 + 1.0 / nx * v
    d["press_ic"], d["temp_ic"] = np.full(nx, P0), np.full(nx, T0)
    d["pres_bc"] = np.array([pres(xc[0] - dx / 2.0)[0], pres(xc[-1] + dx / 2.0)[0]])  # DATA_PRESSURE_BC :1250-1262
    d["temp_bc"] = np.array([temp(xc[0] - dx / 2.0)[0], temp(xc[-1] + dx / 2.0)[0]])
    msrc, hsrc = np.zeros(nx), np.zeros(nx)
    for i, x in enumerate(xc):
        xp, xn = x + pert, x - pert
        k, dk_dx = perm(x)
        P, dP_dx, d2P_dx2 = pres(x)
        T, dT_dx, d2T_dx2 = temp(x)
        mu = phys.viscosity(P, T)
        rho, drho_dP, drho_dT = phys.density(P, T, dens_type)
        rho, drho_dP, drho_dT = rho * FMW, drho_dP * FMW, drho_dT * FMW
        U, H, dU_dT, dH_dT, dU_dP, dH_dP = phys.internal_energy_enthalpy(P, T, iee_type, rho, drho_dT, drho_dP)
        se, dse_dP = phys.vg_sat(P, sat_res, alpha, lam)
        kr, dkr_dP = phys.vg_relperm(P, sat_res, alpha, lam)
        dkr_dx = dkr_dP * dP_dx                                                      # (+ dkr_dse * dse_dp0 * dp0_dx with dse_dp0 = dp0_dx = 0)
        Pp, Tp, Pn, Tn = pres(xp)[0], temp(xp)[0], pres(xn)[0], temp(xn)[0]
        rp, rp_dP, rp_dT = phys.density(Pp, Tp, dens_type)
        rn, rn_dP, rn_dT = phys.density(Pn, Tn, dens_type)
        rp, rn = rp * FMW, rn * FMW
        drho_dx = (rp - rn) / pert / 2.0
        drhoq_dx = -((k * kr / mu) * drho_dx + (rho * kr / mu) * dk_dx + (rho * k / mu) * dkr_dx) * (dP_dx) - (rho * k * kr / mu) * (d2P_dx2)
        msrc[i] = drhoq_dx * dx                                                      # DATA_MASS_SOURCE :1264-1311
        rhoq = -rho * (k * kr / mu * dP_dx)
        Ke = (se + 1.e-6) ** (talpha)
        sep = phys.vg_sat(Pp, sat_res, alpha, lam)[0]; sen = phys.vg_sat(Pn, sat_res, alpha, lam)[0]
        dKe_dx = ((sep + 1.e-6) ** (talpha) - (sen + 1.e-6) ** (talpha)) / pert / 2.0
        kappa = kwet * Ke + kdry * (1.0 - Ke)
        dkappa_dx = 0.0 * Ke + 0.0 * (1.0 - Ke) + (kwet - kdry) * dKe_dx
        Hp = phys.internal_energy_enthalpy(Pp, Tp, iee_type, rp, rp_dT * FMW, rp_dP * FMW)[1]
        Hn = phys.internal_energy_enthalpy(Pn, Tn, iee_type, rn, rn_dT * FMW, rn_dP * FMW)[1]
        dH_dx = (Hp - Hn) / pert / 2.0
        hsrc[i] = -(drhoq_dx * H / FMW + rhoq * dH_dx / FMW - dkappa_dx * dT_dx - kappa * d2T_dx2) * dx     # DATA_HEAT_SOURCE :1353-1449
    d["mass_source"], d["heat_source"] = msrc, hsrc
    return d


def build_th_mms(cls, phys, nx=20, data=None, **kw):
    """`data`: a th_mms_data() result to reuse (parity tests hand both implementations the SAME source arrays: the driver's
    finite-difference formulas, step 1e-6, amplify the last-ulp differences between two EOS implementations to ~1e-5 relative)."""
    d = data if data is not None else th_mms_data(phys, nx)
    p = cls(1, nx, **kw)
    p.set_mesh(K.MESH_HORIZONTAL, np.full((1, nx), d["dx"]), np.array([1.0]))             # connections along x, area = dy*dz = 1 (:205-292)
    ids = {"p0": p.add_condition(1, K.COND_BC, K.COND_DIRICHLET, K.SOIL_TOP_CELLS),       # 'Pressure BC', both ends (:343-355)
           "p1": p.add_condition(1, K.COND_BC, K.COND_DIRICHLET, K.SOIL_BOTTOM_CELLS),
           "msrc": p.add_condition(1, K.COND_SS, K.COND_MASS_RATE, K.SOIL_CELLS),         # 'Source term for MMS', ALL_CELLS (:357-362)
           "t0": p.add_condition(2, K.COND_BC, K.COND_DIRICHLET, K.SOIL_TOP_CELLS),       # 'Temperature BC' (:373-385)
           "t1": p.add_condition(2, K.COND_BC, K.COND_DIRICHLET, K.SOIL_BOTTOM_CELLS),
           "hsrc": p.add_condition(2, K.COND_SS, K.COND_HEAT_RATE, K.SOIL_CELLS)}         # (:387-392)
    # the setter takes ELM's tables: invert VSFMMPPSetSoilsCLM's conversions (as build_mass_and_heat does)
    hksat = d["perm"] / 0.001002 * (1000.0 * K.GRAV) / 0.001
    sucsat = 1.0 / (d["alpha"] * K.GRAV)                                                  # this driver sets alpha directly (:671): invert MPPTHSetSoils exactly, alpha = 1 / (sucsat * grav), MultiPhysicsProbTH.F90:283
    full = lambda v: np.full((1, nx), v)
    p.set_soils(full(0.0), hksat.reshape(1, nx), full(1.0 / d["lam"]), full(sucsat), full(d["sat_res"]), full(0.0), full(d["tkdry"]),
                "van_genuchten", d["density_type"], d["iee_type"])                        # porosity 0, heat capacity 0: a steady state (:1221-1224, :1313-1316)
    p.set_energy_permeability(d["perm"])                                                  # goveq_enthalpy%SetSoilPermeability (:739)
    p.restart(d["press_ic"], d["temp_ic"])
    return p, ids, d


def run_th_mms(p, ids, d):
    """set_source_sink_conditions + set_boundary_conditions + one StepDT(1 s) (run_th_mms_problem :89-141)."""
    p.set_data(K.AUXVAR_SS, K.VAR_BC_SS_CONDITION, ids["msrc"], d["mass_source"], ieqn=1)
    p.set_data(K.AUXVAR_SS, K.VAR_BC_SS_CONDITION, ids["hsrc"], d["heat_source"], ieqn=2)
    for k, key in ((0, "p0"), (1, "p1")):
        p.set_data(K.AUXVAR_BC, K.VAR_BC_SS_CONDITION, ids[key], d["pres_bc"][k:k + 1], ieqn=1)
    for k, key in ((0, "t0"), (1, "t1")):
        p.set_data(K.AUXVAR_BC, K.VAR_BC_SS_CONDITION, ids[key], d["temp_bc"][k:k + 1], ieqn=2)
        p.set_data(K.AUXVAR_BC, K.VAR_PRESSURE, ids[key], d["pres_bc"][k:k + 1], ieqn=2)   # aux_vars_bc%pressure poked by the driver (:858-868)
    conv, reason = p.step_dt(1.0, 1)
    P = p.get_data(K.AUXVAR_INTERNAL, K.VAR_PRESSURE, -1, ieqn=1)
    T = p.get_data(K.AUXVAR_INTERNAL, K.VAR_TEMPERATURE, -1, ieqn=2)
    return conv, reason, P, T


# ---------------------------------------------------------------------------------------------------
# ELM-like batched TH columns -- config #5: VSFM soils + csol, tkdry; Dirichlet temperature
# at the surface (energy equation), mass-rate infiltration (mass equation), heat-rate source (energy equation)
# ---------------------------------------------------------------------------------------------------
def elm_th_inputs(ncol, nlev=15, seed=SEED, satfunc="van_genuchten", density_type=K.DENSITY_TGDPB01,
                  iee_type=K.INT_ENERGY_ENTHALPY_CONSTANT, zwt_min=1.0):
    d = elm_vsfm_inputs(ncol, nlev, seed=seed, satfunc=satfunc, zwt_min=zwt_min)
    rng = np.random.default_rng(seed + 2)
    d["csol"] = rng.uniform(700.0, 900.0, (ncol, nlev))              # J kg^-1 K^-1
    d["tkdry"] = rng.uniform(0.15, 0.3, (ncol, nlev))
    d["temp_ic"] = (283.15 + rng.uniform(-5.0, 15.0, (ncol, nlev))).reshape(-1)
    d["T_top"] = 283.15 + rng.uniform(-5.0, 15.0, ncol)
    d["P_top_bc"] = d["press_ic"].reshape(ncol, nlev)[:, 0].copy()   # the drivers poke aux_vars_bc%pressure (mass_and_heat :616-621)
    d["heat"] = rng.uniform(0.0, 5.0, ncol * nlev) * np.tile(elm_layers(nlev)[2], ncol)   # W per cell
    d["density_type"], d["iee_type"] = density_type, iee_type
    return d


def build_elm_th(cls, d, **kw):
    ncol, nlev = d["ncol"], d["nlev"]
    p = cls(ncol, nlev, **kw)
    p.set_mesh(K.MESH_ALONG_GRAVITY, d["dz"], d["area"])
    ids = {"T_top": p.add_condition(2, K.COND_BC, K.COND_DIRICHLET, K.SOIL_TOP_CELLS),
           "infil": p.add_condition(1, K.COND_SS, K.COND_MASS_RATE, K.SOIL_TOP_CELLS),
           "heat": p.add_condition(2, K.COND_SS, K.COND_HEAT_RATE, K.SOIL_CELLS)}
    p.set_soils(d["watsat"], d["hksat"], d["bsw"], d["sucsat"], d["residual_sat"], d["csol"], d["tkdry"],
                d["satfunc"], d["density_type"], d["iee_type"])
    p.restart(d["press_ic"], d["temp_ic"])
    return p, ids


def elm_th_step(p, ids, d, dt=1800.0, nstep=1):
    p.set_data(K.AUXVAR_BC, K.VAR_BC_SS_CONDITION, ids["T_top"], d["T_top"], ieqn=2)
    p.set_data(K.AUXVAR_BC, K.VAR_PRESSURE, ids["T_top"], d["P_top_bc"], ieqn=2)
    p.set_data(K.AUXVAR_SS, K.VAR_BC_SS_CONDITION, ids["infil"], d["infil"], ieqn=1)
    p.set_data(K.AUXVAR_SS, K.VAR_BC_SS_CONDITION, ids["heat"], d["heat"], ieqn=2)
    p.pre_step_dt()
    conv, reason = p.step_dt(dt, nstep)
    out = {"pressure": p.get_data(K.AUXVAR_INTERNAL, K.VAR_PRESSURE, 1, ieqn=1),
           "temperature": p.get_data(K.AUXVAR_INTERNAL, K.VAR_TEMPERATURE, 1, ieqn=2),
           "sat": p.get_data(K.AUXVAR_INTERNAL, K.VAR_LIQ_SAT, 1, ieqn=1),
           "mass": p.get_data(K.AUXVAR_INTERNAL, K.VAR_MASS, 1, ieqn=1)}
    p.post_step_dt()
    return conv, reason, out
