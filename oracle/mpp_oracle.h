/*
 * mpp_oracle.h -- CPU restatement ("oracle") of the MPP column-physics hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product:
 * only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
 * reference legs may load this library, and only as the checker / CPU baseline.
 * The shipped path is mpp_b200/csrc (CUDA, sm_100a) behind include/mppgpu.h.
 *
 * The reference (MPP-LSM/MPP) is Fortran 2003 + PETSc and cannot be built in
 * this environment (no gfortran / MPI / PETSc), so this is a from-source
 * restatement in plain C.  Every function cites the reference file:line it
 * follows (paths relative to the reference root).  The arithmetic that lives
 * in the un-vendored dependency PETSc (pinned v3.16.2 in README.md:35; git
 * a12052c5 in .ci-scripts/install-petsc.sh:5) -- SNES newtonls, the `bt` line
 * search, SNESConvergedDefault and the KSP/PC solve -- is restated from the
 * published PETSc 3.16 algorithm in snes.c.
 *
 * Pinning: tests/test_oracle_golden.py checks this oracle against the
 * reference's own goldens: src/tests/test_eos_*_density.F90 (EOS known answers),
 * regression_tests/vsfm/vsfm_celia1990.regression.baseline,
 * regression_tests/thermal/thermal_mms.regression.baseline and
 * regression_tests/th/mass_and_heat.regression.baseline.
 */
#ifndef MPP_ORACLE_H
#define MPP_ORACLE_H

#ifdef __cplusplus
extern "C" {
#endif

/* ---- constants: src/mpp/util/MultiPhysicsProbConstants.F90:17-213 ---------- */
#define ORC_PRESSURE_REF      101325.0      /* :199 */
#define ORC_GRAVITY_CONSTANT  9.80665       /* :201 */
#define ORC_FMWH2O            18.01534      /* :202 */
/* src/mpp/util/mpp_varcon.F90:12-28 */
#define ORC_GRAV        9.80616
#define ORC_CPLIQ       4.188e3
#define ORC_CPICE       2.11727e3
#define ORC_DENH2O      1.000e3
#define ORC_DENICE      0.917e3
#define ORC_TKICE       2.290
#define ORC_TKWAT       0.57
#define ORC_THK_BEDROCK 3.0
#define ORC_TFRZ        273.15

/* ids reused verbatim from MultiPhysicsProbConstants.F90 */
enum {
  SOE_RE_ODE = 101, SOE_THERMAL_TBASED = 102, SOE_TH = 104,
  MESH_ALONG_GRAVITY = 311, MESH_AGAINST_GRAVITY = 312, MESH_HORIZONTAL = 313 /* ours: CONN_IN_X_DIR meshes */,
  SOIL_TOP_CELLS = 401, SOIL_BOTTOM_CELLS = 402, SOIL_CELLS = 403,
  COND_NULL = 500, COND_BC = 501, COND_SS = 502, COND_MASS_RATE = 503, COND_MASS_FLUX = 504,
  COND_DIRICHLET = 505, COND_DIRICHLET_FRM_OTR_GOVEQ = 506, COND_HEAT_FLUX = 507,
  COND_SEEPAGE_BC = 509, COND_HEAT_RATE = 511, COND_DOWNREG_MASS_RATE_CAMPBELL = 512, COND_DOWNREG_MASS_RATE_FETCH2 = 513,
  VAR_PRESSURE = 604, VAR_TEMPERATURE = 605, VAR_BC_SS_CONDITION = 607, VAR_LIQ_SAT = 608,
  VAR_MASS = 610, VAR_SOIL_MATRIX_POT = 611, VAR_FRAC_LIQ_SAT = 612,
  VAR_LIQ_AREAL_DEN = 615, VAR_ICE_AREAL_DEN = 617, VAR_FRAC = 618, VAR_SNOW_WATER = 619,
  VAR_NUM_SNOW_LYR = 620, VAR_DHS_DT = 621, VAR_THERMAL_COND = 622, VAR_HEAT_CAP = 623,
  VAR_ACTIVE = 624, VAR_DZ = 627, VAR_DIST_UP = 628, VAR_DIST_DN = 629, VAR_TUNING_FACTOR = 630,
  VAR_POT_MASS_SINK_PRESSURE = 638, VAR_POT_MASS_SINK_EXPONENT = 639, VAR_MASS_FLUX = 644,
  AUXVAR_INTERNAL = 701, AUXVAR_BC = 702, AUXVAR_SS = 703
};

/* src/mpp/util/EOSWaterMod.F90:18-23 */
enum { DENSITY_CONSTANT = 1, DENSITY_TGDPB01 = 2, DENSITY_IFC67 = 3 };
enum { INT_ENERGY_ENTHALPY_CONSTANT = 1, INT_ENERGY_ENTHALPY_IFC67 = 2 };

/* src/mpp/util/SaturationFunction.F90:19-28 */
enum { SAT_FUNC_VAN_GENUCHTEN = 1301, SAT_FUNC_BROOKS_COREY = 1302,
       SAT_FUNC_SMOOTHED_BROOKS_COREY = 1303, RELPERM_FUNC_MUALEM = 1308 };

/* names accepted by VSFMMPPSetSoilsCLM (MultiPhysicsProbVSFM.F90:391-417) */
enum { SATFUNC_NAME_VAN_GENUCHTEN = 0, SATFUNC_NAME_BROOKS_COREY = 1,
       SATFUNC_NAME_SBC_BZ2 = 2, SATFUNC_NAME_SBC_BZ3 = 3 };

/* PETSc SNESConvergedReason values the reference's drivers branch on
 * (MPPVSFMALM_Driver.F90:894-898) */
enum { SNES_CONVERGED_ITERATING = 0, SNES_CONVERGED_FNORM_ABS = 2, SNES_CONVERGED_FNORM_RELATIVE = 3,
       SNES_CONVERGED_SNORM_RELATIVE = 4, SNES_DIVERGED_FUNCTION_COUNT = -2, SNES_DIVERGED_LINEAR_SOLVE = -3,
       SNES_DIVERGED_FNORM_NAN = -4, SNES_DIVERGED_MAX_IT = -5, SNES_DIVERGED_LINE_SEARCH = -6,
       SNES_DIVERGED_DTOL = -9 };

/* ---- EOS: src/mpp/util/EOSWaterMod.F90 -------------------------------------- */
void orc_density(double p, double t_K, int density_itype, double *den, double *dden_dp, double *dden_dT);
void orc_density_constant(double *den, double *dden_dp, double *dden_dT);
void orc_density_tgdpb01(double p, double t_K, double *den, double *dden_dp, double *dden_dT);
void orc_density_ifc67(double t_C, double p, int calc_deriv, double *dw, double *dwmol, double *dwp, double *dwt);
void orc_enthalpy_ifc67(double t_C, double p, int calc_deriv, double *hw, double *hwp, double *hwt);
void orc_viscosity(double p, double t_K, double *vis, double *dvis_dp, double *dvis_dT);
void orc_internal_energy_enthalpy(double P, double t_K, int itype, double den, double dden_dT, double dden_dP,
                                  double *U, double *H, double *dU_dT, double *dH_dT, double *dU_dP, double *dH_dP);

/* ---- saturation functions: src/mpp/util/SaturationFunction.F90 -------------- */
typedef struct {
  int    sat_func_type, relperm_func_type;
  double sat_res, alpha, vg_m, vg_n, bc_lambda, sbc_pu, sbc_ps, sbc_b2, sbc_b3;
} orc_satparams;

int  orc_satfunc_set_vg(orc_satparams *sp, double sat_res, double alpha, double vg_m);
int  orc_satfunc_set_bc(orc_satparams *sp, double sat_res, double alpha, double lambda);
int  orc_satfunc_set_sbc_bz2(orc_satparams *sp, double sat_res, double alpha, double lambda, double ps);
int  orc_satfunc_set_sbc_bz3(orc_satparams *sp, double sat_res, double alpha, double lambda, double ps);
double orc_findgu_sbc_zerocoeff(double lambda, int AA, double gs);
void orc_press_to_sat(const orc_satparams *sp, double press, double *sat, double *dsat_dP);
void orc_press_to_relperm(const orc_satparams *sp, double press, double frac_liq, double *kr, double *dkr_dP);

/* ---- Richards aux var: src/mpp/auxvar/RichardsODEPressureAuxType.F90:18-67 -- */
typedef struct {
  double pressure, temperature, frac_liq_sat, condition_value;
  double perm[3], por;
  int    density_type;
  double vis, kr, sat, den;
  double dpor_dP, dvis_dP, dkr_dP, dsat_dP, dden_dP, dvis_dT, dden_dT;
  orc_satparams satParams;
  double por_base;                 /* PorosityFunctionMod constant model */
} orc_rich_auxvar;

void orc_rich_auxvar_init(orc_rich_auxvar *a);
void orc_rich_auxvar_compute(orc_rich_auxvar *a);

/* src/mpp/dtypes/ConnectionSetType.F90:15-47 */
typedef struct {
  int    id_up, id_dn;             /* 0-based here; -1 = boundary */
  double area, dist_up, dist_dn, unitvec[3];
} orc_conn;

/* src/mpp/ge/RichardsMod.F90:29-340 */
void orc_richards_flux(const orc_rich_auxvar *up, const orc_rich_auxvar *dn, const orc_conn *conn,
                       int compute_deriv, int internal_conn, int swap_order, int cond_type,
                       double *flux, double *dflux_dP_up, double *dflux_dP_dn);

/* src/mpp/ge/RichardsMod.F90:343-648 (true derivatives wrt temperature; swap_order = .false.) */
void orc_richards_flux_dT(const orc_rich_auxvar *up, const orc_rich_auxvar *dn, const orc_conn *conn,
                          int internal_conn, int cond_type, double *flux, double *dflux_dT_up, double *dflux_dT_dn);

/* ---- PETSc-restated nonlinear / linear solver pieces (snes.c) ---------------- */
typedef struct {
  double atol, rtol, stol, divtol;
  int    max_it, max_funcs;
  /* line search bt: alpha, minlambda (= steptol), maxstep, max cubic fits */
  double ls_alpha, ls_minlambda, ls_maxstep;
  int    ls_max_its;
} orc_snes_opts;
void orc_snes_default_opts(orc_snes_opts *o);

typedef struct {
  int reason, its, nfuncs;
  double fnorm0, fnorm, xnorm, ynorm, last_lambda;
} orc_snes_result;

/* generic callbacks over an n-vector; J is block-tridiagonal with block size bs
 * stored as a[n*bs], b[n*bs], c[n*bs] row-major blocks (sub, diag, super), bs in {1,2} */
typedef struct orc_system {
  int n;                           /* unknown count (cells * bs) */
  int bs;                          /* block size */
  int ncell;                       /* n / bs */
  const int *col_start;            /* cell index where each independent tridiagonal chain starts, length nchain+1 */
  int nchain;
  void (*residual)(void *ctx, const double *x, double *f);
  void (*jacobian)(void *ctx, const double *x, double *a, double *b, double *c);
  void *ctx;
} orc_system;

void orc_snes_solve(const orc_system *sys, const orc_snes_opts *o, double *x, orc_snes_result *res);
void orc_tridiag_solve(int n, const double *a, const double *b, const double *c, const double *d, double *x);
void orc_blocktridiag2_solve(int ncell, const double *a, const double *b, const double *c, const double *d, double *x);

/* ---- problem objects (opaque) ------------------------------------------------ */
typedef struct orc_vsfm orc_vsfm;
typedef struct orc_thermal orc_thermal;
typedef struct orc_th orc_th;

/* VSFM: mirrors mpp_vsfm_type / sysofeqns_vsfm_type for 1-D column batches.
 * All 1-D vectors are cell-ordered icell = c*nlev + j (layer fastest), as
 * MultiPhysicsProbVSFM.F90:364.  Soil tables are Fortran (ncol,nlev) column-major. */
orc_vsfm *orc_vsfm_create(int ncol, int nlev);
void      orc_vsfm_destroy(orc_vsfm *p);
/* per_column = 0: one SNES over all columns (what the reference does per MPI rank);
 * per_column = 1: an independent SNES / dt-cut loop per column (the GPU's contract) */
void      orc_vsfm_set_mode(orc_vsfm *p, int per_column, int nthreads);
int       orc_vsfm_set_mesh(orc_vsfm *p, int orientation, const double *dz /*(ncol,nlev) F-order*/,
                            const double *area /*ncol*/, const int *col_active /*ncol or NULL*/);
int       orc_vsfm_add_condition(orc_vsfm *p, int ss_or_bc, int cond_type, int region);
int       orc_vsfm_set_soils(orc_vsfm *p, const double *watsat, const double *hksat, const double *bsw,
                             const double *sucsat, const double *residual_sat, int satfunc_name, int density_type);
int       orc_vsfm_set_soils_direct(orc_vsfm *p, const double *por, const double *perm, const double *alpha,
                                    const double *lambda, const double *sat_res, int satfunc_name, int density_type);
void      orc_vsfm_set_tolerances(orc_vsfm *p, double atol, double rtol, double stol, int max_it, int max_funcs);
int       orc_vsfm_restart(orc_vsfm *p, const double *press /*ncells*/);
int       orc_vsfm_set_data(orc_vsfm *p, int auxvar_type, int var_type, int cond_id /*1-based*/, const double *data, int n);
int       orc_vsfm_get_data(orc_vsfm *p, int auxvar_type, int var_type, int cond_id, double *data, int n);
void      orc_vsfm_pre_step_dt(orc_vsfm *p);
void      orc_vsfm_post_step_dt(orc_vsfm *p);
int       orc_vsfm_step_dt(orc_vsfm *p, double dt, int nstep, int *converged, int *converged_reason);
/* diagnostics: per-column Newton its / reasons / dt cuts of the last StepDT (per_column mode),
 * or the single global values replicated */
void      orc_vsfm_get_stats(orc_vsfm *p, int *newton_its /*ncol*/, int *reasons /*ncol*/, int *ncuts /*ncol*/, int *nfuncs /*ncol*/);
/* raw residual / Jacobian at state x (for unit tests of the kernels' pieces) */
void      orc_vsfm_fill_mailbox(orc_vsfm *p);
/* MPPVSFMALM_Solve for a batch of independent columns (see vsfm.c); returns 0 or -(number of columns that failed all retries) */
int       orc_vsfm_elm_solve(orc_vsfm *p, double dtime_full, int nlevsoi, double watmin, const int *cond_ids,
                             int max_patch_per_col, const int *col_pfti, const int *col_npfts, const int *pft_active, const double *pft_wtcol,
                             const double *rootr_pft, const double *qflx_tran_veg_pft,
                             double *rootr_col, const double *qflx_tran_veg_col, const double *qflx_infl, const double *qflx_dew_snow,
                             const double *qflx_dew_grnd, const double *qflx_sub_snow, const double *frac_h2osfc, const int *snl,
                             double *qflx_drain, double *zwt, const double *zi, const double *dz, double *h2osoi_liq, double *h2osoi_ice,
                             double *mflx_snowlyr_col, const double *mflx_neg_snow, const double *mflx_drain_perched,
                             double *smp_l, double *soilp, double *qcharge, double *abs_mass_error, int *iter_count_out, int *status_out);
void      orc_vsfm_eval(orc_vsfm *p, double dt, const double *x_prev, const double *x, double *f, double *ja, double *jb, double *jc);

/* Thermal (KSP path) */
orc_thermal *orc_thermal_create(int ncol, int nlev);
void      orc_thermal_destroy(orc_thermal *p);
void      orc_thermal_set_threads(orc_thermal *p, int nthreads);
int       orc_thermal_set_mesh(orc_thermal *p, int orientation, const double *dz, const double *area,
                               const double *face_area /* internal conn area per column or NULL */);
int       orc_thermal_set_conn_dist(orc_thermal *p, const double *dist_up /*(ncol,nlev-1) F-order*/, const double *dist_dn);
int       orc_thermal_add_condition(orc_thermal *p, int ss_or_bc, int cond_type, int region);
int       orc_thermal_set_soils(orc_thermal *p, const double *watsat, const double *csol, const double *tkmg,
                                const double *tkdry, const int *lun_type /*ncol*/, int nlevsoi, int istsoil_id);
void      orc_thermal_set_cnfac(orc_thermal *p, double cnfac);
int       orc_thermal_set_soln_prev(orc_thermal *p, const double *T);
int       orc_thermal_set_rdata(orc_thermal *p, int auxvar_type, int var_type, int cond_id, const double *data, int n);
int       orc_thermal_set_idata(orc_thermal *p, int auxvar_type, int var_type, int cond_id, const int *data, int n);
int       orc_thermal_set_bdata(orc_thermal *p, int auxvar_type, int var_type, int cond_id, const int *data, int n);
void      orc_thermal_pre_step_dt(orc_thermal *p);
int       orc_thermal_step_dt(orc_thermal *p, double dt, int nstep, int *converged);
int       orc_thermal_get_soln(orc_thermal *p, double *T);
int       orc_thermal_get_aux(orc_thermal *p, int var_type, double *data);

/* snow + standing surface water + soil thermal system (ELM's real thermal column; parity unpinned, see thermal.c) */
typedef struct orc_thermal3 orc_thermal3;
orc_thermal3 *orc_thermal3_create(int ncol, int nlev, int nsno);
void      orc_thermal3_destroy(orc_thermal3 *p);
void      orc_thermal3_set_threads(orc_thermal3 *p, int nthreads);
void      orc_thermal3_set_cnfac(orc_thermal3 *p, double cnfac);
int       orc_thermal3_set_mesh(orc_thermal3 *p, const double *dz, const double *area, const double *conn_du, const double *conn_dd,
                                const double *soil_top_dist_dn, const double *snow_dz0);
int       orc_thermal3_set_soils(orc_thermal3 *p, const double *watsat, const double *csol, const double *tkmg,
                                 const double *tkdry, const int *lun_type, int nlevsoi, int istsoil_id);
int       orc_thermal3_set_soln_prev(orc_thermal3 *p, const double *T);
int       orc_thermal3_set_rdata(orc_thermal3 *p, int auxvar_type, int var_type, int cond_id, const double *data, int n);
int       orc_thermal3_set_idata(orc_thermal3 *p, int auxvar_type, int var_type, int cond_id, const int *data, int n);
void      orc_thermal3_pre_step_dt(orc_thermal3 *p);
int       orc_thermal3_step_dt(orc_thermal3 *p, double dt, int nstep, int *converged);
int       orc_thermal3_get_soln(orc_thermal3 *p, double *T);

/* TH (coupled Richards + enthalpy) */
orc_th   *orc_th_create(int ncol, int nlev);
void      orc_th_destroy(orc_th *p);
void      orc_th_set_mode(orc_th *p, int per_column, int nthreads);
int       orc_th_set_mesh(orc_th *p, int orientation, const double *dz, const double *area, const double *face_area);
int       orc_th_add_condition(orc_th *p, int ieqn /*1=mass,2=energy*/, int ss_or_bc, int cond_type, int region);
int       orc_th_set_soils(orc_th *p, const double *watsat, const double *hksat, const double *bsw, const double *sucsat,
                           const double *residual_sat, const double *csol /*J/kg/K*/, const double *tkdry,
                           int satfunc_name, int density_type, int int_energy_enthalpy_type);
int       orc_th_set_energy_perm(orc_th *p, const double *perm /*ncells*/);
void      orc_th_set_tolerances(orc_th *p, double atol, double rtol, double stol, int max_it, int max_funcs);
int       orc_th_restart(orc_th *p, const double *press, const double *temp);
int       orc_th_set_data(orc_th *p, int ieqn, int auxvar_type, int var_type, int cond_id, const double *data, int n);
int       orc_th_set_bc_pressure(orc_th *p, int ieqn, int cond_id, const double *data, int n);
int       orc_th_get_data(orc_th *p, int var_type, double *data, int n);
int       orc_th_step_dt(orc_th *p, double dt, int nstep, int *converged, int *converged_reason);
void      orc_th_get_stats(orc_th *p, int *newton_its, int *reasons, int *ncuts, int *nfuncs);
void      orc_th_eval(orc_th *p, double dt, const double *xprev /*2N: P..,T.. interleaved per cell*/, const double *x,
                      double *f, double *ja, double *jb, double *jc);

#ifdef __cplusplus
}
#endif
#endif
