/*
 * vsfm.c -- oracle restatement of the VSFM (Richards equation) system of equations
 * for batches of independent 1-D columns:
 *   src/mpp/ge/GoveqnRichardsODEPressureType.F90  residual / Jacobian / aux-var updates
 *   src/mpp/soe/SystemOfEquationsVSFMType.F90     SoE glue, PreSolve / PostSolve, Set/GetData
 *   src/mpp/soe/SystemOfEquationsBaseType.F90     StepDT_SNES (dt cuts)
 *   src/mpp/mpp/MultiPhysicsProbVSFM.F90          soil-parameter conversion, Restart
 *   src/mpp/dtypes/MeshType.F90                   column meshes + boundary connection sets
 * TEST INFRASTRUCTURE ONLY (see mpp_oracle.h).
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif
#include "mpp_oracle.h"

#define MAXCOND 16

typedef struct {
  int itype, region, nconn, per_cell;   /* per_cell: region SOIL_CELLS (one conn per cell) else one per column */
  orc_conn *conn;
  double *value;            /* cur_cond%value(:)                      */
  double *soe_value;        /* SoE mailbox aux_vars_{bc,ss}%condition_value */
  orc_rich_auxvar *aux;     /* GE aux_vars_bc / aux_vars_ss           */
  double *flux;             /* boundary_flux / ss_flux [kg/s]          */
  double *mass_exc;         /* bnd_mass_exc                            */
  double *pot_pressure, *pot_exponent;   /* aux_vars_ss%pot_mass_sink_{pressure,exponent} (down-regulated sinks) */
} vcond;

struct orc_vsfm {
  int ncol, nlev, ncells, orientation, per_column, nthreads;
  double *vol, *dz, *area_xy;
  int *is_active;
  orc_conn *conn_in;        /* (nlev-1) per column, iconn = c*(nlev-1)+j */
  int nbc, nss;
  vcond bc[MAXCOND], ss[MAXCOND];
  orc_rich_auxvar *aux_in;
  double *internal_flux;
  double dtime, time;
  double *accum_prev, *soln, *soln_prev, *soln_prev_clm;
  /* SoE mailbox (sysofeqns_vsfm_auxvar_type) for internal cells */
  double *soe_frac_liq_sat, *soe_temperature, *soe_liq_sat, *soe_pressure, *soe_mass, *soe_smp;
  orc_snes_opts opts;
  int *stat_its, *stat_reason, *stat_cuts, *stat_nf;
  int *col_start_all;       /* 0, nlev, 2 nlev, ... */
};

orc_vsfm *orc_vsfm_create(int ncol, int nlev)
{
  orc_vsfm *p = (orc_vsfm *)calloc(1, sizeof(*p));
  int i, n = ncol * nlev;
  p->ncol = ncol; p->nlev = nlev; p->ncells = n; p->orientation = MESH_ALONG_GRAVITY;
  p->per_column = 0; p->nthreads = 1;
  p->vol = (double *)calloc(n, sizeof(double)); p->dz = (double *)calloc(n, sizeof(double));
  p->area_xy = (double *)calloc(n, sizeof(double)); p->is_active = (int *)calloc(n, sizeof(int));
  p->conn_in = (orc_conn *)calloc((size_t)ncol * (nlev > 1 ? nlev - 1 : 1), sizeof(orc_conn));
  p->aux_in = (orc_rich_auxvar *)calloc(n, sizeof(orc_rich_auxvar));
  for (i = 0; i < n; i++) orc_rich_auxvar_init(&p->aux_in[i]);
  p->internal_flux = (double *)calloc((size_t)ncol * (nlev > 1 ? nlev - 1 : 1), sizeof(double));
  p->accum_prev = (double *)calloc(n, sizeof(double)); p->soln = (double *)calloc(n, sizeof(double));
  p->soln_prev = (double *)calloc(n, sizeof(double)); p->soln_prev_clm = (double *)calloc(n, sizeof(double));
  p->soe_frac_liq_sat = (double *)malloc(sizeof(double) * n); p->soe_temperature = (double *)malloc(sizeof(double) * n);
  for (i = 0; i < n; i++) { p->soe_frac_liq_sat[i] = 1.0; p->soe_temperature[i] = 273.15 + 25.0; }  /* SystemOfEquationsVSFMAuxType.F90:64-66 */
  p->soe_liq_sat = (double *)calloc(n, sizeof(double)); p->soe_pressure = (double *)calloc(n, sizeof(double));
  p->soe_mass = (double *)calloc(n, sizeof(double)); p->soe_smp = (double *)calloc(n, sizeof(double));
  orc_snes_default_opts(&p->opts);
  p->stat_its = (int *)calloc(ncol, sizeof(int)); p->stat_reason = (int *)calloc(ncol, sizeof(int));
  p->stat_cuts = (int *)calloc(ncol, sizeof(int)); p->stat_nf = (int *)calloc(ncol, sizeof(int));
  p->col_start_all = (int *)malloc(sizeof(int) * (ncol + 1));
  for (i = 0; i <= ncol; i++) p->col_start_all[i] = i * nlev;
  return p;
}

static void vcond_free(vcond *c) { free(c->conn); free(c->value); free(c->soe_value); free(c->aux); free(c->flux); free(c->mass_exc); free(c->pot_pressure); free(c->pot_exponent); }

void orc_vsfm_destroy(orc_vsfm *p)
{
  int i;
  if (!p) return;
  for (i = 0; i < p->nbc; i++) vcond_free(&p->bc[i]);
  for (i = 0; i < p->nss; i++) vcond_free(&p->ss[i]);
  free(p->vol); free(p->dz); free(p->area_xy); free(p->is_active); free(p->conn_in); free(p->aux_in); free(p->internal_flux);
  free(p->accum_prev); free(p->soln); free(p->soln_prev); free(p->soln_prev_clm);
  free(p->soe_frac_liq_sat); free(p->soe_temperature); free(p->soe_liq_sat); free(p->soe_pressure); free(p->soe_mass); free(p->soe_smp);
  free(p->stat_its); free(p->stat_reason); free(p->stat_cuts); free(p->stat_nf); free(p->col_start_all);
  free(p);
}

void orc_vsfm_set_mode(orc_vsfm *p, int per_column, int nthreads) { p->per_column = per_column; p->nthreads = nthreads > 0 ? nthreads : 1; }

void orc_vsfm_set_tolerances(orc_vsfm *p, double atol, double rtol, double stol, int max_it, int max_funcs)
{ p->opts.atol = atol; p->opts.rtol = rtol; p->opts.stol = stol; p->opts.max_it = max_it; p->opts.max_funcs = max_funcs; }

/*
 * Mesh.  ALONG_GRAVITY: MeshType.F90:401-428 (cells top-down, vol = area*dz) and :509-530
 * (vertical connections up=j, dn=j+1, unit vector (0,0,-1), dist = dz/2 each).
 * AGAINST_GRAVITY: MeshCreate1 MeshType.F90:173-269 + mpp_mesh_utils.F90:363-470 (cell 1 at the
 * bottom; unit vector from centroids = (0,0,+1), MeshType.F90:933-938).
 * HORIZONTAL: same with CONN_IN_X_DIR (unit vector (1,0,0), no gravity component).
 * dz is Fortran (ncol,nlev) column-major: dz[j*ncol + c].
 */
int orc_vsfm_set_mesh(orc_vsfm *p, int orientation, const double *dz, const double *area, const int *col_active)
{
  int c, j, ncol = p->ncol, nlev = p->nlev;
  p->orientation = orientation;
  for (c = 0; c < ncol; c++) {
    for (j = 0; j < nlev; j++) {
      int ic = c * nlev + j;
      p->dz[ic] = dz[(size_t)j * ncol + c];
      p->area_xy[ic] = area[c];
      p->vol[ic] = area[c] * p->dz[ic];
      p->is_active[ic] = col_active ? (col_active[c] != 0) : 1;
    }
    for (j = 0; j < nlev - 1; j++) {
      orc_conn *cn = &p->conn_in[c * (nlev - 1) + j];
      cn->id_up = c * nlev + j; cn->id_dn = cn->id_up + 1;
      cn->area = area[c];
      cn->dist_up = 0.5 * p->dz[cn->id_up]; cn->dist_dn = 0.5 * p->dz[cn->id_dn];
      cn->unitvec[0] = 0.0; cn->unitvec[1] = 0.0; cn->unitvec[2] = 0.0;
      if (orientation == MESH_ALONG_GRAVITY)        cn->unitvec[2] = -1.0;
      else if (orientation == MESH_AGAINST_GRAVITY) cn->unitvec[2] = 1.0;
      else                                          cn->unitvec[0] = 1.0;
    }
  }
  return 0;
}

/* AddConditionInGovEqn -> MeshCreateConnectionSet1 (MeshType.F90:648-851).  Returns the 1-based condition id
 * (separate counters for BCs and SSs, like soe_auxvar_id). */
int orc_vsfm_add_condition(orc_vsfm *p, int ss_or_bc, int cond_type, int region)
{
  int c, j, ncol = p->ncol, nlev = p->nlev, i, n;
  vcond *cd;
  if (ss_or_bc == COND_BC) { if (p->nbc >= MAXCOND) return -1; cd = &p->bc[p->nbc++]; }
  else                     { if (p->nss >= MAXCOND) return -1; cd = &p->ss[p->nss++]; }
  memset(cd, 0, sizeof(*cd));
  cd->itype = cond_type; cd->region = region;
  cd->per_cell = (region == SOIL_CELLS);
  n = cd->per_cell ? ncol * nlev : ncol;
  cd->nconn = n;
  cd->conn = (orc_conn *)calloc(n, sizeof(orc_conn));
  cd->value = (double *)calloc(n, sizeof(double)); cd->soe_value = (double *)calloc(n, sizeof(double));
  cd->pot_pressure = (double *)calloc(n, sizeof(double)); cd->pot_exponent = (double *)calloc(n, sizeof(double));
  /* SS aux vars are never computed (RichardsODEPressureUpdateAuxVarsSS returns immediately, GoveqnRichards...:1578),
   * so only BCs carry GE aux vars here */
  cd->aux = ss_or_bc == COND_BC ? (orc_rich_auxvar *)calloc(n, sizeof(orc_rich_auxvar)) : NULL;
  cd->flux = (double *)calloc(n, sizeof(double)); cd->mass_exc = (double *)calloc(n, sizeof(double));
  if (cd->aux) for (i = 0; i < n; i++) orc_rich_auxvar_init(&cd->aux[i]);
  if (cd->per_cell) {
    for (c = 0; c < ncol; c++) for (j = 0; j < nlev; j++) {     /* MeshType.F90:808-838 */
      orc_conn *cn = &cd->conn[c * nlev + j];
      cn->id_up = -1; cn->id_dn = c * nlev + j; cn->area = p->area_xy[cn->id_dn];
      cn->dist_up = 0.0; cn->dist_dn = 0.0;
    }
  } else {
    for (c = 0; c < ncol; c++) {                                 /* MeshType.F90:723-806 */
      orc_conn *cn = &cd->conn[c];
      int first = c * nlev, last = c * nlev + nlev - 1, top_is_first = (p->orientation != MESH_AGAINST_GRAVITY);
      cn->id_up = -1;
      if (region == SOIL_TOP_CELLS) {
        cn->id_dn = top_is_first ? first : last;
        if (p->orientation == MESH_HORIZONTAL) cn->unitvec[0] = 1.0; else cn->unitvec[2] = -1.0;
      } else { /* SOIL_BOTTOM_CELLS */
        cn->id_dn = top_is_first ? last : first;
        if (p->orientation == MESH_HORIZONTAL) cn->unitvec[0] = -1.0; else cn->unitvec[2] = 1.0;
      }
      cn->area = p->area_xy[cn->id_dn];
      cn->dist_up = 0.0;
      cn->dist_dn = 0.5 * p->dz[cn->id_dn];
    }
  }
  return ss_or_bc == COND_BC ? p->nbc : p->nss;
}

static int set_satfunc(orc_satparams *sp, int satfunc_name, double sat_res, double alpha, double lambda)
{
  /* MultiPhysicsProbVSFM.F90:391-420 */
  switch (satfunc_name) {
  case SATFUNC_NAME_BROOKS_COREY:  return orc_satfunc_set_bc(sp, sat_res, alpha, lambda);
  case SATFUNC_NAME_SBC_BZ2:       return orc_satfunc_set_sbc_bz2(sp, sat_res, alpha, lambda, -0.9 / alpha);
  case SATFUNC_NAME_SBC_BZ3:       return orc_satfunc_set_sbc_bz3(sp, sat_res, alpha, lambda, -0.9 / alpha);
  case SATFUNC_NAME_VAN_GENUCHTEN: return orc_satfunc_set_vg(sp, sat_res, alpha, lambda);
  }
  return 9;
}

static void copy_soil_to_cond_aux(orc_vsfm *p, int density_type)
{
  /* MultiPhysicsProbVSFM.F90:424-470: BC/SS aux vars take perm, por, satParams, porParams of the cell they touch;
   * GoveqnRichards...:217-277: density type goes to every aux var */
  int k, i;
  for (k = 0; k < p->nbc; k++) {
    vcond *cd = &p->bc[k];
    for (i = 0; i < cd->nconn; i++) {
      const orc_rich_auxvar *src = &p->aux_in[cd->conn[i].id_dn];
      orc_rich_auxvar *a = &cd->aux[i];
      a->perm[0] = src->perm[0]; a->perm[1] = src->perm[1]; a->perm[2] = src->perm[2];
      a->por = src->por; a->por_base = src->por_base; a->satParams = src->satParams;
      a->density_type = density_type;
    }
  }
}

/* VSFMMPPSetSoilsCLM, MultiPhysicsProbVSFM.F90:249-475.  Tables are (ncol,nlev) Fortran order. */
int orc_vsfm_set_soils(orc_vsfm *p, const double *watsat, const double *hksat, const double *bsw,
                       const double *sucsat, const double *residual_sat, int satfunc_name, int density_type)
{
  const double vish2o = 0.001002;
  int c, j, ncol = p->ncol, nlev = p->nlev, rc = 0;
  for (c = 0; c < ncol; c++) for (j = 0; j < nlev; j++) {
    size_t t = (size_t)j * ncol + c;
    orc_rich_auxvar *a = &p->aux_in[c * nlev + j];
    double perm   = hksat[t] * vish2o / (ORC_DENH2O * ORC_GRAV) * 0.001;    /* :374 */
    double alpha  = 1.0 / (sucsat[t] * ORC_GRAV);                            /* :378 */
    double lambda = 1.0 / bsw[t];                                            /* :381 */
    double sat_res = residual_sat[t], por = watsat[t];
    a->perm[0] = a->perm[1] = a->perm[2] = perm;
    a->por = por; a->por_base = por;
    a->density_type = density_type;
    rc |= set_satfunc(&a->satParams, satfunc_name, sat_res, alpha, lambda);
  }
  copy_soil_to_cond_aux(p, density_type);
  return rc;
}

/* same, with already-converted parameters in cell order (used by unit tests of the kernels' pieces) */
int orc_vsfm_set_soils_direct(orc_vsfm *p, const double *por, const double *perm, const double *alpha,
                              const double *lambda, const double *sat_res, int satfunc_name, int density_type)
{
  int i, rc = 0;
  for (i = 0; i < p->ncells; i++) {
    orc_rich_auxvar *a = &p->aux_in[i];
    a->perm[0] = a->perm[1] = a->perm[2] = perm[i];
    a->por = por[i]; a->por_base = por[i]; a->density_type = density_type;
    rc |= set_satfunc(&a->satParams, satfunc_name, sat_res[i], alpha[i], lambda[i]);
  }
  copy_soil_to_cond_aux(p, density_type);
  return rc;
}

/* VSFMMPPRestart, MultiPhysicsProbVSFM.F90:603-707 */
int orc_vsfm_restart(orc_vsfm *p, const double *press)
{
  size_t nb = sizeof(double) * (size_t)p->ncells;
  memcpy(p->soln, press, nb); memcpy(p->soln_prev, press, nb); memcpy(p->soln_prev_clm, press, nb);
  return 0;
}

static void update_auxvars(orc_vsfm *p, int c0, int c1, const double *X);

/* Not in VSFMMPPRestart: ELM's initialisation leaves the SoE mailbox (mass, saturation, matric potential, pressure) consistent
 * with the initial pressures before the first MPPVSFMALM_Solve reads VAR_MASS (:425).  Mirrors mppgpu_restart, which fills it. */
void orc_vsfm_fill_mailbox(orc_vsfm *p)
{
  int ic;
  update_auxvars(p, 0, p->ncol, p->soln);
  for (ic = 0; ic < p->ncells; ic++) if (p->is_active[ic]) {
    const orc_rich_auxvar *a = &p->aux_in[ic];
    p->soe_liq_sat[ic] = a->sat; p->soe_pressure[ic] = a->pressure;
    p->soe_mass[ic] = a->por * a->den * ORC_FMWH2O * a->sat * p->vol[ic];
    p->soe_smp[ic] = (a->pressure - ORC_PRESSURE_REF) / (a->den * ORC_FMWH2O * ORC_GRAVITY_CONSTANT);
  }
}

/* VSFMSOESetDataFromCLM, SystemOfEquationsVSFMType.F90:663-724 */
int orc_vsfm_set_data(orc_vsfm *p, int auxvar_type, int var_type, int cond_id, const double *data, int n)
{
  int i;
  if (auxvar_type == AUXVAR_INTERNAL) {
    double *dst = NULL;
    if (n > p->ncells) return 1;
    if (var_type == VAR_FRAC_LIQ_SAT)      dst = p->soe_frac_liq_sat;
    else if (var_type == VAR_TEMPERATURE)  dst = p->soe_temperature;
    else if (var_type == VAR_PRESSURE)     dst = p->soe_pressure;
    else return 2;
    for (i = 0; i < n; i++) dst[i] = data[i];
    return 0;
  } else {
    vcond *cd;
    if (auxvar_type == AUXVAR_BC) { if (cond_id < 1 || cond_id > p->nbc) return 3; cd = &p->bc[cond_id - 1]; }
    else if (auxvar_type == AUXVAR_SS) { if (cond_id < 1 || cond_id > p->nss) return 3; cd = &p->ss[cond_id - 1]; }
    else return 4;
    if (n > cd->nconn) return 1;
    /* VSFMMPPSetSourceSinkAuxVarRealValue (MultiPhysicsProbVSFM.F90:1437-1520): straight into the GE's aux_vars_ss */
    if (auxvar_type == AUXVAR_SS && var_type == VAR_POT_MASS_SINK_PRESSURE) { for (i = 0; i < n; i++) cd->pot_pressure[i] = data[i]; return 0; }
    if (auxvar_type == AUXVAR_SS && var_type == VAR_POT_MASS_SINK_EXPONENT) { for (i = 0; i < n; i++) cd->pot_exponent[i] = data[i]; return 0; }
    if (var_type != VAR_BC_SS_CONDITION) return 2;
    for (i = 0; i < n; i++) cd->soe_value[i] = data[i];
    return 0;
  }
}

/* VSFMSOEGetDataForCLM, SystemOfEquationsVSFMType.F90:781-845 */
int orc_vsfm_get_data(orc_vsfm *p, int auxvar_type, int var_type, int cond_id, double *data, int n)
{
  int i;
  if (auxvar_type == 704 /* AUXVAR_CONN_INTERNAL, SystemOfEquationsVSFMType.F90:824 */) {
    int nconn = p->ncol * (p->nlev > 1 ? p->nlev - 1 : 0);
    if (var_type != VAR_MASS_FLUX) return 2;
    if (n > nconn) return 1;
    for (i = 0; i < n; i++) data[i] = p->internal_flux[i];     /* RichardsODEPressureSetDataInSOEAuxVar :1199-1222 */
    return 0;
  }
  if (auxvar_type == AUXVAR_INTERNAL) {
    const double *src = NULL;
    if (n > p->ncells) return 1;
    switch (var_type) {
    case VAR_PRESSURE: src = p->soe_pressure; break;
    case VAR_LIQ_SAT: src = p->soe_liq_sat; break;
    case VAR_MASS: src = p->soe_mass; break;
    case VAR_SOIL_MATRIX_POT: src = p->soe_smp; break;
    case VAR_FRAC_LIQ_SAT: src = p->soe_frac_liq_sat; break;
    case VAR_TEMPERATURE: src = p->soe_temperature; break;
    default: return 2;
    }
    for (i = 0; i < n; i++) data[i] = src[i];
    return 0;
  } else {
    vcond *cd;
    if (auxvar_type == AUXVAR_BC) { if (cond_id < 1 || cond_id > p->nbc) return 3; cd = &p->bc[cond_id - 1]; }
    else if (auxvar_type == AUXVAR_SS) { if (cond_id < 1 || cond_id > p->nss) return 3; cd = &p->ss[cond_id - 1]; }
    else return 4;
    if (n > cd->nconn) return 1;
    if (var_type == VAR_BC_SS_CONDITION) for (i = 0; i < n; i++) data[i] = cd->soe_value[i];
    else if (var_type == VAR_MASS_FLUX)  for (i = 0; i < n; i++) data[i] = cd->flux[i];
    else if (var_type == 614 /*VAR_BC_MASS_EXCHANGED*/) for (i = 0; i < n; i++) data[i] = cd->mass_exc[i];
    else return 2;
    return 0;
  }
}

/* ------------------------------------------------------------------------------------------------
 * Residual / Jacobian over a column range [c0,c1).  X is the full-length solution vector; only the
 * entries of the range are read / written.
 * ------------------------------------------------------------------------------------------------ */

/* VSFMSOEResidual steps 1-2 (SystemOfEquationsVSFMType.F90:127,163-164):
 * RichardsODESavePrmIndepVar (:504-534), UpdateAuxVarsIntrn (:1417-1474), UpdateAuxVarsBC (:1478-1552) */
static void update_auxvars(orc_vsfm *p, int c0, int c1, const double *X)
{
  int c, j, k, nlev = p->nlev;
  for (c = c0; c < c1; c++) for (j = 0; j < nlev; j++) {
    int ic = c * nlev + j;
    p->aux_in[ic].pressure = X[ic];
    if (p->is_active[ic]) orc_rich_auxvar_compute(&p->aux_in[ic]);
  }
  for (k = 0; k < p->nbc; k++) {
    vcond *cd = &p->bc[k];
    int i0 = cd->per_cell ? c0 * nlev : c0, i1 = cd->per_cell ? c1 * nlev : c1, i;
    for (i = i0; i < i1; i++) {
      switch (cd->itype) {
      case COND_DIRICHLET: case COND_SEEPAGE_BC:
        cd->aux[i].pressure = cd->aux[i].condition_value; break;
      case COND_MASS_RATE: case COND_MASS_FLUX:
        cd->aux[i].pressure = p->aux_in[cd->conn[i].id_dn].pressure; break;
      default: break;
      }
      orc_rich_auxvar_compute(&cd->aux[i]);
    }
  }
  /* UpdateAuxVarsSS is a no-op in the reference (early `return`, GoveqnRichards...:1578) */
}

/* RichardsODEPressureAccum, GoveqnRichards...:1603-1634 */
static void accum(orc_vsfm *p, int c0, int c1, double *ff)
{
  int ic, nlev = p->nlev;
  double dtInv = 1.0 / p->dtime;
  for (ic = c0 * nlev; ic < c1 * nlev; ic++) {
    ff[ic] = 0.0;
    if (p->is_active[ic])
      ff[ic] = p->aux_in[ic].por * p->aux_in[ic].den * p->aux_in[ic].sat * p->vol[ic] * dtInv;
  }
}

/* RichardsODEPressureDivergence, GoveqnRichards...:1696-1938 */
static void divergence(orc_vsfm *p, int c0, int c1, double *ff)
{
  int c, j, k, nlev = p->nlev;
  double flux, d1, d2;
  for (c = c0; c < c1; c++) for (j = 0; j < nlev - 1; j++) {
    int iconn = c * (nlev - 1) + j;
    const orc_conn *cn = &p->conn_in[iconn];
    if (!p->is_active[cn->id_up] || !p->is_active[cn->id_dn]) continue;
    orc_richards_flux(&p->aux_in[cn->id_up], &p->aux_in[cn->id_dn], cn, 0, 1, 0, COND_NULL, &flux, &d1, &d2);
    ff[cn->id_up] = ff[cn->id_up] - flux;
    ff[cn->id_dn] = ff[cn->id_dn] + flux;
    p->internal_flux[iconn] = flux * ORC_FMWH2O;
  }
  for (k = 0; k < p->nbc; k++) {
    vcond *cd = &p->bc[k];
    int i0 = cd->per_cell ? c0 * nlev : c0, i1 = cd->per_cell ? c1 * nlev : c1, i;
    for (i = i0; i < i1; i++) {
      int cell = cd->conn[i].id_dn;
      if (!p->is_active[cell]) continue;
      orc_richards_flux(&cd->aux[i], &p->aux_in[cell], &cd->conn[i], 0, 0, 0, cd->itype, &flux, &d1, &d2);
      ff[cell] = ff[cell] + flux;
      cd->flux[i] = flux * ORC_FMWH2O;
    }
  }
  for (k = 0; k < p->nss; k++) {
    vcond *cd = &p->ss[k];
    int i0 = cd->per_cell ? c0 * nlev : c0, i1 = cd->per_cell ? c1 * nlev : c1, i;
    for (i = i0; i < i1; i++) {
      int cell = cd->conn[i].id_dn;
      if (!p->is_active[cell]) continue;          /* active_conn_ids */
      if (cd->itype == COND_MASS_RATE) {
        ff[cell] = ff[cell] - cd->value[i] / ORC_FMWH2O;
        cd->flux[i] = cd->value[i];
      } else if (cd->itype == COND_DOWNREG_MASS_RATE_CAMPBELL) {       /* GoveqnRichards...:1900-1913 */
        double dP = p->aux_in[cell].pressure - ORC_PRESSURE_REF, Pc = cd->pot_pressure[i], n = cd->pot_exponent[i], factor;
        if (dP <= 0.0) factor = 1.0 + pow(dP / Pc, n); else factor = 1.0;
        ff[cell] = ff[cell] - cd->value[i] / factor / ORC_FMWH2O;
        cd->flux[i] = cd->value[i] / factor;
      } else if (cd->itype == COND_DOWNREG_MASS_RATE_FETCH2) {         /* :1915-1928 */
        double dP = p->aux_in[cell].pressure - ORC_PRESSURE_REF, Pc = cd->pot_pressure[i], n = cd->pot_exponent[i], factor;
        if (dP <= 0.0) factor = exp(-pow(dP / Pc, n)); else factor = 1.0;
        ff[cell] = ff[cell] - cd->value[i] * factor / ORC_FMWH2O;
        cd->flux[i] = cd->value[i] * factor;
      }
    }
  }
}

/* RichardsODEComputeResidual, GoveqnRichards...:388-421 */
static void residual_range(orc_vsfm *p, int c0, int c1, const double *X, double *F)
{
  int ic, nlev = p->nlev;
  update_auxvars(p, c0, c1, X);
  accum(p, c0, c1, F);
  for (ic = c0 * nlev; ic < c1 * nlev; ic++) F[ic] = F[ic] - p->accum_prev[ic];
  divergence(p, c0, c1, F);
}

/* RichardsODEComputeJacobian (:425-453) = DivergenceDeriv (:1941-2200) + AccumDeriv (:1638-1693).
 * Uses the aux vars left by the last residual evaluation (VSFMJacobian does not recompute them). */
static void jacobian_range(orc_vsfm *p, int c0, int c1, double *ja, double *jb, double *jc)
{
  int c, j, k, ic, nlev = p->nlev;
  double dummy, Jup, Jdn, dtInv = 1.0 / p->dtime;
  for (ic = c0 * nlev; ic < c1 * nlev; ic++) { ja[ic] = 0.0; jb[ic] = 0.0; jc[ic] = 0.0; }   /* MatZeroEntries */
  for (c = c0; c < c1; c++) for (j = 0; j < nlev - 1; j++) {
    const orc_conn *cn = &p->conn_in[c * (nlev - 1) + j];
    if (!p->is_active[cn->id_up] || !p->is_active[cn->id_dn]) continue;
    orc_richards_flux(&p->aux_in[cn->id_up], &p->aux_in[cn->id_dn], cn, 1, 1, 0, COND_NULL, &dummy, &Jup, &Jdn);
    jb[cn->id_up] += Jup;      /* (up,up) */
    jc[cn->id_up] += Jdn;      /* (up,dn) */
    ja[cn->id_dn] += -Jup;     /* (dn,up) */
    jb[cn->id_dn] += -Jdn;     /* (dn,dn) */
  }
  for (k = 0; k < p->nbc; k++) {
    vcond *cd = &p->bc[k];
    int i0 = cd->per_cell ? c0 * nlev : c0, i1 = cd->per_cell ? c1 * nlev : c1, i;
    for (i = i0; i < i1; i++) {
      int cell = cd->conn[i].id_dn;
      if (!p->is_active[cell]) continue;
      orc_richards_flux(&cd->aux[i], &p->aux_in[cell], &cd->conn[i], 1, 0, 0, cd->itype, &dummy, &Jup, &Jdn);
      jb[cell] += -Jdn;
    }
  }
  /* COND_MASS_RATE source/sinks have no Jacobian contribution (:2150-2151); down-regulated sinks add a diagonal term (:2158-2188) */
  for (k = 0; k < p->nss; k++) {
    vcond *cd = &p->ss[k];
    int i0 = cd->per_cell ? c0 * nlev : c0, i1 = cd->per_cell ? c1 * nlev : c1, i;
    if (cd->itype != COND_DOWNREG_MASS_RATE_CAMPBELL && cd->itype != COND_DOWNREG_MASS_RATE_FETCH2) continue;
    for (i = i0; i < i1; i++) {
      int cell = cd->conn[i].id_dn;
      double dP, Pc, n, factor, val;
      if (!p->is_active[cell]) continue;
      dP = p->aux_in[cell].pressure - ORC_PRESSURE_REF; Pc = cd->pot_pressure[i]; n = cd->pot_exponent[i];
      if (dP > 0.0) continue;
      if (cd->itype == COND_DOWNREG_MASS_RATE_CAMPBELL) {
        factor = 1.0 + pow(dP / Pc, n);
        val = (cd->value[i] / ORC_FMWH2O) * (n * pow(dP / Pc, n)) / (dP * pow(factor, 2.0));
      } else {
        factor = exp(-pow(dP / Pc, n));
        val = (cd->value[i] / ORC_FMWH2O) * (n * pow(dP / Pc, n)) * factor / dP;
      }
      jb[cell] += val;
    }
  }
  for (ic = c0 * nlev; ic < c1 * nlev; ic++) {
    double derivative;
    if (p->is_active[ic]) {
      const orc_rich_auxvar *a = &p->aux_in[ic];
      derivative = (a->dpor_dP * a->den * a->sat + a->por * a->dden_dP * a->sat + a->por * a->den * a->dsat_dP) * p->vol[ic] * dtInv;
    } else derivative = 1.0;
    jb[ic] += derivative;
  }
}

typedef struct { orc_vsfm *p; int c0, c1; double *Xfull, *Ffull, *ja, *jb, *jc; } rctx;

static void cb_residual(void *vctx, const double *x, double *f)
{
  rctx *r = (rctx *)vctx; orc_vsfm *p = r->p;
  int off = r->c0 * p->nlev, n = (r->c1 - r->c0) * p->nlev;
  memcpy(r->Xfull + off, x, sizeof(double) * (size_t)n);
  residual_range(p, r->c0, r->c1, r->Xfull, r->Ffull);
  memcpy(f, r->Ffull + off, sizeof(double) * (size_t)n);
}
static void cb_jacobian(void *vctx, const double *x, double *a, double *b, double *c)
{
  rctx *r = (rctx *)vctx; orc_vsfm *p = r->p;
  int off = r->c0 * p->nlev, n = (r->c1 - r->c0) * p->nlev;
  (void)x;
  jacobian_range(p, r->c0, r->c1, r->ja, r->jb, r->jc);
  memcpy(a, r->ja + off, sizeof(double) * (size_t)n); memcpy(b, r->jb + off, sizeof(double) * (size_t)n);
  memcpy(c, r->jc + off, sizeof(double) * (size_t)n);
}

/* VSFMSOEPreSolve (SystemOfEquationsVSFMType.F90:506-563) + RichardsODEPressurePreSolve (GoveqnRichards...:2747-2768) */
static void pre_solve(orc_vsfm *p, int c0, int c1)
{
  int k, ic, nlev = p->nlev;
  /* GetFromSOEAuxVarsIntrn (:536-583): only frac_liq_sat is copied (temperature deliberately not) */
  for (ic = c0 * nlev; ic < c1 * nlev; ic++) if (p->is_active[ic]) p->aux_in[ic].frac_liq_sat = p->soe_frac_liq_sat[ic];
  /* GetFromSOEAuxVarsBC (:656-763), GetFromSOEAuxVarsSS (:851-945) */
  for (k = 0; k < p->nbc; k++) {
    vcond *cd = &p->bc[k];
    int i0 = cd->per_cell ? c0 * nlev : c0, i1 = cd->per_cell ? c1 * nlev : c1, i;
    for (i = i0; i < i1; i++) cd->aux[i].condition_value = cd->soe_value[i];
  }
  for (k = 0; k < p->nss; k++) {
    vcond *cd = &p->ss[k];
    int i0 = cd->per_cell ? c0 * nlev : c0, i1 = cd->per_cell ? c1 * nlev : c1, i;
    for (i = i0; i < i1; i++) {
      if (!p->is_active[cd->conn[i].id_dn]) continue;
      cd->value[i] = cd->soe_value[i];
    }
  }
  update_auxvars(p, c0, c1, p->soln_prev);   /* SavePrimaryIndependentVar(soln_prev) + UpdateAuxVars */
  accum(p, c0, c1, p->accum_prev);
}

/* VSFMSOEPostSolve (SystemOfEquationsVSFMType.F90:566-660) -> SetDataInSOEAuxVar (GoveqnRichards...:1071-1395) */
static void post_solve(orc_vsfm *p, int c0, int c1)
{
  int k, ic, nlev = p->nlev;
  for (ic = c0 * nlev; ic < c1 * nlev; ic++) {
    p->soln_prev[ic] = p->soln[ic];
    if (p->is_active[ic]) {
      const orc_rich_auxvar *a = &p->aux_in[ic];
      double mass, Pa_to_Meters;
      p->soe_liq_sat[ic]  = a->sat;
      p->soe_pressure[ic] = a->pressure;
      mass = a->por * a->den * ORC_FMWH2O * a->sat * p->vol[ic];           /* :1178-1184 */
      p->soe_mass[ic] = mass;
      Pa_to_Meters = a->den * ORC_FMWH2O * ORC_GRAVITY_CONSTANT;            /* :1188-1192 */
      p->soe_smp[ic] = (a->pressure - ORC_PRESSURE_REF) / Pa_to_Meters;
    }
  }
  for (k = 0; k < p->nbc; k++) {                                            /* :1318-1352 */
    vcond *cd = &p->bc[k];
    int i0 = cd->per_cell ? c0 * nlev : c0, i1 = cd->per_cell ? c1 * nlev : c1, i;
    for (i = i0; i < i1; i++) cd->mass_exc[i] = cd->mass_exc[i] + cd->flux[i] * p->dtime;
  }
}

/* SOEBaseStepDT_SNES, SystemOfEquationsBaseType.F90:368-552, for the column range [c0,c1).
 * (use_dynamic_linesearch is off by default: mpp_varctl.F90) */
static void step_dt_range(orc_vsfm *p, int c0, int c1, double dt, int *converged, int *converged_reason,
                          double *Ffull, double *ja, double *jb, double *jc, int *tot_its, int *tot_nf, int *ncuts)
{
  const int max_num_time_cuts = 20;
  int num_time_cuts = 0, nlev = p->nlev, off = c0 * nlev, n = (c1 - c0) * nlev;
  double time = 0.0, target_time = dt, dt_iter = dt;
  orc_system sys; rctx r; orc_snes_result res;
  double *x = (double *)malloc(sizeof(double) * (size_t)n);

  r.p = p; r.c0 = c0; r.c1 = c1; r.Xfull = p->soln; r.Ffull = Ffull; r.ja = ja; r.jb = jb; r.jc = jc;
  sys.n = n; sys.bs = 1; sys.ncell = n; sys.col_start = NULL; sys.nchain = c1 - c0;
  {
    /* chains relative to the sub-vector */
    int *cs = (int *)malloc(sizeof(int) * (size_t)(c1 - c0 + 1)), i;
    for (i = 0; i <= c1 - c0; i++) cs[i] = i * nlev;
    sys.col_start = cs;
  }
  sys.residual = cb_residual; sys.jacobian = cb_jacobian; sys.ctx = &r;
  *converged = 0; *tot_its = 0; *tot_nf = 0;

  for (;;) {
    p->dtime = dt_iter;                       /* SetDtime */
    pre_solve(p, c0, c1);
    memcpy(x, p->soln + off, sizeof(double) * (size_t)n);
    orc_snes_solve(&sys, &p->opts, x, &res);
    memcpy(p->soln + off, x, sizeof(double) * (size_t)n);
    *converged_reason = res.reason;
    *tot_nf += res.nfuncs;
    if (res.reason < 0) {
      num_time_cuts++;
      dt_iter = 0.5 * dt_iter;
      memcpy(p->soln + off, p->soln_prev + off, sizeof(double) * (size_t)n);
    } else {
      *converged = 1;
      time += dt_iter;
      *tot_its += res.its;
      post_solve(p, c0, c1);
    }
    if (num_time_cuts > max_num_time_cuts) { *converged = 0; break; }
    if (time >= target_time) break;
  }
  *ncuts = num_time_cuts;
  p->time = time;                             /* soe%time: what the failed / finished StepDT did advance (:511) */
  free((void *)sys.col_start); free(x);
}

void orc_vsfm_pre_step_dt(orc_vsfm *p)
{
  /* VSFMSPreStepDT, SystemOfEquationsVSFMType.F90:892-923 (+ GoveqnRichards...:2772-2785) */
  int k, i;
  size_t nb = sizeof(double) * (size_t)p->ncells;
  memcpy(p->soln_prev, p->soln_prev_clm, nb); memcpy(p->soln, p->soln_prev_clm, nb);
  for (k = 0; k < p->nbc; k++) for (i = 0; i < p->bc[k].nconn; i++) p->bc[k].mass_exc[i] = 0.0;
}

void orc_vsfm_post_step_dt(orc_vsfm *p)
{
  /* VSFMSPostStepDT, SystemOfEquationsVSFMType.F90:926-940 */
  memcpy(p->soln_prev_clm, p->soln_prev, sizeof(double) * (size_t)p->ncells);
}

/* sysofeqns%StepDT -> SOEBaseStepDT -> SOEBaseStepDT_SNES */
int orc_vsfm_step_dt(orc_vsfm *p, double dt, int nstep, int *converged, int *converged_reason)
{
  int ncol = p->ncol, n = p->ncells;
  (void)nstep;
  if (!p->per_column) {
    double *w = (double *)malloc(sizeof(double) * (size_t)n * 4);
    int its, nf, cuts, c;
    p->nthreads = 1;
    step_dt_range(p, 0, ncol, dt, converged, converged_reason, w, w + n, w + 2 * n, w + 3 * n, &its, &nf, &cuts);
    for (c = 0; c < ncol; c++) { p->stat_its[c] = its; p->stat_reason[c] = *converged_reason; p->stat_cuts[c] = cuts; p->stat_nf[c] = nf; }
    free(w);
  } else {
    int c, all_conv = 1, worst = 0;
    double *w = (double *)malloc(sizeof(double) * (size_t)n * 4);
    double saved_dtime = p->dtime;
    (void)saved_dtime;
#ifdef _OPENMP
#pragma omp parallel for schedule(dynamic, 64) num_threads(p->nthreads)
#endif
    for (c = 0; c < ncol; c++) {
      /* NB p->dtime is shared: per-column dt cuts make it column-specific, so run on a shallow private copy */
      orc_vsfm q = *p;
      int conv, reason, its, nf, cuts;
      step_dt_range(&q, c, c + 1, dt, &conv, &reason, w, w + n, w + 2 * n, w + 3 * n, &its, &nf, &cuts);
      p->stat_its[c] = its; p->stat_reason[c] = reason; p->stat_cuts[c] = cuts; p->stat_nf[c] = nf;
      if (!conv) {
#ifdef _OPENMP
#pragma omp critical
#endif
        all_conv = 0;
      }
    }
    p->dtime = dt;
    /* report the worst per-column reason (most negative, else the smallest positive) */
    worst = p->stat_reason[0];
    for (c = 1; c < ncol; c++) if (p->stat_reason[c] < worst) worst = p->stat_reason[c];
    *converged = all_conv; *converged_reason = worst;
    free(w);
  }
  return 0;
}

void orc_vsfm_get_stats(orc_vsfm *p, int *newton_its, int *reasons, int *ncuts, int *nfuncs)
{
  int c;
  for (c = 0; c < p->ncol; c++) {
    if (newton_its) newton_its[c] = p->stat_its[c];
    if (reasons) reasons[c] = p->stat_reason[c];
    if (ncuts) ncuts[c] = p->stat_cuts[c];
    if (nfuncs) nfuncs[c] = p->stat_nf[c];
  }
}

/* one residual + Jacobian evaluation at x with accum_prev taken at x_prev (for kernel unit tests) */
void orc_vsfm_eval(orc_vsfm *p, double dt, const double *x_prev, const double *x, double *f, double *ja, double *jb, double *jc)
{
  size_t nb = sizeof(double) * (size_t)p->ncells;
  double *keep = (double *)malloc(nb);
  memcpy(keep, p->soln_prev, nb);
  memcpy(p->soln_prev, x_prev, nb);
  p->dtime = dt;
  pre_solve(p, 0, p->ncol);
  residual_range(p, 0, p->ncol, x, f);
  jacobian_range(p, 0, p->ncol, ja, jb, jc);
  memcpy(p->soln_prev, keep, nb);
  free(keep);
}


/* =====================================================================================================================
 * MPPVSFMALM_Solve (src/driver/alm/MPPVSFMALM_Driver.F90:204-923) for a batch of independent columns: ELM's raw column
 * arrays in, ELM's raw column arrays out.
 *   :204-240   root-fraction weighting of transpiration over the patches of a column           (optional: patch arrays)
 *   :325-372   source/sink packing: ET by root fraction, infiltration, dew, sublimation, drainage distributed over the
 *              layers below the water table and limited by the liquid water there, snow-layer disappearance
 *   :435-450   frac_ice / frac_liq_sat
 *   :552-601   per-column mass and flux totals
 *   :603-923   PreStepDT, the retry loop (<= 10 StepDT calls: a diverged step continues with the remaining time and
 *              stol = 1e-10, a second divergence drops the ice impedance; a converged step whose mass-balance error is
 *              >= 1e-5 kg is redone from the start with rtol or stol tightened tenfold), unpacking (h2osoi_liq/ice, smp_l
 *              in mm, soil pressure, water-table depth by interpolation of the matric potential), PostStepDT
 * The reference runs this loop once per MPI rank (global convergence flag, global maximum of the mass error); here every
 * column is treated as the reference treats a rank holding only that column, in line with the per-column Newton iteration.
 * Lateral / seepage branches (:452-551, :708-800) are not part of the 1-D path.  Cell arrays are cell-ordered (c*nlev + j),
 * zi holds nlev+1 interfaces per column (zi(c,0) first).  PARITY UNPINNED (ELM-only code, no baseline).  TEST INFRASTRUCTURE ONLY.
 * ===================================================================================================================== */
int orc_vsfm_elm_solve(orc_vsfm *p, double dtime_full, int nlevsoi, double watmin, const int *cond_ids /* infil, et, dew, drainage, snow, sublimation */,
                       /* optional patch level (NULL: rootr_col is used as given) */
                       int max_patch_per_col, const int *col_pfti, const int *col_npfts, const int *pft_active, const double *pft_wtcol,
                       const double *rootr_pft /* (npft, nlev) patch-major */, const double *qflx_tran_veg_pft,
                       /* column level */
                       double *rootr_col, const double *qflx_tran_veg_col, const double *qflx_infl, const double *qflx_dew_snow,
                       const double *qflx_dew_grnd, const double *qflx_sub_snow, const double *frac_h2osfc, const int *snl,
                       double *qflx_drain, double *zwt, const double *zi, const double *dz, double *h2osoi_liq, double *h2osoi_ice,
                       double *mflx_snowlyr_col, const double *mflx_neg_snow, const double *mflx_drain_perched,
                       /* outputs */
                       double *smp_l, double *soilp, double *qcharge, double *abs_mass_error, int *iter_count_out, int *status_out)
{
  const int ncol = p->ncol, nlev = p->nlev, n = p->ncells;
  const double area = 1.0, conv = area * ORC_DENH2O * 1.0e-3;          /* flux_unit_conversion [mm/s] -> [kg/s] (:330) */
  const int max_iter_count = 10; const double max_abs_mass_error_col = 1.0e-5, stol_alternate = 1.0e-10;
  vcond *c_inf = &p->ss[cond_ids[0] - 1], *c_et = &p->ss[cond_ids[1] - 1], *c_dew = &p->ss[cond_ids[2] - 1],
        *c_drn = &p->ss[cond_ids[3] - 1], *c_snw = &p->ss[cond_ids[4] - 1], *c_sub = &p->ss[cond_ids[5] - 1];
  double *w = (double *)malloc(sizeof(double) * (size_t)n * 4);
  int c, nfail = 0;
  if (!c_et->per_cell || !c_drn->per_cell || c_inf->per_cell || c_dew->per_cell || c_snw->per_cell || c_sub->per_cell) { free(w); return 1; }
#ifdef _OPENMP
#pragma omp parallel for schedule(dynamic, 16) num_threads(p->nthreads) reduction(+:nfail)
#endif
  for (c = 0; c < ncol; c++) {
    orc_vsfm q = *p;                       /* private dtime / time / tolerances; the arrays are shared, this column's slices only */
    const int off = c * nlev;
    const double *zic = zi + (size_t)c * (nlev + 1);
    double frac_ice[256];
    double tot_et = 0.0, tot_drain = 0.0, mass_beg = 0.0, tot_flux, mass_end, err = 0.0;
    double dtime = dtime_full, rtol = p->opts.rtol, stol = p->opts.stol;
    int j, iter_count = 0, diverged_count = 0, ok = 0, conv_flag, reason, its, nf, cuts;
    /* ---- :204-240 ---- */
    if (col_pfti) {
      double temp = 0.0; int pi;
      for (j = 0; j < nlevsoi; j++) rootr_col[off + j] = 0.0;
      for (pi = 0; pi < max_patch_per_col; pi++) if (pi < col_npfts[c]) {
        int pp = col_pfti[c] + pi;
        if (!pft_active[pp]) continue;
        for (j = 0; j < nlevsoi; j++) rootr_col[off + j] = rootr_col[off + j] + rootr_pft[(size_t)pp * nlev + j] * qflx_tran_veg_pft[pp] * pft_wtcol[pp];
        temp = temp + qflx_tran_veg_pft[pp] * pft_wtcol[pp];
      }
      if (temp != 0.0) for (j = 0; j < nlevsoi; j++) rootr_col[off + j] = rootr_col[off + j] / temp;
    }
    /* ---- :325-372 ---- */
    for (j = 0; j < nlev; j++) { c_et->soe_value[off + j] = 0.0; c_drn->soe_value[off + j] = 0.0; }
    for (j = 0; j < nlevsoi; j++) c_et->soe_value[off + j] = -qflx_tran_veg_col[c] * rootr_col[off + j] * conv;
    c_inf->soe_value[c] = qflx_infl[c] * conv;
    c_dew->soe_value[c] = 0.0; c_sub->soe_value[c] = 0.0;
    if (snl[c] >= 0) {
      c_dew->soe_value[c] = (qflx_dew_snow[c] + qflx_dew_grnd[c]) * (1.0 - frac_h2osfc[c]) * conv;
      c_sub->soe_value[c] = -qflx_sub_snow[c] * (1.0 - frac_h2osfc[c]) * conv;
    }
    if (qflx_drain[c] > 0.0) {
      int jwt = nlev; double dzsum = 0.0, tot = 0.0;          /* 1-based layer numbers as in the reference */
      for (j = 1; j <= nlev; j++) if (zwt[c] <= zic[j]) { jwt = j - 1; break; }
      if (jwt < 1) jwt = 1;
      for (j = jwt; j <= nlev; j++) dzsum = dzsum + dz[off + j - 1];
      for (j = jwt; j <= nlev; j++) {
        double ql = qflx_drain[c] * dz[off + j - 1] / dzsum;
        if (ql * dtime_full > (h2osoi_liq[off + j - 1] - watmin)) ql = (h2osoi_liq[off + j - 1] - watmin) / dtime_full;
        tot = tot + ql;
        c_drn->soe_value[off + j - 1] = -ql * conv;
      }
      qflx_drain[c] = tot;
    }
    c_snw->soe_value[c] = mflx_snowlyr_col[c] * area + mflx_neg_snow[c] * area;
    mflx_snowlyr_col[c] = 0.0;
    for (j = 0; j < nlev; j++) c_drn->soe_value[off + j] = c_drn->soe_value[off + j] + mflx_drain_perched[off + j];     /* :404 */
    /* ---- :435-450 ---- */
    for (j = 0; j < nlev; j++) {
      frac_ice[j] = h2osoi_ice[off + j] / (h2osoi_liq[off + j] + h2osoi_ice[off + j]);
      p->soe_frac_liq_sat[off + j] = 1.0 - frac_ice[j];
    }
    /* ---- :552-601 ---- */
    for (j = 0; j < nlev; j++) {
      tot_et = tot_et + c_et->soe_value[off + j]; tot_drain = tot_drain + c_drn->soe_value[off + j];
      mass_beg = mass_beg + p->soe_mass[off + j];
    }
    tot_flux = tot_et + c_inf->soe_value[c] + c_dew->soe_value[c] + tot_drain + c_snw->soe_value[c] + c_sub->soe_value[c] + 0.0;
    /* ---- PreStepDT (:603) ---- */
    for (j = 0; j < nlev; j++) { p->soln_prev[off + j] = p->soln_prev_clm[off + j]; p->soln[off + j] = p->soln_prev_clm[off + j]; }
    /* ---- retry loop (:628-923) ---- */
    for (;;) {
      iter_count++;
      q.opts.rtol = rtol; q.opts.stol = stol;
      step_dt_range(&q, c, c + 1, dtime, &conv_flag, &reason, w, w + n, w + 2 * n, w + 3 * n, &its, &nf, &cuts);
      p->stat_its[c] = its; p->stat_reason[c] = reason; p->stat_cuts[c] = cuts; p->stat_nf[c] = nf;
      if (!conv_flag) {
        stol = stol_alternate; diverged_count++;
        dtime = dtime - q.time;
        if (diverged_count > 1) for (j = 0; j < nlev; j++) p->soe_frac_liq_sat[off + j] = 1.0;
      } else {
        int jwt = -1;
        mass_end = 0.0;
        for (j = nlev; j >= 1; j--) {
          const int ic = off + j - 1;
          h2osoi_liq[ic] = (1.0 - frac_ice[j - 1]) * p->soe_mass[ic] / area;
          h2osoi_ice[ic] = frac_ice[j - 1] * p->soe_mass[ic] / area;
          mass_end = mass_end + p->soe_mass[ic];
          smp_l[ic] = p->soe_smp[ic] * 1000.0;
          if (jwt == -1) { if (smp_l[ic] < 0.0) jwt = j; }
        }
        err = fabs(mass_beg - mass_end + tot_flux * dtime_full);
        qcharge[c] = 0.0;
        if (jwt == -1 || jwt == nlev) zwt[c] = zic[nlev];
        else {
          double z_dn = (zic[jwt - 1] + zic[jwt]) / 2.0, z_up = (zic[jwt] + zic[jwt + 1]) / 2.0;
          zwt[c] = (0.0 - smp_l[off + jwt - 1]) / (smp_l[off + jwt - 1] - smp_l[off + jwt]) * (z_dn - z_up) + z_dn;
        }
        for (j = 0; j < nlev; j++) soilp[off + j] = p->soe_pressure[off + j];
        if (err >= max_abs_mass_error_col) {
          if (reason == 3) rtol = rtol / 10.0; else if (reason == 4) stol = stol / 10.0;
          dtime = dtime_full;
          for (j = 0; j < nlev; j++) { p->soln_prev[off + j] = p->soln_prev_clm[off + j]; p->soln[off + j] = p->soln_prev_clm[off + j]; }   /* PreStepDT */
        } else ok = 1;
      }
      if (ok) break;
      if (iter_count >= max_iter_count) break;         /* the reference calls endrun here */
    }
    for (j = 0; j < nlev; j++) p->soln_prev_clm[off + j] = p->soln_prev[off + j];     /* PostStepDT (:935) */
    abs_mass_error[c] = err; iter_count_out[c] = iter_count; status_out[c] = ok;
    if (!ok) nfail++;
  }
  free(w);
  return nfail ? -nfail : 0;
}
