/*
 * thermal.c -- oracle restatement of the soil heat-conduction system (T-based, KSP path) for batches of
 * independent 1-D columns:
 *   src/mpp/auxvar/ThermalKSPTemperatureSoilAuxType.F90:71-171   thermal conductivity / heat capacity
 *   src/mpp/ge/GoveqnThermalKSPTemperatureSoilType.F90:646-1229  RHS (Accum, Divergence) and operator assembly
 *   src/mpp/soe/SystemOfEquationsThermalType.F90:171-759         SoE glue (SetSolnPrevCLM, PreSolve, RHS, operators)
 *   src/mpp/soe/SystemOfEquationsBaseType.F90:555-647            StepDT_KSP
 *   src/mpp/mpp/MultiPhysicsProbThermal.F90:76-208               soil-property setter
 * The snow and standing-surface-water governing equations (and their COND_DIRICHLET_FRM_OTR_GOVEQ coupling)
 * are outside this round's scope.
 * TEST INFRASTRUCTURE ONLY (see mpp_oracle.h).
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif
#include "mpp_oracle.h"

#define MAXCOND 8

/* therm_ksp_temp_soil_auxvar_type (+ base), ThermalKSPTemperature{Base,Soil}AuxType.F90 */
typedef struct {
  double temperature; int is_active;
  double frac, dhsdT, dist_up, dist_dn, therm_cond, heat_cap_pva, condition_value;
  double liq_areal_den, ice_areal_den, snow_water; int num_snow_layer; double tuning_factor, dz;
  double por, therm_cond_minerals, therm_cond_dry, heat_cap_minerals_puv; int is_soil_shallow, itype;
} taux;

typedef struct {
  int itype, region, nconn, per_cell;
  orc_conn *conn;
  double *value;                 /* cur_cond%value */
  double *soe_value, *soe_dhsdT, *soe_frac;   /* SoE mailbox for the condition */
  int *soe_active;
  taux *aux;                     /* aux_vars_bc (BCs only) */
} tcond;

struct orc_thermal {
  int ncol, nlev, ncells, orientation, nthreads;
  double *vol, *dz, *area_xy;
  orc_conn *conn_in;
  int nbc, nss;
  tcond bc[MAXCOND], ss[MAXCOND];
  taux *aux_in;
  double dtime, cnfac;
  double *soln, *soln_prev, *soln_prev_clm;
  /* SoE mailbox for internal cells (sysofeqns_thermal_auxvar_type) */
  double *soe_liq, *soe_ice, *soe_snow_water, *soe_tuning, *soe_frac, *soe_dz, *soe_dist_up, *soe_dist_dn;
  int *soe_nsnow, *soe_active;
  int istsoil, istcrop, istice, istice_mec, istwet;
};

static void taux_init(taux *a)
{
  /* ThermKSPTempBaseAuxVarInit :47-67, ThermKSPTempSoilAuxVarInit :42-68 */
  memset(a, 0, sizeof(*a));
  a->temperature = 273.15; a->is_active = 0; a->tuning_factor = 1.0; a->itype = -1;
}

orc_thermal *orc_thermal_create(int ncol, int nlev)
{
  orc_thermal *p = (orc_thermal *)calloc(1, sizeof(*p));
  int i, n = ncol * nlev;
  p->ncol = ncol; p->nlev = nlev; p->ncells = n; p->nthreads = 1; p->orientation = MESH_ALONG_GRAVITY;
  p->cnfac = 0.5;                                                 /* mpp_varcon.F90:28 */
  p->vol = (double *)calloc(n, 8); p->dz = (double *)calloc(n, 8); p->area_xy = (double *)calloc(n, 8);
  p->conn_in = (orc_conn *)calloc((size_t)ncol * (nlev > 1 ? nlev - 1 : 1), sizeof(orc_conn));
  p->aux_in = (taux *)calloc(n, sizeof(taux));
  for (i = 0; i < n; i++) taux_init(&p->aux_in[i]);
  p->soln = (double *)calloc(n, 8); p->soln_prev = (double *)calloc(n, 8); p->soln_prev_clm = (double *)calloc(n, 8);
  p->soe_liq = (double *)calloc(n, 8); p->soe_ice = (double *)calloc(n, 8); p->soe_snow_water = (double *)calloc(n, 8);
  p->soe_tuning = (double *)malloc(8 * (size_t)n); p->soe_frac = (double *)calloc(n, 8); p->soe_dz = (double *)calloc(n, 8);
  p->soe_dist_up = (double *)calloc(n, 8); p->soe_dist_dn = (double *)calloc(n, 8);
  p->soe_nsnow = (int *)calloc(n, sizeof(int)); p->soe_active = (int *)calloc(n, sizeof(int));
  for (i = 0; i < n; i++) p->soe_tuning[i] = 1.0;
  /* ELM's landunit ids (handed to mpp_varcon_init_landunit by the host model) */
  p->istsoil = 1; p->istcrop = 2; p->istice = 3; p->istice_mec = 4; p->istwet = 6;
  return p;
}

static void tcond_free(tcond *c) { free(c->conn); free(c->value); free(c->soe_value); free(c->soe_dhsdT); free(c->soe_frac); free(c->soe_active); free(c->aux); }

void orc_thermal_destroy(orc_thermal *p)
{
  int i;
  if (!p) return;
  for (i = 0; i < p->nbc; i++) tcond_free(&p->bc[i]);
  for (i = 0; i < p->nss; i++) tcond_free(&p->ss[i]);
  free(p->vol); free(p->dz); free(p->area_xy); free(p->conn_in); free(p->aux_in);
  free(p->soln); free(p->soln_prev); free(p->soln_prev_clm);
  free(p->soe_liq); free(p->soe_ice); free(p->soe_snow_water); free(p->soe_tuning); free(p->soe_frac); free(p->soe_dz);
  free(p->soe_dist_up); free(p->soe_dist_dn); free(p->soe_nsnow); free(p->soe_active);
  free(p);
}

void orc_thermal_set_threads(orc_thermal *p, int nthreads) { p->nthreads = nthreads > 0 ? nthreads : 1; }
void orc_thermal_set_cnfac(orc_thermal *p, double cnfac) { p->cnfac = cnfac; }

/* Mesh: as orc_vsfm_set_mesh.  conn_dist_up / conn_dist_dn (optional, (ncol, nlev-1) Fortran order) override the
 * default dz/2 distances, as the ELM driver does (MPPThermalTBasedALM_Initialize.F90:379-381). */
int orc_thermal_set_mesh(orc_thermal *p, int orientation, const double *dz, const double *area, const double *face_area)
{
  int c, j, ncol = p->ncol, nlev = p->nlev;
  (void)face_area;
  p->orientation = orientation;
  for (c = 0; c < ncol; c++) {
    for (j = 0; j < nlev; j++) {
      int ic = c * nlev + j;
      p->dz[ic] = dz[(size_t)j * ncol + c]; p->area_xy[ic] = area[c]; p->vol[ic] = area[c] * p->dz[ic];
    }
    for (j = 0; j < nlev - 1; j++) {
      orc_conn *cn = &p->conn_in[c * (nlev - 1) + j];
      cn->id_up = c * nlev + j; cn->id_dn = cn->id_up + 1; cn->area = area[c];
      cn->dist_up = 0.5 * p->dz[cn->id_up]; cn->dist_dn = 0.5 * p->dz[cn->id_dn];
    }
  }
  return 0;
}

int orc_thermal_set_conn_dist(orc_thermal *p, const double *dist_up, const double *dist_dn)
{
  int c, j, ncol = p->ncol, nlev = p->nlev;
  for (c = 0; c < ncol; c++) for (j = 0; j < nlev - 1; j++) {
    orc_conn *cn = &p->conn_in[c * (nlev - 1) + j];
    cn->dist_up = dist_up[(size_t)j * ncol + c]; cn->dist_dn = dist_dn[(size_t)j * ncol + c];
  }
  return 0;
}

int orc_thermal_add_condition(orc_thermal *p, int ss_or_bc, int cond_type, int region)
{
  int c, j, ncol = p->ncol, nlev = p->nlev, i, n;
  tcond *cd;
  if (ss_or_bc == COND_BC) { if (p->nbc >= MAXCOND) return -1; cd = &p->bc[p->nbc++]; }
  else                     { if (p->nss >= MAXCOND) return -1; cd = &p->ss[p->nss++]; }
  memset(cd, 0, sizeof(*cd));
  cd->itype = cond_type; cd->region = region; cd->per_cell = (region == SOIL_CELLS);
  n = cd->per_cell ? ncol * nlev : ncol;
  cd->nconn = n;
  cd->conn = (orc_conn *)calloc(n, sizeof(orc_conn)); cd->value = (double *)calloc(n, 8);
  cd->soe_value = (double *)calloc(n, 8); cd->soe_dhsdT = (double *)calloc(n, 8); cd->soe_frac = (double *)calloc(n, 8);
  cd->soe_active = (int *)calloc(n, sizeof(int));
  if (ss_or_bc == COND_BC) { cd->aux = (taux *)calloc(n, sizeof(taux)); for (i = 0; i < n; i++) taux_init(&cd->aux[i]); }
  if (cd->per_cell) {
    for (c = 0; c < ncol; c++) for (j = 0; j < nlev; j++) {
      orc_conn *cn = &cd->conn[c * nlev + j];
      cn->id_up = -1; cn->id_dn = c * nlev + j; cn->area = p->area_xy[cn->id_dn];
    }
  } else {
    for (c = 0; c < ncol; c++) {
      orc_conn *cn = &cd->conn[c];
      int first = c * nlev, last = c * nlev + nlev - 1, top_is_first = (p->orientation != MESH_AGAINST_GRAVITY);
      cn->id_up = -1;
      cn->id_dn = (region == SOIL_TOP_CELLS) ? (top_is_first ? first : last) : (top_is_first ? last : first);
      cn->area = p->area_xy[cn->id_dn]; cn->dist_up = 0.0; cn->dist_dn = 0.5 * p->dz[cn->id_dn];
    }
  }
  return ss_or_bc == COND_BC ? p->nbc : p->nss;
}

/* MPPThermalSetSoils, MultiPhysicsProbThermal.F90:76-208 (all columns active: filter_thermal = 1) */
int orc_thermal_set_soils(orc_thermal *p, const double *watsat, const double *csol, const double *tkmg,
                          const double *tkdry, const int *lun_type, int nlevsoi, int istsoil_id)
{
  int c, j, k, i, ncol = p->ncol, nlev = p->nlev;
  if (istsoil_id > 0) p->istsoil = istsoil_id;
  for (c = 0; c < ncol; c++) for (j = 0; j < nlev; j++) {
    size_t t = (size_t)j * ncol + c;
    taux *a = &p->aux_in[c * nlev + j];
    a->is_active = 1;
    a->is_soil_shallow = (j + 1 > nlevsoi) ? 0 : 1;
    a->itype = lun_type[c];
    a->por = watsat[t]; a->therm_cond_minerals = tkmg[t]; a->therm_cond_dry = tkdry[t]; a->heat_cap_minerals_puv = csol[t];
  }
  for (k = 0; k < p->nbc; k++) for (i = 0; i < p->bc[k].nconn; i++) {
    const taux *src = &p->aux_in[p->bc[k].conn[i].id_dn];
    taux *a = &p->bc[k].aux[i];
    /* is_soil_shallow is NOT copied (:195-203): boundary aux vars keep .false. => bedrock conductivity */
    a->itype = src->itype; a->por = src->por; a->therm_cond_minerals = src->therm_cond_minerals;
    a->therm_cond_dry = src->therm_cond_dry; a->heat_cap_minerals_puv = src->heat_cap_minerals_puv;
  }
  for (i = 0; i < p->ncells; i++) p->soe_active[i] = 1;
  return 0;
}

int orc_thermal_set_soln_prev(orc_thermal *p, const double *T)
{ memcpy(p->soln_prev_clm, T, sizeof(double) * (size_t)p->ncells); return 0; }     /* ThermalSOESetSolnPrevCLM :171-199 */

/* SetRDataFromCLM -> SOEThermalAuxSetRData (SystemOfEquationsThermalAuxMod.F90) */
int orc_thermal_set_rdata(orc_thermal *p, int auxvar_type, int var_type, int cond_id, const double *data, int n)
{
  int i;
  if (auxvar_type == AUXVAR_INTERNAL) {
    double *dst;
    if (n > p->ncells) return 1;
    switch (var_type) {
    case VAR_LIQ_AREAL_DEN: dst = p->soe_liq; break;
    case VAR_ICE_AREAL_DEN: dst = p->soe_ice; break;
    case VAR_SNOW_WATER: dst = p->soe_snow_water; break;
    case VAR_TUNING_FACTOR: dst = p->soe_tuning; break;
    case VAR_FRAC: dst = p->soe_frac; break;
    case VAR_DZ: dst = p->soe_dz; break;
    case VAR_DIST_UP: dst = p->soe_dist_up; break;
    case VAR_DIST_DN: dst = p->soe_dist_dn; break;
    case VAR_TEMPERATURE: dst = p->soln_prev_clm; break;
    default: return 2;
    }
    for (i = 0; i < n; i++) dst[i] = data[i];
    return 0;
  } else {
    tcond *cd;
    if (auxvar_type == AUXVAR_BC) { if (cond_id < 1 || cond_id > p->nbc) return 3; cd = &p->bc[cond_id - 1]; }
    else if (auxvar_type == AUXVAR_SS) { if (cond_id < 1 || cond_id > p->nss) return 3; cd = &p->ss[cond_id - 1]; }
    else return 4;
    if (n > cd->nconn) return 1;
    switch (var_type) {
    case VAR_BC_SS_CONDITION: for (i = 0; i < n; i++) cd->soe_value[i] = data[i]; break;
    case VAR_DHS_DT: for (i = 0; i < n; i++) cd->soe_dhsdT[i] = data[i]; break;
    case VAR_FRAC: for (i = 0; i < n; i++) cd->soe_frac[i] = data[i]; if (cd->aux) for (i = 0; i < n; i++) cd->aux[i].frac = data[i]; break;
    case VAR_ACTIVE: if (cd->aux) for (i = 0; i < n; i++) cd->aux[i].is_active = (data[i] != 0.0); break;   /* ThermKSPTempSoilAuxVarSetRValues */
    default: return 2;
    }
    return 0;
  }
}

int orc_thermal_set_idata(orc_thermal *p, int auxvar_type, int var_type, int cond_id, const int *data, int n)
{
  int i;
  (void)cond_id;
  if (auxvar_type != AUXVAR_INTERNAL || n > p->ncells) return 1;
  if (var_type == VAR_NUM_SNOW_LYR) { for (i = 0; i < n; i++) p->soe_nsnow[i] = data[i]; return 0; }
  if (var_type == VAR_ACTIVE) { for (i = 0; i < n; i++) p->soe_active[i] = data[i]; return 0; }
  return 2;
}
int orc_thermal_set_bdata(orc_thermal *p, int auxvar_type, int var_type, int cond_id, const int *data, int n)
{ return orc_thermal_set_idata(p, auxvar_type, var_type, cond_id, data, n); }

/* ThermKSPTempSoilAuxVarCompute, ThermalKSPTemperatureSoilAuxType.F90:71-171 */
static void taux_compute(const orc_thermal *p, taux *a, double dz, double vol)
{
  double satw, fl, dke, dksat;
  (void)vol;
  if (a->itype == p->istsoil || a->itype == p->istcrop) {
    if (a->is_soil_shallow) {
      satw = (a->liq_areal_den / ORC_DENH2O + a->ice_areal_den / ORC_DENICE) / (dz * a->por);
      satw = fmin(1.0, satw);
      if (satw > (double).1e-6f) {                       /* `.1e-6` is a default-REAL literal */
        if (a->temperature >= ORC_TFRZ) dke = fmax(0.0, log10(satw) + 1.0);
        else                            dke = satw;
        fl = (a->liq_areal_den / (ORC_DENH2O * dz)) / (a->liq_areal_den / (ORC_DENH2O * dz) + a->ice_areal_den / (ORC_DENICE * dz));
        dksat = a->therm_cond_minerals * pow(ORC_TKWAT, fl * a->por) * pow(ORC_TKICE, (1.0 - fl) * a->por);
        a->therm_cond = dke * dksat + (1.0 - dke) * a->therm_cond_dry;
      } else {
        a->therm_cond = a->therm_cond_dry;
      }
      a->heat_cap_pva = a->heat_cap_minerals_puv * (1.0 - a->por) * dz + a->ice_areal_den * ORC_CPICE + a->liq_areal_den * ORC_CPLIQ;
      if (a->num_snow_layer == 0) a->heat_cap_pva = a->heat_cap_pva + a->snow_water * ORC_CPICE;
    } else {
      a->therm_cond   = ORC_THK_BEDROCK;
      a->heat_cap_pva = a->heat_cap_minerals_puv * (1.0 - a->por) * dz + a->ice_areal_den * ORC_CPICE + a->liq_areal_den * ORC_CPLIQ;
    }
    a->heat_cap_pva = a->heat_cap_pva / dz;
  } else if (a->itype == p->istwet) {
    if (a->is_soil_shallow) {
      a->therm_cond = (a->temperature < ORC_TFRZ) ? ORC_TKICE : ORC_TKWAT;
      a->heat_cap_pva = a->ice_areal_den * ORC_CPICE + a->liq_areal_den * ORC_CPLIQ;
      if (a->num_snow_layer == 0) a->heat_cap_pva = a->heat_cap_pva + a->snow_water * ORC_CPICE;
      a->heat_cap_pva = a->heat_cap_pva / dz;
    } else {
      a->therm_cond = ORC_THK_BEDROCK; a->heat_cap_pva = a->heat_cap_minerals_puv;
    }
  } else if (a->itype == p->istice || a->itype == p->istice_mec) {
    a->therm_cond = (a->temperature < ORC_TFRZ) ? ORC_TKICE : ORC_TKWAT;
    a->heat_cap_pva = a->ice_areal_den * ORC_CPICE + a->liq_areal_den * ORC_CPLIQ;
    if (a->num_snow_layer == 0) a->heat_cap_pva = a->heat_cap_pva + a->snow_water * ORC_CPICE;
    a->heat_cap_pva = a->heat_cap_pva / dz;
  }
}

/* one StepDT_KSP for columns [c0,c1) */
static void step_range(orc_thermal *p, int c0, int c1, double stale_area_global)
{
  int c, j, k, nlev = p->nlev;
  double dt = p->dtime, cnfac = p->cnfac;
  double *b = (double *)calloc((size_t)nlev * 5, 8), *la = b + nlev, *lb = la + nlev, *lc = lb + nlev, *x = lc + nlev;

  for (c = c0; c < c1; c++) {
    int off = c * nlev;
    double area = stale_area_global, factor = 1.0;     /* see the stale-variable note below */
    /* ---- PreSolve (SystemOfEquationsThermalType.F90:412-481) ---- */
    for (j = 0; j < nlev; j++) {
      taux *a = &p->aux_in[off + j];
      a->temperature = p->soln_prev[off + j];                        /* SavePrimaryIndependentVar(soln_prev) */
      a->liq_areal_den = p->soe_liq[off + j]; a->ice_areal_den = p->soe_ice[off + j];    /* GetFromSOEAuxVarsIntrn :240-272 */
      a->snow_water = p->soe_snow_water[off + j]; a->num_snow_layer = p->soe_nsnow[off + j];
      a->tuning_factor = p->soe_tuning[off + j]; a->frac = p->soe_frac[off + j]; a->dz = p->soe_dz[off + j];
      a->is_active = p->soe_active[off + j];
    }
    for (k = 0; k < p->nbc; k++) {                                    /* GetFromSOEAuxVarsBC :276-366 */
      tcond *cd = &p->bc[k];
      int i = c, cell = cd->conn[i].id_dn;
      if (cd->itype == COND_HEAT_FLUX) {
        cd->aux[i].condition_value = cd->soe_value[i];
        cd->value[i] = cd->soe_value[i] - cd->soe_dhsdT[i] * p->aux_in[cell].temperature;   /* H - dH/dT * T */
        cd->aux[i].dhsdT = cd->soe_dhsdT[i];
        cd->aux[i].frac = cd->soe_frac[i];
      } else if (cd->itype == COND_DIRICHLET) {
        cd->aux[i].condition_value = cd->soe_value[i];
      }
    }
    for (k = 0; k < p->nss; k++) {                                    /* GetFromSOEAuxVarsSS :370-446 */
      tcond *cd = &p->ss[k];
      int i0 = cd->per_cell ? off : c, i1 = cd->per_cell ? off + nlev : c + 1, i;
      for (i = i0; i < i1; i++) cd->value[i] = cd->soe_value[i];
    }
    /* ---- ComputeRHS: UpdateAuxVarsIntrn / BC (:546-649; GoveqnThermalKSP...:556-640) ---- */
    for (j = 0; j < nlev; j++) taux_compute(p, &p->aux_in[off + j], p->dz[off + j], p->vol[off + j]);
    for (k = 0; k < p->nbc; k++) {
      tcond *cd = &p->bc[k];
      int i = c, cell = cd->conn[i].id_dn;
      if (cd->itype == COND_DIRICHLET) cd->aux[i].temperature = cd->aux[i].condition_value;
      else if (cd->itype == COND_HEAT_FLUX) cd->aux[i].temperature = p->aux_in[cell].temperature;
      taux_compute(p, &cd->aux[i], p->dz[cell], p->vol[cell]);
    }
    /* ---- Accum (:671-714) ---- */
    for (j = 0; j < nlev; j++) {
      const taux *a = &p->aux_in[off + j];
      b[j] = 0.0;
      if (a->is_active) b[j] = a->heat_cap_pva * p->vol[off + j] / (dt * a->tuning_factor) * a->temperature;
    }
    /* ---- Divergence (:718-972) ---- */
    for (j = 0; j < nlev - 1; j++) {
      const orc_conn *cn = &p->conn_in[c * (nlev - 1) + j];
      const taux *up = &p->aux_in[cn->id_up], *dn = &p->aux_in[cn->id_dn];
      double therm_cond, flux;
      if (!up->is_active || !dn->is_active) continue;
      /* DiffHeatFlux :976-1003 */
      therm_cond = up->therm_cond * dn->therm_cond * (cn->dist_up + cn->dist_dn) / (up->therm_cond * cn->dist_dn + dn->therm_cond * cn->dist_up);
      flux = -therm_cond * (up->temperature - dn->temperature) / (cn->dist_up + cn->dist_dn);
      b[j]     = b[j]     + cnfac * flux * cn->area * 1.0;
      b[j + 1] = b[j + 1] - cnfac * flux * cn->area * 1.0;
    }
    /* the reference's boundary loop runs after the internal loop over ALL columns: its stale `area` is the one of the
     * mesh's last internal connection (or of the last column's earlier heat-flux BC) */
    area = stale_area_global; factor = 1.0;
    for (k = 0; k < p->nbc; k++) {
      tcond *cd = &p->bc[k];
      int i = c, cell = cd->conn[i].id_dn, jc = cell - off;
      const taux *in = &p->aux_in[cell]; const taux *ba = &cd->aux[i];
      if (!in->is_active) continue;
      if (cd->itype == COND_DIRICHLET) {
        double dist_up, dist_dn, dist, kup, kdn, kav;
        if (!ba->is_active) continue;
        dist_up = cd->conn[i].dist_up; dist_dn = cd->conn[i].dist_dn; dist = dist_up + dist_dn;
        kup = ba->therm_cond; kdn = in->therm_cond;
        kav = kup * kdn * (dist_up + dist_dn) / (kup * dist_dn + kdn * dist_up);
        /* NB `area` and `factor` are NOT assigned in this branch of the reference (:883-908): they keep whatever the
         * previous loop left behind (the last internal connection of the mesh, or an earlier heat-flux BC) */
        b[jc] = b[jc] + kav / dist * ba->temperature * area * factor;
      } else if (cd->itype == COND_HEAT_FLUX) {
        b[jc] = b[jc] + 1.0 * cd->value[i] * ba->frac * cd->conn[i].area;
      }
    }
    for (k = 0; k < p->nss; k++) {
      tcond *cd = &p->ss[k];
      int i0 = cd->per_cell ? off : c, i1 = cd->per_cell ? off + nlev : c + 1, i;
      for (i = i0; i < i1; i++) {
        int cell = cd->conn[i].id_dn;
        if (!p->aux_in[cell].is_active) continue;
        if (cd->itype == COND_HEAT_RATE) b[cell - off] = b[cell - off] + cd->value[i] * 1.0;
      }
    }
    /* ---- ComputeOperatorsDiag (:1007-1229) ---- */
    for (j = 0; j < nlev; j++) {
      const taux *a = &p->aux_in[off + j];
      la[j] = 0.0; lc[j] = 0.0;
      lb[j] = a->is_active ? a->heat_cap_pva * p->vol[off + j] / (dt * a->tuning_factor) : 1.0;
    }
    for (j = 0; j < nlev - 1; j++) {
      const orc_conn *cn = &p->conn_in[c * (nlev - 1) + j];
      const taux *up = &p->aux_in[cn->id_up], *dn = &p->aux_in[cn->id_dn];
      double dist, kav, value;
      if (!up->is_active || !dn->is_active) continue;
      dist = cn->dist_up + cn->dist_dn;
      kav = up->therm_cond * dn->therm_cond * dist / (up->therm_cond * cn->dist_dn + dn->therm_cond * cn->dist_up);
      value = (1.0 - cnfac) * kav / dist * cn->area;
      lb[j]     +=  value * 1.0; lc[j]     += -value * 1.0;
      la[j + 1] += -value * 1.0; lb[j + 1] +=  value * 1.0;
    }
    for (k = 0; k < p->nbc; k++) {
      tcond *cd = &p->bc[k];
      int i = c, cell = cd->conn[i].id_dn, jc = cell - off;
      const taux *in = &p->aux_in[cell]; const taux *ba = &cd->aux[i];
      if (!in->is_active) continue;
      if (cd->itype == COND_DIRICHLET) {
        double dist_up, dist_dn, dist, kav, value;
        if (!ba->is_active) continue;
        dist_up = cd->conn[i].dist_up; dist_dn = cd->conn[i].dist_dn; dist = dist_up + dist_dn;
        kav = ba->therm_cond * in->therm_cond * dist / (ba->therm_cond * dist_dn + in->therm_cond * dist_up);
        value = ba->frac * (1.0 - cnfac) * kav / dist * cd->conn[i].area * 1.0;
        lb[jc] += value;
      } else if (cd->itype == COND_HEAT_FLUX) {
        /* `value = -frac*dhsdT**area*factor` (:1215): ** binds tighter than *, i.e. dhsdT raised to the power `area` */
        double value = -ba->frac * pow(ba->dhsdT, cd->conn[i].area) * 1.0;
        lb[jc] += value;
      }
    }
    /* ---- KSPSolve: tridiagonal => ILU(0) exact ---- */
    orc_tridiag_solve(nlev, la, lb, lc, b, x);
    /* ---- PostSolve (SOEBasePostSolve :650-668) ---- */
    for (j = 0; j < nlev; j++) { p->soln[off + j] = x[j]; p->soln_prev[off + j] = x[j]; p->aux_in[off + j].temperature = x[j]; }
  }
  free(b);
}

void orc_thermal_pre_step_dt(orc_thermal *p)
{
  size_t nb = sizeof(double) * (size_t)p->ncells;               /* ThermalSOEPreStepDT :393-408 */
  memcpy(p->soln_prev, p->soln_prev_clm, nb); memcpy(p->soln, p->soln_prev_clm, nb);
}

int orc_thermal_step_dt(orc_thermal *p, double dt, int nstep, int *converged)
{
  int c, ncol = p->ncol;
  double stale_area = (p->nlev > 1) ? p->conn_in[(size_t)ncol * (p->nlev - 1) - 1].area : p->area_xy[0];
  (void)nstep;
  p->dtime = dt;
#ifdef _OPENMP
#pragma omp parallel for schedule(static) num_threads(p->nthreads)
#endif
  for (c = 0; c < ncol; c++) step_range(p, c, c + 1, stale_area);
  if (converged) *converged = 1;
  return 0;
}

int orc_thermal_get_soln(orc_thermal *p, double *T) { memcpy(T, p->soln, sizeof(double) * (size_t)p->ncells); return 0; }

int orc_thermal_get_aux(orc_thermal *p, int var_type, double *data)
{
  int i;
  for (i = 0; i < p->ncells; i++) {
    if (var_type == VAR_THERMAL_COND) data[i] = p->aux_in[i].therm_cond;
    else if (var_type == VAR_HEAT_CAP) data[i] = p->aux_in[i].heat_cap_pva;
    else return 2;
  }
  return 0;
}
