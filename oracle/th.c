/*
 * th.c -- oracle restatement of the coupled thermal-hydrology (TH) system for batches of independent 1-D columns:
 * Richards equation (mass, ieqn 1) + enthalpy-form energy equation (ieqn 2), two unknowns (P, T) per cell.
 *   src/mpp/soe/SystemOfEquationsTHType.F90:119-302, 533-1004    PreSolve/PostSolve, aux-var exchange, residual, Jacobian
 *   src/mpp/auxvar/ThermalEnthalpySoilAuxType.F90:57-278         energy aux vars
 *   src/mpp/ge/ThermalEnthalpyMod.F90:27-332                     energy flux and derivatives
 *   src/mpp/ge/GoveqnThermalEnthalpySoilType.F90:1174-2377       accumulation, divergence, Jacobian blocks
 *   src/mpp/ge/GoveqnRichardsODEPressureType.F90:2333-2613       d(mass residual)/dT block
 *   src/mpp/mpp/MultiPhysicsProbTH.F90:75-560                    soil-property setters
 *
 * Linear solve: the reference orders unknowns [P_1..P_N | T_1..T_N] and preconditions GMRES with ILU(0), so its Newton
 * method is inexact (KSP rtol 1e-5); here unknowns are interleaved per cell and the 2x2 block-tridiagonal Newton
 * system is solved exactly (block Thomas).  Converged states agree to ~1e-12 relative with the reference baseline
 * (SURVEY.md Appendix C); iteration paths differ.
 *
 * The reference's TH has no freeze-thaw physics (no ice phase in therm_enthalpy_soil_auxvar_type): none is added.
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

/* therm_enthalpy_soil_auxvar_type extends rich_ode_pres_auxvar_type (ThermalEnthalpySoilAuxType.F90:19-50) */
typedef struct {
  orc_rich_auxvar r;
  double ul, hl, dul_dP, dhl_dP, dul_dT, dhl_dT, dkr_dT, dsat_dT;
  double Kel, therm_cond_wet, therm_cond_dry, therm_cond, dtherm_cond_dP, dKel_dp, therm_alpha;
  double den_soil, heat_cap_soil;
  int int_energy_enthalpy_type;
} eaux;

typedef struct {
  int ieqn, itype, region, nconn, per_cell;
  orc_conn *conn;
  double *value, *soe_value;
  orc_rich_auxvar *raux;   /* mass-GE boundary aux vars */
  eaux *eaux_;             /* energy-GE boundary aux vars */
} thcond;

struct orc_th {
  int ncol, nlev, ncells, orientation, per_column, nthreads;
  double *vol, *dz, *area_xy;
  orc_conn *conn_in;
  int nbc, nss;
  thcond bc[MAXCOND], ss[MAXCOND];
  orc_rich_auxvar *maux;   /* mass GE aux_vars_in */
  eaux *ea;                /* energy GE aux_vars_in */
  double dtime;
  double *accum_prev_m, *accum_prev_e;
  double *soln, *soln_prev;          /* interleaved (P,T) per cell, 2*ncells */
  orc_snes_opts opts;
  int *stat_its, *stat_reason, *stat_cuts, *stat_nf;
};

static void eaux_init(eaux *a)
{
  /* ThermEnthalpyAuxVarInit :57-100 */
  memset(a, 0, sizeof(*a));
  orc_rich_auxvar_init(&a->r);
  a->r.perm[0] = a->r.perm[1] = a->r.perm[2] = 8.3913e-12;     /* :93; never overwritten (MultiPhysicsProbTH.F90:293 is commented out) */
  a->int_energy_enthalpy_type = INT_ENERGY_ENTHALPY_CONSTANT;
}

/* ThermEnthalpyAuxVarCompute :219-278 */
static void eaux_compute(eaux *a)
{
  double pressure;
  orc_press_to_sat(&a->r.satParams, a->r.pressure, &a->r.sat, &a->r.dsat_dP);
  orc_press_to_relperm(&a->r.satParams, a->r.pressure, 1.0, &a->r.kr, &a->r.dkr_dP);
  a->r.por = a->r.por_base; a->r.dpor_dP = 0.0;
  pressure = a->r.pressure;
  if (a->r.pressure < ORC_PRESSURE_REF) pressure = ORC_PRESSURE_REF;
  orc_density(pressure, a->r.temperature, a->r.density_type, &a->r.den, &a->r.dden_dP, &a->r.dden_dT);
  orc_viscosity(pressure, a->r.temperature, &a->r.vis, &a->r.dvis_dP, &a->r.dvis_dT);
  orc_internal_energy_enthalpy(pressure, a->r.temperature, a->int_energy_enthalpy_type, a->r.den * ORC_FMWH2O,
                               a->r.dden_dT * ORC_FMWH2O, a->r.dden_dP * ORC_FMWH2O,
                               &a->ul, &a->hl, &a->dul_dT, &a->dhl_dT, &a->dul_dP, &a->dhl_dP);
  a->Kel     = pow(a->r.sat + 1.e-6, a->therm_alpha);
  a->dKel_dp = a->therm_alpha * pow(a->r.sat + 1.e-6, a->therm_alpha - 1.0) * a->r.dsat_dP;
  a->therm_cond     = a->therm_cond_wet * a->Kel + a->therm_cond_dry * (1.0 - a->Kel);
  a->dtherm_cond_dP = (a->therm_cond_wet - a->therm_cond_dry) * a->dKel_dp;
}

/* ThermalEnthalpyFlux (ThermalEnthalpyMod.F90:27-165) and ...DerivativeWrtPressure (:168-332); wrt = 0: T, 1: P */
static void energy_flux(const eaux *up, const eaux *dn, double mflux, double dmflux_up, double dmflux_dn, const orc_conn *conn,
                        int internal_conn, int cond_type, int wrt, double *eflux, double *de_up, double *de_dn)
{
  double T_up = up->r.temperature, T_dn = dn->r.temperature, area = conn->area, dist_up = conn->dist_up, dist_dn = conn->dist_dn;
  double kup = up->therm_cond, kdn = dn->therm_cond, kod, h, dh_up, dh_dn;
  if (internal_conn || cond_type == COND_DIRICHLET_FRM_OTR_GOVEQ) kod = (kup * kdn) / (dist_up * kdn + dist_dn * kup);
  else                                                             kod = kdn / (dist_up + dist_dn);
  h = (mflux <= 0.0) ? up->hl : dn->hl;
  *eflux = mflux * h + (-kod * (T_up - T_dn) * area);
  if (wrt == 0) {
    if (mflux < 0.0) { dh_up = up->dhl_dT; dh_dn = 0.0; } else { dh_up = 0.0; dh_dn = dn->dhl_dT; }
    *de_up = dmflux_up * h + mflux * dh_up + (-kod * area);
    *de_dn = dmflux_dn * h + mflux * dh_dn + (+kod * area);
  } else {
    double dDk_up, dDk_dn;
    if (mflux < 0.0) { dh_up = up->dhl_dP; dh_dn = 0.0; } else { dh_up = 0.0; dh_dn = dn->dhl_dP; }
    if (internal_conn || cond_type == COND_DIRICHLET_FRM_OTR_GOVEQ) {
      dDk_up = pow(kod, 2.0) / pow(kup, 2.0) * dist_up * up->dtherm_cond_dP;
      dDk_dn = pow(kod, 2.0) / pow(kdn, 2.0) * dist_dn * dn->dtherm_cond_dP;
    } else {
      dDk_up = 0.0;
      dDk_dn = 1.0 / (dist_up + dist_dn) * dn->dtherm_cond_dP;
    }
    *de_up = dmflux_up * h + mflux * dh_up + (-dDk_up * (T_up - T_dn) * area);
    *de_dn = dmflux_dn * h + mflux * dh_dn + (-dDk_dn * (T_up - T_dn) * area);
  }
}

orc_th *orc_th_create(int ncol, int nlev)
{
  orc_th *p = (orc_th *)calloc(1, sizeof(*p));
  int i, n = ncol * nlev;
  p->ncol = ncol; p->nlev = nlev; p->ncells = n; p->orientation = MESH_ALONG_GRAVITY; p->per_column = 1; p->nthreads = 1;
  p->vol = (double *)calloc(n, 8); p->dz = (double *)calloc(n, 8); p->area_xy = (double *)calloc(n, 8);
  p->conn_in = (orc_conn *)calloc((size_t)ncol * (nlev > 1 ? nlev - 1 : 1), sizeof(orc_conn));
  p->maux = (orc_rich_auxvar *)calloc(n, sizeof(orc_rich_auxvar)); p->ea = (eaux *)calloc(n, sizeof(eaux));
  for (i = 0; i < n; i++) { orc_rich_auxvar_init(&p->maux[i]); eaux_init(&p->ea[i]); }
  p->accum_prev_m = (double *)calloc(n, 8); p->accum_prev_e = (double *)calloc(n, 8);
  p->soln = (double *)calloc(2 * (size_t)n, 8); p->soln_prev = (double *)calloc(2 * (size_t)n, 8);
  orc_snes_default_opts(&p->opts);
  p->stat_its = (int *)calloc(ncol, sizeof(int)); p->stat_reason = (int *)calloc(ncol, sizeof(int));
  p->stat_cuts = (int *)calloc(ncol, sizeof(int)); p->stat_nf = (int *)calloc(ncol, sizeof(int));
  return p;
}

void orc_th_destroy(orc_th *p)
{
  int k;
  if (!p) return;
  for (k = 0; k < p->nbc + p->nss; k++) {
    thcond *c = k < p->nbc ? &p->bc[k] : &p->ss[k - p->nbc];
    free(c->conn); free(c->value); free(c->soe_value); free(c->raux); free(c->eaux_);
  }
  free(p->vol); free(p->dz); free(p->area_xy); free(p->conn_in); free(p->maux); free(p->ea);
  free(p->accum_prev_m); free(p->accum_prev_e); free(p->soln); free(p->soln_prev);
  free(p->stat_its); free(p->stat_reason); free(p->stat_cuts); free(p->stat_nf);
  free(p);
}

void orc_th_set_mode(orc_th *p, int per_column, int nthreads) { p->per_column = per_column; p->nthreads = nthreads > 0 ? nthreads : 1; }
void orc_th_set_tolerances(orc_th *p, double atol, double rtol, double stol, int max_it, int max_funcs)
{ p->opts.atol = atol; p->opts.rtol = rtol; p->opts.stol = stol; p->opts.max_it = max_it; p->opts.max_funcs = max_funcs; }

int orc_th_set_mesh(orc_th *p, int orientation, const double *dz, const double *area, const double *face_area)
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
      cn->unitvec[0] = cn->unitvec[1] = cn->unitvec[2] = 0.0;
      if (orientation == MESH_ALONG_GRAVITY) cn->unitvec[2] = -1.0;
      else if (orientation == MESH_AGAINST_GRAVITY) cn->unitvec[2] = 1.0;
      else cn->unitvec[0] = 1.0;
    }
  }
  return 0;
}

int orc_th_add_condition(orc_th *p, int ieqn, int ss_or_bc, int cond_type, int region)
{
  int c, j, i, n, ncol = p->ncol, nlev = p->nlev;
  thcond *cd;
  if (ss_or_bc == COND_BC) { if (p->nbc >= MAXCOND) return -1; cd = &p->bc[p->nbc++]; }
  else                     { if (p->nss >= MAXCOND) return -1; cd = &p->ss[p->nss++]; }
  memset(cd, 0, sizeof(*cd));
  cd->ieqn = ieqn; cd->itype = cond_type; cd->region = region; cd->per_cell = (region == SOIL_CELLS);
  n = cd->per_cell ? ncol * nlev : ncol;
  cd->nconn = n;
  cd->conn = (orc_conn *)calloc(n, sizeof(orc_conn)); cd->value = (double *)calloc(n, 8); cd->soe_value = (double *)calloc(n, 8);
  if (ss_or_bc == COND_BC) {
    if (ieqn == 1) { cd->raux = (orc_rich_auxvar *)calloc(n, sizeof(orc_rich_auxvar)); for (i = 0; i < n; i++) orc_rich_auxvar_init(&cd->raux[i]); }
    else           { cd->eaux_ = (eaux *)calloc(n, sizeof(eaux)); for (i = 0; i < n; i++) eaux_init(&cd->eaux_[i]); }
  }
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
      if (region == SOIL_TOP_CELLS) {
        cn->id_dn = top_is_first ? first : last;
        if (p->orientation == MESH_HORIZONTAL) cn->unitvec[0] = 1.0; else cn->unitvec[2] = -1.0;
      } else {
        cn->id_dn = top_is_first ? last : first;
        if (p->orientation == MESH_HORIZONTAL) cn->unitvec[0] = -1.0; else cn->unitvec[2] = 1.0;
      }
      cn->area = p->area_xy[cn->id_dn]; cn->dist_up = 0.0; cn->dist_dn = 0.5 * p->dz[cn->id_dn];
    }
  }
  return ss_or_bc == COND_BC ? p->nbc : p->nss;
}

/* MPPTHSetSoils -> ...ForVSFM (MultiPhysicsProbTH.F90:405-560) and ...ForThermalEnthalpy (:179-401).
 * One set of (ncol,nlev) tables serves both governing equations. */
int orc_th_set_soils(orc_th *p, const double *watsat, const double *hksat, const double *bsw, const double *sucsat,
                     const double *residual_sat, const double *csol, const double *tkdry,
                     int satfunc_name, int density_type, int iee_type)
{
  const double vish2o = 0.001002;
  int c, j, k, i, rc = 0, ncol = p->ncol, nlev = p->nlev;
  for (c = 0; c < ncol; c++) for (j = 0; j < nlev; j++) {
    size_t t = (size_t)j * ncol + c;
    int ic = c * nlev + j;
    orc_rich_auxvar *m = &p->maux[ic]; eaux *e = &p->ea[ic];
    double perm = hksat[t] * vish2o / (ORC_DENH2O * ORC_GRAV) * 0.001, alpha = 1.0 / (sucsat[t] * ORC_GRAV), lambda = 1.0 / bsw[t];
    orc_satparams sp;
    switch (satfunc_name) {
    case SATFUNC_NAME_BROOKS_COREY: rc |= orc_satfunc_set_bc(&sp, residual_sat[t], alpha, lambda); break;
    case SATFUNC_NAME_SBC_BZ2: rc |= orc_satfunc_set_sbc_bz2(&sp, residual_sat[t], alpha, lambda, -0.9 / alpha); break;
    case SATFUNC_NAME_SBC_BZ3: rc |= orc_satfunc_set_sbc_bz3(&sp, residual_sat[t], alpha, lambda, -0.9 / alpha); break;
    default: rc |= orc_satfunc_set_vg(&sp, residual_sat[t], alpha, lambda);
    }
    m->perm[0] = m->perm[1] = m->perm[2] = perm; m->por = m->por_base = watsat[t]; m->satParams = sp; m->density_type = density_type;
    e->r.por = e->r.por_base = watsat[t]; e->r.satParams = sp; e->r.density_type = density_type;     /* perm stays 8.3913e-12 */
    e->int_energy_enthalpy_type = iee_type;
    e->therm_alpha = 0.45; e->therm_cond_wet = 1.3; e->therm_cond_dry = tkdry[t]; e->heat_cap_soil = csol[t]; e->den_soil = 2700.0;  /* :331-335 */
  }
  for (k = 0; k < p->nbc; k++) for (i = 0; i < p->bc[k].nconn; i++) {
    int cell = p->bc[k].conn[i].id_dn;
    if (p->bc[k].raux) {
      orc_rich_auxvar *a = &p->bc[k].raux[i]; const orc_rich_auxvar *s = &p->maux[cell];
      a->perm[0] = s->perm[0]; a->perm[1] = s->perm[1]; a->perm[2] = s->perm[2]; a->por = s->por; a->por_base = s->por_base;
      a->satParams = s->satParams; a->density_type = density_type;
    } else {
      eaux *a = &p->bc[k].eaux_[i]; const eaux *s = &p->ea[cell];
      a->r.por = s->r.por; a->r.por_base = s->r.por_base; a->r.satParams = s->r.satParams; a->r.density_type = density_type;
      a->int_energy_enthalpy_type = iee_type;
      a->therm_alpha = s->therm_alpha; a->therm_cond_wet = s->therm_cond_wet; a->therm_cond_dry = s->therm_cond_dry;
      a->heat_cap_soil = s->heat_cap_soil; a->den_soil = s->den_soil;
    }
  }
  return rc;
}

/* ThermEnthalpySetSoilPermeability -> ThermalEnthalpySoilAuxVarSetAbsPerm (GoveqnThermalEnthalpySoilType.F90:2454-2480,
 * ThermalEnthalpySoilAuxMod.F90): internal aux vars, then every boundary aux var copies its cell's value.  Called by the
 * standalone drivers (th_mms_problem.F90:739); MPPTHSetSoils does not (MultiPhysicsProbTH.F90:293). */
int orc_th_set_energy_perm(orc_th *p, const double *perm)
{
  int i, k;
  for (i = 0; i < p->ncells; i++) p->ea[i].r.perm[0] = p->ea[i].r.perm[1] = p->ea[i].r.perm[2] = perm[i];
  for (k = 0; k < p->nbc; k++) for (i = 0; i < p->bc[k].nconn; i++) {
    if (p->bc[k].raux) continue;
    {
      eaux *a = &p->bc[k].eaux_[i]; const eaux *s = &p->ea[p->bc[k].conn[i].id_dn];
      a->r.perm[0] = s->r.perm[0]; a->r.perm[1] = s->r.perm[1]; a->r.perm[2] = s->r.perm[2];
    }
  }
  return 0;
}

int orc_th_restart(orc_th *p, const double *press, const double *temp)
{
  int i;
  for (i = 0; i < p->ncells; i++) { p->soln[2 * i] = press[i]; p->soln[2 * i + 1] = temp[i]; }
  memcpy(p->soln_prev, p->soln, sizeof(double) * 2 * (size_t)p->ncells);
  return 0;
}

static thcond *find_cond(orc_th *p, int auxvar_type, int cond_id)
{
  if (auxvar_type == AUXVAR_BC) return (cond_id >= 1 && cond_id <= p->nbc) ? &p->bc[cond_id - 1] : NULL;
  if (auxvar_type == AUXVAR_SS) return (cond_id >= 1 && cond_id <= p->nss) ? &p->ss[cond_id - 1] : NULL;
  return NULL;
}

int orc_th_set_data(orc_th *p, int ieqn, int auxvar_type, int var_type, int cond_id, const double *data, int n)
{
  thcond *cd = find_cond(p, auxvar_type, cond_id);
  int i;
  (void)ieqn;
  if (!cd || n > cd->nconn) return 1;
  if (var_type == VAR_BC_SS_CONDITION) { for (i = 0; i < n; i++) cd->soe_value[i] = data[i]; return 0; }
  return 2;
}

/* the reference's drivers poke aux_vars_bc%pressure of the energy equation's Dirichlet conditions directly
 * (mass_and_heat_model_problem.F90:616-621) */
int orc_th_set_bc_pressure(orc_th *p, int ieqn, int cond_id, const double *data, int n)
{
  thcond *cd = find_cond(p, AUXVAR_BC, cond_id);
  int i;
  (void)ieqn;
  if (!cd || !cd->eaux_ || n > cd->nconn) return 1;
  for (i = 0; i < n; i++) cd->eaux_[i].r.pressure = data[i];
  return 0;
}

int orc_th_get_data(orc_th *p, int var_type, double *data, int n)
{
  int i;
  if (n > p->ncells) return 1;
  for (i = 0; i < n; i++) {
    switch (var_type) {
    case VAR_PRESSURE: data[i] = p->soln[2 * i]; break;
    case VAR_TEMPERATURE: data[i] = p->soln[2 * i + 1]; break;
    case VAR_LIQ_SAT: data[i] = p->maux[i].sat; break;
    case VAR_MASS: data[i] = p->maux[i].por * p->maux[i].den * ORC_FMWH2O * p->maux[i].sat * p->vol[i]; break;
    default: return 2;
    }
  }
  return 0;
}

/* SavePrimaryIndependentVar + GovEqnExchangeAuxVars + UpdateAuxVars of both governing equations
 * (SOETHResidual, SystemOfEquationsTHType.F90:776-801) */
static void update_auxvars(orc_th *p, int c0, int c1, const double *X)
{
  int ic, k, nlev = p->nlev;
  for (ic = c0 * nlev; ic < c1 * nlev; ic++) {
    double P = X[2 * ic], T = X[2 * ic + 1];
    p->maux[ic].pressure = P; p->maux[ic].temperature = T;        /* T received from the energy equation */
    p->ea[ic].r.pressure = P; p->ea[ic].r.temperature = T;        /* P received from the mass equation */
    orc_rich_auxvar_compute(&p->maux[ic]);
    eaux_compute(&p->ea[ic]);
  }
  for (k = 0; k < p->nbc; k++) {
    thcond *cd = &p->bc[k];
    int i;
    for (i = c0; i < c1; i++) {
      if (cd->raux) {        /* RichardsODEPressureUpdateAuxVarsBC :1478-1552 */
        if (cd->itype == COND_DIRICHLET || cd->itype == COND_SEEPAGE_BC) cd->raux[i].pressure = cd->raux[i].condition_value;
        orc_rich_auxvar_compute(&cd->raux[i]);
      } else {               /* ThermEnthalpySoilUpdateAuxVarsBC :959-1008 */
        if (cd->itype == COND_DIRICHLET) cd->eaux_[i].r.temperature = cd->eaux_[i].r.condition_value;
        eaux_compute(&cd->eaux_[i]);
      }
    }
  }
}

static void accum_m(orc_th *p, int c0, int c1, double *f)
{
  int ic, nlev = p->nlev; double dtInv = 1.0 / p->dtime;
  for (ic = c0 * nlev; ic < c1 * nlev; ic++) f[ic] = p->maux[ic].por * p->maux[ic].den * p->maux[ic].sat * p->vol[ic] * dtInv;
}
/* ThermalEnthalpySoilAccum, GoveqnThermalEnthalpySoilType.F90:1174-1219 */
static void accum_e(orc_th *p, int c0, int c1, double *f)
{
  int ic, nlev = p->nlev; double dtInv = 1.0 / p->dtime;
  for (ic = c0 * nlev; ic < c1 * nlev; ic++) {
    const eaux *a = &p->ea[ic];
    f[ic] = 0.0 + (a->r.por * a->r.den * a->r.sat * a->ul + (1.0 - a->r.por) * a->den_soil * a->heat_cap_soil * (a->r.temperature - 273.15)) * p->vol[ic] * dtInv;
  }
}

/* F is interleaved (F_P, F_T) per cell */
static void residual_range(orc_th *p, int c0, int c1, const double *X, double *F)
{
  int c, j, k, ic, nlev = p->nlev;
  double *fm = (double *)malloc(sizeof(double) * 2 * (size_t)p->ncells), *fe = fm + p->ncells;
  double flux, d1, d2, mflux, eflux;
  update_auxvars(p, c0, c1, X);
  accum_m(p, c0, c1, fm); accum_e(p, c0, c1, fe);
  for (ic = c0 * nlev; ic < c1 * nlev; ic++) { fm[ic] = fm[ic] - p->accum_prev_m[ic]; fe[ic] = fe[ic] - p->accum_prev_e[ic]; }
  for (c = c0; c < c1; c++) for (j = 0; j < nlev - 1; j++) {
    const orc_conn *cn = &p->conn_in[c * (nlev - 1) + j];
    /* mass equation: RichardsODEPressureDivergence on the mass aux vars */
    orc_richards_flux(&p->maux[cn->id_up], &p->maux[cn->id_dn], cn, 0, 1, 0, COND_NULL, &flux, &d1, &d2);
    fm[cn->id_up] = fm[cn->id_up] - flux; fm[cn->id_dn] = fm[cn->id_dn] + flux;
    /* energy equation: ThermalEnthalpySoilDivergence (:1299-1497) with the mass flux of the ENERGY aux vars */
    orc_richards_flux(&p->ea[cn->id_up].r, &p->ea[cn->id_dn].r, cn, 0, 1, 0, COND_NULL, &mflux, &d1, &d2);
    energy_flux(&p->ea[cn->id_up], &p->ea[cn->id_dn], mflux, 0.0, 0.0, cn, 1, COND_NULL, 0, &eflux, &d1, &d2);
    fe[cn->id_up] = fe[cn->id_up] - eflux; fe[cn->id_dn] = fe[cn->id_dn] + eflux;
  }
  for (k = 0; k < p->nbc; k++) {
    thcond *cd = &p->bc[k];
    int i;
    for (i = c0; i < c1; i++) {
      int cell = cd->conn[i].id_dn;
      if (cd->raux) {
        orc_richards_flux(&cd->raux[i], &p->maux[cell], &cd->conn[i], 0, 0, 0, cd->itype, &flux, &d1, &d2);
        fm[cell] = fm[cell] + flux;
      } else if (cd->itype == COND_DIRICHLET) {
        orc_richards_flux(&cd->eaux_[i].r, &p->ea[cell].r, &cd->conn[i], 0, 0, 0, cd->itype, &mflux, &d1, &d2);
        energy_flux(&cd->eaux_[i], &p->ea[cell], mflux, 0.0, 0.0, &cd->conn[i], 0, cd->itype, 0, &eflux, &d1, &d2);
        fe[cell] = fe[cell] + eflux;
      }
    }
  }
  for (k = 0; k < p->nss; k++) {
    thcond *cd = &p->ss[k];
    int i0 = cd->per_cell ? c0 * nlev : c0, i1 = cd->per_cell ? c1 * nlev : c1, i;
    for (i = i0; i < i1; i++) {
      int cell = cd->conn[i].id_dn;
      if (cd->ieqn == 1 && cd->itype == COND_MASS_RATE) fm[cell] = fm[cell] - cd->value[i] / ORC_FMWH2O;
      else if (cd->ieqn == 2 && cd->itype == COND_HEAT_RATE) fe[cell] = fe[cell] + cd->value[i];    /* sign as :1478 */
    }
  }
  for (ic = c0 * nlev; ic < c1 * nlev; ic++) { F[2 * ic] = fm[ic]; F[2 * ic + 1] = fe[ic]; }
  free(fm);
}

/* 2x2 blocks, row-major [dFP/dP dFP/dT ; dFT/dP dFT/dT]; a = sub (cell-1), b = diag, c = super (cell+1) */
#define BLK(arr, cell, r, cc) arr[4 * (size_t)(cell) + 2 * (r) + (cc)]
static void jacobian_range(orc_th *p, int c0, int c1, double *ja, double *jb, double *jc)
{
  int c, j, k, ic, nlev = p->nlev;
  double dtInv = 1.0 / p->dtime, dummy, Jup, Jdn, mflux, dm_up, dm_dn;
  for (ic = c0 * nlev; ic < c1 * nlev; ic++) for (k = 0; k < 4; k++) { ja[4 * (size_t)ic + k] = 0.0; jb[4 * (size_t)ic + k] = 0.0; jc[4 * (size_t)ic + k] = 0.0; }
  for (c = c0; c < c1; c++) for (j = 0; j < nlev - 1; j++) {
    const orc_conn *cn = &p->conn_in[c * (nlev - 1) + j];
    int up = cn->id_up, dn = cn->id_dn;
    /* dFP/dP: RichardsODEPressureDivergenceDeriv (GoveqnRichards...:2054-2069) */
    orc_richards_flux(&p->maux[up], &p->maux[dn], cn, 1, 1, 0, COND_NULL, &dummy, &Jup, &Jdn);
    BLK(jb, up, 0, 0) += Jup; BLK(jc, up, 0, 0) += Jdn; BLK(ja, dn, 0, 0) += -Jup; BLK(jb, dn, 0, 0) += -Jdn;
    /* dFP/dT: OffDiagJacobian_Temperature_ForInternalAuxVars (GoveqnRichards...:2560-2605) */
    orc_richards_flux_dT(&p->maux[up], &p->maux[dn], cn, 1, COND_NULL, &dummy, &Jup, &Jdn);
    BLK(jb, up, 0, 1) += -Jup; BLK(jc, up, 0, 1) += -Jdn; BLK(ja, dn, 0, 1) += Jup; BLK(jb, dn, 0, 1) += Jdn;
    /* dFT/dT: ThermalEnthalpySoilDivergenceDeriv (:1560-1620) */
    orc_richards_flux_dT(&p->ea[up].r, &p->ea[dn].r, cn, 1, COND_NULL, &mflux, &dm_up, &dm_dn);
    energy_flux(&p->ea[up], &p->ea[dn], mflux, dm_up, dm_dn, cn, 1, COND_NULL, 0, &dummy, &Jup, &Jdn);
    BLK(jb, up, 1, 1) += -Jup; BLK(jc, up, 1, 1) += -Jdn; BLK(ja, dn, 1, 1) += Jup; BLK(jb, dn, 1, 1) += Jdn;
    /* dFT/dP: OffDiagJacobian_Pressure_ForInternalAuxVars (:2180-2230) */
    orc_richards_flux(&p->ea[up].r, &p->ea[dn].r, cn, 1, 1, 0, COND_NULL, &mflux, &dm_up, &dm_dn);
    dm_up = -dm_up; dm_dn = -dm_dn;
    energy_flux(&p->ea[up], &p->ea[dn], mflux, dm_up, dm_dn, cn, 1, COND_NULL, 1, &dummy, &Jup, &Jdn);
    BLK(jb, up, 1, 0) += -Jup; BLK(jc, up, 1, 0) += -Jdn; BLK(ja, dn, 1, 0) += Jup; BLK(jb, dn, 1, 0) += Jdn;
  }
  for (k = 0; k < p->nbc; k++) {
    thcond *cd = &p->bc[k];
    int i;
    for (i = c0; i < c1; i++) {
      int cell = cd->conn[i].id_dn;
      if (cd->raux) {
        orc_richards_flux(&cd->raux[i], &p->maux[cell], &cd->conn[i], 1, 0, 0, cd->itype, &dummy, &Jup, &Jdn);
        BLK(jb, cell, 0, 0) += -Jdn;
      } else if (cd->itype == COND_DIRICHLET) {
        orc_richards_flux_dT(&cd->eaux_[i].r, &p->ea[cell].r, &cd->conn[i], 0, cd->itype, &mflux, &dm_up, &dm_dn);
        energy_flux(&cd->eaux_[i], &p->ea[cell], mflux, dm_up, dm_dn, &cd->conn[i], 0, cd->itype, 0, &dummy, &Jup, &Jdn);
        BLK(jb, cell, 1, 1) += Jdn;
        orc_richards_flux(&cd->eaux_[i].r, &p->ea[cell].r, &cd->conn[i], 1, 0, 0, cd->itype, &mflux, &dm_up, &dm_dn);
        dm_up = -dm_up; dm_dn = -dm_dn;
        energy_flux(&cd->eaux_[i], &p->ea[cell], mflux, dm_up, dm_dn, &cd->conn[i], 0, cd->itype, 1, &dummy, &Jup, &Jdn);
        BLK(jb, cell, 1, 0) += Jdn;
      }
    }
  }
  for (ic = c0 * nlev; ic < c1 * nlev; ic++) {
    const orc_rich_auxvar *m = &p->maux[ic]; const eaux *e = &p->ea[ic];
    double V = p->vol[ic], der;
    BLK(jb, ic, 0, 0) += (m->dpor_dP * m->den * m->sat + m->por * m->dden_dP * m->sat + m->por * m->den * m->dsat_dP) * V * dtInv;
    BLK(jb, ic, 0, 1) += (0.0 * m->den * m->sat + m->por * m->dden_dT * m->sat + m->por * m->den * 0.0) * V * dtInv;
    der = e->r.por * e->r.dden_dT * e->r.sat * e->ul + e->r.por * e->r.den * e->dsat_dT * e->ul + e->r.por * e->r.den * e->r.sat * e->dul_dT;
    der = (der + (1.0 - e->r.por) * e->den_soil * e->heat_cap_soil) * V * dtInv;
    BLK(jb, ic, 1, 1) += der;
    der = e->r.dpor_dP * e->r.den * e->r.sat * e->ul + e->r.por * e->r.dden_dP * e->r.sat * e->ul + e->r.por * e->r.den * e->r.dsat_dP * e->ul
        + e->r.por * e->r.den * e->r.sat * e->dul_dP + (-e->r.dpor_dP * e->den_soil * e->heat_cap_soil * (e->r.temperature - 273.15));
    BLK(jb, ic, 1, 0) += der * V * dtInv;
  }
}

typedef struct { orc_th *p; int c0, c1; double *ja, *jb, *jc, *Ffull; } thctx;

static void cb_residual(void *v, const double *x, double *f)
{
  thctx *r = (thctx *)v; orc_th *p = r->p;
  size_t off = 2 * (size_t)r->c0 * p->nlev, n = 2 * (size_t)(r->c1 - r->c0) * p->nlev;
  memcpy(p->soln + off, x, sizeof(double) * n);
  residual_range(p, r->c0, r->c1, p->soln, r->Ffull);
  memcpy(f, r->Ffull + off, sizeof(double) * n);
}
static void cb_jacobian(void *v, const double *x, double *a, double *b, double *c)
{
  thctx *r = (thctx *)v; orc_th *p = r->p;
  size_t off = 4 * (size_t)r->c0 * p->nlev, n = 4 * (size_t)(r->c1 - r->c0) * p->nlev;
  (void)x;
  jacobian_range(p, r->c0, r->c1, r->ja, r->jb, r->jc);
  memcpy(a, r->ja + off, sizeof(double) * n); memcpy(b, r->jb + off, sizeof(double) * n); memcpy(c, r->jc + off, sizeof(double) * n);
}

/* SOETHPreSolve, SystemOfEquationsTHType.F90:119-272 */
static void pre_solve(orc_th *p, int c0, int c1)
{
  int k, nlev = p->nlev;
  for (k = 0; k < p->nbc; k++) {
    thcond *cd = &p->bc[k]; int i;
    for (i = c0; i < c1; i++) { if (cd->raux) cd->raux[i].condition_value = cd->soe_value[i]; else cd->eaux_[i].r.condition_value = cd->soe_value[i]; }
  }
  for (k = 0; k < p->nss; k++) {
    thcond *cd = &p->ss[k];
    int i0 = cd->per_cell ? c0 * nlev : c0, i1 = cd->per_cell ? c1 * nlev : c1, i;
    for (i = i0; i < i1; i++) cd->value[i] = cd->soe_value[i];
  }
  update_auxvars(p, c0, c1, p->soln_prev);
  accum_m(p, c0, c1, p->accum_prev_m); accum_e(p, c0, c1, p->accum_prev_e);
}

static void step_dt_range(orc_th *p, int c0, int c1, double dt, int *converged, int *reason, double *W,
                          int *tot_its, int *tot_nf, int *ncuts)
{
  const int max_num_time_cuts = 20;
  int num_time_cuts = 0, nlev = p->nlev, ncell = (c1 - c0) * nlev, i;
  size_t off = 2 * (size_t)c0 * nlev, n = 2 * (size_t)ncell;
  double time = 0.0, dt_iter = dt;
  size_t N = (size_t)p->ncells;
  orc_system sys; thctx r; orc_snes_result res;
  double *x = (double *)malloc(sizeof(double) * n);
  int *cs = (int *)malloc(sizeof(int) * (size_t)(c1 - c0 + 1));
  for (i = 0; i <= c1 - c0; i++) cs[i] = i * nlev;
  r.p = p; r.c0 = c0; r.c1 = c1; r.Ffull = W; r.ja = W + 2 * N; r.jb = W + 6 * N; r.jc = W + 10 * N;
  sys.n = (int)n; sys.bs = 2; sys.ncell = ncell; sys.col_start = cs; sys.nchain = c1 - c0;
  sys.residual = cb_residual; sys.jacobian = cb_jacobian; sys.ctx = &r;
  *converged = 0; *tot_its = 0; *tot_nf = 0;
  for (;;) {
    p->dtime = dt_iter;
    pre_solve(p, c0, c1);
    memcpy(x, p->soln + off, sizeof(double) * n);
    orc_snes_solve(&sys, &p->opts, x, &res);
    memcpy(p->soln + off, x, sizeof(double) * n);
    *reason = res.reason; *tot_nf += res.nfuncs;
    if (res.reason < 0) {
      num_time_cuts++; dt_iter = 0.5 * dt_iter;
      memcpy(p->soln + off, p->soln_prev + off, sizeof(double) * n);
    } else {
      *converged = 1; time += dt_iter; *tot_its += res.its;
      memcpy(p->soln_prev + off, p->soln + off, sizeof(double) * n);       /* SOETHPostSolve :275-302 */
      update_auxvars(p, c0, c1, p->soln);
    }
    if (num_time_cuts > max_num_time_cuts) { *converged = 0; break; }
    if (time >= dt) break;
  }
  *ncuts = num_time_cuts;
  free(x); free(cs);
}

int orc_th_step_dt(orc_th *p, double dt, int nstep, int *converged, int *converged_reason)
{
  int ncol = p->ncol, c;
  size_t N = (size_t)p->ncells;
  double *W = (double *)malloc(sizeof(double) * 14 * N);
  (void)nstep;
  if (!p->per_column) {
    int its, nf, cuts;
    step_dt_range(p, 0, ncol, dt, converged, converged_reason, W, &its, &nf, &cuts);
    for (c = 0; c < ncol; c++) { p->stat_its[c] = its; p->stat_reason[c] = *converged_reason; p->stat_cuts[c] = cuts; p->stat_nf[c] = nf; }
  } else {
    int all = 1, worst;
#ifdef _OPENMP
#pragma omp parallel for schedule(dynamic, 16) num_threads(p->nthreads)
#endif
    for (c = 0; c < ncol; c++) {
      orc_th q = *p;
      int conv, reason, its, nf, cuts;
      step_dt_range(&q, c, c + 1, dt, &conv, &reason, W, &its, &nf, &cuts);
      p->stat_its[c] = its; p->stat_reason[c] = reason; p->stat_cuts[c] = cuts; p->stat_nf[c] = nf;
      if (!conv) {
#ifdef _OPENMP
#pragma omp critical
#endif
        all = 0;
      }
    }
    worst = p->stat_reason[0];
    for (c = 1; c < ncol; c++) if (p->stat_reason[c] < worst) worst = p->stat_reason[c];
    *converged = all; *converged_reason = worst;
  }
  p->dtime = dt;
  free(W);
  return 0;
}

void orc_th_get_stats(orc_th *p, int *newton_its, int *reasons, int *ncuts, int *nfuncs)
{
  int c;
  for (c = 0; c < p->ncol; c++) {
    if (newton_its) newton_its[c] = p->stat_its[c];
    if (reasons) reasons[c] = p->stat_reason[c];
    if (ncuts) ncuts[c] = p->stat_cuts[c];
    if (nfuncs) nfuncs[c] = p->stat_nf[c];
  }
}

/* residual + Jacobian blocks at x (interleaved), accumulation taken at x_prev: for kernel unit tests and for
 * checking the analytic blocks against finite differences */
void orc_th_eval(orc_th *p, double dt, const double *xprev, const double *x, double *f, double *ja, double *jb, double *jc)
{
  size_t nb = sizeof(double) * 2 * (size_t)p->ncells;
  double *keep = (double *)malloc(nb), *xs = (double *)malloc(nb);
  memcpy(keep, p->soln_prev, nb);
  memcpy(p->soln_prev, xprev, nb);
  p->dtime = dt;
  pre_solve(p, 0, p->ncol);
  memcpy(xs, x, nb);
  residual_range(p, 0, p->ncol, xs, f);
  jacobian_range(p, 0, p->ncol, ja, jb, jc);
  memcpy(p->soln_prev, keep, nb);
  free(keep); free(xs);
}
