/*
 * richards.c -- oracle restatement of
 *   src/mpp/auxvar/RichardsODEPressureAuxType.F90 (aux-var init/compute) and
 *   src/mpp/ge/RichardsMod.F90 (two-point Darcy flux and its pressure derivatives).
 * TEST INFRASTRUCTURE ONLY (see mpp_oracle.h).
 */
#include <math.h>
#include <string.h>
#include "mpp_oracle.h"

/* RichardsODEPressureAuxType.F90:77-122 */
void orc_rich_auxvar_init(orc_rich_auxvar *a)
{
  memset(a, 0, sizeof(*a));
  a->pressure     = 0.0;
  a->temperature  = 273.15 + 25.0;     /* :92 -- never overwritten by the VSFM SoE (GoveqnRichards...:573-575) */
  a->frac_liq_sat = 1.0;
  a->density_type = DENSITY_CONSTANT;
}

/* RichardsODEPressureAuxType.F90:237-294 */
void orc_rich_auxvar_compute(orc_rich_auxvar *a)
{
  orc_press_to_sat(&a->satParams, a->pressure, &a->sat, &a->dsat_dP);
  orc_press_to_relperm(&a->satParams, a->pressure, a->frac_liq_sat, &a->kr, &a->dkr_dP);
  orc_density(a->pressure, a->temperature, a->density_type, &a->den, &a->dden_dP, &a->dden_dT);
  orc_viscosity(a->pressure, a->temperature, &a->vis, &a->dvis_dP, &a->dvis_dT);
  /* PorosityFunctionMod.F90:125-141 constant model */
  a->por     = a->por_base;
  a->dpor_dP = 0.0;
}

/* RichardsMod.F90:118-340 RichardsFlux_Internal */
static void richards_flux_internal(const orc_rich_auxvar *aux_var_up, const orc_rich_auxvar *aux_var_dn,
                                   const orc_conn *conn, int compute_deriv, int internal_conn, int cond_type,
                                   double *flux, double *dflux_dP_up, double *dflux_dP_dn)
{
  double area = conn->area, dist_up = conn->dist_up, dist_dn = conn->dist_dn;
  const double *dist_unitvec = conn->unitvec;

  double Pres_up = aux_var_up->pressure, kr_up = aux_var_up->kr, dkr_dP_up = aux_var_up->dkr_dP;
  double den_up = aux_var_up->den, dden_dP_up = aux_var_up->dden_dP, vis_up = aux_var_up->vis, dvis_dP_up = aux_var_up->dvis_dP;
  double Pres_dn = aux_var_dn->pressure, kr_dn = aux_var_dn->kr, dkr_dP_dn = aux_var_dn->dkr_dP;
  double den_dn = aux_var_dn->den, dden_dP_dn = aux_var_dn->dden_dP, vis_dn = aux_var_dn->vis, dvis_dP_dn = aux_var_dn->dvis_dP;

  double perm_up, perm_dn, upweight, Dq, grav_vec[3], udist_dot_ugrav, dist_gravity, den_ave, gravityterm, dphi;
  double ukvr, v_darcy, q;
  int seepage_bc_update;

  perm_up = fabs(dist_unitvec[0]) * aux_var_up->perm[0] + fabs(dist_unitvec[1]) * aux_var_up->perm[1] +
            fabs(dist_unitvec[2]) * aux_var_up->perm[2];
  perm_dn = fabs(dist_unitvec[0]) * aux_var_dn->perm[0] + fabs(dist_unitvec[1]) * aux_var_dn->perm[1] +
            fabs(dist_unitvec[2]) * aux_var_dn->perm[2];

  if (internal_conn) {
    upweight = dist_up / (dist_up + dist_dn);
    Dq       = (perm_up * perm_dn) / (dist_up * perm_dn + dist_dn * perm_up);
  } else {
    switch (cond_type) {
    case COND_DIRICHLET: case COND_MASS_FLUX: case COND_SEEPAGE_BC:
      upweight = 0.0;
      Dq       = perm_dn / (dist_up + dist_dn);
      break;
    case COND_DIRICHLET_FRM_OTR_GOVEQ:
      upweight = dist_up / (dist_up + dist_dn);
      Dq       = (perm_up * perm_dn) / (dist_up * perm_dn + dist_dn * perm_up);
      break;
    default:
      *flux = NAN; *dflux_dP_up = NAN; *dflux_dP_dn = NAN; return;
    }
  }

  grav_vec[0] = 0.0; grav_vec[1] = 0.0; grav_vec[2] = -ORC_GRAVITY_CONSTANT;
  udist_dot_ugrav = dist_unitvec[0] * grav_vec[0] + dist_unitvec[1] * grav_vec[1] + dist_unitvec[2] * grav_vec[2];

  dist_gravity = (dist_up + dist_dn) * udist_dot_ugrav;
  den_ave      = upweight * den_up + (1.0 - upweight) * den_dn;
  gravityterm  = (upweight * den_up + (1.0 - upweight) * den_dn) * ORC_FMWH2O * dist_gravity;
  dphi         = Pres_up - Pres_dn + gravityterm;

  seepage_bc_update = 0;
  if (!internal_conn && cond_type == COND_SEEPAGE_BC) {
    if (dphi > 0.0 && Pres_up <= ORC_PRESSURE_REF) seepage_bc_update = 1;
  }
  if (seepage_bc_update) dphi = 0.0;

  if (dphi >= 0.0) ukvr = kr_up / vis_up;
  else             ukvr = kr_dn / vis_dn;

  if (!internal_conn && cond_type == COND_MASS_FLUX) v_darcy = 0.0;
  else                                               v_darcy = -Dq * ukvr * dphi;

  q     = v_darcy * area;
  *flux = q * den_ave;

  if (compute_deriv) {
    double dden_ave_dP_up = upweight * dden_dP_up;
    double dden_ave_dP_dn = (1.0 - upweight) * dden_dP_dn;
    double dgravityterm_dden_up = upweight * dist_gravity * ORC_FMWH2O;
    double dgravityterm_dden_dn = (1.0 - upweight) * dist_gravity * ORC_FMWH2O;
    double dphi_dP_up =  1.0 + dgravityterm_dden_up * dden_dP_up;
    double dphi_dP_dn = -1.0 + dgravityterm_dden_dn * dden_dP_dn;
    double dukvr_dP_up, dukvr_dP_dn, dq_dP_up, dq_dP_dn;

    if (seepage_bc_update) dphi_dP_dn = 0.0;

    if (dphi >= 0) {
      dukvr_dP_up = dkr_dP_up / vis_up - kr_up / (vis_up * vis_up) * dvis_dP_up;
      dukvr_dP_dn = 0.0;
    } else {
      dukvr_dP_up = 0.0;
      dukvr_dP_dn = dkr_dP_dn / vis_dn - kr_dn / (vis_dn * vis_dn) * dvis_dP_dn;
    }

    /* NB sign convention: these are -d(q)/dP, so dflux_dP_* = -d(flux)/dP (:326-334) */
    dq_dP_up = Dq * (dukvr_dP_up * dphi + ukvr * dphi_dP_up) * area;
    dq_dP_dn = Dq * (dukvr_dP_dn * dphi + ukvr * dphi_dP_dn) * area;

    if (!internal_conn && cond_type == COND_MASS_FLUX) {
      *dflux_dP_up = 0.0;
      *dflux_dP_dn = 0.0;
    } else {
      *dflux_dP_up = (dq_dP_up * den_ave - q * dden_ave_dP_up);
      *dflux_dP_dn = (dq_dP_dn * den_ave - q * dden_ave_dP_dn);
    }
  }
}

/* RichardsMod.F90:29-114 RichardsFlux (swap_order wrapper) */
void orc_richards_flux(const orc_rich_auxvar *up, const orc_rich_auxvar *dn, const orc_conn *conn,
                       int compute_deriv, int internal_conn, int swap_order, int cond_type,
                       double *flux, double *dflux_dP_up, double *dflux_dP_dn)
{
  double f = 0, df_up = 0, df_dn = 0;
  if (!swap_order) {
    richards_flux_internal(up, dn, conn, compute_deriv, internal_conn, cond_type, &f, &df_up, &df_dn);
    *flux = f; *dflux_dP_up = df_up; *dflux_dP_dn = df_dn;
  } else {
    richards_flux_internal(dn, up, conn, compute_deriv, internal_conn, cond_type, &f, &df_up, &df_dn);
    *flux = -f; *dflux_dP_up = -df_dn; *dflux_dP_dn = -df_up;
  }
}

/* RichardsMod.F90:343-648 RichardsFluxDerivativeWrtTemperature (non-swapped order only: every caller on the
 * column hot path passes swap_order = .false.).  NB the reference flips the sign at the end (:640-641), so unlike
 * orc_richards_flux these are the TRUE derivatives d(flux)/dT. */
void orc_richards_flux_dT(const orc_rich_auxvar *aux_var_up, const orc_rich_auxvar *aux_var_dn, const orc_conn *conn,
                          int internal_conn, int cond_type, double *flux, double *dflux_dT_up, double *dflux_dT_dn)
{
  double area = conn->area, dist_up = conn->dist_up, dist_dn = conn->dist_dn;
  const double *dist_unitvec = conn->unitvec;
  double Pres_up = aux_var_up->pressure, kr_up = aux_var_up->kr, den_up = aux_var_up->den, dden_dT_up = aux_var_up->dden_dT;
  double vis_up = aux_var_up->vis, dvis_dT_up = aux_var_up->dvis_dT;
  double Pres_dn = aux_var_dn->pressure, kr_dn = aux_var_dn->kr, den_dn = aux_var_dn->den, dden_dT_dn = aux_var_dn->dden_dT;
  double vis_dn = aux_var_dn->vis, dvis_dT_dn = aux_var_dn->dvis_dT;
  double perm_up, perm_dn, upweight, Dq, udist_dot_ugrav, dist_gravity, den_ave, gravityterm, dphi, ukvr, v_darcy, q;
  double dden_ave_dT_up, dden_ave_dT_dn, dgravityterm_dden_up, dgravityterm_dden_dn, dphi_dT_up, dphi_dT_dn;
  double dukvr_dT_up, dukvr_dT_dn, dq_dT_up, dq_dT_dn;
  int seepage_bc_update = 0;

  perm_up = fabs(dist_unitvec[0]) * aux_var_up->perm[0] + fabs(dist_unitvec[1]) * aux_var_up->perm[1] + fabs(dist_unitvec[2]) * aux_var_up->perm[2];
  perm_dn = fabs(dist_unitvec[0]) * aux_var_dn->perm[0] + fabs(dist_unitvec[1]) * aux_var_dn->perm[1] + fabs(dist_unitvec[2]) * aux_var_dn->perm[2];
  if (internal_conn || cond_type == COND_DIRICHLET_FRM_OTR_GOVEQ) {
    upweight = dist_up / (dist_up + dist_dn);
    Dq       = (perm_up * perm_dn) / (dist_up * perm_dn + dist_dn * perm_up);
  } else {
    upweight = 0.0;
    Dq       = perm_dn / (dist_up + dist_dn);
  }
  udist_dot_ugrav = dist_unitvec[0] * 0.0 + dist_unitvec[1] * 0.0 + dist_unitvec[2] * (-ORC_GRAVITY_CONSTANT);
  dist_gravity = (dist_up + dist_dn) * udist_dot_ugrav;
  den_ave      = upweight * den_up + (1.0 - upweight) * den_dn;
  gravityterm  = (upweight * den_up + (1.0 - upweight) * den_dn) * ORC_FMWH2O * dist_gravity;
  dphi         = Pres_up - Pres_dn + gravityterm;
  if (!internal_conn && cond_type == COND_SEEPAGE_BC && dphi > 0.0 && Pres_up <= ORC_PRESSURE_REF) seepage_bc_update = 1;
  if (seepage_bc_update) dphi = 0.0;
  if (dphi >= 0.0) ukvr = kr_up / vis_up; else ukvr = kr_dn / vis_dn;
  if (!internal_conn && cond_type == COND_MASS_FLUX) v_darcy = 0.0; else v_darcy = -Dq * ukvr * dphi;
  q = v_darcy * area;
  *flux = q * den_ave;

  dden_ave_dT_up       = upweight * dden_dT_up;
  dden_ave_dT_dn       = (1.0 - upweight) * dden_dT_dn;
  dgravityterm_dden_up = upweight * dist_gravity * ORC_FMWH2O;
  dgravityterm_dden_dn = (1.0 - upweight) * dist_gravity * ORC_FMWH2O;
  dphi_dT_up           = dgravityterm_dden_up * dden_dT_up;
  dphi_dT_dn           = dgravityterm_dden_dn * dden_dT_dn;
  if (seepage_bc_update) dphi_dT_dn = 0.0;
  if (dphi >= 0) { dukvr_dT_up = -kr_up / (vis_up * vis_up) * dvis_dT_up; dukvr_dT_dn = 0.0; }
  else           { dukvr_dT_up = 0.0; dukvr_dT_dn = -kr_dn / (vis_dn * vis_dn) * dvis_dT_dn; }
  dq_dT_up = Dq * (dukvr_dT_up * dphi + ukvr * dphi_dT_up) * area;
  dq_dT_dn = Dq * (dukvr_dT_dn * dphi + ukvr * dphi_dT_dn) * area;
  if (!internal_conn && cond_type == COND_MASS_FLUX) { *dflux_dT_up = 0.0; *dflux_dT_dn = 0.0; }
  else { *dflux_dT_up = (dq_dT_up * den_ave - q * dden_ave_dT_up); *dflux_dT_dn = (dq_dT_dn * den_ave - q * dden_ave_dT_dn); }
  *dflux_dT_up = -(*dflux_dT_up);
  *dflux_dT_dn = -(*dflux_dT_dn);
}
