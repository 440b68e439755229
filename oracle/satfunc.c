/*
 * satfunc.c -- oracle restatement of src/mpp/util/SaturationFunction.F90
 * (van Genuchten-Mualem, Brooks-Corey, smoothed Brooks-Corey bz2/bz3).
 * TEST INFRASTRUCTURE ONLY (see mpp_oracle.h).
 */
#include <math.h>
#include <string.h>
#include "mpp_oracle.h"

static void satparams_init(orc_satparams *sp) { memset(sp, 0, sizeof(*sp)); }   /* :97-123 */

/* SaturationFunction.F90:127-159 */
int orc_satfunc_set_vg(orc_satparams *sp, double sat_res, double alpha, double vg_m)
{
  if (sat_res < 0.0 || sat_res > 0.5 || alpha <= 0.0 || alpha > 2.0 || vg_m <= 0.0 || vg_m >= 1.0) return 1;
  satparams_init(sp);
  sp->sat_func_type     = SAT_FUNC_VAN_GENUCHTEN;
  sp->sat_res           = sat_res;
  sp->alpha             = alpha;
  sp->vg_m              = vg_m;
  sp->relperm_func_type = RELPERM_FUNC_MUALEM;
  sp->vg_n              = 1.0 / (1.0 - vg_m);
  return 0;
}

/* SaturationFunction.F90:163-192 */
int orc_satfunc_set_bc(orc_satparams *sp, double sat_res, double alpha, double lambda)
{
  if (sat_res < 0.0 || sat_res > 0.5 || alpha <= 0.0 || alpha > 2.0 || lambda <= 0.0 || lambda >= 2.0) return 1;
  satparams_init(sp);
  sp->sat_func_type     = SAT_FUNC_BROOKS_COREY;
  sp->relperm_func_type = RELPERM_FUNC_MUALEM;
  sp->sat_res           = sat_res;
  sp->alpha             = alpha;
  sp->bc_lambda         = lambda;
  return 0;
}

/* SaturationFunction.F90:425-518: bracketed Newton-Raphson for gu */
double orc_findgu_sbc_zerocoeff(double lambda, int AA, double gs)
{
  const double relTol = 1.e-12;
  double guLeft, gu, guRight, deltaGu, resid, dr_dGu, guInv, guToMinusLam, gsOnGu;

  gu = pow((double)AA / ((double)AA + lambda), -1.0 / lambda);

  if (gs > 0.0) {
    guLeft  = 1.0;
    guRight = gu;
    for (;;) {
      if (gu <= guLeft || gu >= guRight) gu = guLeft + 0.5 * (guRight - guLeft);

      guInv        = 1.0 / gu;
      guToMinusLam = pow(gu, -lambda);
      gsOnGu       = gs * guInv;
      resid        = AA - guToMinusLam * (AA + lambda - lambda * gsOnGu);

      if (resid < 0.0) guLeft = gu;
      else             guRight = gu;

      dr_dGu  = (1.0 + lambda) * (1.0 - gsOnGu) + (AA - 1);
      dr_dGu  = lambda * guToMinusLam * guInv * dr_dGu;
      deltaGu = resid / dr_dGu;
      gu      = gu - deltaGu;

      if (fabs(deltaGu) < relTol * fabs(gu)) break;
    }
  }
  return gu;
}

/* SaturationFunction.F90:260-315 */
int orc_satfunc_set_sbc_bz2(orc_satparams *sp, double sat_res, double alpha, double lambda, double ps)
{
  double pu, bcAtPu, lambdaDeltaPuOnPu, oneOnDeltaPu;
  if (sat_res < 0.0 || sat_res > 0.5 || alpha <= 0.0 || alpha > 2.0 || lambda <= 0.0 || lambda >= 2.0
      || ps <= -1.0 / alpha || ps > 0.0) return 1;
  satparams_init(sp);
  sp->sat_func_type     = SAT_FUNC_SMOOTHED_BROOKS_COREY;
  sp->relperm_func_type = RELPERM_FUNC_MUALEM;
  sp->sat_res           = sat_res;
  sp->alpha             = alpha;
  sp->bc_lambda         = lambda;
  sp->sbc_ps            = ps;

  pu         = orc_findgu_sbc_zerocoeff(lambda, 3, -alpha * ps) / (-alpha);
  sp->sbc_pu = pu;

  bcAtPu            = pow(-alpha * pu, -lambda);
  lambdaDeltaPuOnPu = lambda * (1.0 - ps / pu);
  oneOnDeltaPu      = 1.0 / (pu - ps);

  sp->sbc_b2 = 0.0;
  sp->sbc_b3 = (2.0 - bcAtPu * (2.0 + lambdaDeltaPuOnPu)) * oneOnDeltaPu * oneOnDeltaPu * oneOnDeltaPu;
  if (sp->sbc_b3 <= 0.0) return 2;
  return 0;
}

/* SaturationFunction.F90:319-372 */
int orc_satfunc_set_sbc_bz3(orc_satparams *sp, double sat_res, double alpha, double lambda, double ps)
{
  double pu, bcAtPu, lambdaDeltaPuOnPu, oneOnDeltaPu;
  if (sat_res < 0.0 || sat_res > 0.5 || alpha <= 0.0 || alpha > 2.0 || lambda <= 0.0 || lambda >= 2.0
      || ps <= -1.0 / alpha || ps > 0.0) return 1;
  satparams_init(sp);   /* the reference skips Init() here; all fields it leaves untouched are unused */
  sp->sat_func_type     = SAT_FUNC_SMOOTHED_BROOKS_COREY;
  sp->relperm_func_type = RELPERM_FUNC_MUALEM;
  sp->sat_res           = sat_res;
  sp->alpha             = alpha;
  sp->bc_lambda         = lambda;
  sp->sbc_ps            = ps;

  pu         = orc_findgu_sbc_zerocoeff(lambda, 2, -alpha * ps) / (-alpha);
  sp->sbc_pu = pu;

  bcAtPu            = pow(-alpha * pu, -lambda);
  lambdaDeltaPuOnPu = lambda * (1.0 - ps / pu);
  oneOnDeltaPu      = 1.0 / (pu - ps);

  sp->sbc_b2 = -(3.0 - bcAtPu * (3.0 + lambdaDeltaPuOnPu)) * oneOnDeltaPu * oneOnDeltaPu;
  if (sp->sbc_b2 >= 0.0) return 2;
  sp->sbc_b3 = 0.0;
  return 0;
}

/* SaturationFunction.F90:747-795 */
static void pc_to_sat_vg(const orc_satparams *sp, double pc, double *sat, double *dsat_dP)
{
  double sat_res = sp->sat_res, alpha = sp->alpha, mm = sp->vg_m, nn = sp->vg_n;
  if (pc < 0.0) {
    double pc_alpha_n          = pow(-alpha * pc, nn);
    double one_plus_pc_alpha_n = 1.0 + pc_alpha_n;
    double Se                  = pow(one_plus_pc_alpha_n, -mm);
    double AA, dSe_dpc;
    *sat     = sat_res + (1.0 - sat_res) * Se;
    AA       = pc_alpha_n / one_plus_pc_alpha_n;
    dSe_dpc  = -mm * nn * Se * AA / pc;
    *dsat_dP = (1.0 - sat_res) * dSe_dpc;
  } else {
    *sat     = 1.0;
    *dsat_dP = 0.0;
  }
}

/* SaturationFunction.F90:799-857 */
static void pc_to_relperm_vg(const orc_satparams *sp, double pc, double *kr, double *dkr_dP)
{
  double alpha = sp->alpha, mm = sp->vg_m, nn = sp->vg_n;
  if (pc < 0.0) {
    double pc_alpha_n          = pow(-alpha * pc, nn);
    double one_plus_pc_alpha_n = 1.0 + pc_alpha_n;
    double Se                  = pow(one_plus_pc_alpha_n, -mm);
    double AA                  = pc_alpha_n / one_plus_pc_alpha_n;
    double dSe_dpc             = -mm * nn * Se * AA / pc;
    double BB                  = 1.0 - pow(AA, mm);
    double dkr_dSe;
    *kr     = sqrt(Se) * BB * BB;
    dkr_dSe = 0.5 * (*kr) / Se + 2.0 * pow(Se, 1.0 / mm - 0.5) * pow(AA, mm - 1.0) * BB;
    *dkr_dP = dkr_dSe * dSe_dpc;
  } else {
    *kr     = 1.0;
    *dkr_dP = 0.0;
  }
}

/* SaturationFunction.F90:900-938 */
static void pc_to_sat_bc(const orc_satparams *sp, double pc, double *sat, double *dsat_dP)
{
  double sat_res = sp->sat_res, alpha = sp->alpha, lambda = sp->bc_lambda;
  double pc_alpha = -alpha * pc;
  if (pc_alpha > 1.0) {
    double Se      = pow(pc_alpha, -lambda);
    double dSe_dpc = -lambda * Se / pc;
    *sat     = sat_res + (1.0 - sat_res) * Se;
    *dsat_dP = (1.0 - sat_res) * dSe_dpc;
  } else {
    *sat     = 1.0;
    *dsat_dP = 0.0;
  }
}

/* SaturationFunction.F90:942-990 */
static void pc_to_relperm_bc(const orc_satparams *sp, double pc, double frac_liq, double *kr, double *dkr_dP)
{
  double alpha = sp->alpha, lambda = sp->bc_lambda;
  double pc_alpha = -alpha * pc;
  if (pc_alpha > 1.0) {
    double Se      = pow(pc_alpha, -lambda);
    double dSe_dpc = -lambda * Se / pc;
    double dkr_dSe;
    *kr     = pow(Se, 2.5 + 2.0 / lambda);
    dkr_dSe = (2.5 + 2.0 / lambda) * (*kr) / Se;
    *dkr_dP = dkr_dSe * dSe_dpc;
  } else {
    *kr     = 1.0;
    *dkr_dP = 0.0;
  }
  *kr     = frac_liq * (*kr);
  *dkr_dP = frac_liq * (*dkr_dP);
}

/* SaturationFunction.F90:1027-1076 */
static void pc_to_sat_sbc(const orc_satparams *sp, double pc, double *sat, double *dsat_dP)
{
  double sat_res = sp->sat_res, alpha = sp->alpha, lambda = sp->bc_lambda;
  if (pc <= sp->sbc_pu) {
    double Se      = pow(-alpha * pc, -lambda);
    double dSe_dpc = -lambda * Se / pc;
    *sat     = sat_res + (1.0 - sat_res) * Se;
    *dsat_dP = (1.0 - sat_res) * dSe_dpc;
  } else if (pc < sp->sbc_ps) {
    double deltaPc = pc - sp->sbc_ps;
    double Se      = 1.0 + deltaPc * deltaPc * (sp->sbc_b2 + deltaPc * sp->sbc_b3);
    double dSe_dpc = deltaPc * (2 * sp->sbc_b2 + 3 * deltaPc * sp->sbc_b3);
    *sat     = sat_res + (1.0 - sat_res) * Se;
    *dsat_dP = (1.0 - sat_res) * dSe_dpc;
  } else {
    *sat     = 1.0;
    *dsat_dP = 0.0;
  }
}

/* SaturationFunction.F90:1080-1140 */
static void pc_to_relperm_sbc(const orc_satparams *sp, double pc, double *kr, double *dkr_dP)
{
  double alpha = sp->alpha, lambda = sp->bc_lambda;
  if (pc <= sp->sbc_pu) {
    double Se      = pow(-alpha * pc, -lambda);
    double dSe_dpc = -lambda * Se / pc;
    double dkr_dSe;
    *kr     = pow(Se, 2.5 + 2.0 / lambda);
    dkr_dSe = (2.5 + 2.0 / lambda) * (*kr) / Se;
    *dkr_dP = dkr_dSe * dSe_dpc;
  } else if (pc < sp->sbc_ps) {
    double deltaPc = pc - sp->sbc_ps;
    double Se      = 1.0 + deltaPc * deltaPc * (sp->sbc_b2 + deltaPc * sp->sbc_b3);
    double dSe_dpc = deltaPc * (2 * sp->sbc_b2 + 3 * deltaPc * sp->sbc_b3);
    double dkr_dSe;
    *kr     = pow(Se, 2.5 + 2.0 / lambda);
    dkr_dSe = (2.5 + 2.0 / lambda) * (*kr) / Se;
    *dkr_dP = dkr_dSe * dSe_dpc;
  } else {
    *kr     = 1.0;
    *dkr_dP = 0.0;
  }
}

/* SaturationFunction.F90:564-600 */
void orc_press_to_sat(const orc_satparams *sp, double press, double *sat, double *dsat_dP)
{
  double pc = press - ORC_PRESSURE_REF;
  switch (sp->sat_func_type) {
  case SAT_FUNC_VAN_GENUCHTEN:         pc_to_sat_vg(sp, pc, sat, dsat_dP); break;
  case SAT_FUNC_BROOKS_COREY:          pc_to_sat_bc(sp, pc, sat, dsat_dP); break;
  case SAT_FUNC_SMOOTHED_BROOKS_COREY: pc_to_sat_sbc(sp, pc, sat, dsat_dP); break;
  default: *sat = *dsat_dP = NAN;
  }
}

/* SaturationFunction.F90:604-650 (Mualem branch only; Weibull/Campbell are plant-xylem curves, out of scope) */
void orc_press_to_relperm(const orc_satparams *sp, double press, double frac_liq, double *kr, double *dkr_dP)
{
  double pc = press - ORC_PRESSURE_REF;
  switch (sp->sat_func_type) {
  case SAT_FUNC_VAN_GENUCHTEN:         pc_to_relperm_vg(sp, pc, kr, dkr_dP); break;
  case SAT_FUNC_BROOKS_COREY:          pc_to_relperm_bc(sp, pc, frac_liq, kr, dkr_dP); break;
  case SAT_FUNC_SMOOTHED_BROOKS_COREY: pc_to_relperm_sbc(sp, pc, kr, dkr_dP); break;
  default: *kr = *dkr_dP = NAN;
  }
}
