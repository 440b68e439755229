/*
 * eos_water.c -- oracle restatement of src/mpp/util/EOSWaterMod.F90.
 * TEST INFRASTRUCTURE ONLY (see mpp_oracle.h).
 *
 * Units follow the reference: density in kmol m^-3 (FMWH2O = 18.01534 kg/kmol),
 * enthalpy / internal energy in J kmol^-1.
 */
#include <math.h>
#include "mpp_oracle.h"

/* EOSWaterMod.F90:26-27 */
static const double H2O_CRITICAL_TEMPERATURE = 647.3;
static const double H2O_CRITICAL_PRESSURE    = 22.064e6;

/* IFC-67 coefficient tables, EOSWaterMod.F90:237-257 (identical copy at :407-427) */
static const double aa[23] = {
  6.824687741e03, -5.422063673e02, -2.096666205e04, 3.941286787e04,
 -6.733277739e04,  9.902381028e04, -1.093911774e05, 8.590841667e04,
 -4.511168742e04,  1.418138926e04, -2.017271113e03, 7.982692717e00,
 -2.616571843e-2,  1.522411790e-3,  2.284279054e-2, 2.421647003e02,
  1.269716088e-10, 2.074838328e-7,  2.174020350e-8, 1.105710498e-9,
  1.293441934e01,  1.308119072e-5,  6.047626338e-14 };
static const double a1 = 8.438375405e-1, a2 = 5.362162162e-4, a3 = 1.720000000e00, a4 = 7.342278489e-2,
                    a5 = 4.975858870e-2, a6 = 6.537154300e-1, a7 = 1.150000000e-6, a8 = 1.510800000e-5,
                    a9 = 1.418800000e-1, a10 = 7.002753165e00, a11 = 2.995284926e-4, a12 = 2.040000000e-1;

/* EOSWaterMod.F90:80-99 DensityConstant */
void orc_density_constant(double *den, double *dden_dp, double *dden_dT)
{
  *den     = ORC_DENH2O / ORC_FMWH2O;
  *dden_dp = 0.0;
  *dden_dT = 0.0;
}

/* EOSWaterMod.F90:102-178 DensityTGDPB01 (Tanaka et al. 2001) */
void orc_density_tgdpb01(double p, double t_K, double *den, double *dden_dp, double *dden_dT)
{
  const double b1 = -3.983035, b2 = 301.797, b3 = 522528.9, b4 = 69.34881, b5 = 999.974950;
  const double k0 = 50.74e-11, k1 = -0.326e-11, k2 = 0.00416e-11, p0 = 101325.0;
  double t_c, dent, kappa, ddent_dt, ddent_dt_1, ddent_dt_2, ddent_dt_3, ddent_dp, dkappa_dp, dkappa_dt;

  t_c = t_K - 273.15;

  dent = b5 * (1.0 - (pow(t_c + b1, 2.0)) * (t_c + b2) / b3 / (t_c + b4));

  /* compressibility only above the reference pressure, :151-155 */
  if (p > p0) kappa = (1.0 + (k0 + k1 * t_c + k2 * pow(t_c, 2.0)) * (p - p0));
  else        kappa = 1.0;

  *den = dent * kappa / ORC_FMWH2O;

  ddent_dp   = 0.0;
  ddent_dt_1 = -(pow(t_c + b1, 2.0)) / b3 / (t_c + b4);
  ddent_dt_2 = -2.0 * (t_c + b1) * (t_c + b2) / b3 / (t_c + b4);
  ddent_dt_3 = (pow(t_c + b1, 2.0)) * (t_c + b2) / b3 / (pow(t_c + b4, 2.0));
  ddent_dt   = b5 * (ddent_dt_1 + ddent_dt_2 + ddent_dt_3);

  if (p > p0) {
    dkappa_dp = (k0 + k1 * t_c + k2 * pow(t_c, 2.0));
    dkappa_dt = (k1 + 2.0 * k2 * t_c) * (p - p0);
  } else {
    dkappa_dp = 0.0;
    dkappa_dt = 0.0;
  }

  *dden_dT = (ddent_dt * kappa + dent * dkappa_dt) / ORC_FMWH2O;
  *dden_dp = (ddent_dp * kappa + dent * dkappa_dp) / ORC_FMWH2O;
}

/* EOSWaterMod.F90:181-344 DensityIFC67 (t in Celsius, p in Pa) */
void orc_density_ifc67(double t, double p, int calculate_derivatives, double *dw, double *dwmol, double *dwp, double *dwt)
{
  double beta, beta2x, theta, theta2x, theta18, theta20;
  double xx, yy, zz, u0, u1, u2, u3, u4, u5, u6, u7, u8, u9;
  double vr, ypt, zpt, zpp, vrpt, vrpp, cnv;
  double tc1, pc1, vc1, utc1, upc1, vc1mol;

  tc1    = H2O_CRITICAL_TEMPERATURE;
  pc1    = H2O_CRITICAL_PRESSURE;
  vc1    = 0.00317;
  utc1   = 1.0 / tc1;
  upc1   = 1.0 / pc1;
  vc1mol = vc1 * ORC_FMWH2O;

  theta   = (t + 273.15) * utc1;
  theta2x = theta * theta;
  theta18 = pow(theta, 18.0);
  theta20 = theta18 * theta2x;

  beta   = p * upc1;
  beta2x = beta * beta;

  yy = 1.0 - a1 * theta2x - a2 * pow(theta, -6.0);
  xx = a3 * yy * yy - 2.0 * (a4 * theta - a5 * beta);
  if (xx > 0.0) xx = sqrt(xx);
  else          xx = (double)1.e-6f;     /* reference aborts here (:283-288) */
  zz = yy + xx;
  u0 = -5.0 / 17.0;
  u1 = aa[11] * a5 * pow(zz, u0);
  u2 = 1.0 / (a8 + pow(theta, 11.0));
  u3 = aa[17] + (2.0 * aa[18] + 3.0 * aa[19] * beta) * beta;
  u4 = 1.0 / (a7 + theta18 * theta);
  u5 = pow(a10 + beta, -4.0);
  u6 = a11 - 3.0 * u5;
  u7 = aa[20] * theta18 * (a9 + theta2x);
  u8 = aa[15] * pow(a6 - theta, 9.0);

  vr = u1 + aa[12] + theta * (aa[13] + aa[14] * theta) + u8 * (a6 - theta)
     + aa[16] * u4 - u2 * u3 - u6 * u7 + (3.0 * aa[21] * (a12 - theta)
     + 4.0 * aa[22] * beta / theta20) * beta2x;

  *dwmol = 1.0 / (vr * vc1mol);
  *dw    = 1.0 / (vr * vc1);

  ypt = 6.0 * a2 * pow(theta, -7.0) - 2.0 * a1 * theta;

  if (calculate_derivatives) {
    zpt  = ypt + (a3 * yy * ypt - a4) / xx;
    zpp  = a5 / xx;
    u9   = u0 * u1 / zz;
    vrpt = u9 * zpt + aa[13] + 2.0 * aa[14] * theta - 10.0 * u8
         - 19.0 * aa[16] * u4 * u4 * theta18 + 11.0 * u2 * u2 * u3 * pow(theta, 10.0)
         - aa[20] * u6 * (18.0 * a9 * theta18 + 20.0 * theta20) / theta
         - (3.0 * aa[21] + 80.0 * aa[22] * beta / (theta20 * theta)) * beta2x;

    vrpp = u9 * zpp - u2 * (2.0 * aa[18] + 6.0 * aa[19] * beta) - 12.0 * u7 * u5 /
           (a10 + beta) + (6.0 * aa[21] * (a12 - theta) + 12.0 * aa[22] * beta /
           theta20) * beta;

    cnv  = -1.0 / (vc1mol * vr * vr);
    *dwt = cnv * vrpt * utc1;
    *dwp = cnv * vrpp * upc1;
  } else {
    *dwt = 0.0;
    *dwp = 0.0;
  }
}

/* EOSWaterMod.F90:347-565 EnthalpyIFC67 (t in Celsius, p in Pa; hw in J/kmol) */
void orc_enthalpy_ifc67(double t, double p, int calculate_derivatives, double *hw, double *hwp, double *hwt)
{
  int i;
  double beta, beta2x, beta4, theta, utheta, theta2x, theta18, theta20;
  double xx, yy, zz, u0, u1, tempreal;
  double v0_1, v1_1, v2_1, v3_1, v4_1;
  double v1_2, v2_2, v3_2, v4_2, v20_2, v40_2;
  double v1_3, v2_3, v3_3, v4_3;
  double v1_4, v2_4, v3_4;
  double v1_5, v2_5, v1_6;
  double term1, term2, term2t, term3, term3t, term3p, term4, term4t, term4p,
         term5, term5t, term5p, term6, term6t, term6p, term7, term7t, term7p;
  double dv2t, dv2p, dv3t, ypt, yptt, zpt, zpp;
  double tc1, pc1, vc1, utc1, upc1, vc1mol;

  tc1    = H2O_CRITICAL_TEMPERATURE;
  pc1    = H2O_CRITICAL_PRESSURE;
  vc1    = 0.00317;
  utc1   = 1.0 / tc1;
  upc1   = 1.0 / pc1;
  vc1mol = vc1 * ORC_FMWH2O;

  theta   = (t + 273.15) * utc1;
  theta2x = theta * theta;
  theta18 = pow(theta, 18.0);
  theta20 = theta18 * theta2x;

  beta   = p * upc1;
  beta2x = beta * beta;
  beta4  = beta2x * beta2x;

  yy = 1.0 - a1 * theta2x - a2 * pow(theta, -6.0);
  xx = a3 * yy * yy - 2.0 * (a4 * theta - a5 * beta);
  if (xx > 0.0) xx = sqrt(xx);
  else          xx = (double)1.e-6f;
  zz = yy + xx;
  u0 = -5.0 / 17.0;
  u1 = aa[11] * a5 * pow(zz, u0);

  ypt = 6.0 * a2 * pow(theta, -7.0) - 2.0 * a1 * theta;

  utheta = 1.0 / theta;
  term1  = aa[0] * theta;
  term2  = -aa[1];
  term2t = 0.0;
  for (i = 3; i <= 10; i++) {                      /* :461-465 */
    tempreal = (double)(i - 2) * aa[i] * pow(theta, (double)(i - 1));
    term2t   = term2t + tempreal * utheta * (double)(i - 1);
    term2    = term2 + tempreal;
  }

  /* "v" section 1 */
  v0_1  = u1 / a5;
  v2_1  = 17.0 * (zz / 29.0 - yy / 12.0) + 5.0 * theta * ypt / 12.0;
  v3_1  = a4 * theta - (a3 - 1.0) * theta * yy * ypt;
  v1_1  = zz * v2_1 + v3_1;
  term3 = v0_1 * v1_1;

  /* "v" section 2 */
  v1_2  = 9.0 * theta + a6;
  v20_2 = (a6 - theta);
  v2_2  = pow(v20_2, 9.0);
  v3_2  = a7 + 20.0 * pow(theta, 19.0);
  v40_2 = a7 + pow(theta, 19.0);
  v4_2  = 1.0 / (v40_2 * v40_2);
  term4p = aa[12] - aa[14] * theta2x + aa[15] * v1_2 * v2_2 + aa[16] * v3_2 * v4_2;
  term4  = term4p * beta;

  /* "v" section 3 */
  v1_3  = beta * (aa[17] + aa[18] * beta + aa[19] * beta2x);
  v2_3  = 12.0 * pow(theta, 11.0) + a8;
  v4_3  = 1.0 / (a8 + pow(theta, 11.0));
  v3_3  = v4_3 * v4_3;
  term5 = v1_3 * v2_3 * v3_3;

  /* "v" section 4 */
  v1_4  = pow(a10 + beta, -3.0) + a11 * beta;
  v3_4  = (17.0 * a9 + 19.0 * theta2x);
  v2_4  = aa[20] * theta18 * v3_4;
  term6 = v1_4 * v2_4;

  /* "v" section 5 */
  v1_5  = 21.0 * aa[22] / theta20 * beta4;
  v2_5  = aa[21] * a12 * beta2x * beta;
  term7 = v1_5 + v2_5;

  /* "v" section 6 */
  v1_6 = pc1 * vc1mol;
  *hw  = (term1 - term2 + term3 + term4 - term5 + term6 + term7) * v1_6;

  if (calculate_derivatives) {
    zpt = ypt + (a3 * yy * ypt - a4) / xx;
    zpp = a5 / xx;

    /* block 1 */
    yptt   = -2.0 * a1 - 42.0 * a2 / pow(theta, 8.0);
    dv2t   = 17.0 * (zpt / 29.0 - ypt / 12.0) + 5.0 / 12.0 * (ypt + theta * yptt);
    dv3t   = a4 - (a3 - 1.0) * (theta * yy * yptt + yy * ypt + theta * ypt * ypt);
    dv2p   = 17.0 * zpp / 29.0;
    v4_1   = 5.0 * v1_1 / (17.0 * zz);
    term3t = v0_1 * (zz * dv2t + (v2_1 - v4_1) * zpt + dv3t);
    term3p = v0_1 * (zz * dv2p + (v2_1 - v4_1) * zpp);

    /* block 2 */
    term4t = (-2.0 * aa[14] * theta + 9.0 * aa[15] * (v2_2 - v1_2 * v2_2 / v20_2)
            + 38.0 * theta18 * aa[16] * (10.0 * v4_2 - v3_2 * v4_2 / v40_2)) * beta;

    /* block 3 */
    term5p = v3_3 * v2_3 * (aa[17] + 2.0 * aa[18] * beta + 3.0 * aa[19] * beta2x);
    term5t = v1_3 * (132.0 * v3_3 * pow(theta, 10.0) - 22.0 * v2_3 * v3_3 * v4_3 * pow(theta, 10.0));

    /* block 4 */
    term6p = v2_4 * (a11 - 3.0 * pow(a10 + beta, -4.0));
    term6t = v1_4 * aa[20] * theta18 * (18.0 * v3_4 * utheta + 38.0 * theta);

    /* block 5 */
    term7p = beta2x * (3.0 * aa[21] * a12 + 84.0 * aa[22] * beta / theta20);
    term7t = -420.0 * aa[22] * beta4 / (theta20 * theta);

    *hwp = (term3p + term4p - term5p + term6p + term7p) * vc1mol;
    *hwt = (aa[0] - term2t + term3t + term4t - term5t + term6t + term7t) * v1_6 * utc1;
  } else {
    *hwp = 0.0;
    *hwt = 0.0;
  }
}

/* EOSWaterMod.F90:38-77 Density (dispatcher) */
void orc_density(double p, double t_K, int density_itype, double *den, double *dden_dp, double *dden_dT)
{
  double den_kg;
  switch (density_itype) {
  case DENSITY_CONSTANT: orc_density_constant(den, dden_dp, dden_dT); break;
  case DENSITY_TGDPB01:  orc_density_tgdpb01(p, t_K, den, dden_dp, dden_dT); break;
  case DENSITY_IFC67:    orc_density_ifc67(t_K - 273.15, p, 1, &den_kg, den, dden_dp, dden_dT); break;
  default: *den = *dden_dp = *dden_dT = NAN;
  }
}

/* EOSWaterMod.F90:568-586 Viscosity */
void orc_viscosity(double p, double t_K, double *vis, double *dvis_dp, double *dvis_dT)
{
  (void)p; (void)t_K;
  *vis     = 8.904156e-4;
  *dvis_dp = 0.0;
  *dvis_dT = 0.0;
}

/* EOSWaterMod.F90:589-707 InternalEnergyAndEnthalpy{,Constant,IFC67}.
 * NB `u0 = 4.217 * 1.d3` (:658,:689): 4.217 is a default-REAL (single precision)
 * literal promoted to double, i.e. 4.21700000762939453125d0 * 1000. */
void orc_internal_energy_enthalpy(double P, double t_K, int itype, double den, double dden_dT, double dden_dP,
                                  double *U, double *H, double *dU_dT, double *dH_dT, double *dU_dP, double *dH_dP)
{
  const double u0 = (double)4.217f * 1.e3;
  if (itype == INT_ENERGY_ENTHALPY_CONSTANT) {
    /* :629-672; den here is in kg m^-3 */
    *U     = u0 * (t_K - 273.15);
    *dU_dT = u0;
    *dU_dP = 0.0;
    *H     = *U + P / den;
    *dH_dT = *dU_dT - P / (pow(den, 2.0)) * dden_dT;
    *dH_dP = *dU_dP + 1.0 / den - P / (pow(den, 2.0)) * dden_dP;
    *U     = *U * ORC_FMWH2O;
    *H     = *H * ORC_FMWH2O;
    *dU_dT = *dU_dT * ORC_FMWH2O;
    *dH_dT = *dH_dT * ORC_FMWH2O;
    *dH_dP = *dH_dP * ORC_FMWH2O;
    /* dU_dP is NOT multiplied by FMWH2O in the reference (it is 0 anyway) */
  } else {
    /* :675-707 */
    double T_C = t_K - 273.15;
    orc_enthalpy_ifc67(T_C, P, 1, H, dH_dP, dH_dT);
    *U     = *H - P / (den / ORC_FMWH2O);
    *dU_dT = *dH_dT + P / (pow(den / ORC_FMWH2O, 2.0)) * (dden_dT / ORC_FMWH2O);
    *dU_dP = *dH_dP - 1.0 / (den / ORC_FMWH2O) + P / (pow(den / ORC_FMWH2O, 2.0)) * (dden_dP / ORC_FMWH2O);
  }
}
