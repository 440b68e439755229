// physics.cuh -- device constitutive relations of the MPP column hot path (fp64).
//
// What the reference computes per cell per residual evaluation
// (RichODEPressureAuxVarCompute, src/mpp/auxvar/RichardsODEPressureAuxType.F90:237-294):
//   saturation + d/dP   SaturationFunction.F90:747-795 (VG), 900-938 (BC), 1027-1076 (SBC)
//   rel. perm + d/dP    SaturationFunction.F90:799-857 (VG-Mualem), 942-990 (BC), 1080-1140 (SBC)
//   density + d/dP,d/dT EOSWaterMod.F90:80-99 (constant), 102-178 (Tanaka 2001), 181-344 (IFC-67)
//   viscosity           EOSWaterMod.F90:568-586 (constant 8.904156e-4)
//   enthalpy            EOSWaterMod.F90:347-565 (IFC-67), 629-707
//
// B200 design notes.  The path is fp64-issue bound, not HBM bound: the reference's
// 5 pow() per van Genuchten cell are restated as 3 log + 3 exp for the values (shared between
// saturation and permeability, which the reference evaluates twice) and the derivative terms are
// split off (`*_deriv`) so that line-search trial points, which never need a Jacobian, skip them.
// Everything is __host__ __device__ so the CPU test-suite can check these exact functions against
// the oracle without a GPU (tests/test_physics_host.py).
#pragma once
#include <math.h>

#ifdef __CUDACC__
#define MPP_HD __host__ __device__ __forceinline__
#else
#define MPP_HD inline
#endif

namespace mpp {

// MultiPhysicsProbConstants.F90:199-202, mpp_varcon.F90:12-28
constexpr double PRESSURE_REF     = 101325.0;
constexpr double GRAVITY_CONSTANT = 9.80665;
constexpr double FMWH2O           = 18.01534;
constexpr double GRAV_CLM         = 9.80616;
constexpr double DENH2O           = 1000.0;
constexpr double DENICE           = 917.0;
constexpr double CPLIQ            = 4.188e3;
constexpr double CPICE            = 2.11727e3;
constexpr double TKWAT            = 0.57;
constexpr double TKICE            = 2.29;
constexpr double THK_BEDROCK      = 3.0;
constexpr double TFRZ             = 273.15;
constexpr double VISCOSITY        = 8.904156e-4;   // EOSWaterMod.F90:582

enum { DENSITY_CONSTANT = 1, DENSITY_TGDPB01 = 2, DENSITY_IFC67 = 3 };
enum { INT_ENERGY_ENTHALPY_CONSTANT = 1, INT_ENERGY_ENTHALPY_IFC67 = 2 };
enum { SATFUNC_VG = 0, SATFUNC_BC = 1, SATFUNC_SBC = 2 };   // bz2 and bz3 differ only in (pu, b2, b3)

// ------------------------------------------------------------------------------------------------
// Saturation / relative permeability
// ------------------------------------------------------------------------------------------------
struct SatParams {
  double sat_res, alpha, m /* vg_m or bc_lambda */, n /* vg_n */;
  double pu, ps, b2, b3;   // smoothed Brooks-Corey only
};

// values needed by the residual + what the derivative pass re-uses
struct SatState {
  double sat, kr;
  double Se, AA, AAm, L2, pc;   // VG intermediates; for BC/SBC: Se, pc, AA = dSe_dpc
  int    regime;                 // 0 saturated, 1 unsaturated (VG / full BC), 2 SBC cubic
};

template <int SATFUNC>
MPP_HD void sat_values(const SatParams &sp, double press, double frac_liq, SatState &s)
{
  const double pc = press - PRESSURE_REF;
  s.pc = pc;
  if (SATFUNC == SATFUNC_VG) {
    if (pc < 0.0) {
      // Reference: Se = (1 + x^n)^-m, AA = x^n / (1 + x^n), kr = sqrt(Se) (1 - AA^m)^2 with x = -alpha pc
      // (5 pow calls, SaturationFunction.F90:777-836).  Here: 2 log + 3 exp, using AA^m = x^(n m) Se and n m = n - 1.
      const double L1  = log(-sp.alpha * pc);
      const double pcn = exp(sp.n * L1);               // x^n
      const double opn = 1.0 + pcn;
      const double L2  = log(opn);
      const double mL2 = sp.m * L2;
      const double Se  = exp(-mL2);                    // (1 + x^n)^(-m)
      const double AAm = exp((sp.n - 1.0) * L1 - mL2); // AA^m
      const double BB  = 1.0 - AAm;
      const double rS  = sqrt(Se);
      s.sat = sp.sat_res + (1.0 - sp.sat_res) * Se;
      s.kr  = rS * BB * BB;
      s.Se = Se; s.AA = pcn; s.AAm = AAm; s.L2 = rS; s.regime = 1;   // (AA slot carries x^n, L2 slot carries sqrt(Se))
    } else {
      s.sat = 1.0; s.kr = 1.0; s.regime = 0;
    }
  } else if (SATFUNC == SATFUNC_BC) {
    const double pc_alpha = -sp.alpha * pc;
    if (pc_alpha > 1.0) {
      const double L1 = log(pc_alpha);
      const double Se = exp(-sp.m * L1);
      s.sat = sp.sat_res + (1.0 - sp.sat_res) * Se;
      s.kr  = exp((2.5 + 2.0 / sp.m) * (-sp.m * L1));  // Se^(2.5 + 2/lambda)
      s.Se = Se; s.regime = 1;
    } else {
      s.sat = 1.0; s.kr = 1.0; s.regime = 0;
    }
    s.kr = frac_liq * s.kr;                            // SaturationFunction.F90:987
  } else {
    if (pc <= sp.pu) {
      const double L1 = log(-sp.alpha * pc);
      const double Se = exp(-sp.m * L1);
      s.sat = sp.sat_res + (1.0 - sp.sat_res) * Se;
      s.kr  = exp((2.5 + 2.0 / sp.m) * (-sp.m * L1));
      s.Se = Se; s.regime = 1;
    } else if (pc < sp.ps) {
      const double dpc = pc - sp.ps;
      const double Se  = 1.0 + dpc * dpc * (sp.b2 + dpc * sp.b3);
      s.sat = sp.sat_res + (1.0 - sp.sat_res) * Se;
      s.kr  = exp((2.5 + 2.0 / sp.m) * log(Se));
      s.Se = Se; s.regime = 2;
    } else {
      s.sat = 1.0; s.kr = 1.0; s.regime = 0;
    }
  }
}

template <int SATFUNC>
MPP_HD void sat_derivs(const SatParams &sp, const SatState &s, double frac_liq, double &dsat_dP, double &dkr_dP)
{
  if (s.regime == 0) { dsat_dP = 0.0; dkr_dP = 0.0; return; }
  if (SATFUNC == SATFUNC_VG) {
    // dSe/dpc = -m n Se AA / pc with AA = x^n / (1 + x^n)                       (SaturationFunction.F90:790, 829)
    // dkr/dSe = kr / (2 Se) + 2 Se^(1/m - 1/2) AA^(m-1) BB                       (:836-838)
    //         = BB (BB / 2 + 2 AA^m / x^n) / sqrt(Se)      since Se^(1/m) = 1 / (1 + x^n)
    // (same functions, regrouped so that no further exp/pow is needed; derivative round-off only steers the Newton
    //  path, never the converged answer)
    const double pcn = s.AA, rS = s.L2, BB = 1.0 - s.AAm;
    const double dSe_dpc = -sp.m * sp.n * s.Se * pcn / ((1.0 + pcn) * s.pc);
    const double dkr_dSe = BB * (0.5 * BB + 2.0 * s.AAm / pcn) / rS;
    dsat_dP = (1.0 - sp.sat_res) * dSe_dpc;
    dkr_dP  = dkr_dSe * dSe_dpc;
  } else {
    double dSe_dpc;
    if (s.regime == 1) dSe_dpc = -sp.m * s.Se / s.pc;
    else { const double dpc = s.pc - sp.ps; dSe_dpc = dpc * (2.0 * sp.b2 + 3.0 * dpc * sp.b3); }
    // s.kr carries frac_liq for BC; dkr_dSe = (2.5 + 2/lambda) kr / Se
    dsat_dP = (1.0 - sp.sat_res) * dSe_dpc;
    dkr_dP  = (2.5 + 2.0 / sp.m) * s.kr / s.Se * dSe_dpc;
    (void)frac_liq;
  }
}

// runtime-dispatched wrappers (generic kernels, host tests)
MPP_HD void sat_values_rt(int satfunc, const SatParams &sp, double press, double frac_liq, SatState &s)
{
  if (satfunc == SATFUNC_VG) sat_values<SATFUNC_VG>(sp, press, frac_liq, s);
  else if (satfunc == SATFUNC_BC) sat_values<SATFUNC_BC>(sp, press, frac_liq, s);
  else sat_values<SATFUNC_SBC>(sp, press, frac_liq, s);
}
MPP_HD void sat_derivs_rt(int satfunc, const SatParams &sp, const SatState &s, double frac_liq, double &dsat, double &dkr)
{
  if (satfunc == SATFUNC_VG) sat_derivs<SATFUNC_VG>(sp, s, frac_liq, dsat, dkr);
  else if (satfunc == SATFUNC_BC) sat_derivs<SATFUNC_BC>(sp, s, frac_liq, dsat, dkr);
  else sat_derivs<SATFUNC_SBC>(sp, s, frac_liq, dsat, dkr);
}

// findGu_SBC_zeroCoeff, SaturationFunction.F90:425-518 (bracketed Newton-Raphson, setup only)
MPP_HD double find_gu_sbc_zero_coeff(double lambda, int AA, double gs)
{
  const double relTol = 1.e-12;
  double gu = pow((double)AA / ((double)AA + lambda), -1.0 / lambda);
  if (gs > 0.0) {
    double guLeft = 1.0, guRight = gu;
    for (int it = 0; it < 200; ++it) {
      if (gu <= guLeft || gu >= guRight) gu = guLeft + 0.5 * (guRight - guLeft);
      const double guInv = 1.0 / gu, guToMinusLam = pow(gu, -lambda), gsOnGu = gs * guInv;
      const double resid = AA - guToMinusLam * (AA + lambda - lambda * gsOnGu);
      if (resid < 0.0) guLeft = gu; else guRight = gu;
      double dr = (1.0 + lambda) * (1.0 - gsOnGu) + (AA - 1);
      dr = lambda * guToMinusLam * guInv * dr;
      const double dgu = resid / dr;
      gu = gu - dgu;
      if (fabs(dgu) < relTol * fabs(gu)) break;
    }
  }
  return gu;
}

// VSFMMPPSetSoilsCLM parameter conversion (MultiPhysicsProbVSFM.F90:361-420) + SatFunc_Set_* (SaturationFunction.F90:127-372)
// satfunc_name: 0 VG, 1 BC, 2 SBC bz2, 3 SBC bz3.  Returns non-zero on the reference's "bad param" aborts.
MPP_HD int convert_soil(int satfunc_name, double watsat, double hksat, double bsw, double sucsat, double residual_sat,
                        double &por, double &perm, SatParams &sp)
{
  const double vish2o = 0.001002;
  perm = hksat * vish2o / (DENH2O * GRAV_CLM) * 0.001;
  const double alpha = 1.0 / (sucsat * GRAV_CLM), lambda = 1.0 / bsw;
  por = watsat;
  sp.sat_res = residual_sat; sp.alpha = alpha; sp.m = lambda; sp.n = 0.0; sp.pu = sp.ps = sp.b2 = sp.b3 = 0.0;
  int bad = (residual_sat < 0.0 || residual_sat > 0.5 || alpha <= 0.0 || alpha > 2.0);
  if (satfunc_name == 0) {
    bad |= (lambda <= 0.0 || lambda >= 1.0);
    sp.n = 1.0 / (1.0 - lambda);
  } else {
    bad |= (lambda <= 0.0 || lambda >= 2.0);
    if (satfunc_name >= 2) {
      const double ps = -0.9 / alpha;
      bad |= (ps <= -1.0 / alpha || ps > 0.0);
      const int AA = (satfunc_name == 2) ? 3 : 2;
      const double pu = find_gu_sbc_zero_coeff(lambda, AA, -alpha * ps) / (-alpha);
      const double bcAtPu = pow(-alpha * pu, -lambda), lamDelta = lambda * (1.0 - ps / pu), oneOnDelta = 1.0 / (pu - ps);
      sp.ps = ps; sp.pu = pu;
      if (satfunc_name == 2) { sp.b2 = 0.0; sp.b3 = (2.0 - bcAtPu * (2.0 + lamDelta)) * oneOnDelta * oneOnDelta * oneOnDelta; bad |= (sp.b3 <= 0.0); }
      else                   { sp.b3 = 0.0; sp.b2 = -(3.0 - bcAtPu * (3.0 + lamDelta)) * oneOnDelta * oneOnDelta;             bad |= (sp.b2 >= 0.0); }
    }
  }
  return bad;
}

// ------------------------------------------------------------------------------------------------
// Water EOS
// ------------------------------------------------------------------------------------------------
// Tanaka et al. (2001), EOSWaterMod.F90:102-178.  den in kmol m^-3.
MPP_HD void density_tgdpb01(double p, double t_K, double &den, double &dden_dp, double &dden_dT)
{
  const double a1 = -3.983035, a2 = 301.797, a3 = 522528.9, a4 = 69.34881, a5 = 999.974950;
  const double k0 = 50.74e-11, k1 = -0.326e-11, k2 = 0.00416e-11, p0 = 101325.0;
  const double t_c = t_K - 273.15;
  const double s1 = (t_c + a1) * (t_c + a1);
  const double dent = a5 * (1.0 - s1 * (t_c + a2) / a3 / (t_c + a4));
  const double kc = k0 + k1 * t_c + k2 * t_c * t_c;
  double kappa, dkappa_dp, dkappa_dt;
  if (p > p0) { kappa = 1.0 + kc * (p - p0); dkappa_dp = kc; dkappa_dt = (k1 + 2.0 * k2 * t_c) * (p - p0); }
  else        { kappa = 1.0; dkappa_dp = 0.0; dkappa_dt = 0.0; }
  den = dent * kappa / FMWH2O;
  const double d1 = -s1 / a3 / (t_c + a4);
  const double d2 = -2.0 * (t_c + a1) * (t_c + a2) / a3 / (t_c + a4);
  const double d3 = s1 * (t_c + a2) / a3 / ((t_c + a4) * (t_c + a4));
  const double ddent_dt = a5 * (d1 + d2 + d3);
  dden_dT = (ddent_dt * kappa + dent * dkappa_dt) / FMWH2O;
  dden_dp = (dent * dkappa_dp) / FMWH2O;
}

// The VSFM aux vars never receive a temperature (GoveqnRichardsODEPressureType.F90:573-575), so T = 298.15 K
// always (RichardsODEPressureAuxType.F90:92): the T-dependent factors collapse to two per-problem constants.
struct DensityTable { int type; double dent_over_fmw, kcoef; };
MPP_HD DensityTable make_density_table(int density_type, double t_K)
{
  DensityTable t; t.type = density_type; t.dent_over_fmw = DENH2O / FMWH2O; t.kcoef = 0.0;
  if (density_type == DENSITY_TGDPB01) {
    double d0, dp0, dt0, d1, dp1, dt1;
    density_tgdpb01(101325.0, t_K, d0, dp0, dt0);          // kappa = 1
    density_tgdpb01(101325.0 + 1.0, t_K, d1, dp1, dt1);    // dden_dp = dent * kc / FMW
    t.dent_over_fmw = d0; t.kcoef = dp1 / d0;
  }
  return t;
}
MPP_HD void density_fixedT(const DensityTable &t, double p, double &den, double &dden_dp)
{
  if (t.type == DENSITY_TGDPB01 && p > 101325.0) { den = t.dent_over_fmw * (1.0 + t.kcoef * (p - 101325.0)); dden_dp = t.dent_over_fmw * t.kcoef; }
  else { den = t.dent_over_fmw; dden_dp = 0.0; }
}

MPP_HD double ipow(double x, int n) { double r = 1.0; for (int i = 0; i < n; ++i) r *= x; return r; }

namespace ifc67 {
constexpr double aa0 = 6.824687741e03, aa1 = -5.422063673e02, aa2 = -2.096666205e04, aa3 = 3.941286787e04,
  aa4 = -6.733277739e04, aa5 = 9.902381028e04, aa6 = -1.093911774e05, aa7 = 8.590841667e04,
  aa8 = -4.511168742e04, aa9 = 1.418138926e04, aa10 = -2.017271113e03, aa11 = 7.982692717e00,
  aa12 = -2.616571843e-2, aa13 = 1.522411790e-3, aa14 = 2.284279054e-2, aa15 = 2.421647003e02,
  aa16 = 1.269716088e-10, aa17 = 2.074838328e-7, aa18 = 2.174020350e-8, aa19 = 1.105710498e-9,
  aa20 = 1.293441934e01, aa21 = 1.308119072e-5, aa22 = 6.047626338e-14;
constexpr double a1 = 8.438375405e-1, a2 = 5.362162162e-4, a3 = 1.720000000e00, a4 = 7.342278489e-2,
  a5 = 4.975858870e-2, a6 = 6.537154300e-1, a7 = 1.150000000e-6, a8 = 1.510800000e-5,
  a9 = 1.418800000e-1, a10 = 7.002753165e00, a11 = 2.995284926e-4, a12 = 2.040000000e-1;
constexpr double TC1 = 647.3, PC1 = 22.064e6, VC1 = 0.00317;
}

// DensityIFC67, EOSWaterMod.F90:181-344 (t in Celsius).  dwmol [kmol m^-3], dwp [kmol m^-3 Pa^-1], dwt [kmol m^-3 C^-1]
MPP_HD void density_ifc67(double t, double p, double &dwmol, double &dwp, double &dwt)
{
  using namespace ifc67;
  const double utc1 = 1.0 / TC1, upc1 = 1.0 / PC1, vc1mol = VC1 * FMWH2O;
  const double theta = (t + 273.15) * utc1, theta2x = theta * theta;
  const double th4 = theta2x * theta2x, th8 = th4 * th4, th16 = th8 * th8;
  const double theta18 = th16 * theta2x, theta20 = theta18 * theta2x, th10 = th8 * theta2x, th11 = th10 * theta;
  const double beta = p * upc1, beta2x = beta * beta;
  const double th6 = th4 * theta2x;
  const double yy = 1.0 - a1 * theta2x - a2 / th6;
  double xx = a3 * yy * yy - 2.0 * (a4 * theta - a5 * beta);
  xx = (xx > 0.0) ? sqrt(xx) : (double)1.e-6f;
  const double zz = yy + xx;
  const double u0 = -5.0 / 17.0;
  const double u1 = aa11 * a5 * pow(zz, u0);
  const double u2 = 1.0 / (a8 + th11);
  const double u3 = aa17 + (2.0 * aa18 + 3.0 * aa19 * beta) * beta;
  const double u4 = 1.0 / (a7 + theta18 * theta);
  const double ab = a10 + beta, ab2 = ab * ab;
  const double u5 = 1.0 / (ab2 * ab2);
  const double u6 = a11 - 3.0 * u5;
  const double u7 = aa20 * theta18 * (a9 + theta2x);
  const double amt = a6 - theta, amt2 = amt * amt, amt4 = amt2 * amt2;
  const double u8 = aa15 * (amt4 * amt4 * amt);
  const double vr = u1 + aa12 + theta * (aa13 + aa14 * theta) + u8 * amt + aa16 * u4 - u2 * u3 - u6 * u7
                  + (3.0 * aa21 * (a12 - theta) + 4.0 * aa22 * beta / theta20) * beta2x;
  dwmol = 1.0 / (vr * vc1mol);
  const double ypt = 6.0 * a2 / (th6 * theta) - 2.0 * a1 * theta;
  const double zpt = ypt + (a3 * yy * ypt - a4) / xx;
  const double zpp = a5 / xx;
  const double u9 = u0 * u1 / zz;
  const double vrpt = u9 * zpt + aa13 + 2.0 * aa14 * theta - 10.0 * u8 - 19.0 * aa16 * u4 * u4 * theta18
                    + 11.0 * u2 * u2 * u3 * th10 - aa20 * u6 * (18.0 * a9 * theta18 + 20.0 * theta20) / theta
                    - (3.0 * aa21 + 80.0 * aa22 * beta / (theta20 * theta)) * beta2x;
  const double vrpp = u9 * zpp - u2 * (2.0 * aa18 + 6.0 * aa19 * beta) - 12.0 * u7 * u5 / ab
                    + (6.0 * aa21 * (a12 - theta) + 12.0 * aa22 * beta / theta20) * beta;
  const double cnv = -1.0 / (vc1mol * vr * vr);
  dwt = cnv * vrpt * utc1;
  dwp = cnv * vrpp * upc1;
}

// EnthalpyIFC67, EOSWaterMod.F90:347-565 (t in Celsius).  hw [J kmol^-1]
MPP_HD void enthalpy_ifc67(double t, double p, double &hw, double &hwp, double &hwt)
{
  using namespace ifc67;
  const double utc1 = 1.0 / TC1, upc1 = 1.0 / PC1, vc1mol = VC1 * FMWH2O;
  const double theta = (t + 273.15) * utc1, theta2x = theta * theta;
  const double th4 = theta2x * theta2x, th8 = th4 * th4, th16 = th8 * th8, th6 = th4 * theta2x;
  const double theta18 = th16 * theta2x, theta20 = theta18 * theta2x, th10 = th8 * theta2x, th11 = th10 * theta, th19 = theta18 * theta;
  const double beta = p * upc1, beta2x = beta * beta, beta4 = beta2x * beta2x;
  const double yy = 1.0 - a1 * theta2x - a2 / th6;
  double xx = a3 * yy * yy - 2.0 * (a4 * theta - a5 * beta);
  xx = (xx > 0.0) ? sqrt(xx) : (double)1.e-6f;
  const double zz = yy + xx;
  const double u0 = -5.0 / 17.0;
  const double u1 = aa11 * a5 * pow(zz, u0);
  const double ypt = 6.0 * a2 / (th6 * theta) - 2.0 * a1 * theta;
  const double utheta = 1.0 / theta;
  const double term1 = aa0 * theta;
  // do i = 3,10: tempreal = (i-2) aa(i) theta^(i-1); term2t += tempreal (i-1)/theta; term2 += tempreal   (:461-465)
  const double caa[8] = {aa3, aa4, aa5, aa6, aa7, aa8, aa9, aa10};
  double term2 = -aa1, term2t = 0.0, thp = theta2x;   // theta^(i-1), i = 3 -> theta^2
  for (int i = 3; i <= 10; ++i) {
    const double tempreal = (double)(i - 2) * caa[i - 3] * thp;
    term2t += tempreal * utheta * (double)(i - 1);
    term2  += tempreal;
    thp *= theta;
  }
  const double v0_1 = u1 / a5;
  const double v2_1 = 17.0 * (zz / 29.0 - yy / 12.0) + 5.0 * theta * ypt / 12.0;
  const double v3_1 = a4 * theta - (a3 - 1.0) * theta * yy * ypt;
  const double v1_1 = zz * v2_1 + v3_1;
  const double term3 = v0_1 * v1_1;
  const double v1_2 = 9.0 * theta + a6;
  const double v20_2 = a6 - theta, v20_2_2 = v20_2 * v20_2, v20_2_4 = v20_2_2 * v20_2_2;
  const double v2_2 = v20_2_4 * v20_2_4 * v20_2;
  const double v3_2 = a7 + 20.0 * th19;
  const double v40_2 = a7 + th19;
  const double v4_2 = 1.0 / (v40_2 * v40_2);
  const double term4p = aa12 - aa14 * theta2x + aa15 * v1_2 * v2_2 + aa16 * v3_2 * v4_2;
  const double term4 = term4p * beta;
  const double v1_3 = beta * (aa17 + aa18 * beta + aa19 * beta2x);
  const double v2_3 = 12.0 * th11 + a8;
  const double v4_3 = 1.0 / (a8 + th11);
  const double v3_3 = v4_3 * v4_3;
  const double term5 = v1_3 * v2_3 * v3_3;
  const double ab = a10 + beta, ab2 = ab * ab;
  const double v1_4 = 1.0 / (ab2 * ab) + a11 * beta;
  const double v3_4 = 17.0 * a9 + 19.0 * theta2x;
  const double v2_4 = aa20 * theta18 * v3_4;
  const double term6 = v1_4 * v2_4;
  const double v1_5 = 21.0 * aa22 / theta20 * beta4;
  const double v2_5 = aa21 * a12 * beta2x * beta;
  const double term7 = v1_5 + v2_5;
  const double v1_6 = PC1 * vc1mol;
  hw = (term1 - term2 + term3 + term4 - term5 + term6 + term7) * v1_6;

  const double zpt = ypt + (a3 * yy * ypt - a4) / xx;
  const double zpp = a5 / xx;
  const double yptt = -2.0 * a1 - 42.0 * a2 / th8;
  const double dv2t = 17.0 * (zpt / 29.0 - ypt / 12.0) + 5.0 / 12.0 * (ypt + theta * yptt);
  const double dv3t = a4 - (a3 - 1.0) * (theta * yy * yptt + yy * ypt + theta * ypt * ypt);
  const double dv2p = 17.0 * zpp / 29.0;
  const double v4_1 = 5.0 * v1_1 / (17.0 * zz);
  const double term3t = v0_1 * (zz * dv2t + (v2_1 - v4_1) * zpt + dv3t);
  const double term3p = v0_1 * (zz * dv2p + (v2_1 - v4_1) * zpp);
  const double term4t = (-2.0 * aa14 * theta + 9.0 * aa15 * (v2_2 - v1_2 * v2_2 / v20_2)
                       + 38.0 * theta18 * aa16 * (10.0 * v4_2 - v3_2 * v4_2 / v40_2)) * beta;
  const double term5p = v3_3 * v2_3 * (aa17 + 2.0 * aa18 * beta + 3.0 * aa19 * beta2x);
  const double term5t = v1_3 * (132.0 * v3_3 * th10 - 22.0 * v2_3 * v3_3 * v4_3 * th10);
  const double term6p = v2_4 * (a11 - 3.0 / (ab2 * ab2));
  const double term6t = v1_4 * aa20 * theta18 * (18.0 * v3_4 * utheta + 38.0 * theta);
  const double term7p = beta2x * (3.0 * aa21 * a12 + 84.0 * aa22 * beta / theta20);
  const double term7t = -420.0 * aa22 * beta4 / (theta20 * theta);
  hwp = (term3p + term4p - term5p + term6p + term7p) * vc1mol;
  hwt = (aa0 - term2t + term3t + term4t - term5t + term6t + term7t) * v1_6 * utc1;
}

// Density dispatcher, EOSWaterMod.F90:38-77
MPP_HD void density(int itype, double p, double t_K, double &den, double &dden_dp, double &dden_dT)
{
  if (itype == DENSITY_CONSTANT)     { den = DENH2O / FMWH2O; dden_dp = 0.0; dden_dT = 0.0; }
  else if (itype == DENSITY_TGDPB01) density_tgdpb01(p, t_K, den, dden_dp, dden_dT);
  else                               density_ifc67(t_K - 273.15, p, den, dden_dp, dden_dT);
}

// InternalEnergyAndEnthalpy, EOSWaterMod.F90:589-707.  `den` arrives in kg m^-3 (callers multiply by FMWH2O).
// NB u0 = 4.217 * 1.d3 with a single-precision literal (:658,:689).
MPP_HD void internal_energy_enthalpy(int itype, double P, double t_K, double den, double dden_dT, double dden_dP,
                                     double &U, double &H, double &dU_dT, double &dH_dT, double &dU_dP, double &dH_dP)
{
  const double u0 = (double)4.217f * 1.e3;
  if (itype == INT_ENERGY_ENTHALPY_CONSTANT) {
    U = u0 * (t_K - 273.15); dU_dT = u0; dU_dP = 0.0;
    H = U + P / den;
    dH_dT = dU_dT - P / (den * den) * dden_dT;
    dH_dP = dU_dP + 1.0 / den - P / (den * den) * dden_dP;
    U *= FMWH2O; H *= FMWH2O; dU_dT *= FMWH2O; dH_dT *= FMWH2O; dH_dP *= FMWH2O;
  } else {
    enthalpy_ifc67(t_K - 273.15, P, H, dH_dP, dH_dT);
    const double dm = den / FMWH2O;
    U = H - P / dm;
    dU_dT = dH_dT + P / (dm * dm) * (dden_dT / FMWH2O);
    dU_dP = dH_dP - 1.0 / dm + P / (dm * dm) * (dden_dP / FMWH2O);
  }
}

}  // namespace mpp
