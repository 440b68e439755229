// hostcheck.cpp -- exposes the __host__ __device__ constitutive relations of physics.cuh to the CPU test-suite
// (tests/test_physics_host.py) so they can be compared with the oracle without a GPU.  Not part of the product path.
#include "physics.cuh"
using namespace mpp;
extern "C" {
int hc_convert_soil(int satfunc_name, double watsat, double hksat, double bsw, double sucsat, double residual_sat, double *out /*por,perm,sat_res,alpha,m,n,pu,ps,b2,b3*/)
{
  SatParams sp; double por, perm;
  int bad = convert_soil(satfunc_name, watsat, hksat, bsw, sucsat, residual_sat, por, perm, sp);
  out[0] = por; out[1] = perm; out[2] = sp.sat_res; out[3] = sp.alpha; out[4] = sp.m; out[5] = sp.n; out[6] = sp.pu; out[7] = sp.ps; out[8] = sp.b2; out[9] = sp.b3;
  return bad;
}
void hc_sat(int satfunc /*0 VG 1 BC 2 SBC*/, const double *p /*sat_res,alpha,m,n,pu,ps,b2,b3*/, double press, double frac_liq, double *out /*sat,dsat,kr,dkr*/)
{
  SatParams sp; sp.sat_res = p[0]; sp.alpha = p[1]; sp.m = p[2]; sp.n = p[3]; sp.pu = p[4]; sp.ps = p[5]; sp.b2 = p[6]; sp.b3 = p[7];
  SatState s; sat_values_rt(satfunc, sp, press, frac_liq, s);
  double ds, dk; sat_derivs_rt(satfunc, sp, s, frac_liq, ds, dk);
  out[0] = s.sat; out[1] = ds; out[2] = s.kr; out[3] = dk;
}
void hc_log_exp(int n, const double *x, double *lg, double *ex) { for (int i = 0; i < n; ++i) { lg[i] = mpp_log(x[i]); ex[i] = mpp_exp(x[i]); } }
void hc_density(int itype, double p, double t_K, double *out) { density(itype, p, t_K, out[0], out[1], out[2]); }
void hc_density_fixedT(int itype, double t_K, double p, double *out) { DensityTable t = make_density_table(itype, t_K); density_fixedT(t, p, out[0], out[1]); }
void hc_enthalpy_ifc67(double t_C, double p, double *out) { enthalpy_ifc67(t_C, p, out[0], out[1], out[2]); }
void hc_internal_energy_enthalpy(int itype, double P, double t_K, double den, double dden_dT, double dden_dP, double *out)
{ internal_energy_enthalpy(itype, P, t_K, den, dden_dT, dden_dP, out[0], out[1], out[2], out[3], out[4], out[5]); }
}
