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
// 5 pow() per van Genuchten cell are restated as 2 log + 2 exp + 1 reciprocal for the values (shared between
// saturation and permeability, which the reference evaluates twice) and the derivative terms are
// split off (`*_deriv`) so that line-search trial points, which never need a Jacobian, skip them.
// Everything is __host__ __device__ so the CPU test-suite can check these exact functions against
// the oracle without a GPU (tests/test_physics_host.py).
#pragma once
#include <math.h>
#include <string.h>

#ifdef __CUDACC__
#define MPP_HD __host__ __device__ __forceinline__
#else
#define MPP_HD inline
#endif

namespace mpp {

// ------------------------------------------------------------------------------------------------
// Lean fp64 log / exp / reciprocal for the saturation curves.
//
// Why not the CUDA math library here: its log()/exp() materialise every polynomial coefficient with two IMAD.MOV
// (22 % of all executed instructions of the first kernels were constant moves, profiles/r1_vsfm_v4.md), branch on
// denormals / NaN / Inf (BSSY/BSYNC pairs that serialise the two cells a lane owns) and take ~70 instructions per
// call.  The arguments here are always positive, finite and far from the denormal range, so:
//   * coefficients live in __constant__ memory and enter DFMA as constant-bank operands (no move instructions);
//   * no special-case branches; the polynomials are split even/odd (Estrin) so two chains overlap;
//   * the reciprocal is MUFU.RCP64H + two Newton steps.
// Accuracy (tests/test_physics_host.py, checked against libm over the ranges the soil curves produce): < 2 ulp.
// Coefficients were derived for this file by Chebyshev interpolation in 50-digit arithmetic (mpmath):
//   exp(r) = 1 + r + r^2 q(r), q of degree 9 on |r| <= ln2/2        (approximation error 0.15 ulp)
//   log(1+f) = f - f^2/2 + s (f^2/2 + R(s^2)), s = f/(2+f), R of degree 7 in s^2, 1+f in [sqrt(1/2), sqrt(2)]  (0.05 ulp)
// ------------------------------------------------------------------------------------------------
#define MPP_CMATH_TABLE { \
  /* 0..6  log: Lg1..Lg7 */ \
  0x1.5555555555558p-1, 0x1.99999999952b6p-2, 0x1.2492492df62cfp-2, 0x1.c71c62df4373bp-3, 0x1.7462b6717d448p-3, \
  0x1.39fe256b357ffp-3, 0x1.2b5b2383c1004p-3, \
  /* 7,8   ln2 split for log (hi has 21 trailing zero bits) */ 0x1.62e42fee00000p-1, 0x1.a39ef35793c76p-33, \
  /* 9..18 exp: c2..c11 */ \
  0x1.0000000000001p-1, 0x1.5555555555556p-3, 0x1.5555555553d63p-5, 0x1.11111111109b3p-7, 0x1.6c16c1788bd90p-10, \
  0x1.a01a01a7c41d5p-13, 0x1.a019b90d2ae7ap-16, 0x1.71de0dae63bb3p-19, 0x1.289185613a3d6p-22, 0x1.af38a9b0ec855p-26, \
  /* 19..21 1/ln2, ln2 split for exp (hi has 11 trailing zero bits) */ 1.4426950408889634, 0.6931471805598903, 5.497923018708371e-14, \
  /* 22..27 table log: Taylor coefficients of log1p(r) / r - 1 */ -0.5, 1.0 / 3.0, -0.25, 0.2, -1.0 / 6.0, 1.0 / 7.0, \
  /* 28,29 table log: ln2 split */ MPP_LOG_LN2_HI, MPP_LOG_LN2_LO, \
  /* 30..32 table exp: 128/ln2, ln2/128 split */ MPP_EXP_INVLN2N, MPP_EXP_LN2N_HI, MPP_EXP_LN2N_LO, \
  /* 33..36 table exp: 1/2!, 1/3!, 1/4!, 1/5! */ 0.5, 1.0 / 6.0, 1.0 / 24.0, 1.0 / 120.0, \
  /* 37 round-to-integer magic 1.5 * 2^52 */ 6755399441055744.0 }
#include "math_tables.inc"
#ifdef __CUDACC__
static __constant__ double mpp_cmath_dev[38] = MPP_CMATH_TABLE;
#endif
static const double mpp_cmath_host[38] = MPP_CMATH_TABLE;
#ifdef __CUDA_ARCH__
#define MPPC(i) mpp_cmath_dev[i]
#else
#define MPPC(i) mpp_cmath_host[i]
#endif

// Pairs.  A lane of the fast VSFM kernel owns two cells; ptxas schedules two calls of a scalar function one after the
// other (profiles/r1_vsfm_v6.md: every DFMA of the log/exp chains waited its full ~9-cycle latency), so the curve math
// is written once, generically, and instantiated for `double` and for the pair type `d2`, whose element-wise
// operators put the two independent chains next to each other in program order.
struct d2 { double a, b; };
MPP_HD d2 operator+(d2 x, d2 y) { return d2{x.a + y.a, x.b + y.b}; }
MPP_HD d2 operator-(d2 x, d2 y) { return d2{x.a - y.a, x.b - y.b}; }
MPP_HD d2 operator*(d2 x, d2 y) { return d2{x.a * y.a, x.b * y.b}; }
MPP_HD d2 operator-(d2 x) { return d2{-x.a, -x.b}; }
MPP_HD double vfma(double x, double y, double z) { return fma(x, y, z); }
MPP_HD d2 vfma(d2 x, d2 y, d2 z) { return d2{fma(x.a, y.a, z.a), fma(x.b, y.b, z.b)}; }
template <class T> MPP_HD T vbc(double c);
template <> MPP_HD double vbc<double>(double c) { return c; }
template <> MPP_HD d2 vbc<d2>(double c) { return d2{c, c}; }
MPP_HD double vsel(bool pa, bool, double x, double y) { return pa ? x : y; }            // element-wise p ? x : y
MPP_HD d2 vsel(bool pa, bool pb, d2 x, d2 y) { return d2{pa ? x.a : y.a, pb ? x.b : y.b}; }

MPP_HD void mpp_split(double x, int &hi, unsigned &lo)
{
#ifdef __CUDA_ARCH__
  hi = __double2hiint(x); lo = (unsigned)__double2loint(x);
#else
  unsigned long long u; memcpy(&u, &x, 8); hi = (int)(u >> 32); lo = (unsigned)u;
#endif
}
MPP_HD double mpp_join(int hi, unsigned lo)
{
#ifdef __CUDA_ARCH__
  return __hiloint2double(hi, (int)lo);
#else
  unsigned long long u = ((unsigned long long)(unsigned)hi << 32) | lo; double x; memcpy(&x, &u, 8); return x;
#endif
}

// reciprocal seed (MUFU.RCP64H, ~2^-23) refined by two Newton steps to <= 1 ulp
MPP_HD double rcp_seed(double x)
{
#ifdef __CUDA_ARCH__
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
  return r;
#else
  return 1.0 / x;
#endif
}
MPP_HD d2 rcp_seed(d2 x) { return d2{rcp_seed(x.a), rcp_seed(x.b)}; }
template <class T> MPP_HD T rcp(T x)
{
  const T one = vbc<T>(1.0);
  T r = rcp_seed(x);
  T e = vfma(-x, r, one); r = vfma(r, e, r);                    // 2^-46
  e = vfma(-x, r, one);   r = vfma(r, e, r);                    // <= 1 ulp
  return r;
}

// one Newton step only (relative error ~2^-46): for pivots of the Newton linear solves, whose accuracy steers the Newton path
// but never the converged answer
template <class T> MPP_HD T rcp1(T x)
{
  T r = rcp_seed(x);
  const T e = vfma(-x, r, vbc<T>(1.0));
  return vfma(r, e, r);
}

// log: x = 2^k (1 + f), 1 + f in [sqrt(1/2), sqrt(2))
MPP_HD void log_reduce(double x, double &f, double &dk)
{
  int hi; unsigned lo;
  mpp_split(x, hi, lo);
  int k = (hi >> 20) - 1023;
  hi &= 0x000fffff;
  const int i = (hi + 0x95f64) & 0x100000;                    // mantissa >= sqrt(2): halve it, bump the exponent
  hi |= (i ^ 0x3ff00000);
  k += (i >> 20);
  f = mpp_join(hi, lo) - 1.0;
  dk = (double)k;
}
MPP_HD void log_reduce(d2 x, d2 &f, d2 &dk) { log_reduce(x.a, f.a, dk.a); log_reduce(x.b, f.b, dk.b); }

// natural logarithm of positive, finite, normal doubles
template <class T> MPP_HD T mpp_log_poly(T x)
{
  T f, dk;
  log_reduce(x, f, dk);
  const T s = f * rcp(vbc<T>(2.0) + f);
  const T z = s * s, w = z * z;
  const T t1 = w * vfma(w, vfma(w, vbc<T>(MPPC(5)), vbc<T>(MPPC(3))), vbc<T>(MPPC(1)));
  const T t2 = z * vfma(w, vfma(w, vfma(w, vbc<T>(MPPC(6)), vbc<T>(MPPC(4))), vbc<T>(MPPC(2))), vbc<T>(MPPC(0)));
  const T R = t2 + t1;
  const T hfsq = vbc<T>(0.5) * f * f;
  return dk * vbc<T>(MPPC(7)) - ((hfsq - (s * (hfsq + R) + dk * vbc<T>(MPPC(8)))) - f);
}

// exp: clamp to [-708, 708] on the high word (one integer compare + selects), and 2^k from the magic-number sum
MPP_HD double exp_clamp(double x)
{
  int xh; unsigned xl;
  mpp_split(x, xh, xl);
  if ((xh & 0x7fffffff) >= 0x40862000) x = (xh < 0) ? -708.0 : 708.0;     // |x| >= 708
  return x;
}
MPP_HD d2 exp_clamp(d2 x) { return d2{exp_clamp(x.a), exp_clamp(x.b)}; }
MPP_HD double exp_scale(double fn)
{
  int hi; unsigned lo;
  mpp_split(fn, hi, lo);                                       // fn = 1.5 * 2^52 + round(x / ln2): the low word holds the integer
  return mpp_join(((int)lo + 1023) << 20, 0u);
}
MPP_HD d2 exp_scale(d2 fn) { return d2{exp_scale(fn.a), exp_scale(fn.b)}; }

// exponential; the argument is clamped to [-708, 708] (results stay normal and finite)
template <class T> MPP_HD T mpp_exp_poly(T x)
{
  x = exp_clamp(x);
  const T magic = vbc<T>(6755399441055744.0);
  const T fn = vfma(x, vbc<T>(MPPC(19)), magic);
  const T kd = fn - magic;
  T r = vfma(-kd, vbc<T>(MPPC(20)), x);
  r = vfma(-kd, vbc<T>(MPPC(21)), r);                          // |r| <= ln2 / 2
  const T r2 = r * r;
  const T p0 = vfma(r, vbc<T>(MPPC(10)), vbc<T>(MPPC(9))),  p1 = vfma(r, vbc<T>(MPPC(12)), vbc<T>(MPPC(11)));
  const T p2 = vfma(r, vbc<T>(MPPC(14)), vbc<T>(MPPC(13))), p3 = vfma(r, vbc<T>(MPPC(16)), vbc<T>(MPPC(15)));
  const T p4 = vfma(r, vbc<T>(MPPC(18)), vbc<T>(MPPC(17)));
  T q = vfma(r2, p4, p3);
  q = vfma(r2, q, p2); q = vfma(r2, q, p1); q = vfma(r2, q, p0);
  const T e = vfma(r2, q, r) + vbc<T>(1.0);
  return e * exp_scale(fn);
}

// ------------------------------------------------------------------------------------------------
// Table-driven log / exp (round 2).  The polynomial versions above cost 28 / 18 fp64 instructions and a full-precision
// reciprocal inside the log; the step kernels are bound by fp64 issue and by the length of exactly these dependent chains
// (profiles/r1_vsfm_v13.md), and each cell evaluation needs two of each.  With a 128-entry table (tools/gen_math_tables.py):
//   log x:  x = 2^k z, z in [0.6875, 1.375); interval i = top 7 mantissa bits of z; r = z invc_i - 1 (one exact-product FMA,
//           |r| <= 2^-8); log x = k ln2 + logc_i + log1p(r), log1p by its Taylor series through r^7 (truncation < 2^-59 |r|);
//           logc_i = -log(invc_i) for the STORED invc_i, so the identity is exact.  12 fp64 instructions, no reciprocal.
//   exp x:  k = round(128 x / ln2), r = x - k ln2/128 (|r| <= ln2/256), exp x = 2^(k>>7) T[k & 127] (1 + r + r^2/2 + ... + r^5/120)
//           (truncation < 2^-60).  10 fp64 instructions.
// The tables (3 KB) sit in global memory and are read through the read-only L1 path; lanes index them independently.  (A per-block copy
// in shared memory measured slower: 2.21 vs 2.05 ms per Mi columns in the VSFM step kernel, round 2.)
// Accuracy against libm: < 1.5 ulp over the ranges the soil curves produce (tests/test_physics_host.py).
// ------------------------------------------------------------------------------------------------
#ifdef __CUDACC__
static __device__ const double mpp_log_tab_dev[256] = MPP_LOG_TABLE;
static __device__ const double mpp_exp_tab_dev[128] = MPP_EXP_TABLE;
#endif
static const double mpp_log_tab_host[256] = MPP_LOG_TABLE;
static const double mpp_exp_tab_host[128] = MPP_EXP_TABLE;

MPP_HD void log_tab_reduce(double x, double &r, double &w)
{
  int hi; unsigned lo;
  mpp_split(x, hi, lo);
  const int tmp = hi - 0x3fe60000;
  const int i = (tmp >> 13) & 127;
  const int k = tmp >> 20;                                     // arithmetic shift: floor
  const double z = mpp_join(hi - (tmp & (int)0xfff00000), lo);
#ifdef __CUDA_ARCH__
  const double2 t = __ldg(reinterpret_cast<const double2 *>(mpp_log_tab_dev) + i);
  const double invc = t.x, logc = t.y;
#else
  const double invc = mpp_log_tab_host[2 * i], logc = mpp_log_tab_host[2 * i + 1];
#endif
  r = fma(z, invc, -1.0);
  w = fma((double)k, MPPC(28), logc);                          // exact: ln2_hi has 21 trailing zero bits, |k| < 2^11
  w = fma((double)k, MPPC(29), w);
}
MPP_HD void log_tab_reduce(d2 x, d2 &r, d2 &w) { log_tab_reduce(x.a, r.a, w.a); log_tab_reduce(x.b, r.b, w.b); }

// natural logarithm of positive, finite, normal doubles
template <class T> MPP_HD T mpp_log_tab(T x)
{
  T r, w;
  log_tab_reduce(x, r, w);
  const T r2 = r * r;
  const T q01 = vfma(r, vbc<T>(MPPC(23)), vbc<T>(MPPC(22)));
  const T q23 = vfma(r, vbc<T>(MPPC(25)), vbc<T>(MPPC(24)));
  const T q45 = vfma(r, vbc<T>(MPPC(27)), vbc<T>(MPPC(26)));
  T t = vfma(r2, q45, q23);
  t = vfma(r2, t, q01);
  return w + vfma(r2, t, r);
}

MPP_HD double exp_tab_scale(double fn)
{
  int hi; unsigned lo;
  mpp_split(fn, hi, lo);                                       // fn = 1.5 * 2^52 + round(128 x / ln2): the low word holds the integer k
  const int k = (int)lo;
#ifdef __CUDA_ARCH__
  const double T = __ldg(mpp_exp_tab_dev + (k & 127));
#else
  const double T = mpp_exp_tab_host[k & 127];
#endif
  int th; unsigned tl;
  mpp_split(T, th, tl);
  return mpp_join(th + ((k >> 7) << 20), tl);                  // 2^(k >> 7) T[k & 127]; stays normal for |x| <= 708
}
MPP_HD d2 exp_tab_scale(d2 fn) { return d2{exp_tab_scale(fn.a), exp_tab_scale(fn.b)}; }

// exponential; the argument is clamped to [-708, 708] (results stay normal and finite)
template <class T> MPP_HD T mpp_exp_tab(T x)
{
  x = exp_clamp(x);
  const T magic = vbc<T>(MPPC(37));
  const T fn = vfma(x, vbc<T>(MPPC(30)), magic);
  const T kd = fn - magic;
  T r = vfma(-kd, vbc<T>(MPPC(31)), x);                        // exact: ln2/128_hi has 24 trailing zero bits, |k| < 2^18
  r = vfma(-kd, vbc<T>(MPPC(32)), r);
  const T s = exp_tab_scale(fn);
  const T r2 = r * r;
  const T q1 = vfma(r, vbc<T>(MPPC(34)), vbc<T>(MPPC(33)));
  const T q2 = vfma(r, vbc<T>(MPPC(36)), vbc<T>(MPPC(35)));
  const T t = vfma(r2, q2, q1);
  const T p = vfma(r2, t, r);
  return vfma(s, p, s);
}

#ifndef MPP_POLY_MATH
template <class T> MPP_HD T mpp_log(T x) { return mpp_log_tab(x); }
template <class T> MPP_HD T mpp_exp(T x) { return mpp_exp_tab(x); }
#else
template <class T> MPP_HD T mpp_log(T x) { return mpp_log_poly(x); }
template <class T> MPP_HD T mpp_exp(T x) { return mpp_exp_poly(x); }
#endif

// pow() kept out of line for the rare calls on hot paths (the reference's `dhsdT**area` with area != 1): inlined, the CUDA library's pow
// hoists ~50 instructions of argument classification above the guarding branch, which every lane then executes on every call
// (profiles/r2_thermal_v1.md: 10 % of the soil thermal kernel's instructions)
#ifdef __CUDACC__
static __device__ __noinline__ double mpp_pow_rare(double x, double y) { return pow(x, y); }
#endif

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
  double Se, AA, AAm, L2, rx, pc;   // VG intermediates (rx = 1 / (-alpha pc)); for BC/SBC: Se, pc
  int    regime;                 // 0 saturated, 1 unsaturated (VG / full BC), 2 SBC cubic
};

// van Genuchten - Mualem curves, generic over double / d2 (see the pair types above).
// Reference: Se = (1 + x^n)^-m, AA = x^n / (1 + x^n), kr = sqrt(Se) (1 - AA^m)^2 with x = -alpha pc for pc < 0,
// sat = kr = 1 otherwise (5 pow calls, SaturationFunction.F90:776-836).  Here: 2 log + 2 exp + 1 reciprocal, using
//   sqrt(Se) = (1 + x^n)^(-m/2),  AA^m = x^(n m) Se = x^(n-1) Se = x^n Se / x      (n m = n - 1).
// Branch-free on purpose: saturated cells evaluate the chain at x = 1 and discard it, so the chains of the cells a lane
// owns stay in one basic block and overlap.
template <class T> struct VGState { T sat, kr, Se, pcn, AAm, rS, rx; };

template <class T>
MPP_HD void vg_values(T sat_res, T alpha, T m, T n, T pc, bool unsat_a, bool unsat_b, VGState<T> &o)
{
  const T one = vbc<T>(1.0);
  const T x   = vsel(unsat_a, unsat_b, -alpha * pc, one);
  const T L1  = mpp_log(x);
  const T rx  = rcp(x);
  const T pcn = mpp_exp(n * L1);                      // x^n
  const T L2  = mpp_log(one + pcn);
  const T rS  = mpp_exp(vbc<T>(-0.5) * m * L2);       // sqrt(Se)
  const T Se  = rS * rS;
  const T AAm = pcn * Se * rx;                        // AA^m
  const T BB  = one - AAm;
  o.sat = vsel(unsat_a, unsat_b, sat_res + (one - sat_res) * Se, one);
  o.kr  = vsel(unsat_a, unsat_b, rS * BB * BB, one);
  o.Se = Se; o.pcn = pcn; o.AAm = AAm; o.rS = rS; o.rx = rx;
}

// dSe/dpc = -m n Se AA / pc with AA = x^n / (1 + x^n)                       (SaturationFunction.F90:790, 829)
// dkr/dSe = kr / (2 Se) + 2 Se^(1/m - 1/2) AA^(m-1) BB                       (:836-838)
//         = BB (BB / 2 + 2 AA^m / x^n) / sqrt(Se)      since Se^(1/m) = 1 / (1 + x^n)
// (same functions, regrouped so that no further exp/pow is needed; derivative round-off only steers the Newton path,
//  never the converged answer).  1 / pc = -alpha / x and AA^m / x^n = Se / x: two reciprocals, no division.
template <class T>
MPP_HD void vg_derivs(T sat_res, T alpha, T m, T n, bool unsat_a, bool unsat_b, const VGState<T> &o, T &dsat_dP, T &dkr_dP)
{
  const T one = vbc<T>(1.0), zero = vbc<T>(0.0);
  const T BB = one - o.AAm;
  const T dSe_dpc = vsel(unsat_a, unsat_b, (m * n) * o.Se * o.pcn * rcp(one + o.pcn) * (alpha * o.rx), zero);
  const T dkr_dSe = BB * (vbc<T>(0.5) * BB + vbc<T>(2.0) * o.Se * o.rx) * rcp(o.rS);
  dsat_dP = (one - sat_res) * dSe_dpc;
  dkr_dP  = dkr_dSe * dSe_dpc;
}

template <int SATFUNC>
MPP_HD void sat_values(const SatParams &sp, double press, double frac_liq, SatState &s)
{
  const double pc = press - PRESSURE_REF;
  s.pc = pc;
  if (SATFUNC == SATFUNC_VG) {
    VGState<double> o;
    const bool unsat = (pc < 0.0);
    vg_values<double>(sp.sat_res, sp.alpha, sp.m, sp.n, pc, unsat, unsat, o);
    s.sat = o.sat; s.kr = o.kr;
    s.Se = o.Se; s.AA = o.pcn; s.AAm = o.AAm; s.L2 = o.rS; s.rx = o.rx; s.regime = unsat ? 1 : 0;   // (AA slot carries x^n, L2 slot carries sqrt(Se))
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
  if (SATFUNC == SATFUNC_VG) {
    VGState<double> o;
    o.Se = s.Se; o.pcn = s.AA; o.AAm = s.AAm; o.rS = s.L2; o.rx = s.rx;
    vg_derivs<double>(sp.sat_res, sp.alpha, sp.m, sp.n, s.regime != 0, s.regime != 0, o, dsat_dP, dkr_dP);
  } else {
    if (s.regime == 0) { dsat_dP = 0.0; dkr_dP = 0.0; return; }
    double dSe_dpc;
    if (s.regime == 1) dSe_dpc = -sp.m * s.Se / s.pc;
    else { const double dpc = s.pc - sp.ps; dSe_dpc = dpc * (2.0 * sp.b2 + 3.0 * dpc * sp.b3); }
    // s.kr carries frac_liq for BC; dkr_dSe = (2.5 + 2/lambda) kr / Se
    dsat_dP = (1.0 - sp.sat_res) * dSe_dpc;
    dkr_dP  = (2.5 + 2.0 / sp.m) * s.kr / s.Se * dSe_dpc;
    (void)frac_liq;
  }
}

// Brooks-Corey and smoothed Brooks-Corey, generic over double / d2 and branch-free (the fast VSFM kernel): one lean log and two lean
// exp per cell whatever the regime.  With Lse = ln(Se):  regime A (pc <= pu; plain BC: -alpha pc > 1)  Lse = -lambda ln(-alpha pc);
// regime B (SBC cubic, pu < pc < ps)  Se = 1 + dpc^2 (b2 + dpc b3), Lse = ln(Se);  saturated  Lse = 0.  kr = exp((2.5 + 2/lambda) Lse).
// Same functions as sat_values<SATFUNC_BC / SATFUNC_SBC> (SaturationFunction.F90:900-1140), regrouped.
template <class T>
MPP_HD void bc_sbc_values(T sat_res, T alpha, T lam, T ps, T b2, T b3, T pc, bool Aa, bool Ab, bool Ba, bool Bb, T &sat, T &kr, T &Se)
{
  const T one = vbc<T>(1.0);
  const T dpc = pc - ps;
  const T SeB = one + dpc * dpc * (b2 + dpc * b3);
  const T arg = vsel(Aa, Ab, -alpha * pc, vsel(Ba, Bb, SeB, one));
  const T L = mpp_log(arg);
  const T Lse = vsel(Aa, Ab, -lam * L, L);
  Se = vsel(Aa, Ab, mpp_exp(Lse), vsel(Ba, Bb, SeB, one));
  kr = mpp_exp((vbc<T>(2.5) + vbc<T>(2.0) * rcp(lam)) * Lse);
  sat = vsel(Aa || Ba, Ab || Bb, sat_res + (one - sat_res) * Se, one);
}
template <class T>
MPP_HD void bc_sbc_derivs(T sat_res, T lam, T ps, T b2, T b3, T pc, T Se, T kr, bool Aa, bool Ab, bool Ba, bool Bb, T &dsat_dP, T &dkr_dP)
{
  const T one = vbc<T>(1.0), zero = vbc<T>(0.0);
  const T dpc = pc - ps;
  const T pcs = vsel(Aa, Ab, pc, one);                                  // keep the reciprocal finite outside regime A
  const T dSe = vsel(Aa, Ab, -lam * Se * rcp(pcs), vsel(Ba, Bb, dpc * (vbc<T>(2.0) * b2 + vbc<T>(3.0) * dpc * b3), zero));
  dsat_dP = (one - sat_res) * dSe;
  dkr_dP  = (vbc<T>(2.5) + vbc<T>(2.0) * rcp(lam)) * kr * rcp(Se) * dSe;   // kr carries frac_liq for BC
}

// two cells at once (the fast VSFM kernel): van Genuchten goes through the pair instantiation, the others call the scalar code twice
template <int SATFUNC>
MPP_HD void sat_values_pair(const SatParams &pa, const SatParams &pb, double Pa, double Pb, double fla, double flb, SatState &sa, SatState &sb)
{
  if (SATFUNC == SATFUNC_VG) {
    VGState<d2> o;
    const d2 pc = d2{Pa - PRESSURE_REF, Pb - PRESSURE_REF};
    const bool ua = (pc.a < 0.0), ub = (pc.b < 0.0);
    vg_values<d2>(d2{pa.sat_res, pb.sat_res}, d2{pa.alpha, pb.alpha}, d2{pa.m, pb.m}, d2{pa.n, pb.n}, pc, ua, ub, o);
    sa.pc = pc.a; sa.sat = o.sat.a; sa.kr = o.kr.a; sa.Se = o.Se.a; sa.AA = o.pcn.a; sa.AAm = o.AAm.a; sa.L2 = o.rS.a; sa.rx = o.rx.a; sa.regime = ua ? 1 : 0;
    sb.pc = pc.b; sb.sat = o.sat.b; sb.kr = o.kr.b; sb.Se = o.Se.b; sb.AA = o.pcn.b; sb.AAm = o.AAm.b; sb.L2 = o.rS.b; sb.rx = o.rx.b; sb.regime = ub ? 1 : 0;
  } else {
    const d2 pc = d2{Pa - PRESSURE_REF, Pb - PRESSURE_REF};
    bool Aa, Ab, Ba = false, Bb = false;
    if (SATFUNC == SATFUNC_BC) { Aa = (-pa.alpha * pc.a > 1.0); Ab = (-pb.alpha * pc.b > 1.0); }
    else { Aa = (pc.a <= pa.pu); Ab = (pc.b <= pb.pu); Ba = !Aa && (pc.a < pa.ps); Bb = !Ab && (pc.b < pb.ps); }
    d2 sat, kr, Se;
    bc_sbc_values<d2>(d2{pa.sat_res, pb.sat_res}, d2{pa.alpha, pb.alpha}, d2{pa.m, pb.m}, d2{pa.ps, pb.ps}, d2{pa.b2, pb.b2}, d2{pa.b3, pb.b3},
                      pc, Aa, Ab, Ba, Bb, sat, kr, Se);
    if (SATFUNC == SATFUNC_BC) { kr.a = fla * kr.a; kr.b = flb * kr.b; }    // SaturationFunction.F90:987
    sa.pc = pc.a; sa.sat = sat.a; sa.kr = kr.a; sa.Se = Se.a; sa.regime = Aa ? 1 : (Ba ? 2 : 0);
    sb.pc = pc.b; sb.sat = sat.b; sb.kr = kr.b; sb.Se = Se.b; sb.regime = Ab ? 1 : (Bb ? 2 : 0);
  }
}
template <int SATFUNC>
MPP_HD void sat_derivs_pair(const SatParams &pa, const SatParams &pb, const SatState &sa, const SatState &sb, double fla, double flb,
                            double &dsat_a, double &dkr_a, double &dsat_b, double &dkr_b)
{
  if (SATFUNC == SATFUNC_VG) {
    VGState<d2> o;
    o.Se = d2{sa.Se, sb.Se}; o.pcn = d2{sa.AA, sb.AA}; o.AAm = d2{sa.AAm, sb.AAm}; o.rS = d2{sa.L2, sb.L2}; o.rx = d2{sa.rx, sb.rx};
    d2 ds, dk;
    vg_derivs<d2>(d2{pa.sat_res, pb.sat_res}, d2{pa.alpha, pb.alpha}, d2{pa.m, pb.m}, d2{pa.n, pb.n}, sa.regime != 0, sb.regime != 0, o, ds, dk);
    dsat_a = ds.a; dkr_a = dk.a; dsat_b = ds.b; dkr_b = dk.b;
  } else {
    d2 ds, dk;
    bc_sbc_derivs<d2>(d2{pa.sat_res, pb.sat_res}, d2{pa.m, pb.m}, d2{pa.ps, pb.ps}, d2{pa.b2, pb.b2}, d2{pa.b3, pb.b3}, d2{sa.pc, sb.pc},
                      d2{sa.Se, sb.Se}, d2{sa.kr, sb.kr}, sa.regime == 1, sb.regime == 1, sa.regime == 2, sb.regime == 2, ds, dk);
    dsat_a = ds.a; dkr_a = dk.a; dsat_b = ds.b; dkr_b = dk.b;
    (void)fla; (void)flb;
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
// Fast variant for the TH step kernel: split into the temperature part (shared by every pressure at which the density is
// needed -- the TH aux vars evaluate it at P and at max(P, P_ref)) and the pressure part; divisions by constants and by
// (t_c + a4) are reciprocal multiplies (1-ulp differences from the reference order below).
struct TanakaT { double dent, ddent_dt, kc, dkc_dt; };
MPP_HD TanakaT tanaka_T(double t_K)
{
  const double a1 = -3.983035, a2 = 301.797, a3 = 522528.9, a4 = 69.34881, a5 = 999.974950;
  const double k0 = 50.74e-11, k1 = -0.326e-11, k2 = 0.00416e-11;
  const double t_c = t_K - 273.15;
  const double r4 = rcp(t_c + a4), ra3 = 1.0 / a3;
  const double s1 = (t_c + a1) * (t_c + a1);
  const double q  = (t_c + a2) * ra3 * r4;
  TanakaT o;
  o.dent = a5 * (1.0 - s1 * q);
  o.kc = k0 + k1 * t_c + k2 * t_c * t_c;
  o.dkc_dt = k1 + 2.0 * k2 * t_c;
  const double d1 = -s1 * ra3 * r4;
  const double d2 = -2.0 * (t_c + a1) * q;
  const double d3 = s1 * q * r4;
  o.ddent_dt = a5 * (d1 + d2 + d3);
  return o;
}
MPP_HD void tanaka_P(const TanakaT &t, double p, double &den, double &dden_dp, double &dden_dT)
{
  const double p0 = 101325.0, rfmw = 1.0 / FMWH2O;
  const bool comp = (p > p0);                                   // compressible only above P_ref (EOSWaterMod.F90:151-155)
  const double kappa = comp ? 1.0 + t.kc * (p - p0) : 1.0;
  const double dkappa_dp = comp ? t.kc : 0.0, dkappa_dt = comp ? t.dkc_dt * (p - p0) : 0.0;
  den = t.dent * kappa * rfmw;
  dden_dT = (t.ddent_dt * kappa + t.dent * dkappa_dt) * rfmw;
  dden_dp = (t.dent * dkappa_dp) * rfmw;
}
// reference operation order (divisions kept): used wherever the value is computed once -- the VSFM density table, the
// generic kernels, the host-side checks against the oracle, which it matches bit for bit
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
  const bool comp = (t.type == DENSITY_TGDPB01) && (p > 101325.0);          // compressible only above P_ref (EOSWaterMod.F90:151-155)
  den     = comp ? t.dent_over_fmw * (1.0 + t.kcoef * (p - 101325.0)) : t.dent_over_fmw;
  dden_dp = comp ? t.dent_over_fmw * t.kcoef : 0.0;
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

// density at the VSFM aux vars' fixed 298.15 K for every density model; EXT = false compiles the IFC-67 polynomial out (the common
// specialisation of the step kernel), EXT = true dispatches on the run-time type
// On the device the IFC-67 evaluation stays out of line: the step kernel calls this at seven sites inside its Newton loop, and seven
// inlined copies of the polynomial were most of the 60 KB that loop spanned (a 32 KB instruction cache).
#ifdef __CUDACC__
static __device__ __noinline__ void density_ifc67_25C_rare(double p, double &den, double &dden_dp) { double dT; density_ifc67(25.0, p, den, dden_dp, dT); }
#endif
template <bool EXT>
MPP_HD void density_fixedT_x(const DensityTable &t, double p, double &den, double &dden_dp)
{
  if (EXT && t.type == DENSITY_IFC67) {
#ifdef __CUDA_ARCH__
    density_ifc67_25C_rare(p, den, dden_dp);
#else
    double dT; density_ifc67(25.0, p, den, dden_dp, dT);
#endif
  }
  else density_fixedT(t, p, den, dden_dp);
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
    const double r = rcp(den), Pr2 = P * r * r;        // one lean reciprocal instead of four divisions
    H = U + P * r;
    dH_dT = dU_dT - Pr2 * dden_dT;
    dH_dP = dU_dP + r - Pr2 * dden_dP;
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
