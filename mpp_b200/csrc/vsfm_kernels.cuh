// vsfm_kernels.cuh -- fused VSFM (Richards equation) time step for batches of independent soil columns.
//
// One launch = one sysofeqns%StepDT for every column of the batch:
//   SOEBaseStepDT_SNES        src/mpp/soe/SystemOfEquationsBaseType.F90:368-552  (dt cuts, <= 20)
//   VSFMSOEPreSolve/PostSolve src/mpp/soe/SystemOfEquationsVSFMType.F90:506-660
//   VSFMSOEResidual/Jacobian  src/mpp/soe/SystemOfEquationsVSFMType.F90:94-403
//   Richards residual/Jacobian src/mpp/ge/GoveqnRichardsODEPressureType.F90:1603-2200
//   RichardsFlux              src/mpp/ge/RichardsMod.F90:118-340
//   PETSc SNES newtonls + bt line search + SNESConvergedDefault, KSP on a tridiagonal matrix
//
// Mapping (B200-first; see DESIGN.md "VSFM kernel"): one LANE per soil cell, GROUP (16 or 32) lanes per
// column, so a warp advances 2 (or 1) columns.  Per-cell soil parameters, state and the Jacobian row live in
// registers for the whole Newton loop; neighbour cells are reached with warp shuffles; the tridiagonal
// Newton system is solved by parallel cyclic reduction across the group (log2(GROUP) shuffle steps);
// norms are butterfly all-reduces (bitwise identical in every lane of the group, so all control flow is
// group-uniform).  Converged columns drop out (their group idles until the warp's other column is done).
// HBM traffic is the algorithmic minimum: every input array is read once, every output written once,
// in the reference's own cell order (icell = c*nlev + j), fully coalesced.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "physics.cuh"

namespace mpp {

#ifndef VSFM_MIN_BLOCKS
#define VSFM_MIN_BLOCKS 4
#endif
constexpr int MAX_SS = 16;
constexpr int MAX_BC = 2;

// PETSc SNESConvergedReason codes the reference's drivers branch on
enum { SNES_ITERATING = 0, SNES_CONVERGED_FNORM_ABS = 2, SNES_CONVERGED_FNORM_RELATIVE = 3, SNES_CONVERGED_SNORM_RELATIVE = 4,
       SNES_DIVERGED_FUNCTION_COUNT = -2, SNES_DIVERGED_FNORM_NAN = -4, SNES_DIVERGED_MAX_IT = -5,
       SNES_DIVERGED_LINE_SEARCH = -6, SNES_DIVERGED_DTOL = -9 };

enum { REGION_TOP = 401, REGION_BOTTOM = 402, REGION_CELLS = 403 };
enum { CT_MASS_RATE = 503, CT_DIRICHLET = 505, CT_SEEPAGE = 509 };

struct CondDev {
  const double *value;     // SoE mailbox condition_value: per column (top/bottom regions) or per cell (REGION_CELLS)
  int itype, region;
  double *flux;            // boundary_flux [kg/s] (BCs only)
  double *mass_exc;        // boundary mass exchanged [kg] (BCs only)
};

struct SnesOpts {
  double atol, rtol, stol, divtol;
  int max_it, max_funcs;
  double ls_alpha, ls_minlambda, ls_maxstep;
  int ls_max_its;
};

struct VsfmArgs {
  int ncol, nlev;
  double uz;               // z component of the internal-connection unit vector (-1 along gravity, +1 against, 0 horizontal)
  int top_is_first;        // the SOIL_TOP cell is j = 0 (along gravity / horizontal) or j = nlev-1 (against gravity)
  DensityTable dtab;
  // static per-cell soil parameters (cell order)
  const double *por, *perm, *sat_res, *alpha, *lam, *vgn, *pu, *ps, *b2, *b3, *dz;
  const double *area;      // per column
  const int *active;       // per column or nullptr
  const double *frac_liq;  // per cell (SoE mailbox frac_liq_sat)
  const double *x_in;      // soln == soln_prev at StepDT entry
  double *x_out;           // soln after the step (may alias x_in)
  int nss, nbc;
  CondDev ss[MAX_SS], bc[MAX_BC];
  // SoE mailbox outputs (VSFMSOEPostSolve)
  double *liq_sat, *pressure, *mass, *smp;
  // per-column diagnostics
  int *stat_its, *stat_reason, *stat_cuts, *stat_nf;
  double *col_mass;        // in: sum_j mass at the last PostSolve; out: new sum
  double *col_err;         // |m_beg - m_end + q dt| (MPPVSFMALM_Driver.F90:860-863)
  double *col_src;         // sum of mass-rate sources [kg/s]
  double *block_partials;  // gridDim.x * 8 doubles: deterministic two-stage reduction
  double dt;
  SnesOpts so;
};

enum { PH_INIT = 0, PH_NEWTON = 1, PH_LS_FULL = 2, PH_LS_QUAD = 3, PH_LS_CUBIC = 4, PH_DONE = 5 };

// All shuffles in this file use the compile-time full mask and are executed by the whole warp under
// warp-uniform control flow: partial (run-time) masks make nvcc wrap every SHFL in a MATCH/WARPSYNC sequence
// (measured: 26% of all executed instructions in the first version, profiles/r1_vsfm_first.md).
constexpr unsigned FULL_MASK = 0xffffffffu;

template <int GROUP>
__device__ __forceinline__ double group_sum(double v)
{
#pragma unroll
  for (int s = GROUP / 2; s > 0; s >>= 1) v += __shfl_xor_sync(FULL_MASK, v, s, GROUP);
  return v;
}

// Parallel cyclic reduction of a tridiagonal system with one row per lane (identity rows pad the group).
// The a/c couplings that would reach outside the chain are exactly zero at every stage, so the values that
// out-of-range shuffles return are multiplied by zero.
template <int GROUP>
__device__ __forceinline__ double pcr_solve(double a, double b, double c, double d)
{
  const unsigned mask = FULL_MASK;
#pragma unroll
  for (int s = 1; s < GROUP; s <<= 1) {
    const double r   = __drcp_rn(b);
    const double a_m = __shfl_up_sync(mask, a, s, GROUP),   c_m = __shfl_up_sync(mask, c, s, GROUP);
    const double d_m = __shfl_up_sync(mask, d, s, GROUP),   r_m = __shfl_up_sync(mask, r, s, GROUP);
    const double a_p = __shfl_down_sync(mask, a, s, GROUP), c_p = __shfl_down_sync(mask, c, s, GROUP);
    const double d_p = __shfl_down_sync(mask, d, s, GROUP), r_p = __shfl_down_sync(mask, r, s, GROUP);
    const double k1 = a * r_m, k2 = c * r_p;
    b = b - c_m * k1 - a_p * k2;
    d = d - d_m * k1 - d_p * k2;
    a = -a_m * k1;
    c = -c_p * k2;
  }
  return d * __drcp_rn(b);
}

template <int GROUP, int SATFUNC>
__global__ void __launch_bounds__(128, VSFM_MIN_BLOCKS)
vsfm_step_kernel(const VsfmArgs A)
{
  const int tid   = blockIdx.x * blockDim.x + threadIdx.x;
  const int col   = tid / GROUP;
  const int j     = tid % GROUP;                       // lane within the group == layer index
  const int lane  = threadIdx.x & 31;
  constexpr unsigned FULL = FULL_MASK;
  constexpr double RVIS = 1.0 / VISCOSITY, RFMW = 1.0 / FMWH2O;
  const int nlev  = A.nlev;
  const bool col_ok = (col < A.ncol) && (A.active == nullptr || A.active[col] != 0);
  const bool valid  = col_ok && (j < nlev);
  const long long cell = (long long)col * nlev + j;

  // ---- static per-cell data ---------------------------------------------------------------------
  SatParams sp; sp.sat_res = 0.0; sp.alpha = 1.0; sp.m = 0.5; sp.n = 2.0; sp.pu = sp.ps = sp.b2 = sp.b3 = 0.0;
  double por = 0.0, perm = 1.0, dz = 1.0, area = 1.0, frac_liq = 1.0, X = PRESSURE_REF;
  if (valid) {
    por = A.por[cell]; perm = A.perm[cell]; dz = A.dz[cell]; area = A.area[col];
    sp.sat_res = A.sat_res[cell]; sp.alpha = A.alpha[cell]; sp.m = A.lam[cell];
    if (SATFUNC == SATFUNC_VG)  sp.n = A.vgn[cell];
    if (SATFUNC == SATFUNC_SBC) { sp.pu = A.pu[cell]; sp.ps = A.ps[cell]; sp.b2 = A.b2[cell]; sp.b3 = A.b3[cell]; }
    frac_liq = A.frac_liq[cell];
    X = A.x_in[cell];
  }
  const double vol = area * dz;                                       // MeshType.F90:427

  // internal connection j -> j+1, owned by lane j (MeshType.F90:509-530; RichardsMod.F90:257-259,279-285)
  const double perm_dn = __shfl_down_sync(FULL, perm, 1, GROUP);
  const double dz_dn   = __shfl_down_sync(FULL, dz, 1, GROUP);
  const bool has_conn  = valid && (j < nlev - 1);
  const double dist_up = 0.5 * dz, dist_dn = 0.5 * dz_dn;
  const double upw     = dist_up / (dist_up + dist_dn);
  const double Dq      = (perm * perm_dn) / (dist_up * perm_dn + dist_dn * perm);
  const double gfac    = FMWH2O * ((dist_up + dist_dn) * (A.uz * (-GRAVITY_CONSTANT)));   // FMWH2O * dist_gravity

  // boundary conditions (MeshType.F90:723-806): top -> unit vector (0,0,-1), bottom -> (0,0,+1); dist_up = 0
  const int jtop = A.top_is_first ? 0 : nlev - 1, jbot = A.top_is_first ? nlev - 1 : 0;
  double bcP[MAX_BC], bcKr[MAX_BC], bcGfac[MAX_BC], bcMassExc[MAX_BC], bcFlux[MAX_BC];
  bool   bcMine[MAX_BC];
  const double DqBC = perm / (0.0 + 0.5 * dz);
#pragma unroll
  for (int k = 0; k < MAX_BC; ++k) {
    bcMine[k] = false; bcP[k] = PRESSURE_REF; bcKr[k] = 1.0; bcGfac[k] = 0.0; bcMassExc[k] = 0.0; bcFlux[k] = 0.0;
    if (k < A.nbc) {
      const bool top = (A.bc[k].region == REGION_TOP);
      bcMine[k] = valid && (j == (top ? jtop : jbot));
      if (bcMine[k]) {
        const double uzbc = (A.uz == 0.0) ? 0.0 : (top ? -1.0 : 1.0);
        bcGfac[k] = FMWH2O * ((0.0 + 0.5 * dz) * (uzbc * (-GRAVITY_CONSTANT)));
        bcP[k] = A.bc[k].value[col];
        SatState sb;
        sat_values<SATFUNC>(sp, bcP[k], 1.0, sb);       // BC aux vars keep frac_liq_sat = 1 (RichardsODEPressureAuxType.F90:93)
        bcKr[k] = sb.kr;
      }
    }
  }

  // mass-rate source/sinks (GoveqnRichards...:1871-1875): F -= value / FMWH2O
  double src = 0.0, src_kg = 0.0;
  for (int k = 0; k < A.nss; ++k) {
    const CondDev &c = A.ss[k];
    bool mine = false; long long idx = 0;
    if (c.region == REGION_CELLS) { mine = valid; idx = cell; }
    else { mine = valid && (j == (c.region == REGION_TOP ? jtop : jbot)); idx = col; }
    if (mine) { const double v = c.value[idx]; src += v * RFMW; src_kg += v; }
  }

  // ---- time-step / Newton state (group-uniform unless noted) -------------------------------------
  const SnesOpts so = A.so;
  double Xprev = X;                       // soln_prev
  double dt_iter = A.dt, time_done = 0.0;
  double dtInv = 1.0 / dt_iter;
  int    cuts = 0, tot_its = 0, tot_nf = 0, last_reason = 0, converged = 0;
  int    phase = col_ok ? PH_INIT : PH_DONE;
  int    its = 0, nfuncs = 0, ls_count = 0;
  double accum_prev = 0.0;
  double F = 0.0, Y = 0.0, W = X, G = 0.0;
  double fnorm = 0.0, xnorm = 0.0, ynorm = 0.0, ttol = 0.0, rnorm0 = 0.0;
  double f2 = 0.0, initslope = -1.0, lambda = 1.0, lambdaprev = 1.0, gprev = 0.0;
  // aux vars at the accepted point X (per lane)
  double kr = 1.0, den = 1.0, dden = 0.0, sat = 1.0, dsat = 0.0, dkr = 0.0;

  for (;;) {
    // ================= Newton step set-up: Jacobian, linear solve, line-search initialisation =================
    // Executed by the WHOLE warp whenever either of its columns starts a Newton iteration (warp-uniform branch, so
    // every shuffle below is convergent); lanes of a column that is not in PH_NEWTON compute and discard.
    if (__any_sync(FULL, phase == PH_NEWTON)) {
      const bool nw = (phase == PH_NEWTON);
      // neighbour (dn) state of connection j
      const double P_d    = __shfl_down_sync(FULL, X, 1, GROUP);
      const double kr_d   = __shfl_down_sync(FULL, kr, 1, GROUP);
      const double den_d  = __shfl_down_sync(FULL, den, 1, GROUP);
      const double dkr_d  = __shfl_down_sync(FULL, dkr, 1, GROUP);
      const double dden_d = __shfl_down_sync(FULL, dden, 1, GROUP);
      double Jup = 0.0, Jdn = 0.0;
      if (has_conn) {       // RichardsFlux_Internal with compute_deriv (RichardsMod.F90:298-336)
        const double den_ave = upw * den + (1.0 - upw) * den_d;
        const double dphi    = X - P_d + den_ave * gfac;
        const bool   upwind  = (dphi >= 0.0);
        const double ukvr    = (upwind ? kr : kr_d) * RVIS;
        const double q       = (-Dq * ukvr * dphi) * area;
        const double dphi_dP_up =  1.0 + (upw * gfac) * dden;
        const double dphi_dP_dn = -1.0 + ((1.0 - upw) * gfac) * dden_d;
        const double dukvr_up = upwind ? dkr * RVIS : 0.0;
        const double dukvr_dn = upwind ? 0.0 : dkr_d * RVIS;
        const double dq_up = Dq * (dukvr_up * dphi + ukvr * dphi_dP_up) * area;
        const double dq_dn = Dq * (dukvr_dn * dphi + ukvr * dphi_dP_dn) * area;
        Jup = dq_up * den_ave - q * (upw * dden);
        Jdn = dq_dn * den_ave - q * ((1.0 - upw) * dden_d);
      }
      const double Jup_m = __shfl_up_sync(FULL, Jup, 1, GROUP);
      const double Jdn_m = __shfl_up_sync(FULL, Jdn, 1, GROUP);
      // row j of the tridiagonal Jacobian (GoveqnRichards...:2054-2069 insertion order)
      double ja = 0.0, jb = 0.0, jc = 0.0;
      if (valid) {
        if (j > 0) { ja = -Jup_m; jb += -Jdn_m; }
        jb += Jup; jc = Jdn;
#pragma unroll
        for (int k = 0; k < MAX_BC; ++k) if (bcMine[k]) {    // boundary: (dn,dn) -= Jdn  (:2136-2140)
          const double dphi0 = bcP[k] - X + den * bcGfac[k];
          const bool seep = (A.bc[k].itype == CT_SEEPAGE) && (dphi0 > 0.0) && (bcP[k] <= PRESSURE_REF);
          const double dphi = seep ? 0.0 : dphi0;
          const bool upwind = (dphi >= 0.0);
          const double ukvr = (upwind ? bcKr[k] : kr) * RVIS;
          const double q    = (-DqBC * ukvr * dphi) * area;
          const double dphi_dP_dn = seep ? 0.0 : (-1.0 + bcGfac[k] * dden);
          const double dukvr_dn = upwind ? 0.0 : dkr * RVIS;
          const double dq_dn = DqBC * (dukvr_dn * dphi + ukvr * dphi_dP_dn) * area;
          jb += -(dq_dn * den - q * dden);
        }
        jb += (por * dden * sat + por * den * dsat) * vol * dtInv;   // AccumDeriv (:1673-1675), dpor_dP = 0
      } else { jb = 1.0; }

      const double Yn = pcr_solve<GROUP>(ja, jb, jc, valid ? F : 0.0);   // J Y = F
      const double yn2 = group_sum<GROUP>(valid ? Yn * Yn : 0.0);
      const double xn2 = group_sum<GROUP>(valid ? X * X : 0.0);
      // initslope = F . (J Y), forced negative (SNESLineSearchApply_BT)
      const double Y_m = __shfl_up_sync(FULL, Yn, 1, GROUP), Y_p = __shfl_down_sync(FULL, Yn, 1, GROUP);
      double JY = jb * Yn;
      if (j > 0) JY = ja * Y_m + JY;
      if (has_conn) JY += jc * Y_p;
      double slope = group_sum<GROUP>(valid ? F * JY : 0.0);
      if (nw) {
        Y = Yn; ynorm = sqrt(yn2); xnorm = sqrt(xn2);
        if (slope > 0.0) slope = -slope;
        if (slope == 0.0) slope = -1.0;
        initslope = slope;
        lambda = 1.0; f2 = fnorm * fnorm; ls_count = 0;
        if (ynorm == 0.0) {
          // zero step: line search "fails"; stol*xnorm > ynorm => SNES_CONVERGED_SNORM_RELATIVE (ls.c)
          last_reason = (so.stol * xnorm > ynorm) ? SNES_CONVERGED_SNORM_RELATIVE : SNES_DIVERGED_LINE_SEARCH;
          phase = -1;   // SNES finished, handled below
        } else {
          if (ynorm > so.ls_maxstep) { Y *= so.ls_maxstep / ynorm; ynorm = so.ls_maxstep; }
          W = X - lambda * Y;
          phase = PH_LS_FULL;
          if (nfuncs >= so.max_funcs && so.max_funcs >= 0) { last_reason = SNES_DIVERGED_FUNCTION_COUNT; phase = -1; }
        }
      }
    }

    // ================= end-of-SNES bookkeeping (SOEBaseStepDT_SNES :481-536) =================
    if (phase == -1) {
      tot_nf += nfuncs;
      if (last_reason < 0) {
        cuts += 1; dt_iter = 0.5 * dt_iter; dtInv = 1.0 / dt_iter;
        X = Xprev;                                        // VecCopy(soln_prev, soln)
        if (cuts > 20) { converged = 0; phase = PH_DONE; }
        else { W = X; phase = PH_INIT; }
      } else {
        converged = 1; time_done += dt_iter; tot_its += its;
        Xprev = X;                                        // PostSolve: soln -> soln_prev
#pragma unroll
        for (int k = 0; k < MAX_BC; ++k) if (bcMine[k]) bcMassExc[k] += bcFlux[k] * dt_iter;
        if (time_done >= A.dt) phase = PH_DONE;
        else { W = X; phase = PH_INIT; }
      }
      its = 0; nfuncs = 0;
    }

    if (__all_sync(FULL, phase == PH_DONE)) break;

    // ================= residual evaluation at W (VSFMSOEResidual) =================
    SatState st;
    double den_w, dden_w, G_bcflux[MAX_BC];
    sat_values<SATFUNC>(sp, W, frac_liq, st);
    density_fixedT(A.dtab, W, den_w, dden_w);
    {
      const double acc = por * den_w * st.sat * vol * dtInv;          // Accum (:1626-1630)
      if (phase == PH_INIT) accum_prev = acc;                          // PreSolve: accumulation at soln_prev (== W here)
      const double P_d   = __shfl_down_sync(FULL, W, 1, GROUP);
      const double kr_d  = __shfl_down_sync(FULL, st.kr, 1, GROUP);
      const double den_d = __shfl_down_sync(FULL, den_w, 1, GROUP);
      double flux = 0.0;
      if (has_conn) {                                                  // RichardsFlux_Internal (:257-296)
        const double den_ave = upw * den_w + (1.0 - upw) * den_d;
        const double dphi    = W - P_d + den_ave * gfac;
        const double ukvr    = ((dphi >= 0.0) ? st.kr : kr_d) * RVIS;
        flux = ((-Dq * ukvr * dphi) * area) * den_ave;
      }
      const double flux_m = __shfl_up_sync(FULL, flux, 1, GROUP);
      G = acc - accum_prev;
      if (j > 0) G = G + flux_m;                                       // ff(dn) += flux  (:1806)
      G = G - flux;                                                    // ff(up) -= flux  (:1805)
#pragma unroll
      for (int k = 0; k < MAX_BC; ++k) {
        G_bcflux[k] = 0.0;
        if (bcMine[k]) {                                               // boundary connection, upweight = 0 (:262-264)
          double dphi = bcP[k] - W + den_w * bcGfac[k];
          if ((A.bc[k].itype == CT_SEEPAGE) && (dphi > 0.0) && (bcP[k] <= PRESSURE_REF)) dphi = 0.0;
          const double ukvr = ((dphi >= 0.0) ? bcKr[k] : st.kr) * RVIS;
          const double fl = ((-DqBC * ukvr * dphi) * area) * den_w;
          G = G + fl; G_bcflux[k] = fl * FMWH2O;
        }
      }
      G = G - src;
      if (!valid) G = 0.0;
    }
    const double g2 = group_sum<GROUP>(G * G);
    const double w2 = group_sum<GROUP>(valid ? W * W : 0.0);
    nfuncs += 1;

    // ================= after the evaluation: line-search / convergence logic =================
    bool take = false;          // adopt W as the new iterate (and its aux vars)
    const bool g_bad = !(g2 == g2) || (g2 > 1.7e308);                  // NaN or Inf
    const bool out_of_funcs = (nfuncs >= so.max_funcs && so.max_funcs >= 0);
    if (phase == PH_INIT) {
      // SNESSolve_NEWTONLS: F(X0) and the iteration-0 convergence test
      take = true;
    } else if (phase == PH_LS_FULL) {
      if (g_bad) {
        if (lambda <= so.ls_minlambda) { last_reason = SNES_DIVERGED_FNORM_NAN; phase = -1; }
        else if (out_of_funcs)         { last_reason = SNES_DIVERGED_FUNCTION_COUNT; phase = -1; }
        else { lambda = .5 * lambda; W = X - lambda * Y; }
      } else if (.5 * g2 <= .5 * f2 + lambda * so.ls_alpha * initslope) {
        take = true;
      } else if (so.stol * xnorm > ynorm) {
        // "full step didn't work and the step is tiny": line search fails, SNES then sees stol*xnorm > ynorm
        last_reason = SNES_CONVERGED_SNORM_RELATIVE; phase = -1;
      } else if (out_of_funcs) {
        last_reason = SNES_DIVERGED_FUNCTION_COUNT; phase = -1;
      } else {
        double lt = -initslope / (g2 - f2 - 2.0 * lambda * initslope);  // quadratic fit
        lambdaprev = lambda; gprev = g2;
        if (lt > .5 * lambda) lt = .5 * lambda;
        lambda = (lt <= .1 * lambda) ? .1 * lambda : lt;
        W = X - lambda * Y; phase = PH_LS_QUAD; ls_count = 0;
      }
    } else if (phase == PH_LS_QUAD || phase == PH_LS_CUBIC) {
      if (phase == PH_LS_CUBIC) ls_count += 1;                         // cubic trial points evaluated so far
      const int ls_fail = (so.stol * xnorm > ynorm) ? SNES_CONVERGED_SNORM_RELATIVE : SNES_DIVERGED_LINE_SEARCH;
      if (g_bad) {
        last_reason = ls_fail; phase = -1;
      } else if (.5 * g2 < .5 * f2 + lambda * so.ls_alpha * initslope) {
        take = true;
      } else if (ls_count >= so.ls_max_its) {
        take = true;                                                   // PETSc leaves the cubic loop after max_its fits and keeps the last point
      } else if (lambda <= so.ls_minlambda) {
        last_reason = ls_fail; phase = -1;
      } else if (out_of_funcs) {
        last_reason = SNES_DIVERGED_FUNCTION_COUNT; phase = -1;
      } else {
        const double t1 = .5 * (g2 - f2) - lambda * initslope;          // cubic fit
        const double t2 = .5 * (gprev - f2) - lambdaprev * initslope;
        const double a  = (t1 / (lambda * lambda) - t2 / (lambdaprev * lambdaprev)) / (lambda - lambdaprev);
        const double b  = (-lambdaprev * t1 / (lambda * lambda) + lambda * t2 / (lambdaprev * lambdaprev)) / (lambda - lambdaprev);
        double d = b * b - 3 * a * initslope;
        if (d < 0.0) d = 0.0;
        double lt = (a == 0.0) ? -initslope / (2.0 * b) : (-b + sqrt(d)) / (3.0 * a);
        lambdaprev = lambda; gprev = g2;
        if (lt > .5 * lambda) lt = .5 * lambda;
        lambda = (lt <= .1 * lambda) ? .1 * lambda : lt;
        W = X - lambda * Y; phase = PH_LS_CUBIC;
      }
    }

    if (take) {
      // "copy the solution over": X <- W, F <- G; the aux vars of this point feed the next Jacobian / PostSolve
      X = W; F = G;
      kr = st.kr; sat = st.sat; den = den_w; dden = dden_w;
      sat_derivs<SATFUNC>(sp, st, frac_liq, dsat, dkr);
#pragma unroll
      for (int k = 0; k < MAX_BC; ++k) bcFlux[k] = G_bcflux[k];
      fnorm = sqrt(g2);
      int reason = 0;
      if (phase == PH_INIT) {
        its = 0; ttol = fnorm * so.rtol; rnorm0 = fnorm;               // SNESConvergedDefault, it == 0
        if (g_bad)                reason = SNES_DIVERGED_FNORM_NAN;
        else if (fnorm < so.atol) reason = SNES_CONVERGED_FNORM_ABS;
      } else {
        xnorm = sqrt(w2);
        its += 1;
        if (fnorm < so.atol)      reason = SNES_CONVERGED_FNORM_ABS;   // SNESConvergedDefault, it > 0
        else if (out_of_funcs)    reason = SNES_DIVERGED_FUNCTION_COUNT;
        else if (fnorm <= ttol)   reason = SNES_CONVERGED_FNORM_RELATIVE;
        else if (ynorm < so.stol * xnorm) reason = SNES_CONVERGED_SNORM_RELATIVE;
        else if (so.divtol > 0 && fnorm > so.divtol * rnorm0) reason = SNES_DIVERGED_DTOL;
        else if (its >= so.max_it) reason = SNES_DIVERGED_MAX_IT;
      }
      if (reason) { last_reason = reason; phase = -1; } else phase = PH_NEWTON;
    }
  }

  // ---- VSFMSOEPostSolve -> SetDataInSOEAuxVar (GoveqnRichards...:1170-1195) ---------------------------------
  double mass = 0.0;
  if (valid) {
    A.x_out[cell] = X;
    if (converged) {
      A.liq_sat[cell]  = sat;
      A.pressure[cell] = X;
      mass = por * den * FMWH2O * sat * vol;
      A.mass[cell] = mass;
      A.smp[cell]  = (X - PRESSURE_REF) / (den * FMWH2O * GRAVITY_CONSTANT);
#pragma unroll
      for (int k = 0; k < MAX_BC; ++k) if (bcMine[k]) {
        A.bc[k].flux[col] = bcFlux[k];
        A.bc[k].mass_exc[col] += bcMassExc[k];
      }
    }
  }
  const double m_end = group_sum<GROUP>(mass);
  const double q_col = group_sum<GROUP>(src_kg);
  double err = 0.0, m_beg = 0.0;
  if (col_ok && j == 0) {
    A.stat_its[col] = tot_its; A.stat_reason[col] = last_reason; A.stat_cuts[col] = cuts; A.stat_nf[col] = tot_nf;
    m_beg = A.col_mass[col];
    if (converged) {
      err = fabs(m_beg - m_end + q_col * A.dt);
      A.col_mass[col] = m_end;
    }
    A.col_err[col] = err; A.col_src[col] = q_col;
  }

  // ---- block partials for the global mass-balance / convergence reductions (deterministic order) ----------
  __shared__ double red[8][128 / 32];
  const bool leader = col_ok && (j == 0);
  double v[8];
  v[0] = leader ? m_beg : 0.0;                                  // sum mass before
  v[1] = leader ? (converged ? m_end : m_beg) : 0.0;            // sum mass after
  v[2] = leader ? q_col * A.dt : 0.0;                           // sum sources * dt
  v[3] = 0.0;
#pragma unroll
  for (int k = 0; k < MAX_BC; ++k) if (valid && bcMine[k]) v[3] += bcMassExc[k];
  v[4] = leader ? err : 0.0;                                    // max |mass error|
  v[5] = leader ? (double)tot_its : 0.0;                        // max Newton its
  v[6] = leader ? (converged ? 0.0 : 1.0) : 0.0;                // any diverged
  v[7] = leader ? (double)cuts : 0.0;                           // max dt cuts
  // also carry the worst (minimum) reason through slot 3's sign-free neighbour: use a separate int path
  int worst = leader ? last_reason : 0x7fffffff;
#pragma unroll
  for (int s = 16; s > 0; s >>= 1) {
#pragma unroll
    for (int k = 0; k < 4; ++k) v[k] += __shfl_xor_sync(FULL, v[k], s);
#pragma unroll
    for (int k = 4; k < 8; ++k) v[k] = fmax(v[k], __shfl_xor_sync(FULL, v[k], s));
    worst = min(worst, __shfl_xor_sync(FULL, worst, s));
  }
  __shared__ int redw[128 / 32];
  const int warp = threadIdx.x >> 5;
  if (lane == 0) { for (int k = 0; k < 8; ++k) red[k][warp] = v[k]; redw[warp] = worst; }
  __syncthreads();
  if (threadIdx.x == 0) {
    const int nw = blockDim.x >> 5;
    double o[8]; int ow = 0x7fffffff;
    for (int k = 0; k < 8; ++k) o[k] = 0.0;
    for (int w = 0; w < nw; ++w) {
      for (int k = 0; k < 4; ++k) o[k] += red[k][w];
      for (int k = 4; k < 8; ++k) o[k] = fmax(o[k], red[k][w]);
      ow = min(ow, redw[w]);
    }
    double *bp = A.block_partials + (size_t)blockIdx.x * 9;
    for (int k = 0; k < 8; ++k) bp[k] = o[k];
    bp[8] = (double)ow;
  }
}

// Second stage of the deterministic reduction: `gridDim.x` blocks each fold a contiguous slice of the per-block
// partials (fixed order), the last block to finish folds the slice results (fixed order) into out[0..8]:
// out[0..3] sums, out[4..7] maxima, out[8] worst (minimum) SNES reason.  scratch: gridDim.x*9 doubles + 1 counter.
__global__ void reduce_partials_kernel(const double *__restrict__ partials, int nblocks, double *scratch,
                                       unsigned int *counter, double *out)
{
  __shared__ double sh[9][256];
  __shared__ bool last;
  const int per = (nblocks + gridDim.x - 1) / gridDim.x;
  const int b0 = blockIdx.x * per, b1 = min(nblocks, b0 + per);
  double v[9];
  for (int k = 0; k < 8; ++k) v[k] = 0.0;
  v[8] = 2147483647.0;
  for (int b = b0 + threadIdx.x; b < b1; b += blockDim.x) {
    const double *bp = partials + (size_t)b * 9;
    for (int k = 0; k < 4; ++k) v[k] += bp[k];
    for (int k = 4; k < 8; ++k) v[k] = fmax(v[k], bp[k]);
    v[8] = fmin(v[8], bp[8]);
  }
  for (int k = 0; k < 9; ++k) sh[k][threadIdx.x] = v[k];
  __syncthreads();
  for (int s = blockDim.x / 2; s > 0; s >>= 1) {
    if ((int)threadIdx.x < s) {
      for (int k = 0; k < 4; ++k) sh[k][threadIdx.x] += sh[k][threadIdx.x + s];
      for (int k = 4; k < 8; ++k) sh[k][threadIdx.x] = fmax(sh[k][threadIdx.x], sh[k][threadIdx.x + s]);
      sh[8][threadIdx.x] = fmin(sh[8][threadIdx.x], sh[8][threadIdx.x + s]);
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    for (int k = 0; k < 9; ++k) scratch[(size_t)blockIdx.x * 9 + k] = sh[k][0];
    __threadfence();
    last = (atomicAdd(counter, 1u) == gridDim.x - 1);
  }
  __syncthreads();
  if (last && threadIdx.x == 0) {
    __threadfence();
    double o[9];
    for (int k = 0; k < 8; ++k) o[k] = 0.0;
    o[8] = 2147483647.0;
    for (unsigned b = 0; b < gridDim.x; ++b) {
      const volatile double *bp = scratch + (size_t)b * 9;
      for (int k = 0; k < 4; ++k) o[k] += bp[k];
      for (int k = 4; k < 8; ++k) o[k] = fmax(o[k], bp[k]);
      o[8] = fmin(o[8], bp[8]);
    }
    for (int k = 0; k < 9; ++k) out[k] = o[k];
    *counter = 0u;
  }
}

}  // namespace mpp
