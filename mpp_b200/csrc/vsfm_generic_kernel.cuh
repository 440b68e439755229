// vsfm_generic_kernel.cuh -- VSFM time step for columns with more than 32 layers (e.g. the 100-cell
// Celia-1990 column, regression_tests/vsfm/vsfm_celia1990): one warp per column, cells strided over the
// lanes, per-cell state in shared memory, Thomas algorithm by lane 0.  Same algorithm and reference
// citations as vsfm_kernels.cuh; this path is for correctness on tall single columns, not for throughput.
#pragma once
#include "vsfm_kernels.cuh"

namespace mpp {

constexpr int VSFM_GENERIC_WARPS = 2;
constexpr int VSFM_GENERIC_NARR  = 22;     // per-cell shared arrays

__host__ __device__ inline size_t vsfm_generic_smem_bytes(int nlev)
{
  return (size_t)VSFM_GENERIC_WARPS * VSFM_GENERIC_NARR * (size_t)nlev * sizeof(double);
}

__device__ __forceinline__ double warp_sum(double v)
{
#pragma unroll
  for (int s = 16; s > 0; s >>= 1) v += __shfl_xor_sync(0xffffffffu, v, s);
  return v;
}

__global__ void __launch_bounds__(32 * VSFM_GENERIC_WARPS)
vsfm_step_generic_kernel(const VsfmArgs A, const int satfunc)
{
  extern __shared__ double smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int col = blockIdx.x * VSFM_GENERIC_WARPS + warp;
  const int nlev = A.nlev;
  const bool col_ok = (col < A.ncol) && (A.active == nullptr || A.active[col] != 0);
  double *base = smem + (size_t)warp * VSFM_GENERIC_NARR * nlev;
  double *X = base, *Xprev = X + nlev, *accp = Xprev + nlev, *F = accp + nlev, *Y = F + nlev, *W = Y + nlev, *G = W + nlev;
  double *kr = G + nlev, *den = kr + nlev, *dden = den + nlev, *sat = dden + nlev, *dsat = sat + nlev, *dkr = dsat + nlev;
  double *ja = dkr + nlev, *jb = ja + nlev, *jc = jb + nlev, *cp = jc + nlev, *dp = cp + nlev, *flx = dp + nlev;
  double *srcs = flx + nlev, *JY = srcs + nlev, *fl_liq = JY + nlev;
  const long long c0 = (long long)col * nlev;
  const SnesOpts so = A.so;
  const int jtop = A.top_is_first ? 0 : nlev - 1, jbot = A.top_is_first ? nlev - 1 : 0;
  const double area = col_ok ? A.area[col] : 1.0;

  double src_kg_l = 0.0;
  if (col_ok) {
    for (int j = lane; j < nlev; j += 32) {
      X[j] = A.x_in[c0 + j]; Xprev[j] = X[j]; W[j] = X[j]; fl_liq[j] = A.frac_liq[c0 + j];
      double s = 0.0;
      for (int k = 0; k < A.nss; ++k) {
        const CondDev &c = A.ss[k];
        if (c.itype != CT_MASS_RATE) continue;                       // the down-regulated sink depends on the pressure: handled per evaluation
        if (c.region == REGION_CELLS) { const double v = c.value[c0 + j]; s += v / FMWH2O; src_kg_l += v; }
        else if (j == (c.region == REGION_TOP ? jtop : jbot)) { const double v = c.value[col]; s += v / FMWH2O; src_kg_l += v; }
      }
      srcs[j] = s;
    }
  }
  __syncwarp();

  // boundary conditions (all lanes hold the same copies)
  double bcP[MAX_BC], bcKr[MAX_BC], bcGfac[MAX_BC], bcDq[MAX_BC], bcMassExc[MAX_BC], bcFlux[MAX_BC];
  int bcCell[MAX_BC];
  for (int k = 0; k < MAX_BC; ++k) {
    bcP[k] = PRESSURE_REF; bcKr[k] = 1.0; bcGfac[k] = 0.0; bcDq[k] = 0.0; bcMassExc[k] = 0.0; bcFlux[k] = 0.0; bcCell[k] = -1;
    if (col_ok && k < A.nbc) {
      const bool top = (A.bc[k].region == REGION_TOP);
      const int jc_ = top ? jtop : jbot;
      bcCell[k] = jc_;
      const double dzc = A.dz[c0 + jc_];
      const double uzbc = (A.uz == 0.0) ? 0.0 : (top ? -1.0 : 1.0);
      bcGfac[k] = FMWH2O * ((0.0 + 0.5 * dzc) * (uzbc * (-GRAVITY_CONSTANT)));
      bcDq[k] = A.perm[c0 + jc_] / (0.0 + 0.5 * dzc);
      bcP[k] = A.bc[k].value[col];
      SatParams sp; sp.sat_res = A.sat_res[c0 + jc_]; sp.alpha = A.alpha[c0 + jc_]; sp.m = A.lam[c0 + jc_];
      sp.n = A.vgn ? A.vgn[c0 + jc_] : 0.0;
      if (A.pu) { sp.pu = A.pu[c0 + jc_]; sp.ps = A.ps[c0 + jc_]; sp.b2 = A.b2[c0 + jc_]; sp.b3 = A.b3[c0 + jc_]; } else sp.pu = sp.ps = sp.b2 = sp.b3 = 0.0;
      SatState sb; sat_values_rt(satfunc, sp, bcP[k], 1.0, sb);
      bcKr[k] = sb.kr;
    }
  }

  double dt_iter = A.dt, dtInv = 1.0 / A.dt, time_done = 0.0;
  int cuts = 0, tot_its = 0, tot_nf = 0, last_reason = 0, converged = 0;
  int phase = col_ok ? PH_INIT : PH_DONE, its = 0, nfuncs = 0, ls_count = 0;
  double fnorm = 0.0, xnorm = 0.0, ynorm = 0.0, ttol = 0.0, rnorm0 = 0.0;
  double f2 = 0.0, initslope = -1.0, lambda = 1.0, lambdaprev = 1.0, gprev = 0.0;
  double G_bcflux[MAX_BC];

  while (phase != PH_DONE) {
    if (phase == PH_NEWTON) {
      // Jacobian rows from the aux vars of the accepted point
      for (int j = lane; j < nlev; j += 32) { ja[j] = 0.0; jb[j] = 0.0; jc[j] = 0.0; }
      __syncwarp();
      // internal connections j -> j+1: lane-owned, contributions combined per row afterwards via flx/cp scratch
      for (int j = lane; j < nlev - 1; j += 32) {
        const double dzu = A.dz[c0 + j], dzd = A.dz[c0 + j + 1], pu_ = A.perm[c0 + j], pd_ = A.perm[c0 + j + 1];
        const double dist_up = 0.5 * dzu, dist_dn = 0.5 * dzd;
        const double upw = dist_up / (dist_up + dist_dn);
        const double Dq = (pu_ * pd_) / (dist_up * pd_ + dist_dn * pu_);
        const double gfac = FMWH2O * ((dist_up + dist_dn) * (A.uz * (-GRAVITY_CONSTANT)));
        const double den_ave = upw * den[j] + (1.0 - upw) * den[j + 1];
        const double dphi = X[j] - X[j + 1] + den_ave * gfac;
        const bool upwind = (dphi >= 0.0);
        const double ukvr = (upwind ? kr[j] : kr[j + 1]) / VISCOSITY;
        const double q = (-Dq * ukvr * dphi) * area;
        const double dphi_dP_up = 1.0 + (upw * gfac) * dden[j];
        const double dphi_dP_dn = -1.0 + ((1.0 - upw) * gfac) * dden[j + 1];
        const double dukvr_up = upwind ? dkr[j] / VISCOSITY : 0.0;
        const double dukvr_dn = upwind ? 0.0 : dkr[j + 1] / VISCOSITY;
        const double dq_up = Dq * (dukvr_up * dphi + ukvr * dphi_dP_up) * area;
        const double dq_dn = Dq * (dukvr_dn * dphi + ukvr * dphi_dP_dn) * area;
        cp[j] = dq_up * den_ave - q * (upw * dden[j]);                 // Jup of connection j
        dp[j] = dq_dn * den_ave - q * ((1.0 - upw) * dden[j + 1]);     // Jdn of connection j
      }
      __syncwarp();
      for (int j = lane; j < nlev; j += 32) {
        double b = 0.0;
        if (j > 0) { ja[j] = -cp[j - 1]; b += -dp[j - 1]; }
        if (j < nlev - 1) { b += cp[j]; jc[j] = dp[j]; }
        for (int k = 0; k < MAX_BC; ++k) if (bcCell[k] == j) {
          const double dphi0 = bcP[k] - X[j] + den[j] * bcGfac[k];
          const bool seep = (A.bc[k].itype == CT_SEEPAGE) && (dphi0 > 0.0) && (bcP[k] <= PRESSURE_REF);
          const double dphi = seep ? 0.0 : dphi0;
          const bool upwind = (dphi >= 0.0);
          const double ukvr = (upwind ? bcKr[k] : kr[j]) / VISCOSITY;
          const double q = (-bcDq[k] * ukvr * dphi) * area;
          const double dphi_dP_dn = seep ? 0.0 : (-1.0 + bcGfac[k] * dden[j]);
          const double dukvr_dn = upwind ? 0.0 : dkr[j] / VISCOSITY;
          const double dq_dn = bcDq[k] * (dukvr_dn * dphi + ukvr * dphi_dP_dn) * area;
          b += -(dq_dn * den[j] - q * dden[j]);
        }
        const double por = A.por[c0 + j], vol = area * A.dz[c0 + j];
        b += (por * dden[j] * sat[j] + por * den[j] * dsat[j]) * vol * dtInv;
        if (A.dr_type && (A.dr_region == REGION_CELLS || j == (A.dr_region == REGION_TOP ? jtop : jbot))) {
          const long long i = (A.dr_region == REGION_CELLS) ? c0 + j : (long long)col;
          double rate, dj; downreg_sink(A.dr_type, A.dr_value[i], A.dr_pc[i], A.dr_n[i], X[j], rate, dj); b += dj;
        }
        jb[j] = b;
      }
      __syncwarp();
      if (lane == 0) {                                                   // Thomas: J Y = F
        cp[0] = jc[0] / jb[0]; dp[0] = F[0] / jb[0];
        for (int i = 1; i < nlev; ++i) {
          const double m = jb[i] - ja[i] * cp[i - 1];
          cp[i] = jc[i] / m; dp[i] = (F[i] - ja[i] * dp[i - 1]) / m;
        }
        Y[nlev - 1] = dp[nlev - 1];
        for (int i = nlev - 2; i >= 0; --i) Y[i] = dp[i] - cp[i] * Y[i + 1];
      }
      __syncwarp();
      double sy = 0.0, sx = 0.0, ss = 0.0;
      for (int j = lane; j < nlev; j += 32) {
        sy += Y[j] * Y[j]; sx += X[j] * X[j];
        double t = jb[j] * Y[j];
        if (j > 0) t = ja[j] * Y[j - 1] + t;
        if (j < nlev - 1) t += jc[j] * Y[j + 1];
        ss += F[j] * t;
      }
      ynorm = sqrt(warp_sum(sy)); xnorm = sqrt(warp_sum(sx)); initslope = warp_sum(ss);
      if (initslope > 0.0) initslope = -initslope;
      if (initslope == 0.0) initslope = -1.0;
      lambda = 1.0; f2 = fnorm * fnorm; ls_count = 0;
      if (ynorm == 0.0) {
        last_reason = (so.stol * xnorm > ynorm) ? SNES_CONVERGED_SNORM_RELATIVE : SNES_DIVERGED_LINE_SEARCH; phase = -1;
      } else {
        if (ynorm > so.ls_maxstep) { const double s = so.ls_maxstep / ynorm; for (int j = lane; j < nlev; j += 32) Y[j] *= s; ynorm = so.ls_maxstep; }
        for (int j = lane; j < nlev; j += 32) W[j] = X[j] - lambda * Y[j];
        phase = PH_LS_FULL;
        if (nfuncs >= so.max_funcs && so.max_funcs >= 0) { last_reason = SNES_DIVERGED_FUNCTION_COUNT; phase = -1; }
      }
      __syncwarp();
    }

    if (phase == -1) {
      tot_nf += nfuncs;
      if (last_reason < 0) {
        cuts += 1; dt_iter = 0.5 * dt_iter; dtInv = 1.0 / dt_iter;
        for (int j = lane; j < nlev; j += 32) X[j] = Xprev[j];
        if (cuts > 20) { converged = 0; phase = PH_DONE; }
        else { for (int j = lane; j < nlev; j += 32) W[j] = Xprev[j]; phase = PH_INIT; }
      } else {
        converged = 1; time_done += dt_iter; tot_its += its;
        for (int j = lane; j < nlev; j += 32) Xprev[j] = X[j];
        for (int k = 0; k < MAX_BC; ++k) if (bcCell[k] >= 0) bcMassExc[k] += bcFlux[k] * dt_iter;
        if (time_done >= A.dt) phase = PH_DONE;
        else { for (int j = lane; j < nlev; j += 32) W[j] = X[j]; phase = PH_INIT; }
      }
      its = 0; nfuncs = 0;
      __syncwarp();
      if (phase == PH_DONE) break;
    }

    // ---- residual at W; aux vars (with derivatives) of W overwrite the per-cell arrays ----
    for (int j = lane; j < nlev; j += 32) {
      SatParams sp; sp.sat_res = A.sat_res[c0 + j]; sp.alpha = A.alpha[c0 + j]; sp.m = A.lam[c0 + j];
      sp.n = A.vgn ? A.vgn[c0 + j] : 0.0;
      if (A.pu) { sp.pu = A.pu[c0 + j]; sp.ps = A.ps[c0 + j]; sp.b2 = A.b2[c0 + j]; sp.b3 = A.b3[c0 + j]; } else sp.pu = sp.ps = sp.b2 = sp.b3 = 0.0;
      SatState st;
      sat_values_rt(satfunc, sp, W[j], fl_liq[j], st);
      double ds, dk; sat_derivs_rt(satfunc, sp, st, fl_liq[j], ds, dk);
      double dn, ddn; density_fixedT_x<true>(A.dtab, W[j], dn, ddn);
      kr[j] = st.kr; sat[j] = st.sat; dsat[j] = ds; dkr[j] = dk; den[j] = dn; dden[j] = ddn;
    }
    __syncwarp();
    for (int j = lane; j < nlev - 1; j += 32) {
      const double dzu = A.dz[c0 + j], dzd = A.dz[c0 + j + 1], pu_ = A.perm[c0 + j], pd_ = A.perm[c0 + j + 1];
      const double dist_up = 0.5 * dzu, dist_dn = 0.5 * dzd;
      const double upw = dist_up / (dist_up + dist_dn);
      const double Dq = (pu_ * pd_) / (dist_up * pd_ + dist_dn * pu_);
      const double gfac = FMWH2O * ((dist_up + dist_dn) * (A.uz * (-GRAVITY_CONSTANT)));
      const double den_ave = upw * den[j] + (1.0 - upw) * den[j + 1];
      const double dphi = W[j] - W[j + 1] + den_ave * gfac;
      const double ukvr = ((dphi >= 0.0) ? kr[j] : kr[j + 1]) / VISCOSITY;
      flx[j] = ((-Dq * ukvr * dphi) * area) * den_ave;
    }
    __syncwarp();
    for (int k = 0; k < MAX_BC; ++k) G_bcflux[k] = 0.0;
    double sg = 0.0, sw = 0.0;
    for (int j = lane; j < nlev; j += 32) {
      const double por = A.por[c0 + j], vol = area * A.dz[c0 + j];
      const double acc = por * den[j] * sat[j] * vol * dtInv;
      if (phase == PH_INIT) accp[j] = acc;
      double g = acc - accp[j];
      if (j > 0) g = g + flx[j - 1];
      if (j < nlev - 1) g = g - flx[j];
      for (int k = 0; k < MAX_BC; ++k) if (bcCell[k] == j) {
        double dphi = bcP[k] - W[j] + den[j] * bcGfac[k];
        if ((A.bc[k].itype == CT_SEEPAGE) && (dphi > 0.0) && (bcP[k] <= PRESSURE_REF)) dphi = 0.0;
        const double ukvr = ((dphi >= 0.0) ? bcKr[k] : kr[j]) / VISCOSITY;
        const double fl = ((-bcDq[k] * ukvr * dphi) * area) * den[j];
        g = g + fl; G_bcflux[k] = fl * FMWH2O;
      }
      g = g - srcs[j];
      if (A.dr_type && (A.dr_region == REGION_CELLS || j == (A.dr_region == REGION_TOP ? jtop : jbot))) {
        const long long i = (A.dr_region == REGION_CELLS) ? c0 + j : (long long)col;
        double rate, dj; downreg_sink(A.dr_type, A.dr_value[i], A.dr_pc[i], A.dr_n[i], W[j], rate, dj); g = g - rate / FMWH2O;
      }
      G[j] = g; sg += g * g; sw += W[j] * W[j];
    }
    for (int k = 0; k < MAX_BC; ++k) G_bcflux[k] = warp_sum(G_bcflux[k]);   // only the owning lane contributed
    const double g2 = warp_sum(sg), w2 = warp_sum(sw);
    nfuncs += 1;
    __syncwarp();

    bool take = false;
    const bool g_bad = !(g2 == g2) || (g2 > 1.7e308);
    const bool out_of_funcs = (nfuncs >= so.max_funcs && so.max_funcs >= 0);
    if (phase == PH_INIT) {
      take = true;
    } else if (phase == PH_LS_FULL) {
      if (g_bad) {
        if (lambda <= so.ls_minlambda) { last_reason = SNES_DIVERGED_FNORM_NAN; phase = -1; }
        else if (out_of_funcs)         { last_reason = SNES_DIVERGED_FUNCTION_COUNT; phase = -1; }
        else { lambda = .5 * lambda; for (int j = lane; j < nlev; j += 32) W[j] = X[j] - lambda * Y[j]; }
      } else if (.5 * g2 <= .5 * f2 + lambda * so.ls_alpha * initslope) {
        take = true;
      } else if (so.stol * xnorm > ynorm) {
        last_reason = SNES_CONVERGED_SNORM_RELATIVE; phase = -1;
      } else if (out_of_funcs) {
        last_reason = SNES_DIVERGED_FUNCTION_COUNT; phase = -1;
      } else {
        double lt = -initslope / (g2 - f2 - 2.0 * lambda * initslope);
        lambdaprev = lambda; gprev = g2;
        if (lt > .5 * lambda) lt = .5 * lambda;
        lambda = (lt <= .1 * lambda) ? .1 * lambda : lt;
        for (int j = lane; j < nlev; j += 32) W[j] = X[j] - lambda * Y[j];
        phase = PH_LS_QUAD; ls_count = 0;
      }
    } else if (phase == PH_LS_QUAD || phase == PH_LS_CUBIC) {
      if (phase == PH_LS_CUBIC) ls_count += 1;
      const int ls_fail = (so.stol * xnorm > ynorm) ? SNES_CONVERGED_SNORM_RELATIVE : SNES_DIVERGED_LINE_SEARCH;
      if (g_bad) { last_reason = ls_fail; phase = -1; }
      else if (.5 * g2 < .5 * f2 + lambda * so.ls_alpha * initslope) take = true;
      else if (ls_count >= so.ls_max_its) take = true;
      else if (lambda <= so.ls_minlambda) { last_reason = ls_fail; phase = -1; }
      else if (out_of_funcs) { last_reason = SNES_DIVERGED_FUNCTION_COUNT; phase = -1; }
      else {
        const double t1 = .5 * (g2 - f2) - lambda * initslope;
        const double t2 = .5 * (gprev - f2) - lambdaprev * initslope;
        const double a  = (t1 / (lambda * lambda) - t2 / (lambdaprev * lambdaprev)) / (lambda - lambdaprev);
        const double b  = (-lambdaprev * t1 / (lambda * lambda) + lambda * t2 / (lambdaprev * lambdaprev)) / (lambda - lambdaprev);
        double d = b * b - 3 * a * initslope;
        if (d < 0.0) d = 0.0;
        double lt = (a == 0.0) ? -initslope / (2.0 * b) : (-b + sqrt(d)) / (3.0 * a);
        lambdaprev = lambda; gprev = g2;
        if (lt > .5 * lambda) lt = .5 * lambda;
        lambda = (lt <= .1 * lambda) ? .1 * lambda : lt;
        for (int j = lane; j < nlev; j += 32) W[j] = X[j] - lambda * Y[j];
        phase = PH_LS_CUBIC;
      }
    }
    if (take) {
      for (int j = lane; j < nlev; j += 32) { X[j] = W[j]; F[j] = G[j]; }
      for (int k = 0; k < MAX_BC; ++k) bcFlux[k] = G_bcflux[k];
      fnorm = sqrt(g2);
      int reason = 0;
      if (phase == PH_INIT) {
        its = 0; ttol = fnorm * so.rtol; rnorm0 = fnorm;
        if (g_bad)                reason = SNES_DIVERGED_FNORM_NAN;
        else if (fnorm < so.atol) reason = SNES_CONVERGED_FNORM_ABS;
      } else {
        xnorm = sqrt(w2); its += 1;
        if (fnorm < so.atol)      reason = SNES_CONVERGED_FNORM_ABS;
        else if (out_of_funcs)    reason = SNES_DIVERGED_FUNCTION_COUNT;
        else if (fnorm <= ttol)   reason = SNES_CONVERGED_FNORM_RELATIVE;
        else if (ynorm < so.stol * xnorm) reason = SNES_CONVERGED_SNORM_RELATIVE;
        else if (so.divtol > 0 && fnorm > so.divtol * rnorm0) reason = SNES_DIVERGED_DTOL;
        else if (its >= so.max_it) reason = SNES_DIVERGED_MAX_IT;
      }
      if (reason) { last_reason = reason; phase = -1; } else phase = PH_NEWTON;
    }
    __syncwarp();
  }

  // ---- PostSolve outputs ----
  double m_l = 0.0;
  if (col_ok) {
    for (int j = lane; j < nlev; j += 32) {
      A.x_out[c0 + j] = X[j];
      if (converged) {
        const double por = A.por[c0 + j], vol = area * A.dz[c0 + j];
        const double m = por * den[j] * FMWH2O * sat[j] * vol;
        A.liq_sat[c0 + j] = sat[j]; A.pressure[c0 + j] = X[j]; A.mass[c0 + j] = m;
        A.smp[c0 + j] = (X[j] - PRESSURE_REF) / (den[j] * FMWH2O * GRAVITY_CONSTANT);
        m_l += m;
      }
    }
  }
  if (A.dr_type && col_ok) {
    for (int j = lane; j < nlev; j += 32) if (A.dr_region == REGION_CELLS || j == (A.dr_region == REGION_TOP ? jtop : jbot)) {
      const long long i = (A.dr_region == REGION_CELLS) ? c0 + j : (long long)col;
      double rate, dj; downreg_sink(A.dr_type, A.dr_value[i], A.dr_pc[i], A.dr_n[i], X[j], rate, dj); src_kg_l += rate;
    }
  }
  const double m_end = warp_sum(m_l), q_col = warp_sum(src_kg_l);
  double err = 0.0, m_beg = 0.0, bexc = 0.0;
  if (col_ok && lane == 0) {
    A.stat_its[col] = tot_its; A.stat_reason[col] = last_reason; A.stat_cuts[col] = cuts; A.stat_nf[col] = tot_nf;
    m_beg = A.col_mass[col];
    if (converged) {
      err = fabs(m_beg - m_end + q_col * A.dt); A.col_mass[col] = m_end;
      for (int k = 0; k < MAX_BC; ++k) if (bcCell[k] >= 0) { A.bc[k].flux[col] = bcFlux[k]; A.bc[k].mass_exc[col] += bcMassExc[k]; bexc += bcMassExc[k]; }
    }
    A.col_err[col] = err; A.col_src[col] = q_col;
    if (A.t_done) A.t_done[col] = time_done;
  }
  __shared__ double red[9][VSFM_GENERIC_WARPS];
  if (lane == 0) {
    red[0][warp] = col_ok ? m_beg : 0.0;
    red[1][warp] = col_ok ? (converged ? m_end : m_beg) : 0.0;
    red[2][warp] = col_ok ? q_col * A.dt : 0.0;
    red[3][warp] = bexc;
    red[4][warp] = err;
    red[5][warp] = col_ok ? (double)tot_its : 0.0;
    red[6][warp] = col_ok ? (converged ? 0.0 : 1.0) : 0.0;
    red[7][warp] = col_ok ? (double)cuts : 0.0;
    red[8][warp] = col_ok ? (double)last_reason : 2147483647.0;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double *bp = A.block_partials + (size_t)blockIdx.x * 9;
    for (int k = 0; k < 9; ++k) {
      double o = (k == 8) ? 2147483647.0 : 0.0;
      for (int w = 0; w < VSFM_GENERIC_WARPS; ++w)
        o = (k < 4) ? o + red[k][w] : (k < 8 ? fmax(o, red[k][w]) : fmin(o, red[k][w]));
      bp[k] = o;
    }
  }
}

}  // namespace mpp
