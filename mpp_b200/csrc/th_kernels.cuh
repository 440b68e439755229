// th_kernels.cuh -- fused coupled thermal-hydrology (TH) time step: Richards (mass) + enthalpy (energy) equations,
// two unknowns (P, T) per cell, 2x2 block-tridiagonal Newton system per column.
//
// One launch = one sysofeqns%StepDT of the TH system of equations for every column:
//   SOETHPreSolve / Residual / Jacobian / PostSolve   src/mpp/soe/SystemOfEquationsTHType.F90:119-302, 736-1004
//   energy aux vars                                   src/mpp/auxvar/ThermalEnthalpySoilAuxType.F90:219-278
//   energy flux + derivatives                         src/mpp/ge/ThermalEnthalpyMod.F90:27-332
//   accumulation / divergence / Jacobian blocks       src/mpp/ge/GoveqnThermalEnthalpySoilType.F90:1174-2377
//   d(mass residual)/dT                               src/mpp/ge/GoveqnRichardsODEPressureType.F90:2333-2613,
//                                                     src/mpp/ge/RichardsMod.F90:343-648
//   PETSc SNES newtonls + bt + SNESConvergedDefault (as vsfm_kernels.cuh); the Newton system is solved exactly by
//   block Thomas on cell-interleaved unknowns (the reference's GMRES+ILU(0) on segregated unknowns is inexact).
//
// Round-1 mapping: one warp per column, cells strided over the lanes, per-cell state in shared memory, block Thomas
// by lane 0 -- correct for any number of layers (the reference's TH goldens have 100 and 20 cells).  A lane-per-cell
// variant with a register block-PCR (as the VSFM kernel) is the planned next step for the 15-layer batches.
#pragma once
#include "vsfm_kernels.cuh"

namespace mpp {

constexpr int TH_NARR = 56;            // doubles of shared memory per cell

struct THCondDev { const double *value; const double *bc_pressure; int ieqn, itype, region; };

struct THArgs {
  int ncol, nlev;
  double uz; int top_is_first;
  int satfunc, density_type, iee_type;
  const double *por, *perm, *sat_res, *alpha, *lam, *vgn, *pu, *ps, *b2, *b3, *dz, *area;
  const double *tkdry, *csol;
  const double *perm_e;                        // energy equation's own permeability (mppgpu_th_set_energy_permeability); nullptr: the aux-var default
  const double *x_in; double *x_out;           // cell-interleaved (P, T)
  int nbc, nss;
  THCondDev bc[4], ss[4];
  double *liq_sat, *mass;
  int *stat_its, *stat_reason, *stat_cuts, *stat_nf;
  double *block_partials;
  const int *order;                     // launch order of the columns (most expensive of the previous step first) or nullptr: batch order
  double dt;
  SnesOpts so;
  // the only boundary condition is a Dirichlet temperature at the top of a column of <= 15 layers (the ELM-like TH batch): the fast
  // kernel treats the boundary aux var as one more cell on the padding lane (th_kernels2.cuh) instead of a one-lane loop
  int bc_on_pad_lane;
  // kernel unit-test probe (mppgpu_eval): accumulation at x_in, residual + Jacobian blocks at eval_x, no time step
  const double *eval_x; double *eval_f, *eval_ja, *eval_jb, *eval_jc;
};

constexpr int PH_EVAL = 6, PH_EVAL_J = 7;

struct THCell {      // aux vars of one cell at one state (both governing equations)
  double sat, kr, dsat, dkr;                      // identical for the two equations (same P, frac_liq_sat = 1)
  double den_m, ddenP_m, ddenT_m;                 // mass equation: density at (P, T)
  double den_e, ddenP_e, ddenT_e;                 // energy equation: density at (max(P, Pref), T)
  double ul, hl, dulT, dhlT, dulP, dhlP, tc, dtcP;
};

// SF / DT / IEE >= 0: saturation function, density and enthalpy models fixed at compile time (no dispatch branches, so the
// independent chains -- curves, two densities, enthalpy -- share one basic block and overlap); -1: read from A at run time.
template <int SF = -1, int DT = -1, int IEE = -1>
__device__ __forceinline__ void th_cell_compute(const THArgs &A, const SatParams &sp, double tkdry, double P, double T, THCell &c)
{
  const int dtype = (DT >= 0) ? DT : A.density_type, itype = (IEE >= 0) ? IEE : A.iee_type;
  SatState st;
  if (SF == SATFUNC_SBC || SF == SATFUNC_BC) {
    // the lean branch-free Brooks-Corey / smoothed Brooks-Corey curves of the fast VSFM kernel (physics.cuh:bc_sbc_values): one log and
    // two exp whatever the regime; frac_liq_sat = 1 in the TH model
    const double pc = P - PRESSURE_REF;
    const bool rA = (SF == SATFUNC_BC) ? (-sp.alpha * pc > 1.0) : (pc <= sp.pu), rB = (SF == SATFUNC_SBC) && !rA && (pc < sp.ps);
    double Se;
    bc_sbc_values<double>(sp.sat_res, sp.alpha, sp.m, sp.ps, sp.b2, sp.b3, pc, rA, rA, rB, rB, c.sat, c.kr, Se);
    bc_sbc_derivs<double>(sp.sat_res, sp.m, sp.ps, sp.b2, sp.b3, pc, Se, c.kr, rA, rA, rB, rB, c.dsat, c.dkr);
  } else {
    if (SF >= 0) { sat_values<(SF >= 0 ? SF : 0)>(sp, P, 1.0, st); sat_derivs<(SF >= 0 ? SF : 0)>(sp, st, 1.0, c.dsat, c.dkr); }
    else { sat_values_rt(A.satfunc, sp, P, 1.0, st); sat_derivs_rt(A.satfunc, sp, st, 1.0, c.dsat, c.dkr); }
    c.sat = st.sat; c.kr = st.kr;
  }
  const double Pe = (P < PRESSURE_REF) ? PRESSURE_REF : P;                 // ThermalEnthalpySoilAuxType.F90:251-252
  if (DT == DENSITY_TGDPB01) {                                              // one temperature part for both pressures
    const TanakaT tt = tanaka_T(T);
    tanaka_P(tt, P, c.den_m, c.ddenP_m, c.ddenT_m);
    tanaka_P(tt, Pe, c.den_e, c.ddenP_e, c.ddenT_e);
  } else {
    density(dtype, P, T, c.den_m, c.ddenP_m, c.ddenT_m);
    density(dtype, Pe, T, c.den_e, c.ddenP_e, c.ddenT_e);
  }
  internal_energy_enthalpy(itype, Pe, T, c.den_e * FMWH2O, c.ddenT_e * FMWH2O, c.ddenP_e * FMWH2O,
                           c.ul, c.hl, c.dulT, c.dhlT, c.dulP, c.dhlP);
  const double therm_alpha = 0.45, wet = 1.3;                              // MultiPhysicsProbTH.F90:331-332
  // Kel = (sat + 1e-6)^alpha, dKel/dP = alpha (sat + 1e-6)^(alpha - 1) dsat/dP   (ThermalEnthalpySoilAuxType.F90:269-275)
  const double se = c.sat + 1.e-6;
  const double Kel = mpp_exp(therm_alpha * mpp_log(se));
  const double dKel = therm_alpha * Kel * rcp(se) * c.dsat;
  c.tc = wet * Kel + tkdry * (1.0 - Kel);
  c.dtcP = (wet - tkdry) * dKel;
}

// RichardsFlux_Internal on (up, dn) states; returns flux and -d(flux)/dP (reference sign convention) or d(flux)/dT
struct FluxIn { double P, kr, dkr, den, ddenP, ddenT; };
__device__ __forceinline__ void th_rich_flux(const FluxIn &u, const FluxIn &d, double upw, double Dq, double gfac, double area,
                                             double &flux, double &mJup, double &mJdn, double &dT_up, double &dT_dn)
{
  const double den_ave = upw * u.den + (1.0 - upw) * d.den;
  const double dphi = u.P - d.P + den_ave * gfac;
  const bool upwind = (dphi >= 0.0);
  constexpr double RVIS = 1.0 / VISCOSITY;
  const double ukvr = (upwind ? u.kr : d.kr) * RVIS;
  const double q = (-Dq * ukvr * dphi) * area;
  flux = q * den_ave;
  const double dphi_dP_up = 1.0 + (upw * gfac) * u.ddenP, dphi_dP_dn = -1.0 + ((1.0 - upw) * gfac) * d.ddenP;
  const double dukvr_up = upwind ? u.dkr * RVIS : 0.0, dukvr_dn = upwind ? 0.0 : d.dkr * RVIS;
  const double dq_up = Dq * (dukvr_up * dphi + ukvr * dphi_dP_up) * area, dq_dn = Dq * (dukvr_dn * dphi + ukvr * dphi_dP_dn) * area;
  mJup = dq_up * den_ave - q * (upw * u.ddenP);
  mJdn = dq_dn * den_ave - q * ((1.0 - upw) * d.ddenP);
  // RichardsFluxDerivativeWrtTemperature (dvis_dT = 0): true derivatives (sign flipped at RichardsMod.F90:640-641)
  const double dqT_up = Dq * (ukvr * ((upw * gfac) * u.ddenT)) * area, dqT_dn = Dq * (ukvr * (((1.0 - upw) * gfac) * d.ddenT)) * area;
  dT_up = -(dqT_up * den_ave - q * (upw * u.ddenT));
  dT_dn = -(dqT_dn * den_ave - q * ((1.0 - upw) * d.ddenT));
}

#ifndef MPP_STEP_KERNEL_TU   // the generic (one warp per column) kernel is compiled once, in mppgpu.cu
// all per-cell arrays live in shared memory; this view indexes them
struct THView {
  double *P, *T, *Pp, *Tp, *accm, *acce, *Fm, *Fe, *Ym, *Ye, *Wm, *We, *Gm, *Ge;
  double *sat, *kr, *dsat, *dkr, *denm, *dPm, *dTm, *dene, *dPe, *dTe, *ul, *hl, *dulT, *dhlT, *dulP, *dhlP, *tc, *dtcP;
  double *ja, *jb, *jc, *cp, *dp, *fm, *fe, *srcm, *srce;     // ja/jb/jc/cp: 4 per cell; dp: 2 per cell
};

__global__ void __launch_bounds__(32)
th_step_generic_kernel(const THArgs A)
{
  extern __shared__ double smem[];
  const int lane = threadIdx.x, col = blockIdx.x, nlev = A.nlev;
  const bool col_ok = col < A.ncol;
  THView v;
  {
    double *b = smem; const int n = nlev;
    double **p1[] = {&v.P, &v.T, &v.Pp, &v.Tp, &v.accm, &v.acce, &v.Fm, &v.Fe, &v.Ym, &v.Ye, &v.Wm, &v.We, &v.Gm, &v.Ge,
                     &v.sat, &v.kr, &v.dsat, &v.dkr, &v.denm, &v.dPm, &v.dTm, &v.dene, &v.dPe, &v.dTe, &v.ul, &v.hl, &v.dulT, &v.dhlT,
                     &v.dulP, &v.dhlP, &v.tc, &v.dtcP, &v.fm, &v.fe, &v.srcm, &v.srce};
    for (auto pp : p1) { *pp = b; b += n; }
    v.ja = b; b += 4 * n; v.jb = b; b += 4 * n; v.jc = b; b += 4 * n; v.cp = b; b += 4 * n; v.dp = b; b += 2 * n;
  }
  const long long c0 = (long long)col * nlev;
  const SnesOpts so = A.so;
  const int jtop = A.top_is_first ? 0 : nlev - 1, jbot = A.top_is_first ? nlev - 1 : 0;
  const double area = col_ok ? A.area[col] : 1.0;
  // energy-equation aux vars keep the default permeability (ThermalEnthalpySoilAuxType.F90:93) unless the driver set its own
  // (goveq_enthalpy%SetSoilPermeability, th_mms_problem.F90:739)
#define PERM_E_AT(jj) (A.perm_e ? A.perm_e[c0 + (jj)] : 8.3913e-12)

  if (col_ok) {
    for (int j = lane; j < nlev; j += 32) {
      v.P[j] = A.x_in[2 * (c0 + j)]; v.T[j] = A.x_in[2 * (c0 + j) + 1];
      v.Pp[j] = v.P[j]; v.Tp[j] = v.T[j]; v.Wm[j] = v.P[j]; v.We[j] = v.T[j];
      double sm = 0.0, se = 0.0;
      for (int k = 0; k < A.nss; ++k) {
        const THCondDev &c = A.ss[k];
        double val = 0.0; bool mine = false;
        if (c.region == REGION_CELLS) { val = c.value[c0 + j]; mine = true; }
        else if (j == (c.region == REGION_TOP ? jtop : jbot)) { val = c.value[col]; mine = true; }
        if (mine) { if (c.ieqn == 1) sm += val / FMWH2O; else se += val; }
      }
      v.srcm[j] = sm; v.srce[j] = se;
    }
  }
  __syncwarp();

  // boundary conditions: Dirichlet pressure on the mass equation / Dirichlet temperature on the energy equation
  struct BC { int cell, ieqn; double val, P, gfac, Dq, uzsign; FluxIn fin; double hl, dhlT, dhlP, tc, T; };
  BC bcs[4];
  for (int k = 0; k < 4; ++k) {
    bcs[k].cell = -1;
    if (!col_ok || k >= A.nbc) continue;
    const bool top = (A.bc[k].region == REGION_TOP);
    const int jc_ = top ? jtop : jbot;
    bcs[k].cell = jc_; bcs[k].ieqn = A.bc[k].ieqn; bcs[k].val = A.bc[k].value[col];
    const double dzc = A.dz[c0 + jc_];
    const double uzbc = (A.uz == 0.0) ? 0.0 : (top ? -1.0 : 1.0);
    bcs[k].gfac = FMWH2O * ((0.0 + 0.5 * dzc) * (uzbc * (-GRAVITY_CONSTANT)));
    SatParams sp; sp.sat_res = A.sat_res[c0 + jc_]; sp.alpha = A.alpha[c0 + jc_]; sp.m = A.lam[c0 + jc_]; sp.n = A.vgn ? A.vgn[c0 + jc_] : 0.0;
    if (A.pu) { sp.pu = A.pu[c0 + jc_]; sp.ps = A.ps[c0 + jc_]; sp.b2 = A.b2[c0 + jc_]; sp.b3 = A.b3[c0 + jc_]; } else sp.pu = sp.ps = sp.b2 = sp.b3 = 0.0;
    THCell bc;
    if (bcs[k].ieqn == 1) {
      // mass-equation boundary aux var: pressure = condition value, temperature stays at its default 298.15 K
      bcs[k].P = bcs[k].val; bcs[k].T = 273.15 + 25.0;
      bcs[k].Dq = A.perm[c0 + jc_] / (0.0 + 0.5 * dzc);
      th_cell_compute(A, sp, A.tkdry[c0 + jc_], bcs[k].P, bcs[k].T, bc);
      bcs[k].fin = {bcs[k].P, bc.kr, bc.dkr, bc.den_m, bc.ddenP_m, bc.ddenT_m};
    } else {
      // energy-equation boundary aux var: temperature = condition value; pressure is whatever the driver poked
      // (mass_and_heat_model_problem.F90:616-621), default 0 (RichardsODEPressureAuxType.F90:90)
      bcs[k].T = bcs[k].val; bcs[k].P = A.bc[k].bc_pressure ? A.bc[k].bc_pressure[col] : 0.0;
      bcs[k].Dq = PERM_E_AT(jc_) / (0.0 + 0.5 * dzc);
      th_cell_compute(A, sp, A.tkdry[c0 + jc_], bcs[k].P, bcs[k].T, bc);
      bcs[k].fin = {bcs[k].P, bc.kr, bc.dkr, bc.den_e, bc.ddenP_e, bc.ddenT_e};
      bcs[k].hl = bc.hl; bcs[k].dhlT = bc.dhlT; bcs[k].dhlP = bc.dhlP; bcs[k].tc = bc.tc;
    }
  }

  double dt_iter = A.dt, dtInv = 1.0 / A.dt, time_done = 0.0;
  int cuts = 0, tot_its = 0, tot_nf = 0, last_reason = 0, converged = 0;
  int phase = col_ok ? PH_INIT : PH_DONE, its = 0, nfuncs = 0, ls_count = 0;
  double fnorm = 0.0, xnorm = 0.0, ynorm = 0.0, ttol = 0.0, rnorm0 = 0.0;
  double f2 = 0.0, initslope = -1.0, lambda = 1.0, lambdaprev = 1.0, gprev = 0.0;

  while (phase != PH_DONE) {
    if (phase == PH_NEWTON || phase == PH_EVAL_J) {
      for (int j = lane; j < nlev; j += 32) for (int k = 0; k < 4; ++k) { v.ja[4 * j + k] = 0.0; v.jb[4 * j + k] = 0.0; v.jc[4 * j + k] = 0.0; }
      __syncwarp();
      // connection contributions, staged per connection in cp (8 values: as "up" row) ... computed twice (once per side) to avoid races
      for (int j = lane; j < nlev; j += 32) {
        const double por = A.por[c0 + j], vol = area * A.dz[c0 + j];
        double b00 = 0, b01 = 0, b10 = 0, b11 = 0;
        for (int side = 0; side < 2; ++side) {          // side 0: connection j-1 -> j (this cell is dn); side 1: j -> j+1 (this cell is up)
          const int ju = (side == 0) ? j - 1 : j, jd = ju + 1;
          if (ju < 0 || jd >= nlev) continue;
          const double dzu = A.dz[c0 + ju], dzd = A.dz[c0 + jd];
          const double dist_up = 0.5 * dzu, dist_dn = 0.5 * dzd, upw = dist_up / (dist_up + dist_dn);
          const double gfac = FMWH2O * ((dist_up + dist_dn) * (A.uz * (-GRAVITY_CONSTANT)));
          const double pmu = A.perm[c0 + ju], pmd = A.perm[c0 + jd];
          const double Dqm = (pmu * pmd) / (dist_up * pmd + dist_dn * pmu);
          const double peu = PERM_E_AT(ju), ped = PERM_E_AT(jd);
          const double Dqe = (peu * ped) / (dist_up * ped + dist_dn * peu);
          FluxIn um = {v.P[ju], v.kr[ju], v.dkr[ju], v.denm[ju], v.dPm[ju], v.dTm[ju]}, dm = {v.P[jd], v.kr[jd], v.dkr[jd], v.denm[jd], v.dPm[jd], v.dTm[jd]};
          FluxIn ue = {v.P[ju], v.kr[ju], v.dkr[ju], v.dene[ju], v.dPe[ju], v.dTe[ju]}, de = {v.P[jd], v.kr[jd], v.dkr[jd], v.dene[jd], v.dPe[jd], v.dTe[jd]};
          double fl, mJup, mJdn, dTu, dTd;
          th_rich_flux(um, dm, upw, Dqm, gfac, area, fl, mJup, mJdn, dTu, dTd);
          double mfl, eJup, eJdn, edTu, edTd;
          th_rich_flux(ue, de, upw, Dqe, gfac, area, mfl, eJup, eJdn, edTu, edTd);
          // energy flux derivatives (ThermalEnthalpyFlux / ...DerivativeWrtPressure)
          const double ku = v.tc[ju], kd = v.tc[jd];
          const double kod = (ku * kd) / (dist_up * kd + dist_dn * ku);
          const double h = (mfl <= 0.0) ? v.hl[ju] : v.hl[jd];
          const double dhT_u = (mfl < 0.0) ? v.dhlT[ju] : 0.0, dhT_d = (mfl < 0.0) ? 0.0 : v.dhlT[jd];
          const double dhP_u = (mfl < 0.0) ? v.dhlP[ju] : 0.0, dhP_d = (mfl < 0.0) ? 0.0 : v.dhlP[jd];
          const double JTT_u = edTu * h + mfl * dhT_u + (-kod * area), JTT_d = edTd * h + mfl * dhT_d + (+kod * area);
          const double dDk_u = (kod * kod) / (ku * ku) * dist_up * v.dtcP[ju], dDk_d = (kod * kod) / (kd * kd) * dist_dn * v.dtcP[jd];
          const double dTud = v.T[ju] - v.T[jd];
          const double JTP_u = (-eJup) * h + mfl * dhP_u + (-dDk_u * dTud * area), JTP_d = (-eJdn) * h + mfl * dhP_d + (-dDk_d * dTud * area);
          if (side == 1) {      // this cell is "up": rows get (+Jup, +Jdn) for mass/P and (-J) for the true-derivative blocks
            b00 += mJup; v.jc[4 * j + 0] = mJdn;
            b01 += -dTu; v.jc[4 * j + 1] = -dTd;
            b11 += -JTT_u; v.jc[4 * j + 3] = -JTT_d;
            b10 += -JTP_u; v.jc[4 * j + 2] = -JTP_d;
          } else {              // this cell is "dn"
            v.ja[4 * j + 0] = -mJup; b00 += -mJdn;
            v.ja[4 * j + 1] = dTu;   b01 += dTd;
            v.ja[4 * j + 3] = JTT_u; b11 += JTT_d;
            v.ja[4 * j + 2] = JTP_u; b10 += JTP_d;
          }
        }
        for (int k = 0; k < 4; ++k) if (bcs[k].cell == j) {
          const double dzc = A.dz[c0 + j];
          if (bcs[k].ieqn == 1) {
            FluxIn dn = {v.P[j], v.kr[j], v.dkr[j], v.denm[j], v.dPm[j], v.dTm[j]};
            double fl, mJup, mJdn, a1, a2;
            th_rich_flux(bcs[k].fin, dn, 0.0, bcs[k].Dq, bcs[k].gfac, area, fl, mJup, mJdn, a1, a2);
            b00 += -mJdn;
          } else {
            FluxIn dn = {v.P[j], v.kr[j], v.dkr[j], v.dene[j], v.dPe[j], v.dTe[j]};
            double mfl, eJup, eJdn, edTu, edTd;
            th_rich_flux(bcs[k].fin, dn, 0.0, bcs[k].Dq, bcs[k].gfac, area, mfl, eJup, eJdn, edTu, edTd);
            const double kod = v.tc[j] / (0.0 + 0.5 * dzc);
            const double h = (mfl <= 0.0) ? bcs[k].hl : v.hl[j];
            const double dhT_d = (mfl < 0.0) ? 0.0 : v.dhlT[j], dhP_d = (mfl < 0.0) ? 0.0 : v.dhlP[j];
            b11 += edTd * h + mfl * dhT_d + (+kod * area);
            const double dDk_d = 1.0 / (0.0 + 0.5 * dzc) * v.dtcP[j];
            b10 += (-eJdn) * h + mfl * dhP_d + (-dDk_d * (bcs[k].T - v.T[j]) * area);
          }
        }
        // accumulation derivatives (GoveqnRichards...:1673, 2547; GoveqnThermalEnthalpySoilType.F90:1276-1281, 2146-2153)
        b00 += (por * v.dPm[j] * v.sat[j] + por * v.denm[j] * v.dsat[j]) * vol * dtInv;
        b01 += (por * v.dTm[j] * v.sat[j]) * vol * dtInv;
        const double csol = A.csol[c0 + j];
        b11 += ((por * v.dTe[j] * v.sat[j] * v.ul[j] + por * v.dene[j] * v.sat[j] * v.dulT[j]) + (1.0 - por) * 2700.0 * csol) * vol * dtInv;
        b10 += (por * v.dPe[j] * v.sat[j] * v.ul[j] + por * v.dene[j] * v.dsat[j] * v.ul[j] + por * v.dene[j] * v.sat[j] * v.dulP[j]) * vol * dtInv;
        v.jb[4 * j + 0] = b00; v.jb[4 * j + 1] = b01; v.jb[4 * j + 2] = b10; v.jb[4 * j + 3] = b11;
      }
      __syncwarp();
      if (phase == PH_EVAL_J) {
        for (int j = lane; j < nlev; j += 32) for (int k = 0; k < 4; ++k) {
          A.eval_ja[4 * (c0 + j) + k] = v.ja[4 * j + k]; A.eval_jb[4 * (c0 + j) + k] = v.jb[4 * j + k]; A.eval_jc[4 * (c0 + j) + k] = v.jc[4 * j + k];
        }
        break;
      }
      if (lane == 0) {       // block Thomas: J Y = F
        auto inv2 = [](const double *m, double *r) { const double det = m[0] * m[3] - m[1] * m[2]; r[0] = m[3] / det; r[1] = -m[1] / det; r[2] = -m[2] / det; r[3] = m[0] / det; };
        auto mul22 = [](const double *x, const double *y, double *r) { r[0] = x[0] * y[0] + x[1] * y[2]; r[1] = x[0] * y[1] + x[1] * y[3]; r[2] = x[2] * y[0] + x[3] * y[2]; r[3] = x[2] * y[1] + x[3] * y[3]; };
        double binv[4], m[4], t[4], d0, d1;
        inv2(v.jb, binv); mul22(binv, v.jc, v.cp);
        v.dp[0] = binv[0] * v.Fm[0] + binv[1] * v.Fe[0]; v.dp[1] = binv[2] * v.Fm[0] + binv[3] * v.Fe[0];
        for (int i = 1; i < nlev; ++i) {
          mul22(v.ja + 4 * i, v.cp + 4 * (i - 1), t);
          for (int k = 0; k < 4; ++k) m[k] = v.jb[4 * i + k] - t[k];
          inv2(m, binv); mul22(binv, v.jc + 4 * i, v.cp + 4 * i);
          d0 = v.Fm[i] - (v.ja[4 * i] * v.dp[2 * i - 2] + v.ja[4 * i + 1] * v.dp[2 * i - 1]);
          d1 = v.Fe[i] - (v.ja[4 * i + 2] * v.dp[2 * i - 2] + v.ja[4 * i + 3] * v.dp[2 * i - 1]);
          v.dp[2 * i] = binv[0] * d0 + binv[1] * d1; v.dp[2 * i + 1] = binv[2] * d0 + binv[3] * d1;
        }
        v.Ym[nlev - 1] = v.dp[2 * nlev - 2]; v.Ye[nlev - 1] = v.dp[2 * nlev - 1];
        for (int i = nlev - 2; i >= 0; --i) {
          v.Ym[i] = v.dp[2 * i] - (v.cp[4 * i] * v.Ym[i + 1] + v.cp[4 * i + 1] * v.Ye[i + 1]);
          v.Ye[i] = v.dp[2 * i + 1] - (v.cp[4 * i + 2] * v.Ym[i + 1] + v.cp[4 * i + 3] * v.Ye[i + 1]);
        }
      }
      __syncwarp();
      double sy = 0.0, sx = 0.0, ss = 0.0;
      for (int j = lane; j < nlev; j += 32) {
        sy += v.Ym[j] * v.Ym[j] + v.Ye[j] * v.Ye[j]; sx += v.P[j] * v.P[j] + v.T[j] * v.T[j];
        double t0 = v.jb[4 * j] * v.Ym[j] + v.jb[4 * j + 1] * v.Ye[j], t1 = v.jb[4 * j + 2] * v.Ym[j] + v.jb[4 * j + 3] * v.Ye[j];
        if (j > 0) { t0 += v.ja[4 * j] * v.Ym[j - 1] + v.ja[4 * j + 1] * v.Ye[j - 1]; t1 += v.ja[4 * j + 2] * v.Ym[j - 1] + v.ja[4 * j + 3] * v.Ye[j - 1]; }
        if (j < nlev - 1) { t0 += v.jc[4 * j] * v.Ym[j + 1] + v.jc[4 * j + 1] * v.Ye[j + 1]; t1 += v.jc[4 * j + 2] * v.Ym[j + 1] + v.jc[4 * j + 3] * v.Ye[j + 1]; }
        ss += v.Fm[j] * t0 + v.Fe[j] * t1;
      }
      ynorm = sqrt(warp_sum(sy)); xnorm = sqrt(warp_sum(sx)); initslope = warp_sum(ss);
      if (initslope > 0.0) initslope = -initslope;
      if (initslope == 0.0) initslope = -1.0;
      lambda = 1.0; f2 = fnorm * fnorm; ls_count = 0;
      if (ynorm == 0.0) { last_reason = (so.stol * xnorm > ynorm) ? SNES_CONVERGED_SNORM_RELATIVE : SNES_DIVERGED_LINE_SEARCH; phase = -1; }
      else {
        if (ynorm > so.ls_maxstep) { const double s = so.ls_maxstep / ynorm; for (int j = lane; j < nlev; j += 32) { v.Ym[j] *= s; v.Ye[j] *= s; } ynorm = so.ls_maxstep; }
        for (int j = lane; j < nlev; j += 32) { v.Wm[j] = v.P[j] - lambda * v.Ym[j]; v.We[j] = v.T[j] - lambda * v.Ye[j]; }
        phase = PH_LS_FULL;
        if (nfuncs >= so.max_funcs && so.max_funcs >= 0) { last_reason = SNES_DIVERGED_FUNCTION_COUNT; phase = -1; }
      }
      __syncwarp();
    }

    if (phase == -1) {
      tot_nf += nfuncs;
      if (last_reason < 0) {
        cuts += 1; dt_iter = 0.5 * dt_iter; dtInv = 1.0 / dt_iter;
        for (int j = lane; j < nlev; j += 32) { v.P[j] = v.Pp[j]; v.T[j] = v.Tp[j]; }
        if (cuts > 20) { converged = 0; phase = PH_DONE; }
        else { for (int j = lane; j < nlev; j += 32) { v.Wm[j] = v.Pp[j]; v.We[j] = v.Tp[j]; } phase = PH_INIT; }
      } else {
        converged = 1; time_done += dt_iter; tot_its += its;
        for (int j = lane; j < nlev; j += 32) { v.Pp[j] = v.P[j]; v.Tp[j] = v.T[j]; }
        if (time_done >= A.dt) phase = PH_DONE;
        else { for (int j = lane; j < nlev; j += 32) { v.Wm[j] = v.P[j]; v.We[j] = v.T[j]; } phase = PH_INIT; }
      }
      its = 0; nfuncs = 0;
      // optional give-up budget (mppgpu_set_step_budget; not in the reference, off by default): a column that has burnt this
      // many residual evaluations inside one StepDT fails like one that ran out of dt cuts, instead of stalling the batch
      if (so.step_budget > 0 && tot_nf >= so.step_budget && phase != PH_DONE) { converged = 0; last_reason = SNES_DIVERGED_FUNCTION_COUNT; phase = PH_DONE; }
      __syncwarp();
      if (phase == PH_DONE) break;
    }

    // ---- residual at W ----
    for (int j = lane; j < nlev; j += 32) {
      SatParams sp; sp.sat_res = A.sat_res[c0 + j]; sp.alpha = A.alpha[c0 + j]; sp.m = A.lam[c0 + j]; sp.n = A.vgn ? A.vgn[c0 + j] : 0.0;
      if (A.pu) { sp.pu = A.pu[c0 + j]; sp.ps = A.ps[c0 + j]; sp.b2 = A.b2[c0 + j]; sp.b3 = A.b3[c0 + j]; } else sp.pu = sp.ps = sp.b2 = sp.b3 = 0.0;
      THCell c;
      th_cell_compute(A, sp, A.tkdry[c0 + j], v.Wm[j], v.We[j], c);
      v.sat[j] = c.sat; v.kr[j] = c.kr; v.dsat[j] = c.dsat; v.dkr[j] = c.dkr;
      v.denm[j] = c.den_m; v.dPm[j] = c.ddenP_m; v.dTm[j] = c.ddenT_m; v.dene[j] = c.den_e; v.dPe[j] = c.ddenP_e; v.dTe[j] = c.ddenT_e;
      v.ul[j] = c.ul; v.hl[j] = c.hl; v.dulT[j] = c.dulT; v.dhlT[j] = c.dhlT; v.dulP[j] = c.dulP; v.dhlP[j] = c.dhlP; v.tc[j] = c.tc; v.dtcP[j] = c.dtcP;
    }
    __syncwarp();
    for (int j = lane; j < nlev - 1; j += 32) {
      const double dzu = A.dz[c0 + j], dzd = A.dz[c0 + j + 1];
      const double dist_up = 0.5 * dzu, dist_dn = 0.5 * dzd, upw = dist_up / (dist_up + dist_dn);
      const double gfac = FMWH2O * ((dist_up + dist_dn) * (A.uz * (-GRAVITY_CONSTANT)));
      const double pmu = A.perm[c0 + j], pmd = A.perm[c0 + j + 1];
      const double Dqm = (pmu * pmd) / (dist_up * pmd + dist_dn * pmu);
      const double peu = PERM_E_AT(j), ped = PERM_E_AT(j + 1);
      const double Dqe = (peu * ped) / (dist_up * ped + dist_dn * peu);
      FluxIn um = {v.Wm[j], v.kr[j], 0, v.denm[j], 0, 0}, dm = {v.Wm[j + 1], v.kr[j + 1], 0, v.denm[j + 1], 0, 0};
      FluxIn ue = {v.Wm[j], v.kr[j], 0, v.dene[j], 0, 0}, de = {v.Wm[j + 1], v.kr[j + 1], 0, v.dene[j + 1], 0, 0};
      double fl, a1, a2, a3, a4, mfl;
      th_rich_flux(um, dm, upw, Dqm, gfac, area, fl, a1, a2, a3, a4);
      th_rich_flux(ue, de, upw, Dqe, gfac, area, mfl, a1, a2, a3, a4);
      const double ku = v.tc[j], kd = v.tc[j + 1];
      const double kod = (ku * kd) / (dist_up * kd + dist_dn * ku);
      const double h = (mfl <= 0.0) ? v.hl[j] : v.hl[j + 1];
      v.fm[j] = fl;
      v.fe[j] = mfl * h + (-kod * (v.We[j] - v.We[j + 1]) * area);
    }
    __syncwarp();
    double sg = 0.0, sw = 0.0;
    for (int j = lane; j < nlev; j += 32) {
      const double por = A.por[c0 + j], vol = area * A.dz[c0 + j], csol = A.csol[c0 + j];
      const double am = por * v.denm[j] * v.sat[j] * vol * dtInv;
      const double ae = (por * v.dene[j] * v.sat[j] * v.ul[j] + (1.0 - por) * 2700.0 * csol * (v.We[j] - 273.15)) * vol * dtInv;
      if (phase == PH_INIT) { v.accm[j] = am; v.acce[j] = ae; }
      double gm = am - v.accm[j], ge = ae - v.acce[j];
      if (j > 0) { gm = gm + v.fm[j - 1]; ge = ge + v.fe[j - 1]; }
      if (j < nlev - 1) { gm = gm - v.fm[j]; ge = ge - v.fe[j]; }
      for (int k = 0; k < 4; ++k) if (bcs[k].cell == j) {
        const double dzc = A.dz[c0 + j];
        double fl, a1, a2, a3, a4;
        if (bcs[k].ieqn == 1) {
          FluxIn dn = {v.Wm[j], v.kr[j], 0, v.denm[j], 0, 0};
          th_rich_flux(bcs[k].fin, dn, 0.0, bcs[k].Dq, bcs[k].gfac, area, fl, a1, a2, a3, a4);
          gm = gm + fl;
        } else {
          FluxIn dn = {v.Wm[j], v.kr[j], 0, v.dene[j], 0, 0};
          th_rich_flux(bcs[k].fin, dn, 0.0, bcs[k].Dq, bcs[k].gfac, area, fl, a1, a2, a3, a4);
          const double kod = v.tc[j] / (0.0 + 0.5 * dzc);
          const double h = (fl <= 0.0) ? bcs[k].hl : v.hl[j];
          ge = ge + (fl * h + (-kod * (bcs[k].T - v.We[j]) * area));
        }
      }
      gm = gm - v.srcm[j];
      ge = ge + v.srce[j];                      // heat-rate sources ADD to the residual in the reference (:1478)
      v.Gm[j] = gm; v.Ge[j] = ge;
      sg += gm * gm + ge * ge; sw += v.Wm[j] * v.Wm[j] + v.We[j] * v.We[j];
    }
    const double g2 = warp_sum(sg), w2 = warp_sum(sw);
    nfuncs += 1;
    __syncwarp();

    bool take = false;
    const bool g_bad = !(g2 == g2) || (g2 > 1.7e308);
    const bool out_of_funcs = (nfuncs >= so.max_funcs && so.max_funcs >= 0);
    auto new_trial = [&]() { for (int j = lane; j < nlev; j += 32) { v.Wm[j] = v.P[j] - lambda * v.Ym[j]; v.We[j] = v.T[j] - lambda * v.Ye[j]; } };
    if (A.eval_x && phase == PH_INIT) {
      for (int j = lane; j < nlev; j += 32) { v.Wm[j] = A.eval_x[2 * (c0 + j)]; v.We[j] = A.eval_x[2 * (c0 + j) + 1]; }
      phase = PH_EVAL; __syncwarp(); continue;
    } else if (phase == PH_EVAL) {
      for (int j = lane; j < nlev; j += 32) {
        v.P[j] = v.Wm[j]; v.T[j] = v.We[j]; v.Fm[j] = v.Gm[j]; v.Fe[j] = v.Ge[j];
        A.eval_f[2 * (c0 + j)] = v.Gm[j]; A.eval_f[2 * (c0 + j) + 1] = v.Ge[j];
      }
      phase = PH_EVAL_J; __syncwarp(); continue;
    } else if (phase == PH_INIT) {
      take = true;
    } else if (phase == PH_LS_FULL) {
      if (g_bad) {
        if (lambda <= so.ls_minlambda) { last_reason = SNES_DIVERGED_FNORM_NAN; phase = -1; }
        else if (out_of_funcs)         { last_reason = SNES_DIVERGED_FUNCTION_COUNT; phase = -1; }
        else { lambda = .5 * lambda; new_trial(); }
      } else if (.5 * g2 <= .5 * f2 + lambda * so.ls_alpha * initslope) take = true;
      else if (so.stol * xnorm > ynorm) { last_reason = SNES_CONVERGED_SNORM_RELATIVE; phase = -1; }
      else if (out_of_funcs) { last_reason = SNES_DIVERGED_FUNCTION_COUNT; phase = -1; }
      else {
        double lt = -initslope / (g2 - f2 - 2.0 * lambda * initslope);
        lambdaprev = lambda; gprev = g2;
        if (lt > .5 * lambda) lt = .5 * lambda;
        lambda = (lt <= .1 * lambda) ? .1 * lambda : lt;
        new_trial(); phase = PH_LS_QUAD; ls_count = 0;
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
        const double t1 = .5 * (g2 - f2) - lambda * initslope, t2 = .5 * (gprev - f2) - lambdaprev * initslope;
        const double a = (t1 / (lambda * lambda) - t2 / (lambdaprev * lambdaprev)) / (lambda - lambdaprev);
        const double b = (-lambdaprev * t1 / (lambda * lambda) + lambda * t2 / (lambdaprev * lambdaprev)) / (lambda - lambdaprev);
        double d = b * b - 3 * a * initslope;
        if (d < 0.0) d = 0.0;
        double lt = (a == 0.0) ? -initslope / (2.0 * b) : (-b + sqrt(d)) / (3.0 * a);
        lambdaprev = lambda; gprev = g2;
        if (lt > .5 * lambda) lt = .5 * lambda;
        lambda = (lt <= .1 * lambda) ? .1 * lambda : lt;
        new_trial(); phase = PH_LS_CUBIC;
      }
    }
    if (take) {
      for (int j = lane; j < nlev; j += 32) { v.P[j] = v.Wm[j]; v.T[j] = v.We[j]; v.Fm[j] = v.Gm[j]; v.Fe[j] = v.Ge[j]; }
      fnorm = sqrt(g2);
      int reason = 0;
      if (phase == PH_INIT) {
        its = 0; ttol = fnorm * so.rtol; rnorm0 = fnorm;
        if (g_bad) reason = SNES_DIVERGED_FNORM_NAN; else if (fnorm < so.atol) reason = SNES_CONVERGED_FNORM_ABS;
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

  if (col_ok && !A.eval_x) {
    for (int j = lane; j < nlev; j += 32) {
      A.x_out[2 * (c0 + j)] = v.P[j]; A.x_out[2 * (c0 + j) + 1] = v.T[j];
      if (converged) {
        A.liq_sat[c0 + j] = v.sat[j];
        A.mass[c0 + j] = A.por[c0 + j] * v.denm[j] * FMWH2O * v.sat[j] * (area * A.dz[c0 + j]);
      }
    }
    if (lane == 0) {
      A.stat_its[col] = tot_its; A.stat_reason[col] = last_reason; A.stat_cuts[col] = cuts; A.stat_nf[col] = tot_nf;
      double *bp = A.block_partials + (size_t)blockIdx.x * 9;
      for (int k = 0; k < 9; ++k) bp[k] = 0.0;
      bp[5] = (double)tot_its; bp[6] = converged ? 0.0 : 1.0; bp[7] = (double)cuts; bp[8] = (double)last_reason;
    }
  } else if (lane == 0 && !A.eval_x) {
    double *bp = A.block_partials + (size_t)blockIdx.x * 9;
    for (int k = 0; k < 8; ++k) bp[k] = 0.0;
    bp[8] = 2147483647.0;
  }
}

#undef PERM_E_AT

struct THState {
  cudaStream_t stream = nullptr;
  int ncol = 0, nlev = 0, orientation = 311;
  SnesOpts so;
  bool soils_set = false;
  int satfunc_name = 0, density_type = DENSITY_TGDPB01, iee_type = INT_ENERGY_ENTHALPY_CONSTANT;
  const double *d_dz = nullptr, *d_area = nullptr;
  double *por = nullptr, *perm = nullptr, *sat_res = nullptr, *alpha = nullptr, *lam = nullptr, *vgn = nullptr,
         *pu = nullptr, *ps = nullptr, *b2 = nullptr, *b3 = nullptr, *tkdry = nullptr, *csol = nullptr, *perm_e = nullptr;
  double *x = nullptr;              // interleaved (P,T), 2*ncells
  double *Pout = nullptr, *Tout = nullptr, *liq_sat = nullptr, *mass = nullptr;   // de-interleaved views for GetDataForCLM
  bool views_stale = true;
};

#endif  // MPP_STEP_KERNEL_TU

}  // namespace mpp
