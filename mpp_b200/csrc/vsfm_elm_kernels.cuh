// vsfm_elm_kernels.cuh -- the host-model side of one VSFM coupling step, MPPVSFMALM_Solve
// (src/driver/alm/MPPVSFMALM_Driver.F90:204-923), as device kernels: ELM hands over its raw column arrays and gets its raw
// column arrays back; packing, the per-column retry decisions and unpacking never leave the GPU.
//
//   elm_pack_kernel    :204-240 root-fraction weighting of transpiration over a column's patches (optional)
//                      :325-372 source/sink packing -- ET by root fraction, infiltration, dew and sublimation (snow-free columns
//                               only), drainage spread over the layers below the water table in proportion to their thickness
//                               and limited by the liquid water they hold, snow-layer disappearance
//                      :404     + perched drainage;  :435-450 frac_ice / frac_liq_sat;  :552-601 mass and flux totals per column
//   elm_decide_kernel  :628-923 what the retry loop does after each StepDT, per column: a diverged step continues with the
//                               remaining time and stol = 1e-10 (a second divergence drops the ice impedance, frac_liq_sat = 1);
//                               a converged step is unpacked (h2osoi_liq / h2osoi_ice by the ice fraction, smp_l in mm, soil
//                               pressure, water table by interpolating the matric potential, qcharge = 0) and, if its
//                               mass-balance error is >= 1e-5 kg, redone from soln_prev_clm with rtol or stol tightened
//                               tenfold according to the convergence reason; at most 10 StepDT calls
// The reference takes these decisions once per MPI rank (global convergence flag, global maximum of the mass error); here each
// column is its own "rank", consistent with the per-column Newton iteration of the step kernel.  One thread per column: the
// drainage distribution and the water-table search are sequential in the layer index, and the arrays are cell-ordered, so a
// warp's accesses to one layer are 8*nlev bytes apart -- every sector is still used in full over the layer loop (L1), i.e.
// DRAM traffic stays at the algorithmic bytes; these kernels are a few percent of the step.
#pragma once
#include <cuda_runtime.h>
#include "physics.cuh"

namespace mpp {

struct ElmArgs {
  int ncol, nlev, nlevsoi, max_patch_per_col;
  double dtime, watmin, rtol0, stol0;
  const int *active;                       // column filter or nullptr
  // patch level (optional)
  const int *col_pfti, *col_npfts, *pft_active; const double *pft_wtcol, *rootr_pft, *qflx_tran_veg_pft;
  // column level inputs
  double *rootr_col; const double *qflx_tran_veg_col, *qflx_infl, *qflx_dew_snow, *qflx_dew_grnd, *qflx_sub_snow, *frac_h2osfc;
  const int *snl;
  double *qflx_drain, *zwt; const double *zi, *dz;
  double *h2osoi_liq, *h2osoi_ice, *mflx_snowlyr_col; const double *mflx_neg_snow, *mflx_drain_perched;
  // the six COND_MASS_RATE conditions of MPPVSFMALM_Initialize.F90:836-858 and the SoE mailbox
  double *c_infl, *c_et, *c_dew, *c_drain, *c_snow, *c_sub, *frac_liq;
  const double *soe_mass, *soe_smp, *soe_pressure;
  // per-column scratch / retry state
  double *frac_ice, *mass_beg, *tot_flux, *dt_rem, *rtol, *stol; const double *t_done;
  int *iter_count, *diverged, *mask, *status; const int *stat_reason;
  int *pending;                            // number of columns that need another StepDT
  // outputs
  double *smp_l, *soilp, *qcharge, *abs_err;
};

__global__ void elm_pack_kernel(const ElmArgs A)
{
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= A.ncol) return;
  const int nlev = A.nlev, nlevsoi = A.nlevsoi;
  const long long off = (long long)c * nlev;
  const bool on = (A.active == nullptr) || A.active[c] != 0;
  A.iter_count[c] = 0; A.diverged[c] = 0; A.status[c] = 0; A.mask[c] = on ? 1 : 0;
  A.rtol[c] = A.rtol0; A.stol[c] = A.stol0; A.dt_rem[c] = A.dtime; A.abs_err[c] = 0.0;
  if (!on) return;
  const double area = 1.0, conv = area * DENH2O * 1.0e-3;              // flux_unit_conversion [mm/s] -> [kg/s] (:330)
  if (A.col_pfti) {                                                     // :204-240
    double temp = 0.0;
    for (int j = 0; j < nlevsoi; ++j) A.rootr_col[off + j] = 0.0;
    const int np = A.col_npfts[c], p0 = A.col_pfti[c];
    for (int pi = 0; pi < A.max_patch_per_col; ++pi) if (pi < np) {
      const int pp = p0 + pi;
      if (!A.pft_active[pp]) continue;
      const double q = A.qflx_tran_veg_pft[pp], wt = A.pft_wtcol[pp];
      for (int j = 0; j < nlevsoi; ++j) A.rootr_col[off + j] = A.rootr_col[off + j] + A.rootr_pft[(long long)pp * nlev + j] * q * wt;
      temp = temp + q * wt;
    }
    if (temp != 0.0) for (int j = 0; j < nlevsoi; ++j) A.rootr_col[off + j] = A.rootr_col[off + j] / temp;
  }
  double tot_et = 0.0, tot_drain = 0.0, mass_beg = 0.0;
  const double qtran = A.qflx_tran_veg_col[c];
  for (int j = 0; j < nlev; ++j) A.c_drain[off + j] = 0.0;
  const double infl = A.qflx_infl[c] * conv;
  double dew = 0.0, sub = 0.0;
  if (A.snl[c] >= 0) {
    const double wet = 1.0 - A.frac_h2osfc[c];
    dew = (A.qflx_dew_snow[c] + A.qflx_dew_grnd[c]) * wet * conv;
    sub = -A.qflx_sub_snow[c] * wet * conv;
  }
  const double qd = A.qflx_drain[c];
  if (qd > 0.0) {                                                       // :340-372, layer numbers 1-based as in the reference
    const double *zic = A.zi + (long long)c * (nlev + 1);
    const double zw = A.zwt[c];
    int jwt = nlev;
    for (int j = 1; j <= nlev; ++j) if (zw <= zic[j]) { jwt = j - 1; break; }
    if (jwt < 1) jwt = 1;
    double dzsum = 0.0, tot = 0.0;
    for (int j = jwt; j <= nlev; ++j) dzsum = dzsum + A.dz[off + j - 1];
    for (int j = jwt; j <= nlev; ++j) {
      double ql = qd * A.dz[off + j - 1] / dzsum;
      const double avail = A.h2osoi_liq[off + j - 1] - A.watmin;
      if (ql * A.dtime > avail) ql = avail / A.dtime;
      tot = tot + ql;
      A.c_drain[off + j - 1] = -ql * conv;
    }
    A.qflx_drain[c] = tot;
  }
  const double snow = A.mflx_snowlyr_col[c] * area + A.mflx_neg_snow[c] * area;
  A.mflx_snowlyr_col[c] = 0.0;
  for (int j = 0; j < nlev; ++j) {
    const double et = (j < nlevsoi) ? -qtran * A.rootr_col[off + j] * conv : 0.0;
    const double dr = A.c_drain[off + j] + A.mflx_drain_perched[off + j];      // :404
    A.c_et[off + j] = et; A.c_drain[off + j] = dr;
    const double liq = A.h2osoi_liq[off + j], ice = A.h2osoi_ice[off + j];
    const double fi = ice / (liq + ice);                                           // :441
    A.frac_ice[off + j] = fi; A.frac_liq[off + j] = 1.0 - fi;
    tot_et = tot_et + et; tot_drain = tot_drain + dr; mass_beg = mass_beg + A.soe_mass[off + j];
  }
  A.c_infl[c] = infl; A.c_dew[c] = dew; A.c_snow[c] = snow; A.c_sub[c] = sub;
  A.mass_beg[c] = mass_beg;
  A.tot_flux[c] = tot_et + infl + dew + tot_drain + snow + sub + 0.0;             // :583-589 (no lateral flux on the 1-D path)
}

__global__ void elm_decide_kernel(const ElmArgs A)
{
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= A.ncol) return;
  const int m = A.mask[c];
  if (m == 0) return;                                                   // not part of the StepDT that just ran
  const int nlev = A.nlev;
  const long long off = (long long)c * nlev;
  const double area = 1.0;
  const int iter = A.iter_count[c] + 1;
  A.iter_count[c] = iter;
  const int reason = A.stat_reason[c];
  int next = 0, ok = 0;
  if (reason <= 0) {                                                    // .not. converged (:645-660)
    A.stol[c] = 1.0e-10;
    const int dv = A.diverged[c] + 1; A.diverged[c] = dv;
    A.dt_rem[c] = A.dt_rem[c] - A.t_done[c];
    if (dv > 1) for (int j = 0; j < nlev; ++j) A.frac_liq[off + j] = 1.0;
    next = 1;
  } else {                                                              // :662-905
    const double *zic = A.zi + (long long)c * (nlev + 1);
    int jwt = -1;
    double mass_end = 0.0;
    for (int j = nlev; j >= 1; --j) {
      const long long ic = off + j - 1;
      const double fi = A.frac_ice[ic], mass = A.soe_mass[ic];
      A.h2osoi_liq[ic] = (1.0 - fi) * mass / area;
      A.h2osoi_ice[ic] = fi * mass / area;
      mass_end = mass_end + mass;
      const double smp = A.soe_smp[ic] * 1000.0;                        // [m] -> [mm]
      A.smp_l[ic] = smp;
      if (jwt == -1 && smp < 0.0) jwt = j;
      A.soilp[ic] = A.soe_pressure[ic];
    }
    const double err = fabs(A.mass_beg[c] - mass_end + A.tot_flux[c] * A.dtime);
    A.abs_err[c] = err;
    A.qcharge[c] = 0.0;
    if (jwt == -1 || jwt == nlev) A.zwt[c] = zic[nlev];
    else {
      const double z_dn = (zic[jwt - 1] + zic[jwt]) / 2.0, z_up = (zic[jwt] + zic[jwt + 1]) / 2.0;
      const double s0 = A.smp_l[off + jwt - 1], s1 = A.smp_l[off + jwt];
      A.zwt[c] = (0.0 - s0) / (s0 - s1) * (z_dn - z_up) + z_dn;
    }
    if (err >= 1.0e-5) {                                                // max_abs_mass_error_col (:880-897)
      if (reason == SNES_CONVERGED_FNORM_RELATIVE) A.rtol[c] = A.rtol[c] / 10.0;
      else if (reason == SNES_CONVERGED_SNORM_RELATIVE) A.stol[c] = A.stol[c] / 10.0;
      A.dt_rem[c] = A.dtime;
      next = 2;                                                         // PreStepDT: back to soln_prev_clm
    } else ok = 1;
  }
  if (!ok && iter >= 10) next = 0;                                      // max_iter_count: the reference calls endrun here
  A.status[c] = ok;
  A.mask[c] = next;
  if (next) atomicAdd(A.pending, 1);
}

}  // namespace mpp
