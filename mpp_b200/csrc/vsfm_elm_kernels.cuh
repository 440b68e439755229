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
// column is its own "rank", consistent with the per-column Newton iteration of the step kernel.  One lane per cell, G lanes per
// column (every cell array is streamed once, coalesced); the parts that are sequential in the layer index -- the water-table
// search, the thickness sum and the running drainage total, the column totals -- are evaluated redundantly by all lanes of a column
// from shuffled values, in the reference's summation order, so the packed sources are bit-identical to a sequential evaluation.
#pragma once
#include <cuda_runtime.h>
#include "physics.cuh"

namespace mpp {

struct ElmArgs {
  int ncol, nlev, nlevsoi, max_patch_per_col;
  int col0, col_end;                       // the columns this launch covers (a chunk of the pipeline, or the whole batch)
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
  int *retry_list;                         // ... and which ones (compacted, any order)
  // outputs
  double *smp_l, *soilp, *qcharge, *abs_err;
};

template <int G>
__global__ void elm_pack_kernel(const ElmArgs A)
{
  constexpr unsigned FULL = 0xffffffffu;
  const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int c = A.col0 + (int)(tid / G), j = (int)(tid % G);           // lane j owns layer j + 1 of the reference
  const int nlev = A.nlev, nlevsoi = A.nlevsoi;
  const bool col_ok = c < A.col_end;
  const bool on = col_ok && ((A.active == nullptr) || A.active[c] != 0);
  const bool cell = on && j < nlev;
  const long long ic = (long long)c * nlev + j;
  if (col_ok && j == 0) {
    A.iter_count[c] = 0; A.diverged[c] = 0; A.status[c] = 0; A.mask[c] = on ? 1 : 0;
    A.rtol[c] = A.rtol0; A.stol[c] = A.stol0; A.dt_rem[c] = A.dtime; A.abs_err[c] = 0.0;
  }
  const double area = 1.0, conv = area * DENH2O * 1.0e-3;              // flux_unit_conversion [mm/s] -> [kg/s] (:330)
  // ---- loads ----
  double rootr = 0.0, dz = 0.0, liq = 1.0, ice = 0.0, perched = 0.0, mass = 0.0, zi_j = 0.0;
  double qtran = 0.0, qinfl = 0.0, dews = 0.0, dewg = 0.0, subs = 0.0, fh = 0.0, qd = 0.0, zw = 0.0, snowl = 0.0, negs = 0.0;
  int snl = -1;
  if (cell) {
    rootr = A.rootr_col[ic]; dz = A.dz[ic]; liq = A.h2osoi_liq[ic]; ice = A.h2osoi_ice[ic]; perched = A.mflx_drain_perched[ic]; mass = A.soe_mass[ic];
    zi_j = A.zi[(long long)c * (nlev + 1) + j + 1];                    // zi(c, j+1): interface below this layer
  }
  if (on) {
    qtran = A.qflx_tran_veg_col[c]; qinfl = A.qflx_infl[c]; dews = A.qflx_dew_snow[c]; dewg = A.qflx_dew_grnd[c]; subs = A.qflx_sub_snow[c];
    fh = A.frac_h2osfc[c]; qd = A.qflx_drain[c]; zw = A.zwt[c]; snowl = A.mflx_snowlyr_col[c]; negs = A.mflx_neg_snow[c]; snl = A.snl[c];
  }
  // ---- :204-240 root-fraction weighting over the patches of the column ----
  if (A.col_pfti && on) {
    const int np = A.col_npfts[c], p0 = A.col_pfti[c];
    double r = 0.0, temp = 0.0;
    for (int pi = 0; pi < A.max_patch_per_col; ++pi) if (pi < np) {
      const int pp = p0 + pi;
      if (!A.pft_active[pp]) continue;
      const double q = A.qflx_tran_veg_pft[pp], wt = A.pft_wtcol[pp];
      if (cell && j < nlevsoi) r = r + A.rootr_pft[(long long)pp * nlev + j] * q * wt;
      temp = temp + q * wt;
    }
    if (cell && j < nlevsoi) { if (temp != 0.0) r = r / temp; rootr = r; A.rootr_col[ic] = r; }
  }
  // ---- :340-372 drainage below the water table (1-based layer numbers as in the reference) ----
  double drain = 0.0, qd_new = qd;
  {
    // jwt = (first layer with zwt <= zi) - 1, nlev if none; at least 1
    const unsigned hit = __ballot_sync(FULL, cell && (zw <= zi_j));
    const unsigned mine = (hit >> ((threadIdx.x & 31) / G * G)) & ((G == 32) ? 0xffffffffu : ((1u << G) - 1u));
    int jwt = mine ? (__ffs(mine) - 1) : nlev;                         // (j_first) - 1 in 1-based numbering == 0-based index of the first hit
    if (jwt < 1) jwt = 1;
    double dzsum = 0.0, tot = 0.0;
    for (int k = 1; k <= nlev; ++k) { const double v = __shfl_sync(FULL, dz, k - 1, G); if (k >= jwt) dzsum = dzsum + v; }
    double ql = 0.0;
    if (qd > 0.0 && cell && (j + 1) >= jwt) {
      ql = qd * dz / dzsum;
      const double avail = liq - A.watmin;
      if (ql * A.dtime > avail) ql = avail / A.dtime;
      drain = -ql * conv;
    }
    for (int k = 1; k <= nlev; ++k) { const double v = __shfl_sync(FULL, ql, k - 1, G); if (k >= jwt) tot = tot + v; }
    if (qd > 0.0) qd_new = tot;
  }
  // ---- sources, frac_liq_sat, totals ----
  const double et = (cell && j < nlevsoi) ? -qtran * rootr * conv : 0.0;
  const double dr = drain + perched;                                    // :404
  const double infl = qinfl * conv;
  double dew = 0.0, sub = 0.0;
  if (snl >= 0) { const double wet = 1.0 - fh; dew = (dews + dewg) * wet * conv; sub = -subs * wet * conv; }
  const double snow = snowl * area + negs * area;
  if (cell) {
    A.c_et[ic] = et; A.c_drain[ic] = dr;
    const double fi = ice / (liq + ice);                                // :441
    A.frac_ice[ic] = fi; A.frac_liq[ic] = 1.0 - fi;
  }
  double tot_et = 0.0, tot_drain = 0.0, mass_beg = 0.0;
  for (int k = 0; k < nlev; ++k) {
    tot_et = tot_et + __shfl_sync(FULL, et, k, G); tot_drain = tot_drain + __shfl_sync(FULL, dr, k, G); mass_beg = mass_beg + __shfl_sync(FULL, mass, k, G);
  }
  if (on && j == 0) {
    A.qflx_drain[c] = qd_new; A.mflx_snowlyr_col[c] = 0.0;
    A.c_infl[c] = infl; A.c_dew[c] = dew; A.c_snow[c] = snow; A.c_sub[c] = sub;
    A.mass_beg[c] = mass_beg;
    A.tot_flux[c] = tot_et + infl + dew + tot_drain + snow + sub + 0.0;           // :583-589 (no lateral flux on the 1-D path)
  }
}

template <int G>
__global__ void elm_decide_kernel(const ElmArgs A)
{
  constexpr unsigned FULL = 0xffffffffu;
  const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int c = A.col0 + (int)(tid / G), j = (int)(tid % G);
  const int nlev = A.nlev;
  const bool col_ok = c < A.col_end;
  const int m = col_ok ? A.mask[c] : 0;
  if (__all_sync(FULL, m == 0)) return;                                 // no column of this warp took part in the StepDT that just ran
  const bool run = m != 0, cell = run && j < nlev;
  const long long ic = (long long)c * nlev + j;
  const double area = 1.0;
  int reason = 0, iter = 0, dv = 0;
  double dt_rem = 0.0, t_done = 0.0, mass_beg = 0.0, tot_flux = 0.0, rtol = 0.0, stol = 0.0;
  if (run) {
    reason = A.stat_reason[c]; iter = A.iter_count[c] + 1; dv = A.diverged[c]; dt_rem = A.dt_rem[c]; t_done = A.t_done[c];
    mass_beg = A.mass_beg[c]; tot_flux = A.tot_flux[c]; rtol = A.rtol[c]; stol = A.stol[c];
  }
  const bool conv = run && reason > 0;
  double fi = 0.0, mass = 0.0, smp = 0.0, pres = 0.0;
  if (cell && conv) { fi = A.frac_ice[ic]; mass = A.soe_mass[ic]; smp = A.soe_smp[ic] * 1000.0; pres = A.soe_pressure[ic]; }   // [m] -> [mm]
  int next = 0, ok = 0;
  if (run && !conv) {                                                   // .not. converged (:645-660)
    stol = 1.0e-10; dv += 1; dt_rem = dt_rem - t_done;
    if (dv > 1 && cell) A.frac_liq[ic] = 1.0;
    next = 1;
  }
  // converged (:662-905): unpack, water table, mass balance
  if (cell && conv) {
    A.h2osoi_liq[ic] = (1.0 - fi) * mass / area; A.h2osoi_ice[ic] = fi * mass / area;
    A.smp_l[ic] = smp; A.soilp[ic] = pres;
  }
  double mass_end = 0.0;
  for (int k = nlev; k >= 1; --k) mass_end = mass_end + __shfl_sync(FULL, mass, k - 1, G);   // the reference sums from the bottom up
  // jwt: the deepest layer with a negative matric potential (1-based), -1 if none
  const unsigned neg = __ballot_sync(FULL, cell && conv && smp < 0.0);
  const unsigned mine = (neg >> ((threadIdx.x & 31) / G * G)) & ((G == 32) ? 0xffffffffu : ((1u << G) - 1u));
  const int jwt = mine ? (32 - __clz(mine)) : -1;
  const int ja = (jwt >= 1 && jwt < nlev) ? jwt : 1;
  const double s0 = __shfl_sync(FULL, smp, ja - 1, G), s1 = __shfl_sync(FULL, smp, ja, G);
  if (conv && j == 0) {
    const double *zic = A.zi + (long long)c * (nlev + 1);
    const double err = fabs(mass_beg - mass_end + tot_flux * A.dtime);
    A.abs_err[c] = err;
    A.qcharge[c] = 0.0;
    if (jwt == -1 || jwt == nlev) A.zwt[c] = zic[nlev];
    else {
      const double z_dn = (zic[jwt - 1] + zic[jwt]) / 2.0, z_up = (zic[jwt] + zic[jwt + 1]) / 2.0;
      A.zwt[c] = (0.0 - s0) / (s0 - s1) * (z_dn - z_up) + z_dn;
    }
    if (err >= 1.0e-5) {                                                // max_abs_mass_error_col (:880-897)
      if (reason == SNES_CONVERGED_FNORM_RELATIVE) rtol = rtol / 10.0;
      else if (reason == SNES_CONVERGED_SNORM_RELATIVE) stol = stol / 10.0;
      dt_rem = A.dtime;
      next = 2;                                                         // PreStepDT: back to soln_prev_clm
    } else ok = 1;
  }
  if (run && j == 0) {
    if (!ok && iter >= 10) next = 0;                                    // max_iter_count: the reference calls endrun here
    A.iter_count[c] = iter; A.diverged[c] = dv; A.dt_rem[c] = dt_rem; A.rtol[c] = rtol; A.stol[c] = stol;
    A.status[c] = ok; A.mask[c] = next;
    if (next) A.retry_list[atomicAdd(A.pending, 1)] = c;
  }
}

// columns whose solve failed (status 0), not counting the ones the column filter excludes
__global__ void elm_count_failed_kernel(int ncol, const int *active, const int *status, int *nfailed)
{
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  const bool bad = c < ncol && (active == nullptr || active[c] != 0) && status[c] == 0;
  const unsigned m = __ballot_sync(0xffffffffu, bad);
  if ((threadIdx.x & 31) == 0 && m) atomicAdd(nfailed, __popc(m));
}

// Per-block partials of the handle's nine reductions (sums: mass before / after, sources * dt, boundary exchange; maxima: |mass error|,
// Newton iterations, any column failed, dt cuts; worst SNES reason) after an elm_solve, from the per-column arrays -- same
// [nblocks][9] layout as the step kernels write, folded by reduce_partials_kernel, so mppgpu_vsfm_mass_balance and the NCCL
// gather of parallel.py see the whole MPPVSFMALM_Solve (retries included), not its last launch.
__global__ void elm_column_partials_kernel(int ncol, double dtime, const int *active, const double *mass_beg, const double *col_mass, const double *tot_flux,
                                           const double *abs_err, const int *status, const int *stat_its, const int *stat_reason, const int *stat_cuts,
                                           double *partials)
{
  __shared__ double sh[9][256];
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  double v[9];
  for (int k = 0; k < 8; ++k) v[k] = 0.0;
  v[8] = 2147483647.0;
  if (c < ncol && (active == nullptr || active[c] != 0)) {
    v[0] = mass_beg[c]; v[1] = col_mass[c]; v[2] = tot_flux[c] * dtime; v[4] = abs_err[c]; v[5] = (double)stat_its[c];
    v[6] = status[c] ? 0.0 : 1.0; v[7] = (double)stat_cuts[c]; v[8] = (double)stat_reason[c];
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
  if (threadIdx.x == 0) for (int k = 0; k < 9; ++k) partials[(size_t)blockIdx.x * 9 + k] = sh[k][0];
}

}  // namespace mpp
