// thermal_kernels.cuh -- fused soil heat-conduction time step (T-based, KSP path) for batches of independent columns.
//
// One launch = one sysofeqns%StepDT (SOEBaseStepDT_KSP, src/mpp/soe/SystemOfEquationsBaseType.F90:555-647):
//   PreSolve / ComputeRHS / ComputeOperators  src/mpp/soe/SystemOfEquationsThermalType.F90:412-759
//   conductivity + heat capacity              src/mpp/auxvar/ThermalKSPTemperatureSoilAuxType.F90:71-171
//   Accum, Divergence, DiffHeatFlux           src/mpp/ge/GoveqnThermalKSPTemperatureSoilType.F90:671-1003
//   ComputeOperatorsDiag                      src/mpp/ge/GoveqnThermalKSPTemperatureSoilType.F90:1007-1229
//   KSPSolve (GMRES + ILU(0) on a tridiagonal AIJ == exact LU == Thomas)
//
// Mapping (B200-first; DESIGN.md "thermal kernel").  The step is linear and HBM-bound, so the kernel is laid out for
// bandwidth: every array is read once, in the reference's own cell order (icell = c*nlev + j), by a lane-per-cell
// pass (GROUP lanes per column, fully coalesced, neighbour layers via warp shuffles) that turns the twelve inputs
// of a cell into its tridiagonal row (b, c, rhs; the matrix is symmetric so a_j = c_{j-1}) parked in shared
// memory; then a thread-per-column pass runs the Thomas algorithm out of shared memory (odd row stride =>
// conflict-free), and a last lane-per-cell pass streams the new temperatures back, again fully coalesced.
#pragma once
#include <cuda_runtime.h>
#include "physics.cuh"

namespace mpp {

#ifndef MPP_ALLOC_SLACK
#define MPP_ALLOC_SLACK 4096     // bytes behind every device buffer of the library (tile kernels may read past the end of the batch)
#endif
constexpr int TH_TILE = 128;            // columns per block == threads per block
constexpr int TH_MAX_SS = 4;

struct ThermalArgs {
  int ncol, nlev, nlevsoi;
  int istsoil, istcrop, istice, istice_mec, istwet;
  double dt, cnfac;
  // static
  const double *por, *tkmg, *tkdry, *csol, *dz, *area;
  const double *dist_up, *dist_dn;     // per cell: connection j -> j+1 (nullptr: dz/2 of the two cells)
  int dist_uniform;                     // every column has the same distances (ELM's fixed vertical grid): per-layer copies below,
  double lay_du[32], lay_dd[32];        //   read from the kernel's constant bank instead of 16 B per cell from HBM
  const int *lun_type;                  // per column
  // per-step inputs (SoE mailbox)
  const double *T_in, *liq, *ice, *snow_water, *tuning;
  const int *nsnow, *active;
  // boundary conditions: slot 0 = SOIL_TOP_CELLS, slot 1 = SOIL_BOTTOM_CELLS
  int bc_type[2];                       // 0 none, 505 Dirichlet, 507 heat flux
  const double *bc_value[2], *bc_dhsdT[2], *bc_frac[2];
  const double *bc_active[2];           // 0/1 flags as doubles (real-valued VAR_ACTIVE, thermal_mms_problem.F90:633)
  int top_is_first;
  double stale_area;                    // see ThermalKSPTempSoilDivergence's Dirichlet branch (:883-908)
  int nss; const double *ss_value[TH_MAX_SS]; int ss_region[TH_MAX_SS];
  double *T_out;
  double *therm_cond, *heat_cap;        // optional diagnostics (nullptr to skip)
};

// ThermKSPTempSoilAuxVarCompute, ThermalKSPTemperatureSoilAuxType.F90:71-171
__device__ __forceinline__ void thermal_auxvar(const ThermalArgs &A, int itype, bool shallow, double T, double liq, double ice,
                                               double snoww, int nsnow, double por, double tkmg, double tkdry, double csol,
                                               double dz, double &tk, double &hc)
{
  const double LN_TKWAT = -0.56211891815354120;   // ln(0.57)
  const double LN_TKICE = 0.82855181756614820;    // ln(2.29)
  tk = 0.0; hc = 0.0;
  if (itype == A.istsoil || itype == A.istcrop) {
    if (shallow) {
      const double l = liq * (1.0 / DENH2O), i = ice * (1.0 / DENICE);      // water / ice depth [m]
      double satw = (l + i) * rcp(dz * por);
      satw = (satw < 1.0) ? satw : 1.0;
      // branch-free with the lean log / exp / reciprocal of physics.cuh (arguments are positive and normal; a dry cell
      // evaluates them at 1 and discards the result): dke = max(0, log10(satw) + 1) unfrozen, satw frozen
      const bool wet = satw > (double).1e-6f;
      const double sw = wet ? satw : 1.0, li = wet ? (l + i) : 1.0;
      const double lg = mpp_log(sw) * 0.43429448190325182765 + 1.0;
      const double dke = (T >= TFRZ) ? ((lg > 0.0) ? lg : 0.0) : sw;
      const double fl = l * rcp(li);                                        // the reference's common 1/dz cancels
      // tkmg * tkwat^(fl por) * tkice^((1-fl) por)
      const double dksat = tkmg * mpp_exp(por * (fl * LN_TKWAT + (1.0 - fl) * LN_TKICE));
      tk = wet ? dke * dksat + (1.0 - dke) * tkdry : tkdry;
      hc = csol * (1.0 - por) * dz + ice * CPICE + liq * CPLIQ;
      if (nsnow == 0) hc = hc + snoww * CPICE;
    } else {
      tk = THK_BEDROCK;
      hc = csol * (1.0 - por) * dz + ice * CPICE + liq * CPLIQ;
    }
    hc = hc * rcp(dz);
  } else if (itype == A.istwet) {
    if (shallow) {
      tk = (T < TFRZ) ? TKICE : TKWAT;
      hc = ice * CPICE + liq * CPLIQ;
      if (nsnow == 0) hc = hc + snoww * CPICE;
      hc = hc * rcp(dz);
    } else { tk = THK_BEDROCK; hc = csol; }
  } else if (itype == A.istice || itype == A.istice_mec) {
    tk = (T < TFRZ) ? TKICE : TKWAT;
    hc = ice * CPICE + liq * CPLIQ;
    if (nsnow == 0) hc = hc + snoww * CPICE;
    hc = hc * rcp(dz);
  }
}

// Parallel cyclic reduction across a GROUP-lane group (one tridiagonal row per lane, identity rows pad the group), rows
// kept in normalised form (unit diagonal): 3 shuffled doubles per side per stage, one lean reciprocal per stage.
template <int GROUP>
__device__ __forceinline__ double thermal_pcr(double a, double b, double c, double d)
{
  constexpr unsigned FULL = 0xffffffffu;
  double r = rcp(b);
  double al = a * r, ga = c * r, de = d * r;
#pragma unroll
  for (int s = 1; s < GROUP; s <<= 1) {
    const double al_m = __shfl_up_sync(FULL, al, s, GROUP),   ga_m = __shfl_up_sync(FULL, ga, s, GROUP);
    const double de_m = __shfl_up_sync(FULL, de, s, GROUP);
    const double al_p = __shfl_down_sync(FULL, al, s, GROUP), ga_p = __shfl_down_sync(FULL, ga, s, GROUP);
    const double de_p = __shfl_down_sync(FULL, de, s, GROUP);
    r  = rcp(1.0 - al * ga_m - ga * al_p);
    de = (de - al * de_m - ga * de_p) * r;
    al = (-al * al_m) * r;
    ga = (-ga * ga_p) * r;
  }
  return de;
}

// One lane per soil cell, GROUP lanes per column: every array is streamed once in the reference's cell order
// (fully coalesced), neighbour layers come from warp shuffles, the tridiagonal system is solved in registers by
// parallel cyclic reduction, and the new temperature goes straight back to HBM.  No shared memory, no block
// barriers: occupancy is bounded by registers only, which is what hides the HBM latency.
#ifndef THERMAL_MIN_BLOCKS
#define THERMAL_MIN_BLOCKS 8
#endif
template <int GROUP>
__global__ void __launch_bounds__(TH_TILE, THERMAL_MIN_BLOCKS)
thermal_step_kernel(const ThermalArgs A)
{
  constexpr unsigned FULL = 0xffffffffu;
  const int nlev = A.nlev;
  const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int col = (int)(tid / GROUP), j = (int)(tid % GROUP);
  const int jtop = A.top_is_first ? 0 : nlev - 1, jbot = A.top_is_first ? nlev - 1 : 0;
  const double dt = A.dt, cnfac = A.cnfac;
  const bool valid = (col < A.ncol) && (j < nlev);
  const long long cell = (long long)col * nlev + j;

  double T = 0.0, tk = 1.0, hc = 0.0, dz = 1.0, area = 1.0, tf = 1.0, du = 0.5, dd = 0.5;
  double ssv[TH_MAX_SS], bcH[2], bcdH[2], bcfr[2];
  int act = 0;
#pragma unroll
  for (int k = 0; k < TH_MAX_SS; ++k) ssv[k] = 0.0;
#pragma unroll
  for (int k = 0; k < 2; ++k) { bcH[k] = 0.0; bcdH[k] = 0.0; bcfr[k] = 0.0; }
  if (valid) {
    T = A.T_in[cell]; dz = A.dz[cell]; area = A.area[col]; tf = A.tuning[cell]; act = A.active[cell];
    // condition values: issued with the other loads, consumed after the aux-var math
#pragma unroll
    for (int k = 0; k < TH_MAX_SS; ++k) if (k < A.nss) {
      if (A.ss_region[k] == 403) ssv[k] = A.ss_value[k][cell];
      else if (j == (A.ss_region[k] == 401 ? jtop : jbot)) ssv[k] = A.ss_value[k][col];
    }
#pragma unroll
    for (int k = 0; k < 2; ++k) if (A.bc_type[k] == 507 && j == (k == 0 ? jtop : jbot)) {
      bcH[k] = A.bc_value[k][col]; bcdH[k] = A.bc_dhsdT[k][col]; bcfr[k] = A.bc_frac[k][col];
    }
    const double liq = A.liq[cell], ice = A.ice[cell], snoww = A.snow_water[cell];
    const double por = A.por[cell], tkmg = A.tkmg[cell], tkdry = A.tkdry[cell], csol = A.csol[cell];
    const int nsnow = A.nsnow[cell];
    if (A.dist_uniform) { du = A.lay_du[j]; dd = A.lay_dd[j]; }
    else if (A.dist_up) { du = A.dist_up[cell]; dd = A.dist_dn[cell]; }
    thermal_auxvar(A, A.lun_type[col], j < A.nlevsoi, T, liq, ice, snoww, nsnow, por, tkmg, tkdry, csol, dz, tk, hc);
    if (A.therm_cond) { A.therm_cond[cell] = tk; A.heat_cap[cell] = hc; }
  }
  const double vol = area * dz;
  // connection j -> j+1 (owned by lane j)
  const double T_d = __shfl_down_sync(FULL, T, 1, GROUP), tk_d = __shfl_down_sync(FULL, tk, 1, GROUP);
  const double dz_d = __shfl_down_sync(FULL, dz, 1, GROUP);
  const int act_d = __shfl_down_sync(FULL, act, 1, GROUP);
  if (!A.dist_up && !A.dist_uniform) { du = 0.5 * dz; dd = 0.5 * dz_d; }
  double cval = 0.0, flux = 0.0;
  if (valid && j < nlev - 1 && act && act_d) {
    // kav / dist with kav the distance-weighted harmonic mean: tk tk_d (du+dd) / (tk dd + tk_d du) / (du+dd)
    const double kod = tk * tk_d * rcp(tk * dd + tk_d * du) * area;
    flux = -kod * (T - T_d);                                               // DiffHeatFlux * area  (:976-1003)
    cval = (1.0 - cnfac) * kod;                                            // ComputeOperatorsDiag (:1112)
  }
  const double cval_m = __shfl_up_sync(FULL, cval, 1, GROUP), flux_m = __shfl_up_sync(FULL, flux, 1, GROUP);
  double bb, rhs;
  if (act) { bb = hc * vol * rcp(dt * tf); rhs = bb * T; } else { bb = 1.0; rhs = 0.0; }
  rhs = rhs + cnfac * flux; bb += cval;
  if (j > 0) { rhs = rhs - cnfac * flux_m; bb += cval_m; }
  if (valid && act) {
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      if (A.bc_type[k] == 0 || j != (k == 0 ? jtop : jbot)) continue;
      if (A.bc_type[k] == 507) {                    // COND_HEAT_FLUX: value = H - dH/dT * T_cell (GoveqnThermalKSP...:344-348)
        const double H = bcH[k], dH = bcdH[k], fr = bcfr[k];
        rhs = rhs + (H - dH * T) * fr * area;
        bb += -fr * ((area == 1.0) ? dH : mpp_pow_rare(dH, area));   // `-frac*dhsdT**area*factor` (:1215); x**1 == x exactly
      } else if (A.bc_active[k][col] != 0.0) {      // COND_DIRICHLET
        double tkb, hcb;
        const double Tb = A.bc_value[k][col];
        // boundary aux vars never receive is_soil_shallow / water contents (MultiPhysicsProbThermal.F90:195-203)
        thermal_auxvar(A, A.lun_type[col], false, Tb, 0.0, 0.0, 0.0, 0, A.por[cell], A.tkmg[cell], A.tkdry[cell], A.csol[cell], dz, tkb, hcb);
        const double bdu = 0.0, bdd = 0.5 * dz, dist = bdu + bdd;
        const double kav = tkb * tk * dist / (tkb * bdd + tk * bdu);
        rhs = rhs + kav / dist * Tb * A.stale_area;
        bb += A.bc_frac[k][col] * (1.0 - cnfac) * kav / dist * area;
      }
    }
#pragma unroll
    for (int k = 0; k < TH_MAX_SS; ++k) rhs = rhs + ssv[k];      // COND_HEAT_RATE (zero where the condition does not touch this cell)
  }
  if (!valid) { bb = 1.0; rhs = 0.0; }
  // symmetric tridiagonal row: a_j = -cval_{j-1}, c_j = -cval_j   (KSPSolve: exact for a tridiagonal matrix)
  const double x = thermal_pcr<GROUP>((j > 0) ? -cval_m : 0.0, bb, -cval, rhs);
  if (valid) A.T_out[cell] = x;
}

// Two cells per lane, 8 lanes per column, 4 columns per warp (nlev <= 16): half the shuffles and a 3-stage instead of a
// 4-stage reduction for the same cells.  The lane's second row is eliminated in-lane (odd-even step), the remaining
// one-row-per-lane system goes through normalised PCR over 8 lanes, the second unknown is back-substituted in-lane.
// Both cells' aux-var chains (log, exp, reciprocals) sit in one basic block and overlap.
struct ThermalCell {
  double T, tk, hc, dz, tf, du, dd, ssum, bb, rhs;
  int act;
};

// The arrays thermal_step2_body reads and writes: the launch's own (global memory, batch indices) or a tile-local view whose input
// pointers aim at a shared-memory stage filled by bulk-async copies (tile indices).  Scalars and per-layer tables stay in `A`.
struct ThermalPtrs {
  int ncol;
  const double *T_in, *dz, *tuning, *liq, *ice, *snow_water, *por, *tkmg, *tkdry, *csol, *area;
  const int *active, *nsnow, *lun_type;
  const double *bc_value[2], *bc_dhsdT[2], *bc_frac[2], *bc_active[2], *ss_value[TH_MAX_SS];
  double *T_out, *therm_cond, *heat_cap;
};
__device__ __forceinline__ ThermalPtrs thermal_ptrs(const ThermalArgs &A)
{
  ThermalPtrs V;
  V.ncol = A.ncol; V.T_in = A.T_in; V.dz = A.dz; V.tuning = A.tuning; V.liq = A.liq; V.ice = A.ice; V.snow_water = A.snow_water;
  V.por = A.por; V.tkmg = A.tkmg; V.tkdry = A.tkdry; V.csol = A.csol; V.area = A.area; V.active = A.active; V.nsnow = A.nsnow; V.lun_type = A.lun_type;
#pragma unroll
  for (int k = 0; k < 2; ++k) { V.bc_value[k] = A.bc_value[k]; V.bc_dhsdT[k] = A.bc_dhsdT[k]; V.bc_frac[k] = A.bc_frac[k]; V.bc_active[k] = A.bc_active[k]; }
#pragma unroll
  for (int k = 0; k < TH_MAX_SS; ++k) V.ss_value[k] = A.ss_value[k];
  V.T_out = A.T_out; V.therm_cond = A.therm_cond; V.heat_cap = A.heat_cap;
  return V;
}

template <int LPC>
__device__ __forceinline__ double thermal_pcr_unit(double al, double ga, double de)
{
  constexpr unsigned FULL = 0xffffffffu;
#pragma unroll
  for (int s = 1; s < LPC; s <<= 1) {
    const double al_m = __shfl_up_sync(FULL, al, s, LPC),   ga_m = __shfl_up_sync(FULL, ga, s, LPC);
    const double de_m = __shfl_up_sync(FULL, de, s, LPC);
    const double al_p = __shfl_down_sync(FULL, al, s, LPC), ga_p = __shfl_down_sync(FULL, ga, s, LPC);
    const double de_p = __shfl_down_sync(FULL, de, s, LPC);
    const double r = rcp(1.0 - al * ga_m - ga * al_p);
    de = (de - al * de_m - ga * de_p) * r;
    al = (-al * al_m) * r;
    ga = (-ga * ga_p) * r;
  }
  return de;
}

__device__ __forceinline__ void thermal_cell_load(const ThermalArgs &A, const ThermalPtrs &V, ThermalCell &c, bool valid, long long cell, int col, int j, int jtop, int jbot)
{
  c.T = 0.0; c.tk = 1.0; c.hc = 0.0; c.dz = 1.0; c.tf = 1.0; c.du = 0.5; c.dd = 0.5; c.ssum = 0.0; c.act = 0;
  if (valid) {
    c.T = V.T_in[cell]; c.dz = V.dz[cell]; c.tf = V.tuning[cell]; c.act = V.active[cell];
    const double liq = V.liq[cell], ice = V.ice[cell], snoww = V.snow_water[cell];
    const double por = V.por[cell], tkmg = V.tkmg[cell], tkdry = V.tkdry[cell], csol = V.csol[cell];
    const int nsnow = V.nsnow[cell];
    if (A.dist_uniform) { c.du = A.lay_du[j]; c.dd = A.lay_dd[j]; }
    else if (A.dist_up) { c.du = A.dist_up[cell]; c.dd = A.dist_dn[cell]; }
#pragma unroll
    for (int k = 0; k < TH_MAX_SS; ++k) if (k < A.nss) {          // COND_HEAT_RATE
      if (A.ss_region[k] == 403) c.ssum += V.ss_value[k][cell];
      else if (j == (A.ss_region[k] == 401 ? jtop : jbot)) c.ssum += V.ss_value[k][col];
    }
    thermal_auxvar(A, V.lun_type[col], j < A.nlevsoi, c.T, liq, ice, snoww, nsnow, por, tkmg, tkdry, csol, c.dz, c.tk, c.hc);
    if (V.therm_cond) { V.therm_cond[cell] = c.tk; V.heat_cap[cell] = c.hc; }
  }
}

// boundary-condition terms of the cell that sits at the top / bottom of its column (as thermal_step_kernel)
__device__ __forceinline__ void thermal_cell_bc(const ThermalArgs &A, const ThermalPtrs &V, ThermalCell &c, long long cell, int col, int j, int jtop, int jbot, double area)
{
  const double cnfac = A.cnfac;
#pragma unroll
  for (int k = 0; k < 2; ++k) {
    if (A.bc_type[k] == 0 || j != (k == 0 ? jtop : jbot)) continue;
    if (A.bc_type[k] == 507) {                    // COND_HEAT_FLUX: value = H - dH/dT * T_cell (GoveqnThermalKSP...:344-348)
      const double H = V.bc_value[k][col], dH = V.bc_dhsdT[k][col], fr = V.bc_frac[k][col];
      c.rhs = c.rhs + (H - dH * c.T) * fr * area;
      c.bb += -fr * ((area == 1.0) ? dH : mpp_pow_rare(dH, area));   // `-frac*dhsdT**area*factor` (:1215); x**1 == x exactly
    } else if (V.bc_active[k][col] != 0.0) {      // COND_DIRICHLET
      double tkb, hcb;
      const double Tb = V.bc_value[k][col];
      thermal_auxvar(A, V.lun_type[col], false, Tb, 0.0, 0.0, 0.0, 0, V.por[cell], V.tkmg[cell], V.tkdry[cell], V.csol[cell], c.dz, tkb, hcb);
      const double bdu = 0.0, bdd = 0.5 * c.dz, dist = bdu + bdd;
      const double kav = tkb * c.tk * dist / (tkb * bdd + c.tk * bdu);
      c.rhs = c.rhs + kav / dist * Tb * A.stale_area;
      c.bb += V.bc_frac[k][col] * (1.0 - cnfac) * kav / dist * area;
    }
  }
}

// The step of one column group: `V` holds the launch's own arrays (global memory, `col` a batch column) or a tile-local view whose
// input pointers aim at a shared-memory stage filled by bulk-async copies (`col` a column of the tile): the arithmetic is the same code.
template <int LPC>
__device__ __forceinline__ void thermal_step2_body(const ThermalArgs &A, const ThermalPtrs &V, int col, int l)
{
  constexpr unsigned FULL = 0xffffffffu;
  const int nlev = A.nlev;
  const int j0 = 2 * l, j1 = 2 * l + 1;
  const int jtop = A.top_is_first ? 0 : nlev - 1, jbot = A.top_is_first ? nlev - 1 : 0;
  const double dt = A.dt, cnfac = A.cnfac;
  const bool col_ok = col < V.ncol, va = col_ok && j0 < nlev, vb = col_ok && j1 < nlev;
  const long long cell0 = (long long)col * nlev + j0;
  const double area = col_ok ? V.area[col] : 1.0;

  ThermalCell a, b;
  thermal_cell_load(A, V, a, va, cell0, col, j0, jtop, jbot);
  thermal_cell_load(A, V, b, vb, cell0 + 1, col, j1, jtop, jbot);

  // connections 2l -> 2l+1 (in-lane) and 2l+1 -> 2(l+1) (the next lane's first cell)
  const double T_n = __shfl_down_sync(FULL, a.T, 1, LPC), tk_n = __shfl_down_sync(FULL, a.tk, 1, LPC), dz_n = __shfl_down_sync(FULL, a.dz, 1, LPC);
  const int act_n = __shfl_down_sync(FULL, a.act, 1, LPC);
  if (!A.dist_up && !A.dist_uniform) { a.du = 0.5 * a.dz; a.dd = 0.5 * b.dz; b.du = 0.5 * b.dz; b.dd = 0.5 * dz_n; }
  double cv_a = 0.0, fl_a = 0.0, cv_b = 0.0, fl_b = 0.0;
  if (vb && a.act && b.act) {
    const double kod = a.tk * b.tk * rcp(a.tk * a.dd + b.tk * a.du) * area;
    fl_a = -kod * (a.T - b.T); cv_a = (1.0 - cnfac) * kod;                 // DiffHeatFlux * area (:976-1003), ComputeOperatorsDiag (:1112)
  }
  if (vb && j1 < nlev - 1 && b.act && act_n) {
    const double kod = b.tk * tk_n * rcp(b.tk * b.dd + tk_n * b.du) * area;
    fl_b = -kod * (b.T - T_n); cv_b = (1.0 - cnfac) * kod;
  }
  const double cv_p = __shfl_up_sync(FULL, cv_b, 1, LPC), fl_p = __shfl_up_sync(FULL, fl_b, 1, LPC);

  if (a.act) { a.bb = a.hc * (area * a.dz) * rcp(dt * a.tf); a.rhs = a.bb * a.T; } else { a.bb = 1.0; a.rhs = 0.0; }
  if (b.act) { b.bb = b.hc * (area * b.dz) * rcp(dt * b.tf); b.rhs = b.bb * b.T; } else { b.bb = 1.0; b.rhs = 0.0; }
  a.rhs = a.rhs + cnfac * fl_a; a.bb += cv_a;
  if (l > 0) { a.rhs = a.rhs - cnfac * fl_p; a.bb += cv_p; }
  b.rhs = b.rhs + cnfac * fl_b; b.bb += cv_b;
  b.rhs = b.rhs - cnfac * fl_a; b.bb += cv_a;
  if (va && a.act) { thermal_cell_bc(A, V, a, cell0, col, j0, jtop, jbot, area); a.rhs = a.rhs + a.ssum; }
  if (vb && b.act) { thermal_cell_bc(A, V, b, cell0 + 1, col, j1, jtop, jbot, area); b.rhs = b.rhs + b.ssum; }
  if (!va) { a.bb = 1.0; a.rhs = 0.0; }
  if (!vb) { b.bb = 1.0; b.rhs = 0.0; }

  // symmetric tridiagonal rows: sub_a = -cv_p, sup_a = sub_b = -cv_a, sup_b = -cv_b   (KSPSolve: exact for a tridiagonal matrix)
  const double sub_a = (l > 0) ? -cv_p : 0.0, sup_a = -cv_a, sub_b = -cv_a, sup_b = -cv_b;
  const double rb = rcp(b.bb);
  const double bs = sub_b * rb, bu = sup_b * rb, bf = b.rhs * rb;          // y_b = bf - bs y_a(l) - bu y_a(l+1)
  const double bs_p = __shfl_up_sync(FULL, bs, 1, LPC), bu_p = __shfl_up_sync(FULL, bu, 1, LPC), bf_p = __shfl_up_sync(FULL, bf, 1, LPC);
  const double rB = rcp(a.bb - sub_a * bu_p - sup_a * bs);
  const double al = (-sub_a * bs_p) * rB, ga = (-sup_a * bu) * rB, de = (a.rhs - sub_a * bf_p - sup_a * bf) * rB;
  const double xa = thermal_pcr_unit<LPC>(al, ga, de);
  const double xa_n = __shfl_down_sync(FULL, xa, 1, LPC);
  const double xb = bf - bs * xa - bu * xa_n;
  if (va) V.T_out[cell0] = xa;
  if (vb) V.T_out[cell0 + 1] = xb;
}

#ifndef THERMAL2_MIN_BLOCKS
#define THERMAL2_MIN_BLOCKS 8      // 64 registers (a 100-byte spill), 32 warps per SM: 0.376 ms per Mi columns against 0.390 at 6 blocks / 79 registers
#endif
template <int LPC>
__global__ void __launch_bounds__(TH_TILE, THERMAL2_MIN_BLOCKS)
thermal_step2_kernel(const ThermalArgs A)
{
  const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  // (prefetching the late heat-flux boundary values into L1 here measured slower: 0.404 vs 0.376 ms per Mi columns)
  thermal_step2_body<LPC>(A, thermal_ptrs(A), (int)(tid / LPC), (int)(tid % LPC));
}

// ---- bulk-async (1-D TMA) variant: measured, slower, kept as an option --------------------------------------------------------------
// Question: thermal_step2_kernel issues its ~13 input streams as ordinary loads at the top of each warp's life and
// then computes; 40 % of its stall samples sit on the first use of a loaded value and DRAM runs at half its copy rate
// (profiles/r1_thermal_v4.md).  Does taking the loads out of the warps' instruction streams -- persistent blocks, one cp.async.bulk per
// input array and tile into a shared-memory stage, an mbarrier armed with the byte count, the copy of tile i+1 under the arithmetic
// of tile i -- reach the HBM roofline?
// Answer, measured on B200 at 1 Mi columns x 15: no.  This kernel (2 stages per block of 4 warps, 4 blocks per
// SM) needs 0.481 ms against 0.390 ms for the register kernel; a per-warp pipeline (one 5.9 KB stage per warp, the stage handed back
// to the copy engine as soon as the lanes hold their inputs in registers) needs 0.565 ms with 16 warps per SM, 0.59 ms with 20
// (spilling) and 0.68 ms with 12.  Time goes DOWN with the number of resident warps and does not care whether the inputs are already
// in shared memory: the step is bound by the issue / dependent latency of its fp64 chains (log, exp, four reciprocals, three PCR
// stages per lane), and shared-memory staging costs warps (11.8 KB of stage per warp against ~10 KB of registers).
// The kernel stays selectable (mppgpu_thermal_set_bulk_copy(h, 1)) and bit-identical to the register kernel (same device function).
// Supported shape: nlev <= 16, heat-flux or no boundary condition per region, at most TMA_MAX_SS_CELL per-cell and TMA_MAX_SS_COL
// per-column heat-rate sources, default or per-layer-uniform connection distances.  The last tile may hang over the end of the
// batch: every device buffer of the library carries MPP_ALLOC_SLACK bytes of slack, and the over-read columns are computed and dropped.
constexpr int TMA_TILE_COLS = 16, TMA_STAGES = 2, TMA_MAX_SS_CELL = 2, TMA_MAX_SS_COL = 2, TMA_BLOCKS_PER_SM = 4;

__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long *bar, int count)
{ asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(bar)), "r"(count) : "memory"); }
__device__ __forceinline__ void mbar_arrive_expect_tx(unsigned long long *bar, unsigned bytes)
{ asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(bar)), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, unsigned parity)
{
  asm volatile("{\n .reg .pred p;\n WAIT_%=:\n mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n @p bra DONE_%=;\n bra WAIT_%=;\n DONE_%=:\n}"
               :: "r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, unsigned bytes, unsigned long long *bar)
{
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               :: "r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

struct ThermalTmaPlan {            // which arrays travel, in stage order (host-built: thermal_tma_plan)
  int n_cell_d, n_cell_i, n_col_d, n_col_i;      // counts of per-cell double / int and per-column double / int arrays
  const void *cell_d[16], *cell_i[2], *col_d[12], *col_i[1];
  int stage_bytes;
  int ntiles;                      // tiles of TMA_TILE_COLS columns
};

template <int LPC>
__global__ void __launch_bounds__(TH_TILE, TMA_BLOCKS_PER_SM)
thermal_step2_tma_kernel(const ThermalArgs A, const ThermalTmaPlan P)
{
  static_assert(TH_TILE / LPC == TMA_TILE_COLS, "one tile = the columns one block advances");
  extern __shared__ __align__(128) unsigned char tma_smem[];
  __shared__ unsigned long long full_bar[TMA_STAGES];
  const int nlev = A.nlev, cells = TMA_TILE_COLS * nlev;
  const unsigned cd_bytes = (unsigned)cells * 8u, ci_bytes = (unsigned)cells * 4u, kd_bytes = TMA_TILE_COLS * 8u, ki_bytes = TMA_TILE_COLS * 4u;
  if (threadIdx.x == 0) {
    for (int s = 0; s < TMA_STAGES; ++s) mbar_init(&full_bar[s], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  auto issue = [&](int tile, int stage) {                     // thread 0 only
    unsigned char *dst = tma_smem + (size_t)stage * P.stage_bytes;
    const size_t c0 = (size_t)tile * cells, k0 = (size_t)tile * TMA_TILE_COLS;
    mbar_arrive_expect_tx(&full_bar[stage], (unsigned)P.stage_bytes);
    for (int i = 0; i < P.n_cell_d; ++i) { bulk_g2s(dst, (const double *)P.cell_d[i] + c0, cd_bytes, &full_bar[stage]); dst += cd_bytes; }
    for (int i = 0; i < P.n_cell_i; ++i) { bulk_g2s(dst, (const int *)P.cell_i[i] + c0, ci_bytes, &full_bar[stage]); dst += ci_bytes; }
    for (int i = 0; i < P.n_col_d; ++i)  { bulk_g2s(dst, (const double *)P.col_d[i] + k0, kd_bytes, &full_bar[stage]); dst += kd_bytes; }
    for (int i = 0; i < P.n_col_i; ++i)  { bulk_g2s(dst, (const int *)P.col_i[i] + k0, ki_bytes, &full_bar[stage]); dst += ki_bytes; }
  };
  const int first = blockIdx.x, stride = gridDim.x;
  if (threadIdx.x == 0) {
    for (int s = 0; s < TMA_STAGES; ++s) if (first + s * stride < P.ntiles) issue(first + s * stride, s);
  }
  const int col_local = threadIdx.x / LPC, l = threadIdx.x % LPC;
  int it = 0;
  for (int tile = first; tile < P.ntiles; tile += stride, ++it) {
    const int stage = it % TMA_STAGES;
    mbar_wait(&full_bar[stage], (unsigned)((it / TMA_STAGES) & 1));
    // tile-local view of the stage: the same field order as thermal_tma_plan
    ThermalPtrs V = thermal_ptrs(A);
    {
      const unsigned char *q = tma_smem + (size_t)stage * P.stage_bytes;
      auto takeD = [&](unsigned bytes) { const double *r = (const double *)q; q += bytes; return r; };
      auto takeI = [&](unsigned bytes) { const int *r = (const int *)q; q += bytes; return r; };
      V.T_in = takeD(cd_bytes); V.dz = takeD(cd_bytes); V.tuning = takeD(cd_bytes); V.liq = takeD(cd_bytes); V.ice = takeD(cd_bytes);
      V.snow_water = takeD(cd_bytes); V.por = takeD(cd_bytes); V.tkmg = takeD(cd_bytes); V.tkdry = takeD(cd_bytes); V.csol = takeD(cd_bytes);
#pragma unroll
      for (int k = 0; k < TH_MAX_SS; ++k) if (k < A.nss && A.ss_region[k] == 403) V.ss_value[k] = takeD(cd_bytes);
      V.active = takeI(ci_bytes); V.nsnow = takeI(ci_bytes);
      V.area = takeD(kd_bytes);
#pragma unroll
      for (int k = 0; k < 2; ++k) if (A.bc_type[k] == 507) { V.bc_value[k] = takeD(kd_bytes); V.bc_dhsdT[k] = takeD(kd_bytes); V.bc_frac[k] = takeD(kd_bytes); }
#pragma unroll
      for (int k = 0; k < TH_MAX_SS; ++k) if (k < A.nss && A.ss_region[k] != 403) V.ss_value[k] = takeD(kd_bytes);
      V.lun_type = takeI(ki_bytes);
    }
    const long long cbase = (long long)tile * cells;
    V.ncol = min(TMA_TILE_COLS, A.ncol - tile * TMA_TILE_COLS);
    V.T_out = A.T_out + cbase;
    if (A.therm_cond) { V.therm_cond = A.therm_cond + cbase; V.heat_cap = A.heat_cap + cbase; }
    thermal_step2_body<LPC>(A, V, col_local, l);
    __syncthreads();                                            // every thread is done with this stage
    if (threadIdx.x == 0 && tile + TMA_STAGES * stride < P.ntiles) {
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy reads above, async-proxy writes below
      issue(tile + TMA_STAGES * stride, stage);
    }
  }
}

// Any nlev: one thread per column straight from global memory (correctness path for tall columns).
__global__ void thermal_step_generic_kernel(const ThermalArgs A, double *work /* 4 * ncells */)
{
  const int col = blockIdx.x * blockDim.x + threadIdx.x;
  if (col >= A.ncol) return;
  const int nlev = A.nlev;
  const long long c0 = (long long)col * nlev, N = (long long)A.ncol * nlev;
  double *b = work + c0, *c = work + N + c0, *d = work + 2 * N + c0, *tkv = work + 3 * N + c0;
  const int jtop = A.top_is_first ? 0 : nlev - 1, jbot = A.top_is_first ? nlev - 1 : 0;
  const double area = A.area[col], dt = A.dt, cnfac = A.cnfac;
  const int itype = A.lun_type[col];
  for (int j = 0; j < nlev; ++j) {
    const long long cell = c0 + j;
    double tk, hc;
    thermal_auxvar(A, itype, j < A.nlevsoi, A.T_in[cell], A.liq[cell], A.ice[cell], A.snow_water[cell], A.nsnow[cell],
                   A.por[cell], A.tkmg[cell], A.tkdry[cell], A.csol[cell], A.dz[cell], tk, hc);
    if (A.therm_cond) { A.therm_cond[cell] = tk; A.heat_cap[cell] = hc; }
    tkv[j] = tk; c[j] = 0.0;
    if (A.active[cell]) { b[j] = hc * (area * A.dz[cell]) / (dt * A.tuning[cell]); d[j] = b[j] * A.T_in[cell]; }
    else { b[j] = 1.0; d[j] = 0.0; }
  }
  for (int j = 0; j < nlev - 1; ++j) {
    const long long cell = c0 + j;
    if (!A.active[cell] || !A.active[cell + 1]) continue;
    const double du = A.dist_up ? A.dist_up[cell] : 0.5 * A.dz[cell], dd = A.dist_dn ? A.dist_dn[cell] : 0.5 * A.dz[cell + 1];   // (dist_uniform is only used for nlev <= 32)
    const double dist = du + dd;
    const double kav = tkv[j] * tkv[j + 1] * dist / (tkv[j] * dd + tkv[j + 1] * du);
    const double flux = -kav * (A.T_in[cell] - A.T_in[cell + 1]) / dist * area;
    const double cval = (1.0 - cnfac) * kav / dist * area;
    d[j] = d[j] + cnfac * flux; d[j + 1] = d[j + 1] - cnfac * flux;
    b[j] += cval; b[j + 1] += cval; c[j] = -cval;
  }
  for (int k = 0; k < 2; ++k) {
    if (A.bc_type[k] == 0) continue;
    const int j = (k == 0) ? jtop : jbot;
    const long long cell = c0 + j;
    if (!A.active[cell]) continue;
    const double T = A.T_in[cell], dz = A.dz[cell];
    if (A.bc_type[k] == 507) {
      const double H = A.bc_value[k][col], dH = A.bc_dhsdT[k][col], fr = A.bc_frac[k][col];
      d[j] = d[j] + (H - dH * T) * fr * area;
      b[j] += -fr * ((area == 1.0) ? dH : mpp_pow_rare(dH, area));
    } else if (A.bc_active[k][col] != 0.0) {
      double tkb, hcb;
      const double Tb = A.bc_value[k][col];
      thermal_auxvar(A, itype, false, Tb, 0.0, 0.0, 0.0, 0, A.por[cell], A.tkmg[cell], A.tkdry[cell], A.csol[cell], dz, tkb, hcb);
      const double du = 0.0, dd = 0.5 * dz, dist = du + dd;
      const double kav = tkb * tkv[j] * dist / (tkb * dd + tkv[j] * du);
      d[j] = d[j] + kav / dist * Tb * A.stale_area;
      b[j] += A.bc_frac[k][col] * (1.0 - cnfac) * kav / dist * area;
    }
  }
  for (int k = 0; k < A.nss; ++k) {
    if (A.ss_region[k] == 403) { for (int j = 0; j < nlev; ++j) if (A.active[c0 + j]) d[j] = d[j] + A.ss_value[k][c0 + j]; }
    else { const int j = (A.ss_region[k] == 401) ? jtop : jbot; if (A.active[c0 + j]) d[j] = d[j] + A.ss_value[k][col]; }
  }
  // Thomas, a_j = c_{j-1}
  double cprev = c[0], cp = c[0] / b[0], dp = d[0] / b[0];
  c[0] = cp; d[0] = dp;
  for (int i = 1; i < nlev; ++i) {
    const double a_i = cprev;
    cprev = c[i];
    const double m = b[i] - a_i * c[i - 1];
    c[i] = c[i] / m;
    d[i] = (d[i] - a_i * d[i - 1]) / m;
  }
  for (int i = nlev - 2; i >= 0; --i) d[i] = d[i] - c[i] * d[i + 1];
  for (int j = 0; j < nlev; ++j) A.T_out[c0 + j] = d[j];
}

struct ThermalState {
  cudaStream_t stream = nullptr;
  int ncol = 0, nlev = 0, orientation = 311, nlevsoi = 0;
  int istsoil = 1, istcrop = 2, istice = 3, istice_mec = 4, istwet = 6;
  double cnfac = 0.5;                    // mpp_varcon.F90:28
  bool soils_set = false, custom_dist = false, diagnostics = false, dist_uniform = false;
  double lay_du[32], lay_dd[32];
  const double *d_dz = nullptr, *d_area = nullptr;   // owned by the handle
  double *por = nullptr, *tkmg = nullptr, *tkdry = nullptr, *csol = nullptr, *dist_up = nullptr, *dist_dn = nullptr;
  int *lun_type = nullptr;
  double *T_clm = nullptr, *T_work = nullptr, *T_cur = nullptr;
  double *liq = nullptr, *ice = nullptr, *snow_water = nullptr, *tuning = nullptr, *frac = nullptr, *aux_dz = nullptr,
         *aux_dist_up = nullptr, *aux_dist_dn = nullptr, *therm_cond = nullptr, *heat_cap = nullptr, *work = nullptr;
  int *nsnow = nullptr, *active = nullptr;
  double stale_area = 1.0;
  // snow + standing-surface-water coupling (thermal_snow_kernels.cuh): mailbox arrays then hold ncol*(nsno+1+nlev) entries
  bool snow_mode = false, force_two_rows = false; int nsno = 0; size_t nall = 0;
  bool bulk_copy = false, tma_attr_set = false;      // mppgpu_thermal_set_bulk_copy: the bulk-async (1-D TMA) kernel where its shape fits
  double *soil_top_dist_dn = nullptr, *hs[3] = {nullptr, nullptr, nullptr}, *dhs[3] = {nullptr, nullptr, nullptr},
         *frac_soil = nullptr, *sabg_snow = nullptr, *sabg_soil = nullptr;
  int *snow_top_id = nullptr;
  // staging for mppgpu_thermal_elm_solve (ELM's raw column arrays); one allocation, carved up
  double *elm_stage = nullptr; int *elm_snl = nullptr;
  int elm_chunks = 0;                      // column chunks of the ELM solve pipeline (0: default)
  bool elm_static_soil = false, elm_soil_loaded = false;   // soil rows of z / dz / zi go up once (mppgpu_elm_set_pipeline)
};

}  // namespace mpp
