// thermal_snow_kernels.cuh -- ELM's real thermal column in one launch: snow (<= nlevsno layers, variable active count) +
// standing surface water (one cell) + soil, three governing equations coupled through COND_DIRICHLET_FRM_OTR_GOVEQ
// conditions into ONE linear system per column and step.
//
// Reference path (configuration built by src/driver/alm/MPPThermalTBasedALM_Initialize.F90:150-813):
//   ThermalSOEPreSolve / ComputeRHS / ComputeOperators / GovEqnExchangeAuxVars  src/mpp/soe/SystemOfEquationsThermalType.F90:412-915
//   snow   aux vars  src/mpp/auxvar/ThermalKSPTemperatureSnowAuxType.F90:58-84    equation  src/mpp/ge/GoveqnThermalKSPTemperatureSnowType.F90:232-1300
//   ssw    aux vars  src/mpp/auxvar/ThermalKSPTemperatureSSWAuxType.F90:45-66     equation  src/mpp/ge/GoveqnThermalKSPTemperatureSSWType.F90:234-1130
//   soil   coupling branches  src/mpp/ge/GoveqnThermalKSPTemperatureSoilType.F90:820-905, 1150-1190, 1232-1400
// Unknown / mailbox ordering = the reference's SoE vector: [snow cells of all columns | ssw cells | soil cells].
//
// Mapping: one warp per column, one matrix row per lane -- lanes [0, nsno) the snow layers (top to bottom), lanes
// [nsno, nsno+nlev) the soil layers, lane nsno+nlev the standing-water cell.  The column graph is a chain
// (snow - soil) with the standing-water cell hanging off the top soil row: that leaf is folded into the top soil row by
// one Schur step over shuffles, the chain goes through normalised parallel cyclic reduction in registers, and the leaf
// is back-substituted.  Every mailbox array is read once, coalesced per segment; nothing is staged in shared memory.
// The snow -> soil and soil -> snow links are NOT symmetric in the reference (the soil side weights by the snow-cover
// fraction and uses the harmonic conductivity, the snow side uses its own conductivity over its own half thickness),
// so the rows carry separate sub- and super-diagonals.
#pragma once
#include "thermal_kernels.cuh"

namespace mpp {

struct ThermalSnowArgs {
  ThermalArgs S;                        // soil statics (ncol, nlev, nlevsoi, landunit ids, dt, cnfac, tables, distances, stale_area)
  int nsno;
  // SoE mailbox, ncol*(nsno+1+nlev) entries each, SoE order
  const double *T_in, *liq, *ice, *snow_water, *mdz, *dist_up, *dist_dn, *tuning, *frac;
  const int *nsnow, *active;
  const double *hs[3], *dhsdT[3];       // heat-flux conditions: 0 top of snow, 1 standing water, 2 top of soil
  const double *frac_soil;              // VAR_FRAC of the soil heat-flux condition
  const double *sabg_snow, *sabg_soil;  // COND_HEAT_RATE on ALL_CELLS of the snow / soil equation
  const double *soil_top_dist_dn;       // z(c,1) - zi(c,0): dist_dn of the soil's two coupling conditions (Initialize.F90:630-639)
  int *snow_top_id;                     // persistent: cell the snow heat-flux condition points at (UpdateBoundaryConn keeps it while no snow)
  double *T_out;
};

constexpr double THIN_SFCLAYER = 1.0e-6;   // ThermalKSPTemperature{Snow,SSW}AuxType.F90
constexpr double TKAIR = 0.023;            // mpp_varcon.F90:20

__global__ void __launch_bounds__(TH_TILE, 6)
thermal_snow_step_kernel(const ThermalSnowArgs A)
{
  constexpr unsigned FULL = 0xffffffffu;
  const int lane = threadIdx.x & 31;
  const int col = blockIdx.x * (TH_TILE / 32) + (threadIdx.x >> 5);
  if (col >= A.S.ncol) return;                       // whole warp
  const int nsno = A.nsno, nlev = A.S.nlev, ncol = A.S.ncol;
  const int l_so = nsno, l_sw = nsno + nlev, l_bot = nsno - 1;
  const bool is_snow = lane < nsno, is_soil = lane >= l_so && lane < l_sw, is_ssw = lane == l_sw;
  const bool valid = lane <= l_sw;
  const int j = lane - nsno;                          // soil layer
  const double dt = A.S.dt, cnfac = A.S.cnfac, area = A.S.area[col];
  const long long scell = (long long)col * nlev + (is_soil ? j : 0);            // index into the soil-sized static tables
  const long long idx = is_snow ? (long long)col * nsno + lane
                      : is_soil ? (long long)ncol * (nsno + 1) + (long long)col * nlev + j
                                : (long long)ncol * nsno + col;

  // ---- PreSolve: GetFromSOEAuxVarsIntrn of the three equations ----
  double T = 0.0, mdz = 0.0, frac = 0.0, tf = 1.0, adu = 0.0, add_ = 0.0, liq = 0.0, ice = 0.0, snoww = 0.0, src = 0.0;
  int act = 0, nsn = 0;
  if (valid) {
    T = A.T_in[idx]; mdz = A.mdz[idx]; frac = A.frac[idx]; act = A.active[idx];
    if (!is_ssw) {
      tf = A.tuning[idx]; liq = A.liq[idx]; ice = A.ice[idx]; nsn = A.nsnow[idx];
      if (is_snow) { adu = A.dist_up[idx]; add_ = A.dist_dn[idx]; src = A.sabg_snow[(long long)col * nsno + lane]; }
      else { snoww = A.snow_water[idx]; src = A.sabg_soil[scell]; }
    }
  }
  // ---- aux vars: conductivity tk, capacity term cap = heat_cap * vol / (dt * tuning) ----
  double tk = 1.0, cap = 0.0, cdu = 0.5, cdd = 0.5;   // cdu / cdd: distances of the connection lane -> lane+1
  double sw_dzm = THIN_SFCLAYER;
  const double sdz = is_soil ? A.S.dz[scell] : 0.0;  // static soil mesh thickness
  if (is_snow) {
    if (act) {                                        // ThermKSPTempSnowAuxVarCompute; mesh dz = VAR_DZ of an active cell (:268-270)
      const double bw = (ice + liq) * rcp(frac * mdz);
      tk = TKAIR + (7.75e-5 * bw + 1.105e-6 * bw * bw) * (TKICE - TKAIR);
      double hc = THIN_SFCLAYER;
      if (frac > 0.0) { hc = (CPLIQ * liq + CPICE * ice) * rcp(frac); hc = (hc > THIN_SFCLAYER) ? hc : THIN_SFCLAYER; }
      hc = hc * rcp(mdz);
      cap = hc * (area * mdz) * rcp(dt * tf);
    }
    cdu = adu;                                        // UpdateInternalConn: SetDistUp(aux(up)%dist_up), SetDistDn(aux(dn)%dist_dn)
  } else if (is_ssw) {
    if (act) {                                        // SSW UpdateInternalConn (:554-589) + ThermKSPTempSSWAuxVarCompute
      sw_dzm = (mdz * frac * 1.0e3 > THIN_SFCLAYER && frac > THIN_SFCLAYER) ? ((mdz > THIN_SFCLAYER) ? mdz : THIN_SFCLAYER) : THIN_SFCLAYER;
      tk = TKWAT;
      double hc = THIN_SFCLAYER;
      if (sw_dzm * frac * 1.0e3 > THIN_SFCLAYER && frac > THIN_SFCLAYER) { hc = CPLIQ * DENH2O; hc = (hc > THIN_SFCLAYER) ? hc : THIN_SFCLAYER; }
      cap = hc * (area * sw_dzm) * rcp(dt);
    }
  } else if (is_soil) {
    double hc;
    thermal_auxvar(A.S, A.S.lun_type[col], j < A.S.nlevsoi, T, liq, ice, snoww, nsn, A.S.por[scell], A.S.tkmg[scell], A.S.tkdry[scell],
                   A.S.csol[scell], sdz, tk, hc);
    if (act) cap = hc * (area * sdz) * rcp(dt * tf);
    if (A.S.dist_uniform) { cdu = A.S.lay_du[j]; cdd = A.S.lay_dd[j]; }
    else if (A.S.dist_up) { cdu = A.S.dist_up[scell]; cdd = A.S.dist_dn[scell]; }
    else cdu = 0.5 * sdz;
  }
  // neighbour (lane + 1) state
  const double T_d = __shfl_down_sync(FULL, T, 1), tk_d = __shfl_down_sync(FULL, tk, 1), add_d = __shfl_down_sync(FULL, add_, 1);
  const int act_d = __shfl_down_sync(FULL, act, 1);
  const double sdz_d = __shfl_down_sync(FULL, sdz, 1);
  if (is_snow) cdd = add_d;
  if (is_soil && !A.S.dist_uniform && !A.S.dist_up) cdd = 0.5 * sdz_d;
  // ---- rows: accumulation ----
  double bb, rhs, aa = 0.0, cc = 0.0;
  if (act) { bb = cap; rhs = cap * T; } else { bb = 1.0; rhs = 0.0; }
  // ---- internal connections lane -> lane+1 within the same equation (snow GE :817-858 / :1046-1083, soil GE as thermal_step_kernel) ----
  const bool same_eq = (is_snow && lane + 1 < nsno) || (is_soil && lane + 1 < l_sw);
  double cval = 0.0, flux = 0.0;
  if (same_eq && act && act_d) {
    const double kod = tk * tk_d * rcp(tk * cdd + tk_d * cdu) * area;
    flux = -kod * (T - T_d);
    cval = (1.0 - cnfac) * kod;
  }
  const double cval_m = __shfl_up_sync(FULL, cval, 1), flux_m = __shfl_up_sync(FULL, flux, 1);
  rhs = rhs + cnfac * flux; bb += cval; cc = -cval;
  if (lane > 0) { rhs = rhs - cnfac * flux_m; bb += cval_m; aa = -cval_m; }
  // ---- snow: heat flux at the top ACTIVE layer, coupling with the soil at the bottom layer ----
  const double T_m = __shfl_up_sync(FULL, T, 1), tk_m = __shfl_up_sync(FULL, tk, 1), frac_m = __shfl_up_sync(FULL, frac, 1);
  const double adu_m = __shfl_up_sync(FULL, adu, 1);
  const int act_m = __shfl_up_sync(FULL, act, 1);
  if (nsno > 0) {
    const int bot_act = __shfl_sync(FULL, act, l_bot), bot_nsn = __shfl_sync(FULL, nsn, l_bot);
    int top;
    if (bot_act) { top = nsno - bot_nsn; if (lane == 0) A.snow_top_id[col] = top; }     // UpdateBoundaryConn :680-686
    else top = A.snow_top_id[col];
    if (is_snow && lane == top && act) {                                               // COND_HEAT_FLUX (:896-909, :1140-1155)
      const double H = A.hs[0][col], dH = A.dhsdT[0][col];
      rhs = rhs + (H - dH * T) * area;
      bb += -dH * area;
    }
    if (lane == l_bot && act) {
      // boundary aux var = the soil's top cell (T, conductivity); conn dist_up = 0, dist_dn = this cell's dist_up (:688-694)
      const double dd = adu;
      const double kod = tk_d * tk * rcp(tk_d * dd) * area;
      const double fl = -kod * (T_d - T);
      rhs = rhs - cnfac * fl;
      const double v = (1.0 - cnfac) * kod;
      bb += v; cc = -v;
    }
  }
  // ---- standing water <-> top soil cell ----
  const double T_w = __shfl_sync(FULL, T, l_sw), tk_w = __shfl_sync(FULL, tk, l_sw), frac_w = __shfl_sync(FULL, frac, l_sw);
  const double mdz_w = __shfl_sync(FULL, mdz, l_sw);
  const int act_w = __shfl_sync(FULL, act, l_sw);
  const double T_s1 = __shfl_sync(FULL, T, l_so), tk_s1 = __shfl_sync(FULL, tk, l_so);
  double cs = 0.0;                                    // ssw row: coefficient of the top soil unknown
  if (is_ssw && act) {
    const double H = A.hs[1][col], dH = A.dhsdT[1][col];
    rhs = rhs + (H - dH * T) * area; bb += -dH * area;                                  // COND_HEAT_FLUX (:1086-1097, :790-803)
    // coupling: conn dist_up = 0, conn dist_dn = mesh dz / 2 (-> dist), conductivity averaged with aux dz / 2 (:1052-1075)
    const double dist = 0.5 * sw_dzm, dd = 0.5 * mdz;
    const double k = tk_s1 * tk * dd * rcp(tk_s1 * dd);
    const double kod = k * rcp(dist) * area;
    const double fl = -kod * (T_s1 - T);
    rhs = rhs - cnfac * fl;
    const double v = (1.0 - cnfac) * kod;
    bb += v; cs = -v;
  }
  double a1s = 0.0;                                   // top soil row: coefficient of the ssw unknown
  if (lane == l_so && act) {
    const double dd_conn = A.soil_top_dist_dn[col];
    {                                                 // COND_HEAT_FLUX at the top of the soil (soil GE :886-903, :1196-1215)
      const double H = A.hs[2][col], dH = A.dhsdT[2][col], fr = A.frac_soil[col];
      rhs = rhs + (H - dH * T) * fr * area;
      bb += -fr * ((area == 1.0) ? dH : pow(dH, area));
    }
    if (nsno > 0 && act_m) {                          // coupling with the bottom snow layer (lane - 1): conn dist_up = its dist_up
      const double du = adu_m;
      const double kod = tk_m * tk * rcp(tk_m * dd_conn + tk * du);
      const double fl = -kod * (T_m - T);
      rhs = rhs - frac_m * cnfac * fl * A.S.stale_area;               // `area` is stale in this branch of the reference (:843-876)
      const double v = frac_m * (1.0 - cnfac) * kod * area;
      bb += v; aa = -v;
    }
    if (act_w) {                                      // coupling with the standing water (is_bc_sh2o branches)
      const double du = 0.5 * mdz_w, dd = 0.5 * mdz;  // Divergence uses aux dz / 2 of the soil cell, the operators the connection's dist_dn
      const double half = ((du * 2.0 > 1.0e-6) ? du * 2.0 : 1.0e-6) * 0.5;
      const double k1 = tk_w * tk * (du + dd) * rcp(tk_w * dd + tk * du);
      const double fl = -k1 * (T_w - T) * rcp(dd + half);
      rhs = rhs - frac_w * cnfac * fl * A.S.stale_area;
      const double k2 = tk_w * tk * (du + dd_conn) * rcp(tk_w * dd_conn + tk * du);
      const double v = frac_w * (1.0 - cnfac) * k2 * rcp(dd_conn + half) * area;
      bb += v; a1s = -v;
    }
  }
  if (act && !is_ssw) rhs = rhs + src;                // COND_HEAT_RATE on ALL_CELLS (absorbed solar radiation)
  // ---- fold the standing-water leaf into the top soil row ----
  const double bs = __shfl_sync(FULL, bb, l_sw), ds = __shfl_sync(FULL, rhs, l_sw), cs_b = __shfl_sync(FULL, cs, l_sw);
  if (lane == l_so) {
    const double f = a1s * rcp(bs);
    bb -= f * cs_b; rhs -= f * ds;
  }
  // ---- chain solve (the ssw lane and the padding lanes are identity rows) ----
  const bool chain = lane < l_sw;
  const double x = thermal_pcr<32>(chain ? aa : 0.0, chain ? bb : 1.0, chain ? cc : 0.0, chain ? rhs : 0.0);
  const double x1 = __shfl_sync(FULL, x, l_so);
  if (chain) A.T_out[idx] = x;
  else if (is_ssw) A.T_out[idx] = (rhs - cs * x1) * rcp(bb);
}

}  // namespace mpp
