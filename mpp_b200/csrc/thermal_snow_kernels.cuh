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
// Mapping (thermal_snow_step_kernel; thermal_snow_step3_kernel below holds three rows per lane on 8 lanes and is the one ELM's
// 5 + 15 layout runs on): 16 lanes per column (two columns per warp), two matrix rows per lane -- rows [0, nsno) the snow layers (top to
// bottom), rows [nsno, nsno+nlev) the soil layers.  The column graph is a chain (snow - soil) with the standing-water cell
// hanging off the top soil row: that leaf is computed by the lane that owns the top soil row and folded into it by one
// Schur step, the chain's odd rows are eliminated in-lane and its even rows go through normalised parallel cyclic
// reduction in registers (4 stages), then the odd rows and the leaf are back-substituted.  Every mailbox array is read
// once, coalesced per segment; nothing is staged in shared memory.
// The snow -> soil and soil -> snow links are NOT symmetric in the reference (the soil side weights by the snow-cover
// fraction and uses the harmonic conductivity, the snow side uses its own conductivity over its own half thickness),
// so the rows carry separate sub- and super-diagonals.
#pragma once
#include "thermal_kernels.cuh"

namespace mpp {

struct ThermalSnowArgs {
  ThermalArgs S;                        // soil statics (ncol, nlev, nlevsoi, landunit ids, dt, cnfac, tables, distances, stale_area)
  int nsno;
  int col0, col_end;                    // the columns this launch covers (a chunk of the ELM solve pipeline, or the whole batch)
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

// One matrix row of the chain (snow layers top to bottom, then soil layers): the state a neighbouring row needs travels by
// shuffle (T, tk, act, du, x2, frac); everything else stays in the lane.
struct SnowRow {
  double T, tk, cap, du, x2, frac, mdz, src, bb, rhs, sub, sup;   // du: snow aux dist_up | soil connection dist_up; x2: snow aux dist_dn | soil connection dist_dn
  int act, nsn;
};

__device__ __forceinline__ void snow_row_clear(SnowRow &r)
{
  r.T = 0.0; r.tk = 1.0; r.cap = 0.0; r.du = 0.5; r.x2 = 0.5; r.frac = 0.0; r.mdz = 0.0; r.src = 0.0; r.act = 0; r.nsn = 0;
  r.sub = 0.0; r.sup = 0.0; r.bb = 1.0; r.rhs = 0.0;
}

// PreSolve + aux vars of snow layer s (0 = topmost of the nsno slots) of column `col`
__device__ __forceinline__ void snow_layer_load(const ThermalSnowArgs &A, SnowRow &r, int s, int col, double area, double &liq, double &ice, double &tf)
{
  const long long idx = (long long)col * A.nsno + s;
  r.T = A.T_in[idx]; r.mdz = A.mdz[idx]; r.frac = A.frac[idx]; r.act = A.active[idx]; r.nsn = A.nsnow[idx];
  tf = A.tuning[idx]; liq = A.liq[idx]; ice = A.ice[idx];
  r.du = A.dist_up[idx]; r.x2 = A.dist_dn[idx]; r.src = A.sabg_snow[idx];
}
__device__ __forceinline__ void snow_layer_aux(const ThermalSnowArgs &A, SnowRow &r, double area, double liq, double ice, double tf)
{
  if (r.act) {                                        // ThermKSPTempSnowAuxVarCompute; mesh dz = VAR_DZ of an active cell (:268-270)
    const double bw = (ice + liq) * rcp(r.frac * r.mdz);
    r.tk = TKAIR + (7.75e-5 * bw + 1.105e-6 * bw * bw) * (TKICE - TKAIR);
    double hc = THIN_SFCLAYER;
    if (r.frac > 0.0) { hc = (CPLIQ * liq + CPICE * ice) * rcp(r.frac); hc = (hc > THIN_SFCLAYER) ? hc : THIN_SFCLAYER; }
    hc = hc * rcp(r.mdz);
    r.cap = hc * (area * r.mdz) * rcp(A.S.dt * tf);
    r.bb = r.cap; r.rhs = r.cap * r.T;
  }
}

struct SoilIn { double liq, ice, snoww, tf, por, tkmg, tkdry, csol, sdz; };
__device__ __forceinline__ void soil_layer_load(const ThermalSnowArgs &A, SnowRow &r, SoilIn &in, int j, int col)
{
  const int nlev = A.S.nlev;
  const long long scell = (long long)col * nlev + j, idx = (long long)A.S.ncol * (A.nsno + 1) + scell;
  r.T = A.T_in[idx]; r.mdz = A.mdz[idx]; r.frac = A.frac[idx]; r.act = A.active[idx]; r.nsn = A.nsnow[idx];
  in.tf = A.tuning[idx]; in.liq = A.liq[idx]; in.ice = A.ice[idx]; in.snoww = A.snow_water[idx];
  r.src = A.sabg_soil[scell];
  in.sdz = A.S.dz[scell]; in.por = A.S.por[scell]; in.tkmg = A.S.tkmg[scell]; in.tkdry = A.S.tkdry[scell]; in.csol = A.S.csol[scell];
  if (A.S.dist_uniform) { r.du = A.S.lay_du[j]; r.x2 = A.S.lay_dd[j]; }
  else if (A.S.dist_up) { r.du = A.S.dist_up[scell]; r.x2 = A.S.dist_dn[scell]; }
  else { r.du = 0.5 * in.sdz; r.x2 = (j + 1 < nlev) ? 0.5 * A.S.dz[scell + 1] : 0.5; }
}
__device__ __forceinline__ void soil_layer_aux(const ThermalSnowArgs &A, SnowRow &r, const SoilIn &in, int j, int lun, double area)
{
  double hc;
  thermal_auxvar(A.S, lun, j < A.S.nlevsoi, r.T, in.liq, in.ice, in.snoww, r.nsn, in.por, in.tkmg, in.tkdry, in.csol, in.sdz, r.tk, hc);
  if (r.act) { r.cap = hc * (area * in.sdz) * rcp(A.S.dt * in.tf); r.bb = r.cap; r.rhs = r.cap * r.T; }
}

// Contributions of the connection between an upper chain row U and the row D below it to both rows:
//   kind 1 / 2  snow-snow / soil-soil   snow GE :817-858 / :1046-1083, soil GE as thermal_step_kernel (symmetric)
//   kind 3      snow bottom | soil top  the two COND_DIRICHLET_FRM_OTR_GOVEQ conditions (snow GE :861-893, :1096-1137, :1202-1300;
//                                       soil GE :843-876, :1150-1190, :1232-1340) -- not symmetric
__device__ __forceinline__ void snow_connect(const ThermalSnowArgs &A, int kind, double area, double dd_top,
                                             double UT, double Utk, int Uact, double Udu, double Ux2, double Ufrac,
                                             double DT, double Dtk, int Dact, double Dx2,
                                             double &U_rhs, double &U_bb, double &U_sup, double &D_rhs, double &D_bb, double &D_sub)
{
  const double cnfac = A.S.cnfac;
  U_rhs = 0.0; U_bb = 0.0; U_sup = 0.0; D_rhs = 0.0; D_bb = 0.0; D_sub = 0.0;
  if (kind == 3) {                                    // snow bottom (U) | soil top (D)
    if (Uact) {
      // snow side: boundary aux var = the soil's top cell; conn dist_up = 0, dist_dn = the snow cell's dist_up (:688-694)
      const double kod = Dtk * Utk * rcp(Dtk * Udu) * area;
      const double fl = -kod * (DT - UT);
      U_rhs = -cnfac * fl;
      U_bb = (1.0 - cnfac) * kod; U_sup = -U_bb;
      if (Dact) {
        // soil side: conn dist_up = the snow cell's dist_up, dist_dn = z(c,1) - zi(c,0); weighted by the snow-cover fraction;
        // `area` of the flux term is stale in this branch of the reference (:843-876)
        const double kos = Utk * Dtk * rcp(Utk * dd_top + Dtk * Udu);
        const double fs = -kos * (UT - DT);
        D_rhs = -Ufrac * cnfac * fs * A.S.stale_area;
        D_bb = Ufrac * (1.0 - cnfac) * kos * area; D_sub = -D_bb;
      }
    }
  } else if (kind != 0 && Uact && Dact) {
    const double dd = (kind == 1) ? Dx2 : Ux2;        // snow: SetDistDn(aux(dn)%dist_dn); soil: the connection's own dist_dn
    const double kod = Utk * Dtk * rcp(Utk * dd + Dtk * Udu) * area;
    const double fl = -kod * (UT - DT);
    const double cv = (1.0 - cnfac) * kod;
    U_rhs = cnfac * fl; U_bb = cv; U_sup = -cv;
    D_rhs = -cnfac * fl; D_bb = cv; D_sub = -cv;
  }
}

// LPC lanes per column, two chain rows per lane, 32/LPC columns per warp.  A lane is either a snow lane or a soil lane (the snow
// block is padded at the TOP to an even number of rows), so the two rows of a lane share one code path and their loads are
// issued together; the bottom snow layer is always the second row of the last snow lane and the top soil layer the first row
// of the first soil lane, which also owns the standing-water cell.
#ifndef SNOW_MIN_BLOCKS
#define SNOW_MIN_BLOCKS 4
#endif
template <int LPC>
__global__ void __launch_bounds__(TH_TILE, SNOW_MIN_BLOCKS)
thermal_snow_step_kernel(const ThermalSnowArgs A)
{
  constexpr unsigned FULL = 0xffffffffu;
  const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int col = A.col0 + (int)(tid / LPC), l = (int)(tid % LPC);
  const int nsno = A.nsno, nlev = A.S.nlev, ncol = A.S.ncol;
  const int pad = nsno & 1, ns2 = (nsno + pad) >> 1;
  const bool col_ok = col < A.col_end;
  const bool snow_lane = l < ns2;
  const int sa = 2 * l - pad, sb = sa + 1;            // snow layers of a snow lane (sa = -1: the padding row)
  const int ja = 2 * (l - ns2), jb = ja + 1;          // soil layers of a soil lane
  const double cnfac = A.S.cnfac, dt = A.S.dt;
  const double area = col_ok ? A.S.area[col] : 1.0;
  const double dd_top = col_ok ? A.soil_top_dist_dn[col] : 1.0;

  // per-column scalars (conditions, standing-water cell, persistent top-of-snow id): issued up front with the row loads so the
  // kernel pays one DRAM latency, not one per dependent stage; the 16 lanes of a column read the same sectors
  const long long widx = (long long)ncol * nsno + col;
  double H0 = 0.0, dH0 = 0.0, H1 = 0.0, dH1 = 0.0, H2 = 0.0, dH2 = 0.0, frs = 0.0, wT = 0.0, wmdz = 0.0, wfrac = 0.0;
  int wact = 0, top_old = 0;
  if (col_ok) {
    H0 = A.hs[0][col]; dH0 = A.dhsdT[0][col]; H1 = A.hs[1][col]; dH1 = A.dhsdT[1][col]; H2 = A.hs[2][col]; dH2 = A.dhsdT[2][col];
    frs = A.frac_soil[col]; top_old = A.snow_top_id[col];
    wact = A.active[widx]; wT = A.T_in[widx]; wmdz = A.mdz[widx]; wfrac = A.frac[widx];
  }
  SnowRow a, b;
  snow_row_clear(a); snow_row_clear(b);
  bool va = false, vb = false;                        // the row exists
  if (col_ok) {
    if (snow_lane) {
      double liq_a = 0.0, ice_a = 0.0, tf_a = 1.0, liq_b, ice_b, tf_b;
      va = sa >= 0; vb = true;
      if (va) snow_layer_load(A, a, sa, col, area, liq_a, ice_a, tf_a);
      snow_layer_load(A, b, sb, col, area, liq_b, ice_b, tf_b);
      if (va) snow_layer_aux(A, a, area, liq_a, ice_a, tf_a);
      snow_layer_aux(A, b, area, liq_b, ice_b, tf_b);
    } else if (ja < nlev) {
      SoilIn ia, ib;
      const int lun = A.S.lun_type[col];
      va = true; vb = jb < nlev;
      soil_layer_load(A, a, ia, ja, col);
      if (vb) soil_layer_load(A, b, ib, jb, col);
      soil_layer_aux(A, a, ia, ja, lun, area);
      if (vb) soil_layer_aux(A, b, ib, jb, lun, area);
    }
  }

  // ---- connections: a|b in-lane, b|next lane's a ----
  double u0, u1, u2, d0, d1, d2;
  snow_connect(A, (va && vb) ? (snow_lane ? 1 : 2) : 0, area, dd_top, a.T, a.tk, a.act, a.du, a.x2, a.frac, b.T, b.tk, b.act, b.x2, u0, u1, u2, d0, d1, d2);
  a.rhs += u0; a.bb += u1; a.sup = u2; b.rhs += d0; b.bb += d1; b.sub = d2;
  {
    // the connection (this lane's b | next lane's a) is evaluated once, by the upper lane; the lower lane receives its share
    const double nT = __shfl_down_sync(FULL, a.T, 1, LPC), ntk = __shfl_down_sync(FULL, a.tk, 1, LPC), nx2 = __shfl_down_sync(FULL, a.x2, 1, LPC);
    const int nact = __shfl_down_sync(FULL, a.act, 1, LPC), nva = __shfl_down_sync(FULL, (int)va, 1, LPC);
    const int kn = (l + 1 < LPC && vb && nva) ? ((l + 1 < ns2) ? 1 : (l + 1 == ns2 ? 3 : 2)) : 0;
    snow_connect(A, kn, area, dd_top, b.T, b.tk, b.act, b.du, b.x2, b.frac, nT, ntk, nact, nx2, u0, u1, u2, d0, d1, d2);
    b.rhs += u0; b.bb += u1; b.sup = u2;
    const double p0 = __shfl_up_sync(FULL, d0, 1, LPC), p1 = __shfl_up_sync(FULL, d1, 1, LPC), p2 = __shfl_up_sync(FULL, d2, 1, LPC);
    if (l > 0) { a.rhs += p0; a.bb += p1; a.sub = p2; }
  }
  // ---- heat flux at the top ACTIVE snow layer (UpdateBoundaryConn :680-686; :896-909, :1140-1155) ----
  if (nsno > 0) {
    const int bot_act = __shfl_sync(FULL, b.act, ns2 - 1, LPC), bot_nsn = __shfl_sync(FULL, b.nsn, ns2 - 1, LPC);
    if (col_ok && snow_lane) {
      int top = top_old;
      if (bot_act) { top = nsno - bot_nsn; if (l == 0 && top != top_old) A.snow_top_id[col] = top; }
      if (va && sa == top && a.act) { a.rhs += (H0 - dH0 * a.T) * area; a.bb += -dH0 * area; }
      if (sb == top && b.act) { b.rhs += (H0 - dH0 * b.T) * area; b.bb += -dH0 * area; }
    }
  }
  // ---- top soil row (first row of the first soil lane): heat-flux condition and the standing-water leaf ----
  const bool top_lane = col_ok && l == ns2 && va;
  double w_bb = 1.0, w_rhs = 0.0, w_cs = 0.0;         // standing-water row: diagonal, right-hand side, coefficient of the top soil unknown
  if (top_lane) {
    double t_rhs = 0.0, t_bb = 0.0;
    if (a.act) {                                      // COND_HEAT_FLUX at the top of the soil (soil GE :886-903, :1196-1215)
      t_rhs = (H2 - dH2 * a.T) * frs * area;
      t_bb = -frs * ((area == 1.0) ? dH2 : mpp_pow_rare(dH2, area));
    }
    double a1s = 0.0;
    if (wact) {
      // SSW UpdateInternalConn (:554-589) + ThermKSPTempSSWAuxVarCompute
      const bool thick = wmdz * wfrac * 1.0e3 > THIN_SFCLAYER && wfrac > THIN_SFCLAYER;
      const double dzm = thick ? ((wmdz > THIN_SFCLAYER) ? wmdz : THIN_SFCLAYER) : THIN_SFCLAYER;
      double hc = THIN_SFCLAYER;
      if (dzm * wfrac * 1.0e3 > THIN_SFCLAYER && wfrac > THIN_SFCLAYER) { hc = CPLIQ * DENH2O; hc = (hc > THIN_SFCLAYER) ? hc : THIN_SFCLAYER; }
      const double cap = hc * (area * dzm) * rcp(dt);
      w_bb = cap - dH1 * area; w_rhs = cap * wT + (H1 - dH1 * wT) * area;                 // Accum + COND_HEAT_FLUX (:1086-1097, :790-803)
      {                                               // coupling, ssw side: conn dist_up = 0, dist = mesh dz / 2, averaging with aux dz / 2 (:1052-1075)
        const double dist = 0.5 * dzm, dd = 0.5 * wmdz;
        const double k = a.tk * TKWAT * dd * rcp(a.tk * dd);
        const double kod = k * rcp(dist) * area;
        const double fl = -kod * (a.T - wT);
        w_rhs = w_rhs - cnfac * fl;
        const double v = (1.0 - cnfac) * kod;
        w_bb += v; w_cs = -v;
      }
      if (a.act) {                                    // coupling, soil side (is_bc_sh2o branches): Divergence averages with the soil cell's
        const double du = 0.5 * wmdz, dd = 0.5 * a.mdz; // aux dz / 2, the operators with the connection's dist_dn
        const double half = ((du * 2.0 > 1.0e-6) ? du * 2.0 : 1.0e-6) * 0.5;
        const double k1 = TKWAT * a.tk * (du + dd) * rcp(TKWAT * dd + a.tk * du);
        const double fl = -k1 * (wT - a.T) * rcp(dd + half);
        t_rhs -= wfrac * cnfac * fl * A.S.stale_area;
        const double k2 = TKWAT * a.tk * (du + dd_top) * rcp(TKWAT * dd_top + a.tk * du);
        const double v = wfrac * (1.0 - cnfac) * k2 * rcp(dd_top + half) * area;
        t_bb += v; a1s = -v;
      }
    }
    // fold the leaf into the top soil row (Schur step; identity when there is no standing water)
    const double f = a1s * rcp(w_bb);
    t_bb -= f * w_cs; t_rhs -= f * w_rhs;
    a.rhs += t_rhs; a.bb += t_bb;
  }
  if (a.act) a.rhs += a.src;                          // COND_HEAT_RATE on ALL_CELLS (absorbed solar radiation)
  if (b.act) b.rhs += b.src;

  // ---- chain solve: second rows eliminated in-lane, first rows by normalised PCR over LPC lanes, second rows back-substituted ----
  const double rbb = rcp(b.bb);
  const double bs = b.sub * rbb, bu = b.sup * rbb, bf = b.rhs * rbb;         // x_b = bf - bs x_a(l) - bu x_a(l+1)
  const double bs_p = __shfl_up_sync(FULL, bs, 1, LPC), bu_p = __shfl_up_sync(FULL, bu, 1, LPC), bf_p = __shfl_up_sync(FULL, bf, 1, LPC);
  const double sub_a = (l > 0) ? a.sub : 0.0;
  const double rB = rcp(a.bb - sub_a * bu_p - a.sup * bs);
  const double al = (-sub_a * bs_p) * rB, ga = (-a.sup * bu) * rB, de = (a.rhs - sub_a * bf_p - a.sup * bf) * rB;
  const double xa = thermal_pcr_unit<LPC>(al, ga, de);
  const double xa_n = __shfl_down_sync(FULL, xa, 1, LPC);
  const double xb = bf - bs * xa - ((l + 1 < LPC) ? bu * xa_n : 0.0);
  if (col_ok) {
    if (snow_lane) {
      if (va) A.T_out[(long long)col * nsno + sa] = xa;
      A.T_out[(long long)col * nsno + sb] = xb;
    } else {
      const long long base = (long long)ncol * (nsno + 1) + (long long)col * nlev;
      if (va) A.T_out[base + ja] = xa;
      if (vb) A.T_out[base + jb] = xb;
    }
    if (top_lane) A.T_out[widx] = (w_rhs - w_cs * xa) * rcp(w_bb);
  }
}

// Three chain rows per lane (a, m, b), LPC lanes per column, 32/LPC columns per warp: ELM's 5 snow + 15 soil layers fill 7 of 8 lanes
// (snow block padded at the top to a multiple of three), four columns per warp.  The middle row of every lane is eliminated
// in-lane first; what remains is exactly the two-rows-per-lane system of thermal_snow_step_kernel (second rows eliminated in-lane,
// first rows by a 3-stage normalised PCR), then the middle rows are back-substituted.
#ifndef SNOW3_MIN_BLOCKS
#define SNOW3_MIN_BLOCKS 4
#endif
template <int LPC>
__global__ void __launch_bounds__(TH_TILE, SNOW3_MIN_BLOCKS)
thermal_snow_step3_kernel(const ThermalSnowArgs A)
{
  constexpr unsigned FULL = 0xffffffffu;
  const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int col = A.col0 + (int)(tid / LPC), l = (int)(tid % LPC);
  const int nsno = A.nsno, nlev = A.S.nlev, ncol = A.S.ncol;
  const int ns3 = (nsno + 2) / 3, pad = 3 * ns3 - nsno;
  const bool col_ok = col < A.col_end;
  const bool snow_lane = l < ns3;
  const int s0 = 3 * l - pad;                          // snow layers s0, s0+1, s0+2 of a snow lane (negative: padding rows)
  const int j0 = 3 * (l - ns3);                        // soil layers j0, j0+1, j0+2 of a soil lane
  const double cnfac = A.S.cnfac, dt = A.S.dt;
  const double area = col_ok ? A.S.area[col] : 1.0;
  const double dd_top = col_ok ? A.soil_top_dist_dn[col] : 1.0;

  const long long widx = (long long)ncol * nsno + col;
  double H0 = 0.0, dH0 = 0.0, H1 = 0.0, dH1 = 0.0, H2 = 0.0, dH2 = 0.0, frs = 0.0, wT = 0.0, wmdz = 0.0, wfrac = 0.0;
  int wact = 0, top_old = 0;
  if (col_ok) {
    H0 = A.hs[0][col]; dH0 = A.dhsdT[0][col]; H1 = A.hs[1][col]; dH1 = A.dhsdT[1][col]; H2 = A.hs[2][col]; dH2 = A.dhsdT[2][col];
    frs = A.frac_soil[col]; top_old = A.snow_top_id[col];
    wact = A.active[widx]; wT = A.T_in[widx]; wmdz = A.mdz[widx]; wfrac = A.frac[widx];
  }
  SnowRow a, m, b;
  snow_row_clear(a); snow_row_clear(m); snow_row_clear(b);
  bool va = false, vm = false, vb = false;             // the row exists
  if (col_ok) {
    if (snow_lane) {
      double la = 0.0, ia = 0.0, ta = 1.0, lm = 0.0, im = 0.0, tm = 1.0, lb, ib, tb;
      va = s0 >= 0; vm = s0 + 1 >= 0; vb = true;
      if (va) snow_layer_load(A, a, s0, col, area, la, ia, ta);
      if (vm) snow_layer_load(A, m, s0 + 1, col, area, lm, im, tm);
      snow_layer_load(A, b, s0 + 2, col, area, lb, ib, tb);
      if (va) snow_layer_aux(A, a, area, la, ia, ta);
      if (vm) snow_layer_aux(A, m, area, lm, im, tm);
      snow_layer_aux(A, b, area, lb, ib, tb);
    } else if (j0 < nlev) {
      SoilIn ia, im, ib;
      const int lun = A.S.lun_type[col];
      va = true; vm = j0 + 1 < nlev; vb = j0 + 2 < nlev;
      soil_layer_load(A, a, ia, j0, col);
      if (vm) soil_layer_load(A, m, im, j0 + 1, col);
      if (vb) soil_layer_load(A, b, ib, j0 + 2, col);
      soil_layer_aux(A, a, ia, j0, lun, area);
      if (vm) soil_layer_aux(A, m, im, j0 + 1, lun, area);
      if (vb) soil_layer_aux(A, b, ib, j0 + 2, lun, area);
    }
  }

  // ---- connections: a|m and m|b in-lane, b|next lane's a ----
  const int kin = snow_lane ? 1 : 2;
  double u0, u1, u2, d0, d1, d2;
  snow_connect(A, (va && vm) ? kin : 0, area, dd_top, a.T, a.tk, a.act, a.du, a.x2, a.frac, m.T, m.tk, m.act, m.x2, u0, u1, u2, d0, d1, d2);
  a.rhs += u0; a.bb += u1; a.sup = u2; m.rhs += d0; m.bb += d1; m.sub = d2;
  snow_connect(A, (vm && vb) ? kin : 0, area, dd_top, m.T, m.tk, m.act, m.du, m.x2, m.frac, b.T, b.tk, b.act, b.x2, u0, u1, u2, d0, d1, d2);
  m.rhs += u0; m.bb += u1; m.sup = u2; b.rhs += d0; b.bb += d1; b.sub = d2;
  {
    const double nT = __shfl_down_sync(FULL, a.T, 1, LPC), ntk = __shfl_down_sync(FULL, a.tk, 1, LPC), nx2 = __shfl_down_sync(FULL, a.x2, 1, LPC);
    const int nact = __shfl_down_sync(FULL, a.act, 1, LPC), nva = __shfl_down_sync(FULL, (int)va, 1, LPC);
    const int kn = (l + 1 < LPC && vb && nva) ? ((l + 1 < ns3) ? 1 : (l + 1 == ns3 ? 3 : 2)) : 0;
    snow_connect(A, kn, area, dd_top, b.T, b.tk, b.act, b.du, b.x2, b.frac, nT, ntk, nact, nx2, u0, u1, u2, d0, d1, d2);
    b.rhs += u0; b.bb += u1; b.sup = u2;
    const double p0 = __shfl_up_sync(FULL, d0, 1, LPC), p1 = __shfl_up_sync(FULL, d1, 1, LPC), p2 = __shfl_up_sync(FULL, d2, 1, LPC);
    if (l > 0) { a.rhs += p0; a.bb += p1; a.sub = p2; }
  }
  // ---- heat flux at the top ACTIVE snow layer (UpdateBoundaryConn :680-686; :896-909, :1140-1155) ----
  if (nsno > 0) {
    const int bot_act = __shfl_sync(FULL, b.act, ns3 - 1, LPC), bot_nsn = __shfl_sync(FULL, b.nsn, ns3 - 1, LPC);
    if (col_ok && snow_lane) {
      int top = top_old;
      if (bot_act) { top = nsno - bot_nsn; if (l == 0 && top != top_old) A.snow_top_id[col] = top; }
      if (va && s0 == top && a.act) { a.rhs += (H0 - dH0 * a.T) * area; a.bb += -dH0 * area; }
      if (vm && s0 + 1 == top && m.act) { m.rhs += (H0 - dH0 * m.T) * area; m.bb += -dH0 * area; }
      if (s0 + 2 == top && b.act) { b.rhs += (H0 - dH0 * b.T) * area; b.bb += -dH0 * area; }
    }
  }
  // ---- top soil row (first row of the first soil lane): heat-flux condition and the standing-water leaf ----
  const bool top_lane = col_ok && l == ns3 && va;
  double w_bb = 1.0, w_rhs = 0.0, w_cs = 0.0;
  if (top_lane) {
    double t_rhs = 0.0, t_bb = 0.0;
    if (a.act) {
      t_rhs = (H2 - dH2 * a.T) * frs * area;
      t_bb = -frs * ((area == 1.0) ? dH2 : mpp_pow_rare(dH2, area));
    }
    double a1s = 0.0;
    if (wact) {
      const bool thick = wmdz * wfrac * 1.0e3 > THIN_SFCLAYER && wfrac > THIN_SFCLAYER;
      const double dzm = thick ? ((wmdz > THIN_SFCLAYER) ? wmdz : THIN_SFCLAYER) : THIN_SFCLAYER;
      double hc = THIN_SFCLAYER;
      if (dzm * wfrac * 1.0e3 > THIN_SFCLAYER && wfrac > THIN_SFCLAYER) { hc = CPLIQ * DENH2O; hc = (hc > THIN_SFCLAYER) ? hc : THIN_SFCLAYER; }
      const double cap = hc * (area * dzm) * rcp(dt);
      w_bb = cap - dH1 * area; w_rhs = cap * wT + (H1 - dH1 * wT) * area;
      {
        const double dist = 0.5 * dzm, dd = 0.5 * wmdz;
        const double k = a.tk * TKWAT * dd * rcp(a.tk * dd);
        const double kod = k * rcp(dist) * area;
        const double fl = -kod * (a.T - wT);
        w_rhs = w_rhs - cnfac * fl;
        const double v = (1.0 - cnfac) * kod;
        w_bb += v; w_cs = -v;
      }
      if (a.act) {
        const double du = 0.5 * wmdz, dd = 0.5 * a.mdz;
        const double half = ((du * 2.0 > 1.0e-6) ? du * 2.0 : 1.0e-6) * 0.5;
        const double k1 = TKWAT * a.tk * (du + dd) * rcp(TKWAT * dd + a.tk * du);
        const double fl = -k1 * (wT - a.T) * rcp(dd + half);
        t_rhs -= wfrac * cnfac * fl * A.S.stale_area;
        const double k2 = TKWAT * a.tk * (du + dd_top) * rcp(TKWAT * dd_top + a.tk * du);
        const double v = wfrac * (1.0 - cnfac) * k2 * rcp(dd_top + half) * area;
        t_bb += v; a1s = -v;
      }
    }
    const double f = a1s * rcp(w_bb);
    t_bb -= f * w_cs; t_rhs -= f * w_rhs;
    a.rhs += t_rhs; a.bb += t_bb;
  }
  if (a.act) a.rhs += a.src;
  if (m.act) m.rhs += m.src;
  if (b.act) b.rhs += b.src;

  // ---- eliminate the middle row: x_m = mf - ms x_a - mu x_b ----
  const double rmb = rcp(m.bb);
  const double ms = m.sub * rmb, mu = m.sup * rmb, mf = m.rhs * rmb;
  const double a_bb = a.bb - a.sup * ms, a_sup = -a.sup * mu, a_rhs = a.rhs - a.sup * mf;
  const double b_sub = -b.sub * ms, b_bb = b.bb - b.sub * mu, b_rhs = b.rhs - b.sub * mf;
  // ---- two rows per lane (a', b'): second rows eliminated in-lane, first rows by normalised PCR, back-substitution ----
  const double rbb = rcp(b_bb);
  const double bs = b_sub * rbb, bu = b.sup * rbb, bf = b_rhs * rbb;          // x_b = bf - bs x_a(l) - bu x_a(l+1)
  const double bs_p = __shfl_up_sync(FULL, bs, 1, LPC), bu_p = __shfl_up_sync(FULL, bu, 1, LPC), bf_p = __shfl_up_sync(FULL, bf, 1, LPC);
  const double sub_a = (l > 0) ? a.sub : 0.0;
  const double rB = rcp(a_bb - sub_a * bu_p - a_sup * bs);
  const double al = (-sub_a * bs_p) * rB, ga = (-a_sup * bu) * rB, de = (a_rhs - sub_a * bf_p - a_sup * bf) * rB;
  const double xa = thermal_pcr_unit<LPC>(al, ga, de);
  const double xa_n = __shfl_down_sync(FULL, xa, 1, LPC);
  const double xb = bf - bs * xa - ((l + 1 < LPC) ? bu * xa_n : 0.0);
  const double xm = mf - ms * xa - mu * xb;
  if (col_ok) {
    if (snow_lane) {
      const long long base = (long long)col * nsno + s0;
      if (va) A.T_out[base] = xa;
      if (vm) A.T_out[base + 1] = xm;
      A.T_out[base + 2] = xb;
    } else {
      const long long base = (long long)ncol * (nsno + 1) + (long long)col * nlev + j0;
      if (va) A.T_out[base] = xa;
      if (vm) A.T_out[base + 1] = xm;
      if (vb) A.T_out[base + 2] = xb;
    }
    if (top_lane) A.T_out[widx] = (w_rhs - w_cs * xa) * rcp(w_bb);
  }
}

// ---- MPPThermalTBasedALM_Solve (src/driver/alm/MPPThermalTBasedALM_Driver.F90:150-452) on the device -------------------------------
// ELM's own column arrays -- Fortran (c, j) order, layer-major, j = -nlevsno+1 .. nlevgrnd (zi and tvector from -nlevsno) -- in and out;
// the packing into the SoE mailbox (:204-330) and the unpacking of the solution (:460-505) are kernels, so nothing is packed on the CPU.
struct ElmThermalArgs {
  int ncol, nlev, nsno;
  int col0, ncols;                           // the columns this launch covers
  double capr;                               // mpp_varcon.F90:30
  const int *active_col;                     // column filter (col%active and not lake / urban) or nullptr
  const int *snl;
  const double *z, *dz, *zi, *t_soisno, *h2osoi_liq, *h2osoi_ice;
  const double *frac_sno_eff, *h2osno, *h2osfc, *frac_h2osfc, *t_h2osfc, *sabg_lyr, *dhsdT, *hs_soil, *hs_top_snow, *hs_h2osfc;
  // SoE mailbox (see ThermalSnowArgs)
  double *T, *liq, *ice, *snow_water, *mdz, *dist_up, *dist_dn, *tuning, *frac; int *nsnow, *active;
  double *hs[3], *dhs[3], *frac_soil, *sabg_snow, *sabg_soil;
  const double *T_out; double *tvector;
};

// one thread per (row, column), column fastest: ELM's layer-major arrays are read coalesced
__global__ void elm_thermal_pack_kernel(const ElmThermalArgs A)
{
  const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int ncol = A.ncol, nsno = A.nsno, nlev = A.nlev, nrow = nsno + 1 + nlev;
  if (tid >= (long long)A.ncols * nrow) return;
  const int c = A.col0 + (int)(tid % A.ncols), r = (int)(tid / A.ncols);
  const bool on = (A.active_col == nullptr) || A.active_col[c] != 0;
  const int snl = A.snl[c];
  // ELM layer j lives at array column (j + nsno - 1) of the (-nsno+1 : nlev) arrays and at (j + nsno) of zi
#define EL(a, j)  (a)[(long long)((j) + nsno - 1) * ncol + c]
#define ZI(j)     A.zi[(long long)((j) + nsno) * ncol + c]
  if (r < nsno) {                                                        // snow layer j = r - nsno + 1 (:204-240)
    const int j = r - nsno + 1;
    const long long idx = (long long)c * nsno + r;
    double T = 273.15, liq = 0.0, ice = 0.0, dz = 0.0, du = 0.0, dd = 0.0, frac = 1.0, tun = 1.0, sabg = 0.0; int ns = 0, act = 0;
    if (on && j >= snl + 1) {
      T = EL(A.t_soisno, j); dz = EL(A.dz, j); liq = EL(A.h2osoi_liq, j); ice = EL(A.h2osoi_ice, j); ns = -snl; act = 1;
      du = ZI(j) - EL(A.z, j); dd = EL(A.z, j) - ZI(j - 1); frac = A.frac_sno_eff[c];
      if (j != snl + 1) sabg = A.sabg_lyr[(long long)(j + nsno - 1) * ncol + c];
      else tun = EL(A.dz, j) / (0.5 * __dadd_rn(EL(A.z, j) - ZI(j - 1), __dmul_rn(A.capr, EL(A.z, j + 1) - ZI(j - 1))));   // no FMA contraction: bit-identical to the driver
    }
    A.T[idx] = T; A.liq[idx] = liq; A.ice[idx] = ice; A.snow_water[idx] = 0.0; A.mdz[idx] = dz; A.dist_up[idx] = du; A.dist_dn[idx] = dd;
    A.frac[idx] = frac; A.tuning[idx] = tun; A.nsnow[idx] = ns; A.active[idx] = act; A.sabg_snow[idx] = sabg;
  } else if (r == nsno) {                                                // standing surface water (:243-262) + the per-column conditions
    const long long idx = (long long)ncol * nsno + c;
    double T = 273.15, dz = 0.0, frac = 1.0, du = 0.0; int act = 0;
    double frac_soil = 1.0, hs0 = 0.0, dh0 = 0.0, hs1 = 0.0, dh1 = 0.0, hs2 = 0.0, dh2 = 0.0;
    if (on) {
      if (snl < 0) { hs0 = A.hs_top_snow[c]; dh0 = A.dhsdT[c]; frac_soil = frac_soil - A.frac_sno_eff[c]; }     // at j == snl+1 (:229-233)
      const double fw = A.frac_h2osfc[c];
      if (fw > 0.0) {
        T = A.t_h2osfc[c]; dz = 1.0e-3 * A.h2osfc[c]; act = 1; frac = fw; du = dz / 2.0;
        frac_soil = frac_soil - fw; dh1 = A.dhsdT[c]; hs1 = A.hs_h2osfc[c];
      }
      hs2 = A.hs_soil[c]; dh2 = A.dhsdT[c];                              // at soil layer 1 (:322-323)
    }
    A.T[idx] = T; A.liq[idx] = 0.0; A.ice[idx] = 0.0; A.snow_water[idx] = 0.0; A.mdz[idx] = dz; A.dist_up[idx] = du; A.dist_dn[idx] = du;
    A.frac[idx] = frac; A.tuning[idx] = 1.0; A.nsnow[idx] = 0; A.active[idx] = act;
    A.hs[0][c] = hs0; A.dhs[0][c] = dh0; A.hs[1][c] = hs1; A.dhs[1][c] = dh1; A.hs[2][c] = hs2; A.dhs[2][c] = dh2; A.frac_soil[c] = frac_soil;
  } else {                                                               // soil layer j = r - nsno (:268-327)
    const int j = r - nsno;
    const long long sc = (long long)c * nlev + (j - 1), idx = (long long)ncol * (nsno + 1) + sc;
    double T = 273.15, liq = 0.0, ice = 0.0, dz = 0.0, du = 0.0, frac = 1.0, tun = 1.0, sabg = 0.0, sw = 0.0; int ns = 0, act = 0;
    if (on) {
      T = EL(A.t_soisno, j); dz = EL(A.dz, j); act = 1; liq = EL(A.h2osoi_liq, j); ice = EL(A.h2osoi_ice, j);
      du = ZI(j) - EL(A.z, j);
      if (j == 1) {
        dz = EL(A.z, j) * 2.0; ns = -snl;
        if (snl != 0) { sabg = A.frac_sno_eff[c] * A.sabg_lyr[(long long)(j + nsno - 1) * ncol + c]; sw = A.h2osno[c]; }
        else tun = EL(A.dz, j) / (0.5 * __dadd_rn(EL(A.z, j) - ZI(j - 1), __dmul_rn(A.capr, EL(A.z, j + 1) - ZI(j - 1))));   // no FMA contraction: bit-identical to the driver
      }
    }
    A.T[idx] = T; A.liq[idx] = liq; A.ice[idx] = ice; A.snow_water[idx] = sw; A.mdz[idx] = dz; A.dist_up[idx] = du; A.dist_dn[idx] = du;
    A.frac[idx] = frac; A.tuning[idx] = tun; A.nsnow[idx] = ns; A.active[idx] = act; A.sabg_soil[sc] = sabg;
  }
}

// tvector(c, j-1) = snow layer j (active layers), tvector(c, 0) = standing water (when present), tvector(c, j) = soil layer j (:460-505)
__global__ void elm_thermal_unpack_kernel(const ElmThermalArgs A)
{
  const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int ncol = A.ncol, nsno = A.nsno, nlev = A.nlev, nrow = nsno + 1 + nlev;
  if (tid >= (long long)A.ncols * nrow) return;
  const int c = A.col0 + (int)(tid % A.ncols), r = (int)(tid / A.ncols);
  if (A.active_col != nullptr && A.active_col[c] == 0) return;
  const int snl = A.snl[c];
  if (r < nsno) {
    const int j = r - nsno + 1;
    if (j >= snl + 1) A.tvector[(long long)(j - 1 + nsno) * ncol + c] = A.T_out[(long long)c * nsno + r];
  } else if (r == nsno) {
    if (A.frac_h2osfc[c] > 0.0) A.tvector[(long long)(0 + nsno) * ncol + c] = A.T_out[(long long)ncol * nsno + c];
  } else {
    const int j = r - nsno;
    A.tvector[(long long)(j + nsno) * ncol + c] = A.T_out[(long long)ncol * (nsno + 1) + (long long)c * nlev + (j - 1)];
  }
#undef EL
#undef ZI
}

}  // namespace mpp
