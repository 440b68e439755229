// vsfm_kernels2.cuh -- fused VSFM (Richards equation) time step, two soil cells per lane.
//
// One launch = one sysofeqns%StepDT for every column of the batch (same reference citations as vsfm_kernels.cuh):
//   SOEBaseStepDT_SNES         src/mpp/soe/SystemOfEquationsBaseType.F90:368-552  (dt cuts, <= 20)
//   VSFMSOEPreSolve/PostSolve  src/mpp/soe/SystemOfEquationsVSFMType.F90:506-660
//   VSFMSOEResidual/Jacobian   src/mpp/soe/SystemOfEquationsVSFMType.F90:94-403
//   Richards residual/Jacobian src/mpp/ge/GoveqnRichardsODEPressureType.F90:1603-2200
//   RichardsFlux               src/mpp/ge/RichardsMod.F90:118-340
//   PETSc SNES newtonls + bt line search + SNESConvergedDefault, KSP on a tridiagonal matrix
//
// Mapping (DESIGN.md "VSFM kernel").  The path is bound by fp64 issue LATENCY, not by HBM: the first lane-per-cell
// version ran at 0.45 IPC per scheduler with the fp64 pipe 29 % busy (profiles/r1_vsfm_v2.md) because every lane
// carried one serial log -> exp -> log -> exp chain and 128 registers allowed only 4 warps per scheduler.  Here LPC
// (8 or 16) lanes own one column and each lane owns TWO adjacent cells (2l, 2l+1):
//   * two independent transcendental chains per lane (ILP 2) hide the DFMA latency;
//   * a warp advances 4 (or 2) columns, so the per-warp control flow, norms and shuffles are amortised over twice
//     as many cells; the connection between a lane's two cells needs no shuffle at all;
//   * the tridiagonal Newton system is reduced in-lane to one row per lane (odd-even elimination of the lane's second
//     cell), solved by parallel cyclic reduction over LPC lanes in normalised form (3 shuffled doubles per side per
//     stage instead of 4, log2(LPC) stages), then back-substituted in-lane;
//   * all norms are carried squared (no sqrt in the Newton loop) and are bitwise identical in every lane of a column,
//     so control flow is uniform per column; converged columns idle until their warp's slowest column is done.
// HBM traffic is the algorithmic minimum: every input array is read once, every output written once, in the
// reference's own cell order (icell = c*nlev + j), fully coalesced.
#pragma once
#include "vsfm_kernels.cuh"

namespace mpp {

// Threads per block and resident blocks per SM.  A block cannot retire before its slowest column has converged, and the
// iteration counts have a long tail (mean 3.4, max 40+), so small blocks matter: with one warp per block a finished warp
// frees its registers at once instead of idling next to a straggler.
#ifndef VSFM2_THREADS
#define VSFM2_THREADS 32
#endif
#ifndef VSFM2_MIN_BLOCKS
#define VSFM2_MIN_BLOCKS 16
#endif
// the instances with boundary conditions / per-column retries / the evaluation probe carry more live state: at 16 blocks per SM (128
// registers) they spilled 170-200 bytes; 12 blocks (168 registers, no spills) measured 7 % faster on 1 Mi columns with a Dirichlet head
// at the bottom (2.65 vs 2.85 ms per step)
#ifndef VSFM2_MIN_BLOCKS_BC
#define VSFM2_MIN_BLOCKS_BC 12
#endif

template <int LPC>
__device__ __forceinline__ double col_sum(double v)
{
#pragma unroll
  for (int s = LPC / 2; s > 0; s >>= 1) v += __shfl_xor_sync(FULL_MASK, v, s, LPC);
  return v;
}

// Parallel cyclic reduction over LPC lanes of a tridiagonal system in normalised form (unit diagonal):
//   al x[l-1] + x[l] + ga x[l+1] = de.   The couplings that would reach outside the chain are exactly zero at every
// stage, so whatever an out-of-range shuffle returns (the lane's own finite value) is multiplied by zero.
template <int LPC>
__device__ __forceinline__ double pcr_unit_diag(double al, double ga, double de)
{
#pragma unroll
  for (int s = 1; s < LPC; s <<= 1) {
    const double al_m = __shfl_up_sync(FULL_MASK, al, s, LPC),   ga_m = __shfl_up_sync(FULL_MASK, ga, s, LPC);
    const double de_m = __shfl_up_sync(FULL_MASK, de, s, LPC);
    const double al_p = __shfl_down_sync(FULL_MASK, al, s, LPC), ga_p = __shfl_down_sync(FULL_MASK, ga, s, LPC);
    const double de_p = __shfl_down_sync(FULL_MASK, de, s, LPC);
    const double r = rcp1(1.0 - al * ga_m - ga * al_p);
    de = (de - al * de_m - ga * de_p) * r;
    al = (-al * al_m) * r;
    ga = (-ga * ga_p) * r;
  }
  return de;
}

// Newton state of one soil cell (registers).  Its static data -- curve parameters, porosity, volume, net source, the
// connection to the next cell (upwind weight, Dq, gravity factor: MeshType.F90:509-530; RichardsMod.F90:257-259, 279-285)
// -- and the two values that change only once per sub-step (soln_prev, accumulation at soln_prev) are parked in shared
// memory, [field][thread] (conflict-free), and re-read where they are used: 48 registers less of persistent state per
// lane, which the register allocator spends on overlapping the two cells' dependency chains instead of spilling
// (once the residual and the aux vars of the accepted iterate moved here too, 122 registers without spills: 16 one-warp blocks per
// SM; squeezing further -- 18+ blocks, 96 registers -- spills and is slower again).
template <int SATFUNC>
struct Cell2 {
  double X, W;
  bool valid, has_conn;
};
// shared-memory fields per cell: static data, then what changes once per sub-step (soln_prev, accumulation there), then what changes
// once per accepted iterate (residual and aux vars at X: the Jacobian reads them, and the next cell's, from here)
enum { PI_SATRES, PI_ALPHA, PI_M, PI_N, PI_POR, PI_VOL, PI_SRC, PI_UPW, PI_DQ, PI_GFAC, PI_XPREV, PI_ACCP, PI_F, PI_KR, PI_DKR, PI_SAT, PI_DSAT, PI_Y,
       PI_FLIQ, PI_PU, PI_PS, PI_B2, PI_B3 };
template <int SATFUNC> struct ParCount { static constexpr int value = (SATFUNC == SATFUNC_VG) ? 18 : (SATFUNC == SATFUNC_BC ? 19 : 23); };

template <int SATFUNC>
__device__ __forceinline__ void cell_load(const VsfmArgs &A, Cell2<SATFUNC> &c, double (*par)[VSFM2_THREADS], bool valid, long long cell, double area, double &perm, double &dz,
                                          const double *x_src)
{
  const int t = threadIdx.x;
  c.valid = valid;
  double sat_res = 0.0, alpha = 1.0, m = 0.5, n = 2.0, pu = 0.0, ps = 0.0, b2 = 0.0, b3 = 0.0, por = 0.0, fl = 1.0;
  c.X = PRESSURE_REF; perm = 1.0; dz = 1.0;
  if (valid) {
    por = A.por[cell]; perm = A.perm[cell]; dz = A.dz[cell];
    sat_res = A.sat_res[cell]; alpha = A.alpha[cell]; m = A.lam[cell];
    if (SATFUNC == SATFUNC_VG)  n = A.vgn[cell];
    if (SATFUNC == SATFUNC_SBC) { pu = A.pu[cell]; ps = A.ps[cell]; b2 = A.b2[cell]; b3 = A.b3[cell]; }
    if (SATFUNC != SATFUNC_VG)  fl = A.frac_liq[cell];            // only the Brooks-Corey k_r reads it (SaturationFunction.F90:987)
    c.X = x_src[cell];
  }
  par[PI_SATRES][t] = sat_res; par[PI_ALPHA][t] = alpha; par[PI_M][t] = m; par[PI_N][t] = n;
  par[PI_POR][t] = por; par[PI_VOL][t] = area * dz;               // MeshType.F90:427
  par[PI_SRC][t] = 0.0; par[PI_XPREV][t] = c.X; par[PI_ACCP][t] = 0.0;
  if (SATFUNC != SATFUNC_VG)  par[PI_FLIQ][t] = fl;
  if (SATFUNC == SATFUNC_SBC) { par[PI_PU][t] = pu; par[PI_PS][t] = ps; par[PI_B2][t] = b2; par[PI_B3][t] = b3; }
  c.W = c.X;
  par[PI_Y][t] = 0.0; par[PI_F][t] = 0.0; par[PI_KR][t] = 1.0; par[PI_DKR][t] = 0.0; par[PI_SAT][t] = 1.0; par[PI_DSAT][t] = 0.0;
}

template <int SATFUNC>
__device__ __forceinline__ SatParams par_sp(double (*par)[VSFM2_THREADS])
{
  const int t = threadIdx.x;
  SatParams sp;
  sp.sat_res = par[PI_SATRES][t]; sp.alpha = par[PI_ALPHA][t]; sp.m = par[PI_M][t]; sp.n = par[PI_N][t];
  sp.pu = sp.ps = sp.b2 = sp.b3 = 0.0;
  if (SATFUNC == SATFUNC_SBC) { sp.pu = par[PI_PU][t]; sp.ps = par[PI_PS][t]; sp.b2 = par[PI_B2][t]; sp.b3 = par[PI_B3][t]; }
  return sp;
}

__device__ __forceinline__ void conn_setup(double perm_up, double dz_up, double perm_dn, double dz_dn, double uz, double &upw, double &Dq, double &gfac)
{
  const double dist_up = 0.5 * dz_up, dist_dn = 0.5 * dz_dn;
  // (lean reciprocals, <= 1 ulp: the last bit of a static coefficient is far below the 1e-10 parity bar)
  upw  = dist_up * rcp(dist_up + dist_dn);
  Dq   = (perm_up * perm_dn) * rcp(dist_up * perm_dn + dist_dn * perm_up);
  gfac = FMWH2O * ((dist_up + dist_dn) * (uz * (-GRAVITY_CONSTANT)));       // FMWH2O * dist_gravity
}

// RichardsFlux_Internal, residual part (RichardsMod.F90:257-296)
__device__ __forceinline__ double rich_flux(double P_u, double kr_u, double den_u, double P_d, double kr_d, double den_d,
                                            double upw, double Dq, double gfac, double area)
{
  constexpr double RVIS = 1.0 / VISCOSITY;
  const double den_ave = upw * den_u + (1.0 - upw) * den_d;
  const double dphi    = P_u - P_d + den_ave * gfac;
  const double ukvr    = ((dphi >= 0.0) ? kr_u : kr_d) * RVIS;
  return ((-Dq * ukvr * dphi) * area) * den_ave;
}

// RichardsFlux_Internal with compute_deriv (RichardsMod.F90:298-336): Jup = -d flux / dP_up, Jdn = -d flux / dP_dn
__device__ __forceinline__ void rich_flux_deriv(double P_u, double kr_u, double dkr_u, double den_u, double dden_u,
                                                double P_d, double kr_d, double dkr_d, double den_d, double dden_d,
                                                double upw, double Dq, double gfac, double area, double &Jup, double &Jdn)
{
  constexpr double RVIS = 1.0 / VISCOSITY;
  const double den_ave = upw * den_u + (1.0 - upw) * den_d;
  const double dphi    = P_u - P_d + den_ave * gfac;
  const bool   upwind  = (dphi >= 0.0);
  const double ukvr    = (upwind ? kr_u : kr_d) * RVIS;
  const double q       = (-Dq * ukvr * dphi) * area;
  const double dphi_dP_up =  1.0 + (upw * gfac) * dden_u;
  const double dphi_dP_dn = -1.0 + ((1.0 - upw) * gfac) * dden_d;
  const double dukvr_up = upwind ? dkr_u * RVIS : 0.0;
  const double dukvr_dn = upwind ? 0.0 : dkr_d * RVIS;
  const double dq_up = Dq * (dukvr_up * dphi + ukvr * dphi_dP_up) * area;
  const double dq_dn = Dq * (dukvr_dn * dphi + ukvr * dphi_dP_dn) * area;
  Jup = dq_up * den_ave - q * (upw * dden_u);
  Jdn = dq_dn * den_ave - q * ((1.0 - upw) * dden_d);
}

// HAS_BC: the batch has boundary conditions and / or a down-regulated sink (compiled out for plain ELM-like batches)
// RETRY: per-column remaining time, tolerances, start vector and a run mask (the retry loop of MPPVSFMALM_Solve); compiled as a
// separate specialisation so that the common path keeps its register budget
// EVAL: the residual / Jacobian probe (mppgpu_eval): evaluate at x_in (accumulation of the start of the step), then at eval_x, run the
// Newton set-up once, write F and the assembled Jacobian rows, and leave -- the SAME fused assembly code the time step runs
constexpr int PH_VEVAL = 6;
template <int LPC, int SATFUNC, bool HAS_BC, bool RETRY = false, bool EVAL = false>
__global__ void __launch_bounds__(VSFM2_THREADS, (HAS_BC || RETRY || EVAL) ? VSFM2_MIN_BLOCKS_BC : VSFM2_MIN_BLOCKS)
vsfm_step2_kernel(const VsfmArgs A)
{
  constexpr unsigned FULL = FULL_MASK;
  constexpr double RVIS = 1.0 / VISCOSITY, RFMW = 1.0 / FMWH2O;
  constexpr int NBC = HAS_BC ? MAX_BC : 1;
  const int tid  = blockIdx.x * blockDim.x + threadIdx.x;
  int col = tid / LPC;
  if (RETRY) col = (col < A.nretry) ? A.retry_list[col] : A.ncol;
  else if (A.order) col = (col < A.ncol) ? A.order[col] : A.ncol;
  const int l    = tid % LPC;                        // lane within the column; owns layers 2l and 2l+1
  const int lane = threadIdx.x & 31;
  const int nlev = A.nlev;
  const int j0 = 2 * l, j1 = 2 * l + 1;
  int rmask = 1;
  if (RETRY) rmask = (col < A.ncol) ? A.retry_mask[col] : 0;
  const bool col_ok = (col < A.ncol) && (A.active == nullptr || A.active[col] != 0) && (!RETRY || rmask != 0);
  const double *x_src = (RETRY && rmask == 2) ? A.x_redo : A.x_in;
  const long long cell0 = (long long)col * nlev + j0;
  const double area = col_ok ? A.area[col] : 1.0;

  // ---- static per-cell data -----------------------------------------------------------------------
  __shared__ double s_par[2][ParCount<SATFUNC>::value][VSFM2_THREADS];
  double (*const pa)[VSFM2_THREADS] = s_par[0], (*const pb)[VSFM2_THREADS] = s_par[1];
  const int tx = threadIdx.x;
#define PA(i) pa[i][tx]
#define PB(i) pb[i][tx]
  Cell2<SATFUNC> a, b;
  double perm0, dz0, perm1, dz1;
  cell_load<SATFUNC>(A, a, pa, col_ok && j0 < nlev, cell0, area, perm0, dz0, x_src);
  cell_load<SATFUNC>(A, b, pb, col_ok && j1 < nlev, cell0 + 1, area, perm1, dz1, x_src);
  // mass-rate source/sinks (GoveqnRichards...:1871-1875): F -= value / FMWH2O.  The launcher has sorted the conditions by region
  // (vsfm_compact_sources), so no type / region logic runs here; the first NC per-cell, NT top and NB bottom conditions (ELM: 2 + 4 + 0)
  // are loaded up front under uniform predicates, any further ones in a plain loop.
  const int jtop = A.top_is_first ? 0 : nlev - 1, jbot = A.top_is_first ? nlev - 1 : 0;
  double src_kg = 0.0;
  {
    constexpr int NC = 4, NT = 6, NB = 2;
    const bool top_a = a.valid && j0 == jtop, top_b = b.valid && j1 == jtop, bot_a = a.valid && j0 == jbot, bot_b = b.valid && j1 == jbot;
    const bool own_top = top_a || top_b, own_bot = bot_a || bot_b;
    double c0[NC], c1[NC], tp[NT], bt[NB];
#pragma unroll
    for (int k = 0; k < NC; ++k) {
      c0[k] = 0.0; c1[k] = 0.0;
      if (k < A.nss_cell) { const double *p = A.ss_cell[k] + cell0; if (a.valid) c0[k] = __ldg(p); if (b.valid) c1[k] = __ldg(p + 1); }
    }
#pragma unroll
    for (int k = 0; k < NT; ++k) { tp[k] = 0.0; if (k < A.nss_top && own_top) tp[k] = __ldg(A.ss_top[k] + col); }
#pragma unroll
    for (int k = 0; k < NB; ++k) { bt[k] = 0.0; if (k < A.nss_bot && own_bot) bt[k] = __ldg(A.ss_bot[k] + col); }
    double sa_ = 0.0, sb_ = 0.0, st_ = 0.0, sbt_ = 0.0;
#pragma unroll
    for (int k = 0; k < NC; ++k) { sa_ += c0[k] * RFMW; sb_ += c1[k] * RFMW; src_kg += c0[k] + c1[k]; }
#pragma unroll
    for (int k = 0; k < NT; ++k) { st_ += tp[k] * RFMW; src_kg += tp[k]; }
#pragma unroll
    for (int k = 0; k < NB; ++k) { sbt_ += bt[k] * RFMW; src_kg += bt[k]; }
    for (int k = NC; k < A.nss_cell; ++k) {
      const double *p = A.ss_cell[k] + cell0;
      const double v0 = a.valid ? __ldg(p) : 0.0, v1 = b.valid ? __ldg(p + 1) : 0.0;
      sa_ += v0 * RFMW; sb_ += v1 * RFMW; src_kg += v0 + v1;
    }
    for (int k = NT; k < A.nss_top; ++k) { const double v = own_top ? __ldg(A.ss_top[k] + col) : 0.0; st_ += v * RFMW; src_kg += v; }
    for (int k = NB; k < A.nss_bot; ++k) { const double v = own_bot ? __ldg(A.ss_bot[k] + col) : 0.0; sbt_ += v * RFMW; src_kg += v; }
    sa_ += (top_a ? st_ : 0.0) + (bot_a ? sbt_ : 0.0);
    sb_ += (top_b ? st_ : 0.0) + (bot_b ? sbt_ : 0.0);
    PA(PI_SRC) = sa_; PB(PI_SRC) = sb_;
  }

  // connections (after the source loads above were issued: the shuffles below wait for the soil tables, and anything placed behind them
  // would start a second trip to DRAM)
  {
    const double perm_n = __shfl_down_sync(FULL, perm0, 1, LPC), dz_n = __shfl_down_sync(FULL, dz0, 1, LPC);
    a.has_conn = b.valid;                            // 2l -> 2l+1
    b.has_conn = b.valid && (j1 < nlev - 1);         // 2l+1 -> 2(l+1)
    double upw, Dq, gfac;
    conn_setup(perm0, dz0, perm1, dz1, A.uz, upw, Dq, gfac);   PA(PI_UPW) = upw; PA(PI_DQ) = Dq; PA(PI_GFAC) = gfac;
    conn_setup(perm1, dz1, perm_n, dz_n, A.uz, upw, Dq, gfac); PB(PI_UPW) = upw; PB(PI_DQ) = Dq; PB(PI_GFAC) = gfac;
  }

  // optional down-regulated sink: which of this lane's cells it touches, and where its per-connection data sits
  bool dr_a = false, dr_b = false; long long dr_ia = 0, dr_ib = 0;
  if (HAS_BC && A.dr_type) {
    const bool percell = (A.dr_region == REGION_CELLS);
    const int jown = (A.dr_region == REGION_TOP) ? jtop : jbot;
    dr_a = a.valid && (percell || j0 == jown); dr_b = b.valid && (percell || j1 == jown);
    dr_ia = percell ? cell0 : (long long)col; dr_ib = percell ? cell0 + 1 : (long long)col;
  }

  // boundary conditions (MeshType.F90:723-806): top -> unit vector (0,0,-1), bottom -> (0,0,+1); dist_up = 0
  double bcP[NBC], bcKr[NBC], bcGfac[NBC], bcDq[NBC], bcMassExc[NBC], bcFlux[NBC];
  int    bcOwn[NBC];                                 // 0: not on this lane, 1: cell a, 2: cell b
#pragma unroll
  for (int k = 0; k < NBC; ++k) {
    bcOwn[k] = 0; bcP[k] = PRESSURE_REF; bcKr[k] = 1.0; bcGfac[k] = 0.0; bcDq[k] = 0.0; bcMassExc[k] = 0.0; bcFlux[k] = 0.0;
    if (HAS_BC && k < A.nbc) {
      const bool top = (A.bc[k].region == REGION_TOP);
      const int jown = top ? jtop : jbot;
      if (a.valid && j0 == jown) bcOwn[k] = 1;
      if (b.valid && j1 == jown) bcOwn[k] = 2;
      if (bcOwn[k]) {
        const double dzc = (bcOwn[k] == 1) ? dz0 : dz1, permc = (bcOwn[k] == 1) ? perm0 : perm1;
        const double uzbc = (A.uz == 0.0) ? 0.0 : (top ? -1.0 : 1.0);
        bcGfac[k] = FMWH2O * ((0.0 + 0.5 * dzc) * (uzbc * (-GRAVITY_CONSTANT)));
        bcDq[k] = permc / (0.0 + 0.5 * dzc);
        bcP[k] = A.bc[k].value[col];
        SatState sb;
        sat_values<SATFUNC>((bcOwn[k] == 1) ? par_sp<SATFUNC>(pa) : par_sp<SATFUNC>(pb), bcP[k], 1.0, sb);   // BC aux vars keep frac_liq_sat = 1 (RichardsODEPressureAuxType.F90:93)
        bcKr[k] = sb.kr;
      }
    }
  }

  // ---- time-step / Newton state (uniform per column unless noted) ---------------------------------
  const SnesOpts so = A.so;
  const double atol2 = so.atol * so.atol, rtol2_u = so.rtol * so.rtol, stol2_u = so.stol * so.stol;
  const double divtol2 = so.divtol * so.divtol, maxstep2 = so.ls_maxstep * so.ls_maxstep;
  // per-column scalars that are touched once per sub-step or only while back-tracking live in shared memory as well
  __shared__ double s_sc[RETRY ? 6 : 3][VSFM2_THREADS];
#define SC_TDONE   s_sc[0][tx]
#define SC_LAMPREV s_sc[1][tx]
#define SC_GPREV   s_sc[2][tx]
  SC_TDONE = 0.0; SC_LAMPREV = 1.0; SC_GPREV = 0.0;
  if (RETRY) {
    s_sc[RETRY ? 3 : 0][tx] = col_ok ? A.dt_col[col] : A.dt;
    const double rt = col_ok ? A.rtol_col[col] : so.rtol, st = col_ok ? A.stol_col[col] : so.stol;
    s_sc[RETRY ? 4 : 0][tx] = rt * rt; s_sc[RETRY ? 5 : 0][tx] = st * st;
  }
#define DT_TOT (RETRY ? s_sc[RETRY ? 3 : 0][tx] : A.dt)
#define rtol2  (RETRY ? s_sc[RETRY ? 4 : 0][tx] : rtol2_u)
#define stol2  (RETRY ? s_sc[RETRY ? 5 : 0][tx] : stol2_u)
  double dt_iter = DT_TOT, dtInv = 1.0 / dt_iter;
  int    cuts = 0, tot_its = 0, tot_nf = 0, last_reason = 0, converged = 0;
  int    phase = col_ok ? PH_INIT : PH_DONE;
  int    its = 0, nfuncs = 0, ls_count = 0;
  double f2 = 0.0, x2 = 0.0, y2 = 0.0, f2_0 = 0.0;                   // squared norms ||F||^2, ||X||^2, ||Y||^2, ||F0||^2
  double initslope = -1.0, lambda = 1.0;

#ifdef VSFM2_PROFILE
  long long pt_newton = 0, pt_eval = 0, pt_logic = 0, pn_newton = 0, pn_eval = 0, pt_curves = 0, pt_red = 0, pt_nasm = 0, pt_nelim = 0, pt_npcr = 0; const long long pt_start = clock64();
#endif
  for (;;) {
#ifdef VSFM2_PROFILE
    const long long pt0 = clock64();
#endif
    // ================= Newton step set-up: Jacobian, linear solve, line-search initialisation =================
    // Executed by the WHOLE warp whenever any of its columns starts a Newton iteration (warp-uniform branch, so every
    // shuffle below is convergent); lanes of a column that is not in PH_NEWTON compute and discard.
    if (__any_sync(FULL, phase == PH_NEWTON)) {
      const bool nw = (phase == PH_NEWTON);
      __syncwarp();                                  // (also keeps the compiler from hoisting the shared-memory reads out of the loop)
      double den_a, dden_a, den_b, dden_b;
      density_fixedT_x<HAS_BC>(A.dtab, a.X, den_a, dden_a);
      density_fixedT_x<HAS_BC>(A.dtab, b.X, den_b, dden_b);
      // aux vars at X of this lane's cells, and of the first cell of the next lane (dn side of connection b) straight from
      // shared memory; only the pressure itself lives in a register and needs a shuffle
      const double kr_a = PA(PI_KR), dkr_a_ = PA(PI_DKR), sat_a = PA(PI_SAT), dsat_a_ = PA(PI_DSAT);
      const double kr_b = PB(PI_KR), dkr_b_ = PB(PI_DKR), sat_b = PB(PI_SAT), dsat_b_ = PB(PI_DSAT);
      const int txn = b.has_conn ? tx + 1 : tx;
      const double krn = pa[PI_KR][txn], dkrn = pa[PI_DKR][txn];
      const double Xn = __shfl_down_sync(FULL, a.X, 1, LPC);
      double denn, ddenn;
      density_fixedT_x<HAS_BC>(A.dtab, Xn, denn, ddenn);
      // both connections evaluated unconditionally (all inputs are finite on padding lanes) and masked afterwards: no divergent
      // branch, and the two derivative chains overlap
      double Jup_a, Jdn_a, Jup_b, Jdn_b;
      rich_flux_deriv(a.X, kr_a, dkr_a_, den_a, dden_a, b.X, kr_b, dkr_b_, den_b, dden_b, PA(PI_UPW), PA(PI_DQ), PA(PI_GFAC), area, Jup_a, Jdn_a);
      rich_flux_deriv(b.X, kr_b, dkr_b_, den_b, dden_b, Xn, krn, dkrn, denn, ddenn, PB(PI_UPW), PB(PI_DQ), PB(PI_GFAC), area, Jup_b, Jdn_b);
      Jup_a = a.has_conn ? Jup_a : 0.0; Jdn_a = a.has_conn ? Jdn_a : 0.0;
      Jup_b = b.has_conn ? Jup_b : 0.0; Jdn_b = b.has_conn ? Jdn_b : 0.0;
      const double Jup_p = __shfl_up_sync(FULL, Jup_b, 1, LPC), Jdn_p = __shfl_up_sync(FULL, Jdn_b, 1, LPC);   // connection (2l-1) -> 2l
      // rows 2l and 2l+1 of the tridiagonal Jacobian (GoveqnRichards...:2054-2069 insertion order)
      // (padding cells: identity rows; their couplings are already zero because has_conn is false around them)
      const double sub_a = (a.valid && l > 0) ? -Jup_p : 0.0, sup_a = Jdn_a;
      double dia_a = a.valid ? ((l > 0) ? -Jdn_p : 0.0) + Jup_a : 1.0;
      const double sub_b = -Jup_a, sup_b = Jdn_b;
      double dia_b = b.valid ? -Jdn_a + Jup_b : 1.0;
      if (HAS_BC) {
#pragma unroll
        for (int k = 0; k < NBC; ++k) if (bcOwn[k]) {                // boundary: (dn,dn) -= Jdn  (:2136-2140)
          const bool onA = (bcOwn[k] == 1);
          const double Xc = onA ? a.X : b.X, krc = onA ? kr_a : kr_b, dkrc = onA ? dkr_a_ : dkr_b_;
          const double denc = onA ? den_a : den_b, ddenc = onA ? dden_a : dden_b;
          const double dphi0 = bcP[k] - Xc + denc * bcGfac[k];
          const bool seep = (A.bc[k].itype == CT_SEEPAGE) && (dphi0 > 0.0) && (bcP[k] <= PRESSURE_REF);
          const double dphi = seep ? 0.0 : dphi0;
          const bool upwind = (dphi >= 0.0);
          const double ukvr = (upwind ? bcKr[k] : krc) * RVIS;
          const double q    = (-bcDq[k] * ukvr * dphi) * area;
          const double dphi_dP_dn = seep ? 0.0 : (-1.0 + bcGfac[k] * ddenc);
          const double dukvr_dn = upwind ? 0.0 : dkrc * RVIS;
          const double dq_dn = bcDq[k] * (dukvr_dn * dphi + ukvr * dphi_dP_dn) * area;
          const double t = -(dq_dn * denc - q * ddenc);
          if (onA) dia_a += t; else dia_b += t;
        }
      }
      {                                             // AccumDeriv (:1673-1675), dpor_dP = 0
        const double pora = PA(PI_POR), porb = PB(PI_POR);
        const double da = (pora * dden_a * sat_a + pora * den_a * dsat_a_) * PA(PI_VOL) * dtInv;
        const double db = (porb * dden_b * sat_b + porb * den_b * dsat_b_) * PB(PI_VOL) * dtInv;
        dia_a += a.valid ? da : 0.0; dia_b += b.valid ? db : 0.0;
      }

      if (HAS_BC && A.dr_type) {                               // uniform branch; rare path, data re-read from HBM / L2
        double rate, dj;
        if (dr_a) { downreg_sink(A.dr_type, A.dr_value[dr_ia], A.dr_pc[dr_ia], A.dr_n[dr_ia], a.X, rate, dj); dia_a += dj; }
        if (dr_b) { downreg_sink(A.dr_type, A.dr_value[dr_ib], A.dr_pc[dr_ib], A.dr_n[dr_ib], b.X, rate, dj); dia_b += dj; }
      }

      if (EVAL) {
        if (a.valid) { A.eval_f[cell0] = PA(PI_F); A.eval_ja[cell0] = sub_a; A.eval_jb[cell0] = dia_a; A.eval_jc[cell0] = sup_a; }
        if (b.valid) { A.eval_f[cell0 + 1] = PB(PI_F); A.eval_ja[cell0 + 1] = sub_b; A.eval_jb[cell0 + 1] = dia_b; A.eval_jc[cell0 + 1] = sup_b; }
        return;                                       // every column of the warp reaches this point in the same pass
      }
#ifdef VSFM2_PROFILE
      const long long pn1 = clock64() + (long long)(1e-300 * (dia_a + dia_b)); pt_nasm += pn1 - pt0;
#endif
      // ---- J Y = F: eliminate this lane's second unknown, PCR over the first unknowns, back-substitute ----
      const double Fa = a.valid ? PA(PI_F) : 0.0, Fb = b.valid ? PB(PI_F) : 0.0;
      const double rb = rcp1(dia_b);
      const double bs = sub_b * rb, bu = sup_b * rb, bf = Fb * rb;   // y_b = bf - bs y_a(l) - bu y_a(l+1)
      const double bs_p = __shfl_up_sync(FULL, bs, 1, LPC), bu_p = __shfl_up_sync(FULL, bu, 1, LPC), bf_p = __shfl_up_sync(FULL, bf, 1, LPC);
      // reduced row l: sub_a y_b(l-1) + dia_a y_a(l) + sup_a y_b(l) = Fa   (sub_a = 0 on lane 0)
      const double rB = rcp1(dia_a - sub_a * bu_p - sup_a * bs);
      const double al = (-sub_a * bs_p) * rB, ga = (-sup_a * bu) * rB, de = (Fa - sub_a * bf_p - sup_a * bf) * rB;
#ifdef VSFM2_PROFILE
      const long long pn2 = clock64() + (long long)(1e-300 * (al + ga + de)); pt_nelim += pn2 - pn1;
#endif
      const double Ya = pcr_unit_diag<LPC>(al, ga, de);
#ifdef VSFM2_PROFILE
      const long long pn3 = clock64() + (long long)(1e-300 * Ya); pt_npcr += pn3 - pn2;
#endif
      const double Ya_n = __shfl_down_sync(FULL, Ya, 1, LPC);
      const double Yb = bf - bs * Ya - bu * Ya_n;                    // bu = 0 where there is no next cell
      const double Yb_p = __shfl_up_sync(FULL, Yb, 1, LPC);
      // initslope = F . (J Y), forced negative (SNESLineSearchApply_BT)
      double JYa = dia_a * Ya + sup_a * Yb, JYb = sub_b * Ya + dia_b * Yb;
      if (l > 0) JYa += sub_a * Yb_p;
      if (b.has_conn) JYb += sup_b * Ya_n;
      const double yn2 = col_sum<LPC>((a.valid ? Ya * Ya : 0.0) + (b.valid ? Yb * Yb : 0.0));
      double slope = col_sum<LPC>(Fa * JYa + Fb * JYb);
      // line-search initialisation: the common case (a usable step, full step tried first) is written with selects; the rare
      // exits (zero step, step clipped at maxstep, function budget) take the branch
      if (slope > 0.0) slope = -slope;
      if (slope == 0.0) slope = -1.0;
      if (nw) { PA(PI_Y) = Ya; PB(PI_Y) = Yb; }
      y2 = nw ? yn2 : y2;                                             // x2 = ||X||^2 was taken when X was accepted
      initslope = nw ? slope : initslope;
      lambda = nw ? 1.0 : lambda; ls_count = nw ? 0 : ls_count;
      a.W = nw ? a.X - Ya : a.W; b.W = nw ? b.X - Yb : b.W;           // W = X - lambda Y with lambda = 1
      phase = nw ? PH_LS_FULL : phase;
      if (nw && (yn2 == 0.0 || yn2 > maxstep2 || (nfuncs >= so.max_funcs && so.max_funcs >= 0))) {
        if (y2 == 0.0) {
          // zero step: line search "fails"; stol*xnorm > ynorm => SNES_CONVERGED_SNORM_RELATIVE (ls.c)
          last_reason = (stol2 * x2 > y2) ? SNES_CONVERGED_SNORM_RELATIVE : SNES_DIVERGED_LINE_SEARCH;
          phase = -1;   // SNES finished, handled below
        } else {
          if (y2 > maxstep2) { const double sc = so.ls_maxstep / sqrt(y2); PA(PI_Y) = Ya * sc; PB(PI_Y) = Yb * sc; y2 = maxstep2; }
          a.W = fma(-lambda, PA(PI_Y), a.X); b.W = fma(-lambda, PB(PI_Y), b.X);
          if (nfuncs >= so.max_funcs && so.max_funcs >= 0) { last_reason = SNES_DIVERGED_FUNCTION_COUNT; phase = -1; }
        }
      }
    }

    // ================= end-of-SNES bookkeeping (SOEBaseStepDT_SNES :481-536) =================
    if (phase == -1) {
      tot_nf += nfuncs;
      if (last_reason < 0) {
        cuts += 1; dt_iter = 0.5 * dt_iter; dtInv = 1.0 / dt_iter;
        a.X = PA(PI_XPREV); b.X = PB(PI_XPREV);             // VecCopy(soln_prev, soln)
        if (cuts > 20) { converged = 0; phase = PH_DONE; }
        else { a.W = a.X; b.W = b.X; phase = PH_INIT; }
      } else {
        converged = 1; const double time_done = SC_TDONE + dt_iter; SC_TDONE = time_done; tot_its += its;
        PA(PI_XPREV) = a.X; PB(PI_XPREV) = b.X;             // PostSolve: soln -> soln_prev
        if (HAS_BC) {
#pragma unroll
          for (int k = 0; k < NBC; ++k) if (bcOwn[k]) bcMassExc[k] += bcFlux[k] * dt_iter;
        }
        if (time_done >= DT_TOT) phase = PH_DONE;
        else { a.W = a.X; b.W = b.X; phase = PH_INIT; }
      }
      its = 0; nfuncs = 0;
      // optional give-up budget (mppgpu_set_step_budget; not in the reference, off by default): a column that has burnt this
      // many residual evaluations inside one StepDT fails like one that ran out of dt cuts, instead of stalling the batch
      if (so.step_budget > 0 && tot_nf >= so.step_budget && phase != PH_DONE) { converged = 0; last_reason = SNES_DIVERGED_FUNCTION_COUNT; phase = PH_DONE; }
    }

    if (__all_sync(FULL, phase == PH_DONE)) break;
#ifdef VSFM2_PROFILE
    const long long pt1 = clock64(); if (pt1 - pt0 > 200) { pt_newton += pt1 - pt0; pn_newton++; }
#endif

    // ================= residual evaluation at W (VSFMSOEResidual) =================
    __syncwarp();
    SatState sa, sb;
    const SatParams spa = par_sp<SATFUNC>(pa), spb = par_sp<SATFUNC>(pb);
    const double fla = (SATFUNC != SATFUNC_VG) ? PA(PI_FLIQ) : 1.0, flb = (SATFUNC != SATFUNC_VG) ? PB(PI_FLIQ) : 1.0;
    sat_values_pair<SATFUNC>(spa, spb, a.W, b.W, fla, flb, sa, sb);
#ifdef VSFM2_PROFILE
    const long long pt1b = clock64(); pt_curves += pt1b - pt1 + (long long)(1e-300 * (sa.kr + sb.kr));
#endif
    double dena, ddena, denb, ddenb, Ga, Gb, G_bcflux[NBC];
    density_fixedT_x<HAS_BC>(A.dtab, a.W, dena, ddena);
    density_fixedT_x<HAS_BC>(A.dtab, b.W, denb, ddenb);
    {
      const double acc_a = PA(PI_POR) * dena * sa.sat * PA(PI_VOL) * dtInv;       // Accum (:1626-1630)
      const double acc_b = PB(PI_POR) * denb * sb.sat * PB(PI_VOL) * dtInv;
      if (phase == PH_INIT) { PA(PI_ACCP) = acc_a; PB(PI_ACCP) = acc_b; }   // PreSolve: accumulation at soln_prev (== W here)
      const double Wn = __shfl_down_sync(FULL, a.W, 1, LPC), krn = __shfl_down_sync(FULL, sa.kr, 1, LPC), denn = __shfl_down_sync(FULL, dena, 1, LPC);
      const double fa_ = rich_flux(a.W, sa.kr, dena, b.W, sb.kr, denb, PA(PI_UPW), PA(PI_DQ), PA(PI_GFAC), area);
      const double fb_ = rich_flux(b.W, sb.kr, denb, Wn, krn, denn, PB(PI_UPW), PB(PI_DQ), PB(PI_GFAC), area);
      const double flux_a = a.has_conn ? fa_ : 0.0, flux_b = b.has_conn ? fb_ : 0.0;        // masked, not branched
      const double flux_p = __shfl_up_sync(FULL, flux_b, 1, LPC);
      Ga = acc_a - ((phase == PH_INIT) ? acc_a : PA(PI_ACCP));
      if (l > 0) Ga = Ga + flux_p;                                      // ff(dn) += flux  (:1806)
      Ga = Ga - flux_a;                                                 // ff(up) -= flux  (:1805)
      Gb = acc_b - ((phase == PH_INIT) ? acc_b : PB(PI_ACCP));
      Gb = Gb + flux_a;
      Gb = Gb - flux_b;
#pragma unroll
      for (int k = 0; k < NBC; ++k) {
        G_bcflux[k] = 0.0;
        if (HAS_BC && bcOwn[k]) {                                       // boundary connection, upweight = 0 (:262-264)
          const bool onA = (bcOwn[k] == 1);
          const double Wc = onA ? a.W : b.W, krc = onA ? sa.kr : sb.kr, denc = onA ? dena : denb;
          double dphi = bcP[k] - Wc + denc * bcGfac[k];
          if ((A.bc[k].itype == CT_SEEPAGE) && (dphi > 0.0) && (bcP[k] <= PRESSURE_REF)) dphi = 0.0;
          const double ukvr = ((dphi >= 0.0) ? bcKr[k] : krc) * RVIS;
          const double fl = ((-bcDq[k] * ukvr * dphi) * area) * denc;
          if (onA) Ga = Ga + fl; else Gb = Gb + fl;
          G_bcflux[k] = fl * FMWH2O;
        }
      }
      Ga = Ga - PA(PI_SRC); Gb = Gb - PB(PI_SRC);
      if (HAS_BC && A.dr_type) {
        double rate, dj;
        if (dr_a) { downreg_sink(A.dr_type, A.dr_value[dr_ia], A.dr_pc[dr_ia], A.dr_n[dr_ia], a.W, rate, dj); Ga = Ga - rate * RFMW; }
        if (dr_b) { downreg_sink(A.dr_type, A.dr_value[dr_ib], A.dr_pc[dr_ib], A.dr_n[dr_ib], b.W, rate, dj); Gb = Gb - rate * RFMW; }
      }
      if (!a.valid) Ga = 0.0;
      if (!b.valid) Gb = 0.0;
    }
    // derivative terms of the curves for this point: independent of the accept / reject decision, computed here so that their
    // two reciprocal chains run in the shadow of the flux + norm reductions (discarded for rejected line-search trial points)
    double dsat_a, dkr_a, dsat_b, dkr_b;
    sat_derivs_pair<SATFUNC>(spa, spb, sa, sb, fla, flb, dsat_a, dkr_a, dsat_b, dkr_b);
#ifdef VSFM2_PROFILE
    const long long pt1c = clock64() + (long long)(1e-300 * (Ga + Gb));
#endif
    const double g2 = col_sum<LPC>(Ga * Ga + Gb * Gb);
    const double w2 = col_sum<LPC>((a.valid ? a.W * a.W : 0.0) + (b.valid ? b.W * b.W : 0.0));
    nfuncs += 1;

#ifdef VSFM2_PROFILE
    const long long pt2 = clock64() + (long long)(1e-300 * (g2 + w2)); pt_eval += pt2 - pt1; pn_eval++; pt_red += pt2 - pt1c;
#endif
    // ================= after the evaluation: line-search / convergence logic (squared norms) =================
    // Fast path first (PETSc's order of tests, restated as predicates + selects so that the common outcome -- trial point
    // accepted, Newton continues or converges -- costs one branch): `take` = adopt W as the new iterate and its aux vars.
    const bool g_bad = !(g2 == g2) || (g2 > 1.7e308);                   // NaN or Inf
    const bool out_of_funcs = (nfuncs >= so.max_funcs && so.max_funcs >= 0);
    const bool tiny_step = (stol2 * x2 > y2);                           // stol * xnorm > ynorm
    const bool is_init = (phase == PH_INIT), is_full = (phase == PH_LS_FULL), is_bt = (phase == PH_LS_QUAD || phase == PH_LS_CUBIC);
    ls_count += (phase == PH_LS_CUBIC) ? 1 : 0;                         // cubic trial points evaluated so far
    const double ls_rhs = .5 * f2 + lambda * so.ls_alpha * initslope;   // sufficient decrease: <= for the full step, < afterwards
    const bool suff = !g_bad && (is_full ? (.5 * g2 <= ls_rhs) : (.5 * g2 < ls_rhs));
    // (PETSc leaves the cubic loop after max_its fits and keeps the last point)
    const bool take = is_init || (EVAL && phase == PH_VEVAL) || (is_full && suff) || (is_bt && !g_bad && (suff || ls_count >= so.ls_max_its));

    if (take) {
      // "copy the solution over": X <- W, F <- G; the aux vars of this point feed the next Jacobian / PostSolve
      a.X = a.W; b.X = b.W; PA(PI_F) = Ga; PB(PI_F) = Gb;
      PA(PI_KR) = sa.kr; PA(PI_SAT) = sa.sat; PB(PI_KR) = sb.kr; PB(PI_SAT) = sb.sat;
      PA(PI_DSAT) = dsat_a; PA(PI_DKR) = dkr_a; PB(PI_DSAT) = dsat_b; PB(PI_DKR) = dkr_b;
      if (HAS_BC) {
#pragma unroll
        for (int k = 0; k < NBC; ++k) bcFlux[k] = G_bcflux[k];
      }
      f2 = g2; x2 = w2;
      // SNESConvergedDefault: it == 0 sets ttol = fnorm * rtol and only tests NaN / atol; it > 0 tests, in this order,
      // atol, function count, rtol, stol, divergence, max_it (lowest priority assigned first)
      its = is_init ? 0 : its + 1;
      f2_0 = is_init ? g2 : f2_0;                                       // it == 0: ttol = fnorm * rtol
      int reason = (its >= so.max_it) ? SNES_DIVERGED_MAX_IT : 0;
      reason = (so.divtol > 0 && g2 > divtol2 * f2_0) ? SNES_DIVERGED_DTOL : reason;
      reason = (y2 < stol2 * x2) ? SNES_CONVERGED_SNORM_RELATIVE : reason;
      reason = (g2 <= f2_0 * rtol2) ? SNES_CONVERGED_FNORM_RELATIVE : reason;
      reason = out_of_funcs ? SNES_DIVERGED_FUNCTION_COUNT : reason;
      reason = (g2 < atol2) ? SNES_CONVERGED_FNORM_ABS : reason;
      if (is_init) reason = g_bad ? SNES_DIVERGED_FNORM_NAN : ((g2 < atol2) ? SNES_CONVERGED_FNORM_ABS : 0);
      last_reason = reason ? reason : last_reason;
      phase = reason ? -1 : PH_NEWTON;
      if (EVAL) {                                     // probe: x_in -> eval_x -> Newton set-up, whatever the norms say
        if (is_init) { if (a.valid) a.W = A.eval_x[cell0]; if (b.valid) b.W = A.eval_x[cell0 + 1]; phase = PH_VEVAL; }
        else phase = PH_NEWTON;
      }
    } else if (is_full) {
      if (g_bad) {
        if (lambda <= so.ls_minlambda) { last_reason = SNES_DIVERGED_FNORM_NAN; phase = -1; }
        else if (out_of_funcs)         { last_reason = SNES_DIVERGED_FUNCTION_COUNT; phase = -1; }
        else { lambda = .5 * lambda; a.W = fma(-lambda, PA(PI_Y), a.X); b.W = fma(-lambda, PB(PI_Y), b.X); }
      } else if (tiny_step) {
        // "full step didn't work and the step is tiny": line search fails, SNES then sees stol*xnorm > ynorm
        last_reason = SNES_CONVERGED_SNORM_RELATIVE; phase = -1;
      } else if (out_of_funcs) {
        last_reason = SNES_DIVERGED_FUNCTION_COUNT; phase = -1;
      } else {
        double lt = -initslope / (g2 - f2 - 2.0 * lambda * initslope);  // quadratic fit
        SC_LAMPREV = lambda; SC_GPREV = g2;
        if (lt > .5 * lambda) lt = .5 * lambda;
        lambda = (lt <= .1 * lambda) ? .1 * lambda : lt;
        a.W = fma(-lambda, PA(PI_Y), a.X); b.W = fma(-lambda, PB(PI_Y), b.X); phase = PH_LS_QUAD; ls_count = 0;
      }
    } else if (is_bt) {
      const int ls_fail = tiny_step ? SNES_CONVERGED_SNORM_RELATIVE : SNES_DIVERGED_LINE_SEARCH;
      if (g_bad) {
        last_reason = ls_fail; phase = -1;
      } else if (lambda <= so.ls_minlambda) {
        last_reason = ls_fail; phase = -1;
      } else if (out_of_funcs) {
        last_reason = SNES_DIVERGED_FUNCTION_COUNT; phase = -1;
      } else {
        const double t1 = .5 * (g2 - f2) - lambda * initslope;           // cubic fit
        const double lambdaprev = SC_LAMPREV, gprev = SC_GPREV;
        const double t2 = .5 * (gprev - f2) - lambdaprev * initslope;
        const double rl2 = __drcp_rn(lambda * lambda), rp2 = __drcp_rn(lambdaprev * lambdaprev), rd = __drcp_rn(lambda - lambdaprev);
        const double ca = (t1 * rl2 - t2 * rp2) * rd;
        const double cb = (-lambdaprev * t1 * rl2 + lambda * t2 * rp2) * rd;
        double d = cb * cb - 3 * ca * initslope;
        if (d < 0.0) d = 0.0;
        double lt = (ca == 0.0) ? -initslope / (2.0 * cb) : (-cb + sqrt(d)) / (3.0 * ca);
        SC_LAMPREV = lambda; SC_GPREV = g2;
        if (lt > .5 * lambda) lt = .5 * lambda;
        lambda = (lt <= .1 * lambda) ? .1 * lambda : lt;
        a.W = fma(-lambda, PA(PI_Y), a.X); b.W = fma(-lambda, PB(PI_Y), b.X); phase = PH_LS_CUBIC;
      }
    }
#ifdef VSFM2_PROFILE
    pt_logic += clock64() - pt2;
#endif
  }
#ifdef VSFM2_PROFILE
  if (A.prof && lane == 0) {
    atomicAdd((unsigned long long *)A.prof + 0, (unsigned long long)pt_newton); atomicAdd((unsigned long long *)A.prof + 1, (unsigned long long)pn_newton);
    atomicAdd((unsigned long long *)A.prof + 2, (unsigned long long)pt_eval);   atomicAdd((unsigned long long *)A.prof + 3, (unsigned long long)pn_eval);
    atomicAdd((unsigned long long *)A.prof + 4, (unsigned long long)pt_logic);  atomicAdd((unsigned long long *)A.prof + 5, (unsigned long long)(clock64() - pt_start));
    atomicAdd((unsigned long long *)A.prof + 6, 1ull);
    atomicAdd((unsigned long long *)A.prof + 7, (unsigned long long)pt_curves); atomicAdd((unsigned long long *)A.prof + 8, (unsigned long long)pt_red);
    atomicAdd((unsigned long long *)A.prof + 9, (unsigned long long)pt_nasm); atomicAdd((unsigned long long *)A.prof + 10, (unsigned long long)pt_nelim);
    atomicAdd((unsigned long long *)A.prof + 11, (unsigned long long)pt_npcr);
  }
#endif

  // ---- VSFMSOEPostSolve -> SetDataInSOEAuxVar (GoveqnRichards...:1170-1195) ---------------------------------
  const bool leader = col_ok && (l == 0);
  const double m_beg = leader ? A.col_mass[col] : 0.0;        // issued here so that the round trip hides behind the mailbox stores
  double mass = 0.0;
  if (a.valid) {
    A.x_out[cell0] = a.X;
    if (converged) {
      double den, dden; density_fixedT_x<HAS_BC>(A.dtab, a.X, den, dden);
      const double sat = PA(PI_SAT);
      const double m = PA(PI_POR) * den * FMWH2O * sat * PA(PI_VOL);
      A.liq_sat[cell0] = sat; A.pressure[cell0] = a.X; A.mass[cell0] = m;
      A.smp[cell0] = (a.X - PRESSURE_REF) * rcp(den * FMWH2O * GRAVITY_CONSTANT);
      mass += m;
    }
  }
  if (b.valid) {
    A.x_out[cell0 + 1] = b.X;
    if (converged) {
      double den, dden; density_fixedT_x<HAS_BC>(A.dtab, b.X, den, dden);
      const double sat = PB(PI_SAT);
      const double m = PB(PI_POR) * den * FMWH2O * sat * PB(PI_VOL);
      A.liq_sat[cell0 + 1] = sat; A.pressure[cell0 + 1] = b.X; A.mass[cell0 + 1] = m;
      A.smp[cell0 + 1] = (b.X - PRESSURE_REF) * rcp(den * FMWH2O * GRAVITY_CONSTANT);
      mass += m;
    }
  }
  double bc_exc = 0.0;
  if (HAS_BC && converged) {
#pragma unroll
    for (int k = 0; k < NBC; ++k) if (bcOwn[k]) {
      A.bc[k].flux[col] = bcFlux[k];
      A.bc[k].mass_exc[col] += bcMassExc[k];
      bc_exc += bcMassExc[k];
    }
  }
  if (HAS_BC && A.dr_type) {                                 // the rate actually withdrawn at the end-of-step state enters the column's balance
    double rate, dj;
    if (dr_a) { downreg_sink(A.dr_type, A.dr_value[dr_ia], A.dr_pc[dr_ia], A.dr_n[dr_ia], a.X, rate, dj); src_kg += rate; }
    if (dr_b) { downreg_sink(A.dr_type, A.dr_value[dr_ib], A.dr_pc[dr_ib], A.dr_n[dr_ib], b.X, rate, dj); src_kg += rate; }
  }
  const double m_end = col_sum<LPC>(mass);
  const double q_col = col_sum<LPC>(src_kg);
  if (HAS_BC) bc_exc = col_sum<LPC>(bc_exc);
  double err = 0.0;
  if (leader) {
    A.stat_its[col] = tot_its; A.stat_reason[col] = last_reason; A.stat_cuts[col] = cuts; A.stat_nf[col] = tot_nf;
    if (converged) {
      err = fabs(m_beg - m_end + q_col * A.dt);
      A.col_mass[col] = m_end;
    }
    A.col_err[col] = err; A.col_src[col] = q_col;
    if (A.t_done) A.t_done[col] = SC_TDONE;
  }

  // ---- block partials for the global mass-balance / convergence reductions (deterministic order) ----------
  // only the column leaders (lanes 0, LPC, 2 LPC, ...) carry values.  Sums: log2(32 / LPC) shuffle stages in a fixed order.  Maxima and
  // the worst reason: integer warp reductions (REDUX); the mass error is non-negative, so its IEEE bit pattern orders like an integer.
  double v[4];
  v[0] = leader ? m_beg : 0.0;                                  // sum mass before
  v[1] = leader ? (converged ? m_end : m_beg) : 0.0;            // sum mass after
  v[2] = leader ? q_col * A.dt : 0.0;                           // sum sources * dt
  v[3] = leader ? bc_exc : 0.0;                                 // sum boundary mass exchanged
#pragma unroll
  for (int s = 16; s >= LPC; s >>= 1) {
#pragma unroll
    for (int k = 0; k < 4; ++k) v[k] += __shfl_xor_sync(FULL, v[k], s);
  }
  const unsigned ehi = leader ? (unsigned)__double2hiint(err) : 0u, elo = leader ? (unsigned)__double2loint(err) : 0u;
  const unsigned mhi = __reduce_max_sync(FULL, ehi);
  const unsigned mlo = __reduce_max_sync(FULL, (ehi == mhi) ? elo : 0u);
  const int mits = __reduce_max_sync(FULL, leader ? tot_its : 0);
  const int mdiv = __reduce_max_sync(FULL, (leader && !converged) ? 1 : 0);
  const int mcut = __reduce_max_sync(FULL, leader ? cuts : 0);
  const int worst = __reduce_min_sync(FULL, leader ? last_reason : 0x7fffffff);
  static_assert(VSFM2_THREADS == 32, "one warp per block: the block partials are the warp's");
  if (lane == 0) {
    double *bp = A.block_partials + (size_t)blockIdx.x * 9;
    bp[0] = v[0]; bp[1] = v[1]; bp[2] = v[2]; bp[3] = v[3];
    bp[4] = __hiloint2double((int)mhi, (int)mlo);               // max |mass error|
    bp[5] = (double)mits;                                       // max Newton its
    bp[6] = (double)mdiv;                                       // any diverged
    bp[7] = (double)mcut;                                       // max dt cuts
    bp[8] = (double)worst;
  }
#undef PA
#undef PB
#undef DT_TOT
#undef rtol2
#undef stol2
#undef SC_TDONE
#undef SC_LAMPREV
#undef SC_GPREV
}

}  // namespace mpp
