// th_kernels2.cuh -- fused coupled thermal-hydrology (TH) time step for columns of up to 16 layers: one lane per cell,
// 16 lanes per column, two columns per warp, state in registers, accepted-point aux vars in shared memory.
//
// Same step, same reference citations and the same device functions (th_cell_compute, th_rich_flux) as the generic
// one-warp-per-column kernel in th_kernels.cuh, which stays the correctness path for tall columns (the reference's TH
// goldens have 100 and 20 cells).  What changes is the mapping (DESIGN.md "TH kernel"):
//   * the aux vars of both governing equations at the accepted iterate (18 doubles per cell) are parked in shared memory,
//     where the Jacobian also reads the next cell's copy; residual evaluations reach the neighbour with 7 warp shuffles;
//   * the connection j -> j+1 is owned by lane j, which evaluates its four 2x2 derivative blocks once and hands the
//     down-row half to lane j+1 with 8 shuffles;
//   * the 2x2 block-tridiagonal Newton system is solved by block parallel cyclic reduction across the 16 lanes, rows kept
//     normalised (unit diagonal block): 10 shuffled doubles per side per stage, 4 stages;
//   * the line-search slope F.(J Y) is taken as -||F||^2: the block solve is direct, so J Y = F to round-off (the
//     reference's GMRES+ILU(0) is inexact there anyway, see th_kernels.cuh);
//   * all norms are carried squared and are bitwise identical in the 16 lanes of a column, so control flow is uniform
//     per column; the Newton set-up is entered on __any_sync so every shuffle is convergent with a full mask.
#pragma once
#include "th_kernels.cuh"

namespace mpp {

// One warp per block (two columns): a warp whose columns are done retires at once and frees its registers / shared memory for the
// next block instead of idling at the block's final barrier behind a column that is cutting dt.
#ifndef TH2_THREADS
#define TH2_THREADS 32
#endif
#ifndef TH2_MIN_BLOCKS
#define TH2_MIN_BLOCKS (384 / TH2_THREADS)
#endif
#ifndef TH2_MIN_BLOCKS_PAD
#define TH2_MIN_BLOCKS_PAD TH2_MIN_BLOCKS
#endif

struct M2 { double a, b, c, d; };     // row-major 2x2: [a b; c d]
__device__ __forceinline__ M2 m2mul(const M2 &x, const M2 &y) { return M2{x.a * y.a + x.b * y.c, x.a * y.b + x.b * y.d, x.c * y.a + x.d * y.c, x.c * y.b + x.d * y.d}; }
__device__ __forceinline__ M2 m2inv(const M2 &m) { const double r = rcp(m.a * m.d - m.b * m.c); return M2{m.d * r, -m.b * r, -m.c * r, m.a * r}; }
template <int G> __device__ __forceinline__ M2 m2up(const M2 &m, int s) { return M2{__shfl_up_sync(FULL_MASK, m.a, s, G), __shfl_up_sync(FULL_MASK, m.b, s, G), __shfl_up_sync(FULL_MASK, m.c, s, G), __shfl_up_sync(FULL_MASK, m.d, s, G)}; }
template <int G> __device__ __forceinline__ M2 m2dn(const M2 &m, int s) { return M2{__shfl_down_sync(FULL_MASK, m.a, s, G), __shfl_down_sync(FULL_MASK, m.b, s, G), __shfl_down_sync(FULL_MASK, m.c, s, G), __shfl_down_sync(FULL_MASK, m.d, s, G)}; }

template <int G>
__device__ __forceinline__ double grp_sum(double v)
{
#pragma unroll
  for (int s = G / 2; s > 0; s >>= 1) v += __shfl_xor_sync(FULL_MASK, v, s, G);
  return v;
}

// Block PCR: A y[l-1] + B y[l] + C y[l+1] = (d0, d1), one block row per lane; identity rows pad the group.
// TH2_PCR_ROLLED=1 runs the stages before the last as ONE loop body with a run-time shuffle distance (-240 instructions of a hot loop
// that is larger than the 32 KB second-level instruction cache).  Measured SLOWER on 1 Mi columns (6.46 vs 6.15 ms per step): the
// unrolled stages overlap their shuffles with the previous stage's arithmetic, and that is worth more than the fetch misses.
#ifndef TH2_PCR_ROLLED
#define TH2_PCR_ROLLED 0
#endif
template <int G>
__device__ __forceinline__ void block_pcr(M2 A, M2 B, M2 C, double d0, double d1, double &y0, double &y1)
{
  M2 Bi = m2inv(B);
  A = m2mul(Bi, A); C = m2mul(Bi, C);
  double e0 = Bi.a * d0 + Bi.b * d1, e1 = Bi.c * d0 + Bi.d * d1;
#if TH2_PCR_ROLLED
#pragma unroll 1
#else
#pragma unroll
#endif
  for (int s = 1; s < G / 2; s <<= 1) {
    const M2 Am = m2up<G>(A, s), Cm = m2up<G>(C, s), Ap = m2dn<G>(A, s), Cp = m2dn<G>(C, s);
    const double e0m = __shfl_up_sync(FULL_MASK, e0, s, G), e1m = __shfl_up_sync(FULL_MASK, e1, s, G);
    const double e0p = __shfl_down_sync(FULL_MASK, e0, s, G), e1p = __shfl_down_sync(FULL_MASK, e1, s, G);
    const M2 ACm = m2mul(A, Cm), CAp = m2mul(C, Ap);
    const M2 Bn{1.0 - ACm.a - CAp.a, -ACm.b - CAp.b, -ACm.c - CAp.c, 1.0 - ACm.d - CAp.d};
    const double f0 = e0 - (A.a * e0m + A.b * e1m) - (C.a * e0p + C.b * e1p);
    const double f1 = e1 - (A.c * e0m + A.d * e1m) - (C.c * e0p + C.d * e1p);
    const M2 An = m2mul(A, Am), Cn = m2mul(C, Cp);
    Bi = m2inv(Bn);
    A = m2mul(Bi, M2{-An.a, -An.b, -An.c, -An.d}); C = m2mul(Bi, M2{-Cn.a, -Cn.b, -Cn.c, -Cn.d});
    e0 = Bi.a * f0 + Bi.b * f1; e1 = Bi.c * f0 + Bi.d * f1;
  }
  {
    // last stage (distance G/2): rows l and l +- G/2 decouple from everything else; only the right-hand side is carried on
    constexpr int s = G / 2;
    const M2 Cm = m2up<G>(C, s), Ap = m2dn<G>(A, s);
    const double e0m = __shfl_up_sync(FULL_MASK, e0, s, G), e1m = __shfl_up_sync(FULL_MASK, e1, s, G);
    const double e0p = __shfl_down_sync(FULL_MASK, e0, s, G), e1p = __shfl_down_sync(FULL_MASK, e1, s, G);
    const M2 ACm = m2mul(A, Cm), CAp = m2mul(C, Ap);
    const M2 Bn{1.0 - ACm.a - CAp.a, -ACm.b - CAp.b, -ACm.c - CAp.c, 1.0 - ACm.d - CAp.d};
    const double f0 = e0 - (A.a * e0m + A.b * e1m) - (C.a * e0p + C.b * e1p);
    const double f1 = e1 - (A.c * e0m + A.d * e1m) - (C.c * e0p + C.d * e1p);
    Bi = m2inv(Bn);
    e0 = Bi.a * f0 + Bi.b * f1; e1 = Bi.c * f0 + Bi.d * f1;
  }
  y0 = e0; y1 = e1;
}

__device__ __forceinline__ void ax_store(double (*s)[TH2_THREADS], int t, const THCell &c)
{
  s[0][t] = c.sat; s[1][t] = c.kr; s[2][t] = c.dsat; s[3][t] = c.dkr; s[4][t] = c.den_m; s[5][t] = c.ddenP_m; s[6][t] = c.ddenT_m;
  s[7][t] = c.den_e; s[8][t] = c.ddenP_e; s[9][t] = c.ddenT_e; s[10][t] = c.ul; s[11][t] = c.hl; s[12][t] = c.dulT; s[13][t] = c.dhlT;
  s[14][t] = c.dulP; s[15][t] = c.dhlP; s[16][t] = c.tc; s[17][t] = c.dtcP;
}
__device__ __forceinline__ void ax_load(double (*s)[TH2_THREADS], int t, THCell &c)
{
  c.sat = s[0][t]; c.kr = s[1][t]; c.dsat = s[2][t]; c.dkr = s[3][t]; c.den_m = s[4][t]; c.ddenP_m = s[5][t]; c.ddenT_m = s[6][t];
  c.den_e = s[7][t]; c.ddenP_e = s[8][t]; c.ddenT_e = s[9][t]; c.ul = s[10][t]; c.hl = s[11][t]; c.dulT = s[12][t]; c.dhlT = s[13][t];
  c.dulP = s[14][t]; c.dhlP = s[15][t]; c.tc = s[16][t]; c.dtcP = s[17][t];
}

__device__ __forceinline__ SatParams th_load_sp(double (*s)[TH2_THREADS], int t)
{
  SatParams p;
  p.sat_res = s[14][t]; p.alpha = s[15][t]; p.m = s[16][t]; p.n = s[17][t]; p.pu = s[18][t]; p.ps = s[19][t]; p.b2 = s[20][t]; p.b3 = s[21][t];
  return p;
}

// PADBC: the boundary connection is KNOWN to ride on the padding lane (the host launches this instance only when A.bc_on_pad_lane is set): the
// per-lane boundary-condition loops and their local-memory array are compiled out, which takes ~5 KB out of a hot loop that is larger than the
// 32 KB second-level instruction cache.  PADBC = false keeps both paths behind the run-time flag.
template <int G, int SF, int DT, int IEE, bool PADBC = false>
__global__ void __launch_bounds__(TH2_THREADS, PADBC ? TH2_MIN_BLOCKS_PAD : TH2_MIN_BLOCKS)
th_step2_kernel(const THArgs A)
{
  const bool bc_on_pad = PADBC || (A.bc_on_pad_lane != 0);
  constexpr unsigned FULL = FULL_MASK;
  // energy-equation aux vars keep the default permeability (ThermalEnthalpySoilAuxType.F90:93) unless the driver set its own
  // (goveq_enthalpy%SetSoilPermeability -> mppgpu_th_set_energy_permeability)
  constexpr double PERM_E_DEFAULT = 8.3913e-12;
  const int tid = blockIdx.x * blockDim.x + threadIdx.x;
  const int slot = tid / G, j = tid % G, lane = threadIdx.x & 31;
  const int nlev = A.nlev;
  const bool col_ok = slot < A.ncol;
  // the warp's two columns: batch order, or grouped by the cost of their previous StepDT (mppgpu_set_column_ordering; as vsfm_step2_kernel)
  const int col = (col_ok && A.order) ? A.order[slot] : slot;
  const bool valid = col_ok && j < nlev;
  // Boundary connection on the padding lane (A.bc_on_pad_lane: one Dirichlet temperature at the top, nlev <= 15, top cell first).  A
  // Dirichlet face is an internal connection with dist_up = 0 whose up side is the boundary aux var (ThermalEnthalpyFlux,
  // ThermalEnthalpyMod.F90:27-165; upweight 0, RichardsMod.F90:262-264).  Lane G-1 has no cell: it carries that aux var as its "cell"
  // (state = the condition's T and poked P, static data = the top cell's) and owns the connection boundary -> cell 0, so the boundary
  // flux and its four derivative blocks are computed by the same instructions, at the same time, as every interior connection --
  // instead of a loop that one lane in sixteen runs alone out of a local-memory array.
  const bool pad = bc_on_pad && col_ok && (j == G - 1);
  const bool has_conn = (valid && j < nlev - 1) || pad;
  const long long cell = (long long)col * nlev + (pad ? 0 : j);      // the padding lane reads the top cell's static data
  const int src_dn = pad ? lane - (G - 1) : ((j == G - 1) ? lane : lane + 1);                      // lane of the connection's dn cell
  const int src_up = (bc_on_pad && j == 0) ? lane + (G - 1) : ((j == 0) ? lane : lane - 1);   // lane that owns the connection above
  const bool has_up = (j > 0) || bc_on_pad;
  const int jtop = A.top_is_first ? 0 : nlev - 1, jbot = A.top_is_first ? nlev - 1 : 0;
  const SnesOpts so = A.so;

  // ---- static per-cell data ------------------------------------------------------------------------------------------
  SatParams sp; sp.sat_res = 0.0; sp.alpha = 1.0; sp.m = 0.5; sp.n = 2.0; sp.pu = sp.ps = sp.b2 = sp.b3 = 0.0;
  double por = 0.5, perm = 1.0, dz = 1.0, area = 1.0, tkdry = 1.0, csol = 1.0, P = PRESSURE_REF, T = 283.15, srcm = 0.0, srce = 0.0, perm_e = PERM_E_DEFAULT;
  if (valid || pad) {
    por = A.por[cell]; perm = A.perm[cell]; dz = A.dz[cell]; area = A.area[col]; tkdry = A.tkdry[cell]; csol = A.csol[cell];
    if (A.perm_e) perm_e = A.perm_e[cell];
    sp.sat_res = A.sat_res[cell]; sp.alpha = A.alpha[cell]; sp.m = A.lam[cell]; sp.n = A.vgn ? A.vgn[cell] : 0.0;
    if (A.pu) { sp.pu = A.pu[cell]; sp.ps = A.ps[cell]; sp.b2 = A.b2[cell]; sp.b3 = A.b3[cell]; }
    if (pad) { T = A.bc[0].value[col]; P = A.bc[0].bc_pressure ? A.bc[0].bc_pressure[col] : 0.0; }     // energy-equation boundary aux var
    else     { P = A.x_in[2 * cell]; T = A.x_in[2 * cell + 1]; }
    for (int k = 0; k < (pad ? 0 : A.nss); ++k) {
      const THCondDev &c = A.ss[k];
      double val = 0.0; bool mine = false;
      if (c.region == REGION_CELLS) { val = c.value[cell]; mine = true; }
      else if (j == (c.region == REGION_TOP ? jtop : jbot)) { val = c.value[col]; mine = true; }
      if (mine) { if (c.ieqn == 1) srcm += val / FMWH2O; else srce += val; }
    }
  }
  const double vol = area * dz;
  // connection j -> j+1 (owned by lane j)
  const double perm_d = __shfl_sync(FULL, perm, src_dn), dz_d = __shfl_sync(FULL, dz, src_dn);
  const double perm_e_d = __shfl_sync(FULL, perm_e, src_dn);
  double dist_up = 0.5 * dz, dist_dn = 0.5 * dz_d, upw = dist_up / (dist_up + dist_dn);
  double gfac = FMWH2O * ((dist_up + dist_dn) * (A.uz * (-GRAVITY_CONSTANT)));
  double Dqm = (perm * perm_d) / (dist_up * perm_d + dist_dn * perm);
  double Dqe = (perm_e * perm_e_d) / (dist_up * perm_e_d + dist_dn * perm_e);
  if (pad) {
    // boundary connection (MeshType.F90:723-806): dist_up = 0, dist_dn = dz/2 of the top cell, unit vector (0,0,-1); no mass-equation
    // boundary exists in this configuration, so the Richards flux of the mass equation through this face is switched off
    dist_up = 0.0; dist_dn = 0.5 * dz; upw = 0.0;
    gfac = FMWH2O * ((0.0 + 0.5 * dz) * (((A.uz == 0.0) ? 0.0 : -1.0) * (-GRAVITY_CONSTANT)));
    Dqm = 0.0; Dqe = perm_e / (0.0 + 0.5 * dz);
    por = 0.0; csol = 0.0;
  }

  // boundary conditions owned by this lane (at most one per region and equation)
  struct BCL { int ieqn; double P, T, bgf, Dq; FluxIn fin; double hl, tc; };
  BCL bcs[4]; int nmybc = 0;
  for (int k = 0; k < ((PADBC || bc_on_pad) ? 0 : A.nbc); ++k) {
    const bool top = (A.bc[k].region == REGION_TOP);
    if (!valid || j != (top ? jtop : jbot) || nmybc >= 4) continue;
    BCL &b = bcs[nmybc++];
    b.ieqn = A.bc[k].ieqn;
    const double uzbc = (A.uz == 0.0) ? 0.0 : (top ? -1.0 : 1.0);
    b.bgf = FMWH2O * ((0.0 + 0.5 * dz) * (uzbc * (-GRAVITY_CONSTANT)));
    THCell c;
    if (b.ieqn == 1) {       // mass equation: pressure = condition value, temperature stays at its default 298.15 K
      b.P = A.bc[k].value[col]; b.T = 273.15 + 25.0; b.Dq = perm / (0.0 + 0.5 * dz);
      th_cell_compute<SF, DT, IEE>(A, sp, tkdry, b.P, b.T, c);
      b.fin = FluxIn{b.P, c.kr, c.dkr, c.den_m, c.ddenP_m, c.ddenT_m}; b.hl = 0.0; b.tc = 0.0;
    } else {                 // energy equation: temperature = condition value; pressure as poked by the driver (default 0)
      b.T = A.bc[k].value[col]; b.P = A.bc[k].bc_pressure ? A.bc[k].bc_pressure[col] : 0.0; b.Dq = perm_e / (0.0 + 0.5 * dz);
      th_cell_compute<SF, DT, IEE>(A, sp, tkdry, b.P, b.T, c);
      b.fin = FluxIn{b.P, c.kr, c.dkr, c.den_e, c.ddenP_e, c.ddenT_e}; b.hl = c.hl; b.tc = c.tc;
    }
  }

  // Static per-cell data goes to shared memory ([field][thread]) and is re-read where used: ~45 registers less per lane, i.e. a
  // third (and fourth) block per SM for this latency-bound kernel.  The macros below shadow the set-up variables from here on.
  __shared__ double s_tp[22][TH2_THREADS];
  {
    const int t = threadIdx.x;
    s_tp[0][t] = por; s_tp[1][t] = vol; s_tp[2][t] = csol; s_tp[3][t] = tkdry; s_tp[4][t] = srcm; s_tp[5][t] = srce; s_tp[6][t] = upw;
    s_tp[7][t] = Dqm; s_tp[8][t] = Dqe; s_tp[9][t] = gfac; s_tp[10][t] = dist_up; s_tp[11][t] = dist_dn; s_tp[12][t] = area; s_tp[13][t] = dz;
    s_tp[14][t] = sp.sat_res; s_tp[15][t] = sp.alpha; s_tp[16][t] = sp.m; s_tp[17][t] = sp.n;
    s_tp[18][t] = sp.pu; s_tp[19][t] = sp.ps; s_tp[20][t] = sp.b2; s_tp[21][t] = sp.b3;
  }
#define por     s_tp[0][threadIdx.x]
#define vol     s_tp[1][threadIdx.x]
#define csol    s_tp[2][threadIdx.x]
#define tkdry   s_tp[3][threadIdx.x]
#define srcm    s_tp[4][threadIdx.x]
#define srce    s_tp[5][threadIdx.x]
#define upw     s_tp[6][threadIdx.x]
#define Dqm     s_tp[7][threadIdx.x]
#define Dqe     s_tp[8][threadIdx.x]
#define gfac    s_tp[9][threadIdx.x]
#define dist_up s_tp[10][threadIdx.x]
#define dist_dn s_tp[11][threadIdx.x]
#define area    s_tp[12][threadIdx.x]
#define dz      s_tp[13][threadIdx.x]
#define sp      th_load_sp(s_tp, threadIdx.x)

  // ---- time-step / Newton state (uniform per column) -------------------------------------------------------------------
  const double atol2 = so.atol * so.atol, rtol2 = so.rtol * so.rtol, stol2 = so.stol * so.stol;
  const double divtol2 = so.divtol * so.divtol, maxstep2 = so.ls_maxstep * so.ls_maxstep;
  double Pp = P, Tp = T, Wm = P, We = T, Fm = 0.0, Fe = 0.0, Ym = 0.0, Ye = 0.0, accm = 0.0, acce = 0.0;
  // aux vars of both equations at the accepted iterate: 18 doubles per cell parked in shared memory ([field][thread],
  // conflict-free); the Jacobian reads the cell's own copy and the next cell's (thread + 1) without any shuffle
  __shared__ double s_ax[18][TH2_THREADS];
  {
    THCell z;
    z.sat = 1.0; z.kr = 1.0; z.dsat = 0.0; z.dkr = 0.0; z.den_m = 55.0; z.ddenP_m = 0.0; z.ddenT_m = 0.0; z.den_e = 55.0; z.ddenP_e = 0.0;
    z.ddenT_e = 0.0; z.ul = 0.0; z.hl = 0.0; z.dulT = 0.0; z.dhlT = 0.0; z.dulP = 0.0; z.dhlP = 0.0; z.tc = 1.0; z.dtcP = 0.0;
    ax_store(s_ax, threadIdx.x, z);
  }
  double dt_iter = A.dt, dtInv = 1.0 / A.dt, time_done = 0.0;
  int cuts = 0, tot_its = 0, tot_nf = 0, last_reason = 0, converged = 0;
  int phase = col_ok ? PH_INIT : PH_DONE, its = 0, nfuncs = 0, ls_count = 0;
  double f2 = 0.0, x2 = 0.0, y2 = 0.0, ttol2 = 0.0, f2_0 = 0.0, initslope = -1.0, lambda = 1.0, lambdaprev = 1.0, gprev = 0.0;

  for (;;) {
    // ================= Newton step set-up: Jacobian blocks, block PCR, line-search initialisation =================
    if (__any_sync(FULL, phase == PH_NEWTON || (!PADBC && phase == PH_EVAL_J))) {      // (the PADBC instances never serve mppgpu_eval)
      const bool nw = (phase == PH_NEWTON);
      __syncwarp();                              // aux vars were stored under per-column control flow
      THCell ax, ad;                             // this cell and the dn side of connection j (the next lane's cell)
      ax_load(s_ax, threadIdx.x, ax);
      ax_load(s_ax, has_conn ? threadIdx.x + (src_dn - lane) : threadIdx.x, ad);
      const double Pd = __shfl_sync(FULL, P, src_dn), Td = __shfl_sync(FULL, T, src_dn);
      const double krd = ad.kr, dkrd = ad.dkr, denmd = ad.den_m, dPmd = ad.ddenP_m, dTmd = ad.ddenT_m;
      const double dened = ad.den_e, dPed = ad.ddenP_e, dTed = ad.ddenT_e, hld = ad.hl, dhlTd = ad.dhlT, dhlPd = ad.dhlP;
      const double tcd = ad.tc, dtcPd = ad.dtcP;
      // derivative blocks of connection j -> j+1 (same expressions as th_step_generic_kernel)
      double mJup = 0.0, mJdn = 0.0, dTu = 0.0, dTd = 0.0, JTT_u = 0.0, JTT_d = 0.0, JTP_u = 0.0, JTP_d = 0.0;
      if (has_conn) {
        const FluxIn um = {P, ax.kr, ax.dkr, ax.den_m, ax.ddenP_m, ax.ddenT_m}, dm = {Pd, krd, dkrd, denmd, dPmd, dTmd};
        const FluxIn ue = {P, ax.kr, ax.dkr, ax.den_e, ax.ddenP_e, ax.ddenT_e}, de = {Pd, krd, dkrd, dened, dPed, dTed};
        double fl, mfl, eJup, eJdn, edTu, edTd;
        th_rich_flux(um, dm, upw, Dqm, gfac, area, fl, mJup, mJdn, dTu, dTd);
        th_rich_flux(ue, de, upw, Dqe, gfac, area, mfl, eJup, eJdn, edTu, edTd);
        const double ku = ax.tc, kd = tcd;
        const double kod = (ku * kd) * rcp(dist_up * kd + dist_dn * ku);
        const double h = (mfl <= 0.0) ? ax.hl : hld;
        const double dhT_u = (mfl < 0.0) ? ax.dhlT : 0.0, dhT_d = (mfl < 0.0) ? 0.0 : dhlTd;
        const double dhP_u = (mfl < 0.0) ? ax.dhlP : 0.0, dhP_d = (mfl < 0.0) ? 0.0 : dhlPd;
        JTT_u = edTu * h + mfl * dhT_u + (-kod * area); JTT_d = edTd * h + mfl * dhT_d + (+kod * area);
        const double rku = kod * rcp(ku), rkd = kod * rcp(kd);
        const double dDk_u = rku * rku * dist_up * ax.dtcP, dDk_d = rkd * rkd * dist_dn * dtcPd;
        const double dTud = T - Td;
        JTP_u = (-eJup) * h + mfl * dhP_u + (-dDk_u * dTud * area); JTP_d = (-eJdn) * h + mfl * dhP_d + (-dDk_d * dTud * area);
      }
      // this cell as "up" of connection j, as "dn" of connection j-1 (values handed down by lane j-1)
      M2 Ja{0.0, 0.0, 0.0, 0.0}, Jb{1.0, 0.0, 0.0, 1.0}, Jc{0.0, 0.0, 0.0, 0.0};
      const double p_mJup = __shfl_sync(FULL, mJup, src_up), p_mJdn = __shfl_sync(FULL, mJdn, src_up);
      const double p_dTu = __shfl_sync(FULL, dTu, src_up), p_dTd = __shfl_sync(FULL, dTd, src_up);
      const double p_JTTu = __shfl_sync(FULL, JTT_u, src_up), p_JTTd = __shfl_sync(FULL, JTT_d, src_up);
      const double p_JTPu = __shfl_sync(FULL, JTP_u, src_up), p_JTPd = __shfl_sync(FULL, JTP_d, src_up);
      if (valid) {
        double b00 = mJup, b01 = -dTu, b10 = -JTP_u, b11 = -JTT_u;                 // zero where there is no connection j -> j+1
        Jc = M2{mJdn, -dTd, -JTP_d, -JTT_d};
        if (has_up) {
          if (j > 0) Ja = M2{-p_mJup, p_dTu, p_JTPu, p_JTTu};      // (j = 0: the up side is the boundary aux var, not an unknown)
          b00 += -p_mJdn; b01 += p_dTd; b11 += p_JTTd; b10 += p_JTPd;
        }
        for (int k = 0; k < nmybc; ++k) {
          const BCL &b = bcs[k];
          if (b.ieqn == 1) {
            const FluxIn dn = {P, ax.kr, ax.dkr, ax.den_m, ax.ddenP_m, ax.ddenT_m};
            double fl, bJup, bJdn, a1, a2;
            th_rich_flux(b.fin, dn, 0.0, b.Dq, b.bgf, area, fl, bJup, bJdn, a1, a2);
            b00 += -bJdn;
          } else {
            const FluxIn dn = {P, ax.kr, ax.dkr, ax.den_e, ax.ddenP_e, ax.ddenT_e};
            double mfl, eJup, eJdn, edTu, edTd;
            th_rich_flux(b.fin, dn, 0.0, b.Dq, b.bgf, area, mfl, eJup, eJdn, edTu, edTd);
            const double kod = ax.tc / (0.0 + 0.5 * dz);
            const double h = (mfl <= 0.0) ? b.hl : ax.hl;
            const double dhT_d = (mfl < 0.0) ? 0.0 : ax.dhlT, dhP_d = (mfl < 0.0) ? 0.0 : ax.dhlP;
            b11 += edTd * h + mfl * dhT_d + (+kod * area);
            const double dDk_d = 1.0 / (0.0 + 0.5 * dz) * ax.dtcP;
            b10 += (-eJdn) * h + mfl * dhP_d + (-dDk_d * (b.T - T) * area);
          }
        }
        // accumulation derivatives (GoveqnRichards...:1673, 2547; GoveqnThermalEnthalpySoilType.F90:1276-1281, 2146-2153)
        b00 += (por * ax.ddenP_m * ax.sat + por * ax.den_m * ax.dsat) * vol * dtInv;
        b01 += (por * ax.ddenT_m * ax.sat) * vol * dtInv;
        b11 += ((por * ax.ddenT_e * ax.sat * ax.ul + por * ax.den_e * ax.sat * ax.dulT) + (1.0 - por) * 2700.0 * csol) * vol * dtInv;
        b10 += (por * ax.ddenP_e * ax.sat * ax.ul + por * ax.den_e * ax.dsat * ax.ul + por * ax.den_e * ax.sat * ax.dulP) * vol * dtInv;
        Jb = M2{b00, b01, b10, b11};
      }
      if (!PADBC && phase == PH_EVAL_J) {     // kernel unit-test probe: dump the blocks, no solve
        if (valid) {
          double *ja = A.eval_ja + 4 * cell, *jb = A.eval_jb + 4 * cell, *jc = A.eval_jc + 4 * cell;
          ja[0] = Ja.a; ja[1] = Ja.b; ja[2] = Ja.c; ja[3] = Ja.d; jb[0] = Jb.a; jb[1] = Jb.b; jb[2] = Jb.c; jb[3] = Jb.d;
          jc[0] = Jc.a; jc[1] = Jc.b; jc[2] = Jc.c; jc[3] = Jc.d;
        }
        phase = PH_DONE;
      }
      double Yn0, Yn1;
      block_pcr<G>(Ja, Jb, Jc, valid ? Fm : 0.0, valid ? Fe : 0.0, Yn0, Yn1);
      const double yn2 = grp_sum<G>(valid ? Yn0 * Yn0 + Yn1 * Yn1 : 0.0);
      if (nw) {
        Ym = Yn0; Ye = Yn1; y2 = yn2;
        initslope = (f2 > 0.0) ? -f2 : -1.0;     // F.(J Y) with J Y = F (direct block solve), forced negative
        lambda = 1.0; ls_count = 0;
        if (y2 == 0.0) {
          last_reason = (stol2 * x2 > y2) ? SNES_CONVERGED_SNORM_RELATIVE : SNES_DIVERGED_LINE_SEARCH; phase = -1;
        } else {
          if (y2 > maxstep2) { const double sc = so.ls_maxstep / sqrt(y2); Ym *= sc; Ye *= sc; y2 = maxstep2; }
          Wm = fma(-lambda, Ym, P); We = fma(-lambda, Ye, T);
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
        P = Pp; T = Tp;
        if (cuts > 20) { converged = 0; phase = PH_DONE; } else { Wm = P; We = T; phase = PH_INIT; }
      } else {
        converged = 1; time_done += dt_iter; tot_its += its;
        Pp = P; Tp = T;
        if (time_done >= A.dt) phase = PH_DONE; else { Wm = P; We = T; phase = PH_INIT; }
      }
      its = 0; nfuncs = 0;
      // optional give-up budget (mppgpu_set_step_budget; not in the reference, off by default): a column that has burnt this
      // many residual evaluations inside one StepDT fails like one that ran out of dt cuts, instead of stalling the batch
      if (so.step_budget > 0 && tot_nf >= so.step_budget && phase != PH_DONE) { converged = 0; last_reason = SNES_DIVERGED_FUNCTION_COUNT; phase = PH_DONE; }
    }
    if (__all_sync(FULL, phase == PH_DONE)) break;

    // ================= residual evaluation at W (SOETHResidual) =================
    THCell c;
    th_cell_compute<SF, DT, IEE>(A, sp, tkdry, Wm, We, c);
    double Gm, Ge;
    {
      const double Wmd = __shfl_sync(FULL, Wm, src_dn), Wed = __shfl_sync(FULL, We, src_dn);
      const double krd = __shfl_sync(FULL, c.kr, src_dn), denmd = __shfl_sync(FULL, c.den_m, src_dn), dened = __shfl_sync(FULL, c.den_e, src_dn);
      const double hld = __shfl_sync(FULL, c.hl, src_dn), tcd = __shfl_sync(FULL, c.tc, src_dn);
      double fm = 0.0, fe = 0.0;
      if (has_conn) {
        const FluxIn um = {Wm, c.kr, 0, c.den_m, 0, 0}, dm = {Wmd, krd, 0, denmd, 0, 0};
        const FluxIn ue = {Wm, c.kr, 0, c.den_e, 0, 0}, de = {Wmd, krd, 0, dened, 0, 0};
        double a1, a2, a3, a4, mfl;
        th_rich_flux(um, dm, upw, Dqm, gfac, area, fm, a1, a2, a3, a4);
        th_rich_flux(ue, de, upw, Dqe, gfac, area, mfl, a1, a2, a3, a4);
        const double kod = (c.tc * tcd) * rcp(dist_up * tcd + dist_dn * c.tc);
        const double h = (mfl <= 0.0) ? c.hl : hld;
        fe = mfl * h + (-kod * (We - Wed) * area);
      }
      const double fm_p = __shfl_sync(FULL, fm, src_up), fe_p = __shfl_sync(FULL, fe, src_up);
      const double am = por * c.den_m * c.sat * vol * dtInv;
      const double ae = (por * c.den_e * c.sat * c.ul + (1.0 - por) * 2700.0 * csol * (We - 273.15)) * vol * dtInv;
      if (phase == PH_INIT) { accm = am; acce = ae; }
      Gm = am - accm; Ge = ae - acce;
      if (has_up) { Gm = Gm + fm_p; Ge = Ge + fe_p; }
      if (has_conn) { Gm = Gm - fm; Ge = Ge - fe; }
      for (int k = 0; k < nmybc; ++k) {
        const BCL &b = bcs[k];
        double fl, a1, a2, a3, a4;
        if (b.ieqn == 1) {
          const FluxIn dn = {Wm, c.kr, 0, c.den_m, 0, 0};
          th_rich_flux(b.fin, dn, 0.0, b.Dq, b.bgf, area, fl, a1, a2, a3, a4);
          Gm = Gm + fl;
        } else {
          const FluxIn dn = {Wm, c.kr, 0, c.den_e, 0, 0};
          th_rich_flux(b.fin, dn, 0.0, b.Dq, b.bgf, area, fl, a1, a2, a3, a4);
          const double kod = c.tc / (0.0 + 0.5 * dz);
          const double h = (fl <= 0.0) ? b.hl : c.hl;
          Ge = Ge + (fl * h + (-kod * (b.T - We) * area));
        }
      }
      Gm = Gm - srcm;
      Ge = Ge + srce;                          // heat-rate sources ADD to the residual in the reference (:1478)
      if (!valid) { Gm = 0.0; Ge = 0.0; }
    }
    const double g2 = grp_sum<G>(Gm * Gm + Ge * Ge);
    const double w2 = grp_sum<G>(valid ? Wm * Wm + We * We : 0.0);
    nfuncs += 1;

    // ================= line-search / convergence logic (squared norms; as vsfm_step2_kernel) =================
    bool take = false;
    const bool g_bad = !(g2 == g2) || (g2 > 1.7e308);
    const bool out_of_funcs = (nfuncs >= so.max_funcs && so.max_funcs >= 0);
    const bool tiny_step = (stol2 * x2 > y2);
    if (!PADBC && A.eval_x && phase == PH_INIT) {
      if (valid) { Wm = A.eval_x[2 * cell]; We = A.eval_x[2 * cell + 1]; }
      phase = PH_EVAL;
    } else if (!PADBC && phase == PH_EVAL) {
      P = Wm; T = We; Fm = Gm; Fe = Ge; ax_store(s_ax, threadIdx.x, c);
      if (valid) { A.eval_f[2 * cell] = Gm; A.eval_f[2 * cell + 1] = Ge; }
      phase = PH_EVAL_J;
    } else if (phase == PH_INIT) {
      take = true;
    } else if (phase == PH_LS_FULL) {
      if (g_bad) {
        if (lambda <= so.ls_minlambda) { last_reason = SNES_DIVERGED_FNORM_NAN; phase = -1; }
        else if (out_of_funcs)         { last_reason = SNES_DIVERGED_FUNCTION_COUNT; phase = -1; }
        else { lambda = .5 * lambda; Wm = fma(-lambda, Ym, P); We = fma(-lambda, Ye, T); }
      } else if (.5 * g2 <= .5 * f2 + lambda * so.ls_alpha * initslope) take = true;
      else if (tiny_step) { last_reason = SNES_CONVERGED_SNORM_RELATIVE; phase = -1; }
      else if (out_of_funcs) { last_reason = SNES_DIVERGED_FUNCTION_COUNT; phase = -1; }
      else {
        double lt = -initslope / (g2 - f2 - 2.0 * lambda * initslope);
        lambdaprev = lambda; gprev = g2;
        if (lt > .5 * lambda) lt = .5 * lambda;
        lambda = (lt <= .1 * lambda) ? .1 * lambda : lt;
        Wm = fma(-lambda, Ym, P); We = fma(-lambda, Ye, T); phase = PH_LS_QUAD; ls_count = 0;
      }
    } else if (phase == PH_LS_QUAD || phase == PH_LS_CUBIC) {
      if (phase == PH_LS_CUBIC) ls_count += 1;
      const int ls_fail = tiny_step ? SNES_CONVERGED_SNORM_RELATIVE : SNES_DIVERGED_LINE_SEARCH;
      if (g_bad) { last_reason = ls_fail; phase = -1; }
      else if (.5 * g2 < .5 * f2 + lambda * so.ls_alpha * initslope) take = true;
      else if (ls_count >= so.ls_max_its) take = true;
      else if (lambda <= so.ls_minlambda) { last_reason = ls_fail; phase = -1; }
      else if (out_of_funcs) { last_reason = SNES_DIVERGED_FUNCTION_COUNT; phase = -1; }
      else {
        const double t1 = .5 * (g2 - f2) - lambda * initslope, t2 = .5 * (gprev - f2) - lambdaprev * initslope;
        const double ca = (t1 / (lambda * lambda) - t2 / (lambdaprev * lambdaprev)) / (lambda - lambdaprev);
        const double cb = (-lambdaprev * t1 / (lambda * lambda) + lambda * t2 / (lambdaprev * lambdaprev)) / (lambda - lambdaprev);
        double d = cb * cb - 3 * ca * initslope;
        if (d < 0.0) d = 0.0;
        double lt = (ca == 0.0) ? -initslope / (2.0 * cb) : (-cb + sqrt(d)) / (3.0 * ca);
        lambdaprev = lambda; gprev = g2;
        if (lt > .5 * lambda) lt = .5 * lambda;
        lambda = (lt <= .1 * lambda) ? .1 * lambda : lt;
        Wm = fma(-lambda, Ym, P); We = fma(-lambda, Ye, T); phase = PH_LS_CUBIC;
      }
    }
    if (take) {
      P = Wm; T = We; Fm = Gm; Fe = Ge; ax_store(s_ax, threadIdx.x, c);
      f2 = g2; x2 = w2;
      int reason = 0;
      if (phase == PH_INIT) {
        its = 0; ttol2 = g2 * rtol2; f2_0 = g2;
        if (g_bad) reason = SNES_DIVERGED_FNORM_NAN; else if (g2 < atol2) reason = SNES_CONVERGED_FNORM_ABS;
      } else {
        its += 1;
        if (g2 < atol2)           reason = SNES_CONVERGED_FNORM_ABS;
        else if (out_of_funcs)    reason = SNES_DIVERGED_FUNCTION_COUNT;
        else if (g2 <= ttol2)     reason = SNES_CONVERGED_FNORM_RELATIVE;
        else if (y2 < stol2 * x2) reason = SNES_CONVERGED_SNORM_RELATIVE;
        else if (so.divtol > 0 && g2 > divtol2 * f2_0) reason = SNES_DIVERGED_DTOL;
        else if (its >= so.max_it) reason = SNES_DIVERGED_MAX_IT;
      }
      if (reason) { last_reason = reason; phase = -1; } else phase = PH_NEWTON;
    }
  }

  if (!PADBC && A.eval_x) return;
  // ---- SOETHPostSolve: solution, mailbox, statistics ------------------------------------------------------------
  if (valid) {
    A.x_out[2 * cell] = P; A.x_out[2 * cell + 1] = T;
    if (converged) {
      const double sat = s_ax[0][threadIdx.x], den_m = s_ax[4][threadIdx.x];
      A.liq_sat[cell] = sat;
      A.mass[cell] = por * den_m * FMWH2O * sat * vol;
    }
  }
  const bool leader = col_ok && j == 0;
  if (leader) { A.stat_its[col] = tot_its; A.stat_reason[col] = last_reason; A.stat_cuts[col] = cuts; A.stat_nf[col] = tot_nf; }
  // block partials (maxima + worst reason only; the TH SoE keeps no mass-balance sums)
  double vits = leader ? (double)tot_its : 0.0, vdiv = leader ? (converged ? 0.0 : 1.0) : 0.0, vcut = leader ? (double)cuts : 0.0;
  int worst = leader ? last_reason : 0x7fffffff;
#pragma unroll
  for (int s = 16; s >= G; s >>= 1) {
    vits = fmax(vits, __shfl_xor_sync(FULL, vits, s)); vdiv = fmax(vdiv, __shfl_xor_sync(FULL, vdiv, s)); vcut = fmax(vcut, __shfl_xor_sync(FULL, vcut, s));
    worst = min(worst, __shfl_xor_sync(FULL, worst, s));
  }
  __shared__ double red[3][TH2_THREADS / 32];
  __shared__ int redw[TH2_THREADS / 32];
  const int warp = threadIdx.x >> 5;
  if (lane == 0) { red[0][warp] = vits; red[1][warp] = vdiv; red[2][warp] = vcut; redw[warp] = worst; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double o0 = 0.0, o1 = 0.0, o2 = 0.0; int ow = 0x7fffffff;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) { o0 = fmax(o0, red[0][w]); o1 = fmax(o1, red[1][w]); o2 = fmax(o2, red[2][w]); ow = min(ow, redw[w]); }
    double *bp = A.block_partials + (size_t)blockIdx.x * 9;
    for (int k = 0; k < 5; ++k) bp[k] = 0.0;
    bp[5] = o0; bp[6] = o1; bp[7] = o2; bp[8] = (double)ow;
  }
#undef por
#undef vol
#undef csol
#undef tkdry
#undef srcm
#undef srce
#undef upw
#undef Dqm
#undef Dqe
#undef gfac
#undef dist_up
#undef dist_dn
#undef area
#undef dz
#undef sp
}

}  // namespace mpp
