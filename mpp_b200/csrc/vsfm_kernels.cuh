// vsfm_kernels.cuh -- types shared by the VSFM step kernels (argument block, SNES options, condition descriptors,
// phase codes) and the second stage of the deterministic block reduction.  The step kernels themselves are in
// vsfm_kernels2.cuh (nlev <= 32: two cells per lane) and vsfm_generic_kernel.cuh (taller columns).
//
// Reference anchors of the step: SOEBaseStepDT_SNES src/mpp/soe/SystemOfEquationsBaseType.F90:368-552,
// VSFMSOEResidual/Jacobian src/mpp/soe/SystemOfEquationsVSFMType.F90:94-403, PETSc SNES newtonls + bt.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "physics.cuh"

namespace mpp {

constexpr int MAX_SS = 16;
constexpr int MAX_BC = 2;

// PETSc SNESConvergedReason codes the reference's drivers branch on
enum { SNES_ITERATING = 0, SNES_CONVERGED_FNORM_ABS = 2, SNES_CONVERGED_FNORM_RELATIVE = 3, SNES_CONVERGED_SNORM_RELATIVE = 4,
       SNES_DIVERGED_FUNCTION_COUNT = -2, SNES_DIVERGED_FNORM_NAN = -4, SNES_DIVERGED_MAX_IT = -5,
       SNES_DIVERGED_LINE_SEARCH = -6, SNES_DIVERGED_DTOL = -9 };

enum { REGION_TOP = 401, REGION_BOTTOM = 402, REGION_CELLS = 403 };
enum { CT_MASS_RATE = 503, CT_DIRICHLET = 505, CT_SEEPAGE = 509, CT_DOWNREG_CAMPBELL = 512, CT_DOWNREG_FETCH2 = 513 };

struct CondDev {
  const double *value;     // SoE mailbox condition_value: per column (top/bottom regions) or per cell (REGION_CELLS)
  int itype, region;
  double *flux;            // boundary_flux [kg/s] (BCs only)
  double *mass_exc;        // boundary mass exchanged [kg] (BCs only)
};

struct SnesOpts {
  double atol, rtol, stol, divtol;
  int max_it, max_funcs;
  int step_budget;         // > 0: give up on a column once one StepDT has spent this many residual evaluations (not in the reference)
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
  // the COND_MASS_RATE members of ss[] once more, sorted by region (filled by the launcher, vsfm_compact_sources): the fast kernel
  // loads them without looking at types and regions
  int nss_cell, nss_top, nss_bot;
  const double *ss_cell[MAX_SS], *ss_top[MAX_SS], *ss_bot[MAX_SS];
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
  // one optional down-regulated sink (COND_DOWNREG_MASS_RATE_CAMPBELL / _FETCH2, GoveqnRichards...:1900-1927, 2158-2188)
  int dr_type, dr_region; const double *dr_value, *dr_pc, *dr_n;
  long long *prof;         // development aid: per-section cycle counters (nullptr in production)
  double *t_done;          // optional per-column output: time this StepDT did advance (soe%time, SystemOfEquationsBaseType.F90:511)
  // RETRY specialisation only (the per-column retry loop of mppgpu_vsfm_elm_solve, MPPVSFMALM_Driver.F90:628-923):
  const int *retry_mask;   // 0 skip the column, 1 continue from x_in (remaining time), 2 redo from x_redo (soln_prev_clm)
  const int *retry_list; int nretry;   // compacted column indices: the retry launch is sized by the columns that need it
  const double *dt_col, *rtol_col, *stol_col, *x_redo;
  // EVAL specialisation only (mppgpu_eval, the residual / Jacobian probe of the unit tests): accumulation taken at x_in, residual and the
  // three Jacobian bands (sub, diagonal, super; cell order) at eval_x, no time step
  const double *eval_x; double *eval_f, *eval_ja, *eval_jb, *eval_jc;
  // optional launch order (column indices of this launch range, most expensive first by the previous step's cost); nullptr = batch order
  const int *order;
};

inline void vsfm_compact_sources(VsfmArgs &A)
{
  A.nss_cell = A.nss_top = A.nss_bot = 0;
  for (int k = 0; k < A.nss; ++k) {
    if (A.ss[k].itype != CT_MASS_RATE) continue;
    if (A.ss[k].region == REGION_CELLS)    A.ss_cell[A.nss_cell++] = A.ss[k].value;
    else if (A.ss[k].region == REGION_TOP) A.ss_top[A.nss_top++] = A.ss[k].value;
    else                                   A.ss_bot[A.nss_bot++] = A.ss[k].value;
  }
}

// Down-regulated mass sink: actual rate [kg/s] and the Jacobian diagonal term it adds (GoveqnRichards...:1900-1927, 2158-2188)
// (out of line: inlined at its six call sites in the step kernel, its pow and exp were ~10 KB of a Newton loop that has to fit a 32 KB
// instruction cache; the sink is a rare configuration and sits behind a uniform branch)
static __device__ __noinline__ void downreg_sink(int type, double value, double Pc, double n, double P, double &rate, double &djac)
{
  const double dP = P - PRESSURE_REF;
  rate = value; djac = 0.0;
  if (dP <= 0.0) {
    const double r = pow(dP / Pc, n);
    if (type == CT_DOWNREG_CAMPBELL) { const double factor = 1.0 + r; rate = value / factor; djac = (value / FMWH2O) * (n * r) / (dP * (factor * factor)); }
    else                             { const double factor = exp(-r); rate = value * factor; djac = (value / FMWH2O) * (n * r) * factor / dP; }
  }
}

enum { PH_INIT = 0, PH_NEWTON = 1, PH_LS_FULL = 2, PH_LS_QUAD = 3, PH_LS_CUBIC = 4, PH_DONE = 5 };

// All shuffles of the step kernels use the compile-time full mask and are executed by the whole warp under
// warp-uniform control flow: partial (run-time) masks make nvcc wrap every SHFL in a MATCH/WARPSYNC sequence
// (measured: 26% of all executed instructions in the first version, profiles/r1_vsfm_first.md).
constexpr unsigned FULL_MASK = 0xffffffffu;

#ifndef MPP_STEP_KERNEL_TU   // the support kernels below are compiled once, in mppgpu.cu

// Internal-connection mass fluxes [kg/s] of the committed state, on demand (GetDataForCLM(AUXVAR_CONN_INTERNAL, VAR_MASS_FLUX):
// SystemOfEquationsVSFMType.F90:824, internal_flux = flux * FMWH2O GoveqnRichards...:1809, 1199-1222).  One thread per connection
// j -> j+1 of column c, output index c * (nlev - 1) + j (the connection set's order, MeshType.F90:509-530); the aux vars are
// re-evaluated from the mailbox pressure with the same device functions the step kernel uses (frac_liq from the mailbox).
// Identical to the reference's value except after a stagnation exit (reason 4 through a failed line search), where the reference keeps
// the flux of the rejected trial point, O(stol) away from the iterate it returns.
__device__ __forceinline__ void conn_coeffs(double perm_up, double dz_up, double perm_dn, double dz_dn, double uz, double &upw, double &Dq, double &gfac)
{
  const double dist_up = 0.5 * dz_up, dist_dn = 0.5 * dz_dn;
  upw  = dist_up / (dist_up + dist_dn);
  Dq   = (perm_up * perm_dn) / (dist_up * perm_dn + dist_dn * perm_up);
  gfac = FMWH2O * ((dist_up + dist_dn) * (uz * (-GRAVITY_CONSTANT)));
}
__global__ void vsfm_conn_flux_kernel(const VsfmArgs A, int satfunc, const double *__restrict__ press, double *__restrict__ out)
{
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int nc = A.nlev - 1;
  if (nc <= 0 || i >= (long long)A.ncol * nc) return;
  const long long col = i / nc; const int j = (int)(i % nc);
  const long long cu = col * A.nlev + j, cd = cu + 1;
  double den[2], kr[2], P[2];
  for (int s = 0; s < 2; ++s) {
    const long long c = s ? cd : cu;
    SatParams sp; sp.sat_res = A.sat_res[c]; sp.alpha = A.alpha[c]; sp.m = A.lam[c]; sp.n = (satfunc == SATFUNC_VG) ? A.vgn[c] : 0.0;
    sp.pu = sp.ps = sp.b2 = sp.b3 = 0.0;
    if (satfunc == SATFUNC_SBC) { sp.pu = A.pu[c]; sp.ps = A.ps[c]; sp.b2 = A.b2[c]; sp.b3 = A.b3[c]; }
    SatState st; P[s] = press[c];
    sat_values_rt(satfunc, sp, P[s], A.frac_liq[c], st);
    kr[s] = st.kr;
    double dd; density_fixedT(A.dtab, P[s], den[s], dd);
  }
  double upw, Dq, gfac;
  conn_coeffs(A.perm[cu], A.dz[cu], A.perm[cd], A.dz[cd], A.uz, upw, Dq, gfac);
  const double den_ave = upw * den[0] + (1.0 - upw) * den[1];
  const double dphi = P[0] - P[1] + den_ave * gfac;
  const double ukvr = ((dphi >= 0.0) ? kr[0] : kr[1]) * (1.0 / VISCOSITY);
  out[i] = (((-Dq * ukvr * dphi) * A.area[col]) * den_ave) * FMWH2O;
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

// ---- launch order: columns grouped by the cost of their previous StepDT -------------------------------------------------------
// A warp of the step kernel advances 4 columns and is busy until the slowest of them has converged: with the columns in batch order
// that wastes 9-16 % of the warp-iterations of the benchmark batch (sum over warps of 4 max(nf) against sum of nf); the number of
// residual evaluations a column needs changes slowly from one time step to the next (92 % of the columns repeat it exactly), so a
// stable counting sort on the previous step's count, most expensive first (longest jobs start first), brings the waste to ~2 %.
// Results do not depend on the order (columns are independent); only the summation order of the block partials does.
constexpr int ORDER_BUCKETS = 12, ORDER_BLOCK = 1024;
__device__ __forceinline__ int order_bucket(int nf)
{
  // 0: <= 3, 1: >= 48, 2: 24-47, 3: 14-23, 4: 10-13, 5: 9, 6: 8, 7: 7, 8: 6, 9: 5, 10: 4
  // The cheapest columns go FIRST, not last: a column that converged at once in the previous step is the likeliest to stall in this one
  // (TH benchmark batch: 3 % of the columns needed <= 3 evaluations, and half of the columns that then need 60 - 2400 come from them --
  // tools/th_stragglers.py), and a straggler that starts with the last wave of the launch is all tail.
  if (nf <= 3) return 0;
  if (nf >= 10) return (nf >= 48) ? 1 : (nf >= 24 ? 2 : (nf >= 14 ? 3 : 4));
  return 14 - nf;
}
// pass 1: per-block bucket counts, counts[bucket * nblocks + block]
__global__ void order_count_kernel(const int *__restrict__ nf, int n, int *__restrict__ counts)
{
  __shared__ int hist[ORDER_BUCKETS];
  if (threadIdx.x < ORDER_BUCKETS) hist[threadIdx.x] = 0;
  __syncthreads();
  const int i = blockIdx.x * ORDER_BLOCK + threadIdx.x;
  if (i < n) atomicAdd(&hist[order_bucket(nf[i])], 1);
  __syncthreads();
  if (threadIdx.x < ORDER_BUCKETS) counts[threadIdx.x * gridDim.x + blockIdx.x] = hist[threadIdx.x];
}
// pass 2: exclusive scan of the m = ORDER_BUCKETS * nblocks counts in place (one block)
__global__ void order_scan_kernel(int *counts, int m)
{
  __shared__ int part[1024];
  const int per = (m + 1023) / 1024, lo = min(m, (int)threadIdx.x * per), hi = min(m, lo + per);
  int s = 0;
  for (int i = lo; i < hi; ++i) s += counts[i];
  part[threadIdx.x] = s;
  __syncthreads();
  for (int d = 1; d < 1024; d <<= 1) {
    const int v = (threadIdx.x >= (unsigned)d) ? part[threadIdx.x - d] : 0;
    __syncthreads();
    part[threadIdx.x] += v;
    __syncthreads();
  }
  int run = part[threadIdx.x] - s;
  for (int i = lo; i < hi; ++i) { const int c = counts[i]; counts[i] = run; run += c; }
}
// pass 3: stable scatter of the (range-local) column indices
__global__ void order_scatter_kernel(const int *__restrict__ nf, int n, const int *__restrict__ offsets, int *__restrict__ order)
{
  __shared__ int wcount[ORDER_BUCKETS][ORDER_BLOCK / 32];
  const int i = blockIdx.x * ORDER_BLOCK + threadIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int b = (i < n) ? order_bucket(nf[i]) : -1;
  int rank = 0;
#pragma unroll
  for (int k = 0; k < ORDER_BUCKETS; ++k) {
    const unsigned m = __ballot_sync(0xffffffffu, b == k);
    if (lane == 0) wcount[k][warp] = __popc(m);
    if (b == k) rank = __popc(m & ((1u << lane) - 1u));
  }
  __syncthreads();
  if (b >= 0) {
    int before = 0;
    for (int w = 0; w < warp; ++w) before += wcount[b][w];
    order[offsets[b * gridDim.x + blockIdx.x] + before + rank] = i;
  }
}

#endif  // MPP_STEP_KERNEL_TU

}  // namespace mpp
