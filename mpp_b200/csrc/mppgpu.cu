// mppgpu.cu -- C ABI (include/mppgpu.h) and host-side runtime of libmppgpu.so.
//
// Host logic mirrors the reference's system-of-equations objects for 1-D column batches:
//   sysofeqns_vsfm_type     src/mpp/soe/SystemOfEquationsVSFMType.F90
//   sysofeqns_thermal_type  src/mpp/soe/SystemOfEquationsThermalType.F90
//   sysofeqns_th_type       src/mpp/soe/SystemOfEquationsTHType.F90
// but owns device-resident structure-of-arrays state and launches the fused kernels in
// vsfm_kernels.cuh / thermal_kernels.cuh / th_kernels.cuh.  No CPU fallback exists: every entry point
// either runs on the GPU or fails loudly.
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <algorithm>
#include <string>
#include <vector>

#include "../../include/mppgpu.h"
#include "physics.cuh"
#include "vsfm_kernels.cuh"
#include "vsfm_kernels2.cuh"
#include "vsfm_generic_kernel.cuh"
#include "thermal_kernels.cuh"
#include "thermal_snow_kernels.cuh"
#include "vsfm_elm_kernels.cuh"
#include "th_kernels.cuh"
#include "th_kernels2.cuh"
#include "step_launch.h"

using namespace mpp;

// ---- ids (MultiPhysicsProbConstants.F90) -------------------------------------------------------------
enum { COND_BC = 501, COND_SS = 502, COND_MASS_RATE = 503, COND_MASS_FLUX = 504, COND_DIRICHLET = 505,
       COND_HEAT_FLUX = 507, COND_SEEPAGE_BC = 509, COND_HEAT_RATE = 511, COND_DOWNREG_MASS_RATE_CAMPBELL = 512, COND_DOWNREG_MASS_RATE_FETCH2 = 513 };
enum { VAR_PRESSURE = 604, VAR_TEMPERATURE = 605, VAR_BC_SS_CONDITION = 607, VAR_LIQ_SAT = 608, VAR_MASS = 610,
       VAR_SOIL_MATRIX_POT = 611, VAR_FRAC_LIQ_SAT = 612, VAR_BC_MASS_EXCHANGED = 614, VAR_LIQ_AREAL_DEN = 615,
       VAR_ICE_AREAL_DEN = 617, VAR_FRAC = 618, VAR_SNOW_WATER = 619, VAR_NUM_SNOW_LYR = 620, VAR_DHS_DT = 621,
       VAR_THERMAL_COND = 622, VAR_HEAT_CAP = 623, VAR_ACTIVE = 624, VAR_DZ = 627, VAR_DIST_UP = 628,
       VAR_DIST_DN = 629, VAR_TUNING_FACTOR = 630, VAR_POT_MASS_SINK_PRESSURE = 638, VAR_POT_MASS_SINK_EXPONENT = 639, VAR_MASS_FLUX = 644 };
enum { AUXVAR_INTERNAL = 701, AUXVAR_BC = 702, AUXVAR_SS = 703, AUXVAR_CONN_INTERNAL = 704 };

static thread_local std::string g_err;
static int fail(const char *fmt, ...)
{
  char buf[1024];
  va_list ap; va_start(ap, fmt); vsnprintf(buf, sizeof(buf), fmt, ap); va_end(ap);
  g_err = buf;
  return 1;
}
#define CK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) return fail("%s:%d CUDA error %s: %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_)); } while (0)
// Every entry point runs on the handle's device and hands the caller's current device back on return (a host model, or torch in the
// same process, may be working on another one).
struct DeviceScope {
  int prev = -1, cur = -1; cudaError_t err = cudaSuccess;
  explicit DeviceScope(int dev) : cur(dev) {
    err = cudaGetDevice(&prev);
    if (err == cudaSuccess && prev != dev) err = cudaSetDevice(dev);
  }
  ~DeviceScope() { if (prev >= 0 && prev != cur) (void)cudaSetDevice(prev); }
};
#define CHECK_H(h) if (!(h)) return fail("null handle"); DeviceScope device_scope_((h)->device); \
                   if (device_scope_.err != cudaSuccess) return fail("cannot select device %d: %s", (h)->device, cudaGetErrorString(device_scope_.err))

// every device buffer carries MPP_ALLOC_SLACK bytes of slack: the tile kernels (thermal_step2_tma_kernel) copy whole tiles and may
// read up to one tile past the end of the batch
#ifndef MPP_ALLOC_SLACK
#define MPP_ALLOC_SLACK 4096
#endif
static inline cudaError_t mpp_dmalloc(void **p, size_t bytes) { return cudaMalloc(p, bytes + MPP_ALLOC_SLACK); }

template <class T> struct DevBuf {
  T *p = nullptr; size_t n = 0;
  cudaError_t alloc(size_t count) { release(); n = count; if (!count) return cudaSuccess; return mpp_dmalloc((void **)&p, count * sizeof(T)); }
  void release() { if (p) cudaFree(p); p = nullptr; n = 0; }
  ~DevBuf() { release(); }
};

struct HostCond {
  int ieqn, ss_or_bc, itype, region;
  size_t n;                 // entries: ncol (top/bottom) or ncells (SOIL_CELLS)
  DevBuf<double> value, flux, mass_exc, dhsdT, frac, pot_pressure, pot_exponent;
};

struct mppgpu_soe {
  int soe_itype, ncol, nlev, device;
  size_t ncells;
  cudaStream_t stream = nullptr; bool own_stream = true;
  cudaStream_t copy_in = nullptr, copy_out = nullptr;      // coupled-step pipeline
  std::vector<cudaEvent_t> ev_in, ev_comp; cudaEvent_t ev_out_done = nullptr, ev_start = nullptr;
  long long launches = 0;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  float last_ms = 0.f;
  // mesh
  int orientation = MPPGPU_MESH_ALONG_GRAVITY; bool mesh_set = false, soils_set = false;
  DevBuf<double> dz, area; DevBuf<int> active; bool has_active = false;
  // conditions
  std::vector<HostCond *> bcs, sss;
  // solver options
  SnesOpts so;
  // ---- VSFM ----
  int satfunc_name = 0, density_type = DENSITY_CONSTANT;
  DevBuf<double> por, perm, sat_res, alpha, lam, vgn, pu, ps, b2, b3;
  DevBuf<double> frac_liq, temperature, liq_sat, pressure, mass, smp;
  DevBuf<double> xA, xB; double *x_committed = nullptr, *x_current = nullptr;   // soln_prev_clm / soln
  DevBuf<int> stat_its, stat_reason, stat_cuts, stat_nf;
  DevBuf<double> col_mass, col_err, col_src, block_partials, red_out, red_scratch; DevBuf<unsigned int> red_counter;
  double *h_red = nullptr;     // pinned mirror of red_out (9 doubles)
  // launch order of the step kernel (vsfm_kernels.cuh "launch order"): built after every StepDT from its per-column cost
  int sm_count = 148;
  DevBuf<double> conn_flux;          // internal-connection mass fluxes, filled on demand (vsfm_conn_flux_kernel)
  DevBuf<int> order, order_counts; bool order_valid = false; int order_chunks = 0; long long order_per = 0; int ordering = 1;
  int nblocks_last = 0;
  bool result_pending = false;
  // ---- thermal / TH state lives in their own structs ----
  ThermalState *thermal = nullptr;
  struct ElmState *elm = nullptr;          // mppgpu_vsfm_elm_solve (vsfm_elm_host.inl)
  THState *th = nullptr;
  struct CommState *comm = nullptr; bool comm_pending = false;   // comm_host.inl
};
static inline bool &c_pending(mppgpu_soe *h) { return h->comm_pending; }
static void comm_destroy(struct CommState *c);

// host-side pieces of the thermal / TH systems (thermal_host.inl, th_host.inl)
static int thermal_create(ThermalState *t, int ncol, int nlev, cudaStream_t s);
static void thermal_destroy(ThermalState *t);
static int thermal_set_mesh(ThermalState *t, int orientation, const double *d_dz, const double *d_area);
static int thermal_set_temperature(ThermalState *t, const double *T, bool restart);
static int thermal_field(mppgpu_soe *h, ThermalState *t, int auxvar_type, int var_type, int cond_id, bool for_set, double **p, size_t *cap);
static int thermal_set_idata(mppgpu_soe *h, ThermalState *t, int auxvar_type, int var_type, int cond_id, const int *data, int n);
static int thermal_pre_step_dt(ThermalState *t);
static int thermal_post_step_dt(ThermalState *t);
static int thermal_step(mppgpu_soe *h, ThermalState *t, double dt);
static int thermal_add_snow_ssw(mppgpu_soe *h, ThermalState *t, int nlevsno, const double *soil_top_dist_dn);
static void elm_destroy(struct ElmState *e);
static int elm_need(mppgpu_soe *h);
static void elm_set_chunks(struct ElmState *e, int nchunks);
static int thermal_elm_solve(mppgpu_soe *h, ThermalState *t, double dtime, const mppgpu_elm_thermal_columns *cols, double capr);
static int thermal_set_soils(mppgpu_soe *h, ThermalState *t, const double *watsat, const double *csol, const double *tkmg,
                             const double *tkdry, const int *lun_type, int nlevsoi, int istsoil);
static int th_create(THState *t, int ncol, int nlev, cudaStream_t s);
static void th_destroy(THState *t);
static int th_set_mesh(THState *t, int orientation, const double *d_dz, const double *d_area);
static int th_restart(THState *t, const double *x);
static int th_field(mppgpu_soe *h, THState *t, int ieqn, int auxvar_type, int var_type, int cond_id, bool for_set, double **p, size_t *cap);
static int th_pre_step_dt(THState *t);
static int th_post_step_dt(THState *t);
static int th_step(mppgpu_soe *h, THState *t, double dt);
static int th_eval(mppgpu_soe *h, THState *t, double dt, const double *x_prev, const double *x, double *f, double *ja, double *jb, double *jc);
static int th_set_soils(mppgpu_soe *h, THState *t, const double *watsat, const double *hksat, const double *bsw, const double *sucsat,
                        const double *residual_sat, const double *csol, const double *tkdry, int satfunc_type, int density_type, int iee_type);

static int vsfm_fill_args(mppgpu_soe *h, VsfmArgs &A, double dt);

static void default_snes(SnesOpts &so)
{
  so.atol = 1.e-50; so.rtol = 1.e-8; so.stol = 1.e-10; so.divtol = 1.e4;      // MultiPhysicsProbBaseType.F90:1110-1114 + PETSc defaults
  so.max_it = 50; so.max_funcs = 10000; so.step_budget = 0;
  so.ls_alpha = 1.e-4; so.ls_minlambda = 1.e-12; so.ls_maxstep = 1.e8; so.ls_max_its = 40;
}

extern "C" const char *mppgpu_last_error(void) { return g_err.c_str(); }
extern "C" int mppgpu_version(void) { return 100; }
extern "C" int mppgpu_device_count(void) { int n = 0; if (cudaGetDeviceCount(&n) != cudaSuccess) return 0; return n; }

// ---- small utility kernels ------------------------------------------------------------------------------
// (ncol,nlev) Fortran-order table -> cell order
__global__ void transpose_to_cells_kernel(const double *__restrict__ t, double *__restrict__ out, int ncol, int nlev)
{
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long n = (long long)ncol * nlev;
  if (i < n) { const int c = (int)(i / nlev), j = (int)(i % nlev); out[i] = t[(size_t)j * ncol + c]; }
}
__global__ void fill_kernel(double *p, double v, long long n)
{
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = v;
}
__global__ void convert_soils_kernel(int satfunc_name, const double *watsat, const double *hksat, const double *bsw,
                                     const double *sucsat, const double *residual_sat, int ncol, int nlev,
                                     double *por, double *perm, double *sat_res, double *alpha, double *lam, double *vgn,
                                     double *pu, double *ps, double *b2, double *b3, int *bad_flag)
{
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long n = (long long)ncol * nlev;
  if (i >= n) return;
  const int c = (int)(i / nlev), j = (int)(i % nlev);
  const size_t t = (size_t)j * ncol + c;
  SatParams sp; double po, pe;
  const int bad = convert_soil(satfunc_name, watsat[t], hksat[t], bsw[t], sucsat[t], residual_sat[t], po, pe, sp);
  por[i] = po; perm[i] = pe; sat_res[i] = sp.sat_res; alpha[i] = sp.alpha; lam[i] = sp.m;
  if (vgn) vgn[i] = sp.n;
  if (pu) { pu[i] = sp.pu; ps[i] = sp.ps; b2[i] = sp.b2; b3[i] = sp.b3; }
  if (bad) atomicExch(bad_flag, 1);
}

// Aux vars of the restart state (VSFMMPPRestart then the first GetDataForCLM of the ELM driver, MPPVSFMALM_Driver.F90:556-601):
// fills the SoE mailbox (pressure, liq_sat, mass, smp) and the per-column mass that the next StepDT's balance starts from.
// GROUP lanes per column (16 or 32; one cell per lane, coalesced, column mass by a fixed-order butterfly); GROUP = 0: one
// thread per column for taller columns.
template <int GROUP>
__global__ void vsfm_restart_mailbox_kernel(int satfunc, VsfmArgs A)
{
  const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int G = (GROUP > 0) ? GROUP : 1;
  const int col = (int)(tid / G), j0 = (int)(tid % G);
  const bool col_ok = (col < A.ncol) && (A.active == nullptr || A.active[col] != 0);
  const double area = col_ok ? A.area[col] : 1.0;
  double msum = 0.0;
  for (int j = j0; j < ((GROUP > 0) ? j0 + 1 : A.nlev); ++j) {
    if (!col_ok || j >= A.nlev) break;
    const long long cell = (long long)col * A.nlev + j;
    SatParams sp; sp.sat_res = A.sat_res[cell]; sp.alpha = A.alpha[cell]; sp.m = A.lam[cell]; sp.n = A.vgn ? A.vgn[cell] : 0.0;
    sp.pu = sp.ps = sp.b2 = sp.b3 = 0.0;
    if (A.pu) { sp.pu = A.pu[cell]; sp.ps = A.ps[cell]; sp.b2 = A.b2[cell]; sp.b3 = A.b3[cell]; }
    const double X = A.x_in[cell];
    SatState st; double den, dden;
    sat_values_rt(satfunc, sp, X, A.frac_liq[cell], st);
    density_fixedT_x<true>(A.dtab, X, den, dden);
    const double mass = A.por[cell] * den * FMWH2O * st.sat * (area * A.dz[cell]);
    A.liq_sat[cell] = st.sat; A.pressure[cell] = X; A.mass[cell] = mass;
    A.smp[cell] = (X - PRESSURE_REF) / (den * FMWH2O * GRAVITY_CONSTANT);
    msum += mass;
  }
  if (GROUP > 0) {
#pragma unroll
    for (int s = G / 2; s > 0; s >>= 1) msum += __shfl_xor_sync(0xffffffffu, msum, s, G);
  }
  if (col_ok && j0 == 0) A.col_mass[col] = msum;
}

constexpr int REDUCE_BLOCKS = 148;
static inline int nblk(long long n, int bs) { return (int)((n + bs - 1) / bs); }

static int upload_table(mppgpu_soe *h, const double *host, DevBuf<double> &tmp)
{
  CK(tmp.alloc(h->ncells));
  CK(cudaMemcpyAsync(tmp.p, host, h->ncells * sizeof(double), cudaMemcpyHostToDevice, h->stream));
  return 0;
}

// ---- life cycle ---------------------------------------------------------------------------------------------
extern "C" int mppgpu_destroy(mppgpu_handle h);
extern "C" int mppgpu_create(int soe_itype, int ncol, int nlev, int device, mppgpu_handle *out)
{
  if (!out) return fail("mppgpu_create: out is null");
  *out = nullptr;
  if (soe_itype != MPPGPU_SOE_RE_ODE && soe_itype != MPPGPU_SOE_THERMAL_TBASED && soe_itype != MPPGPU_SOE_TH)
    return fail("mppgpu_create: unknown soe_itype %d", soe_itype);
  if (ncol <= 0 || nlev <= 0) return fail("mppgpu_create: ncol and nlev must be positive");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
    return fail("mppgpu_create: no CUDA device available (this library has no CPU fallback)");
  if (device < 0 || device >= ndev) return fail("mppgpu_create: device %d out of range (0..%d)", device, ndev - 1);
  DeviceScope device_scope_(device);
  if (device_scope_.err != cudaSuccess) return fail("mppgpu_create: cannot select device %d: %s", device, cudaGetErrorString(device_scope_.err));
  mppgpu_soe *h = new mppgpu_soe();
  struct Guard { mppgpu_soe *h; ~Guard() { if (h) mppgpu_destroy(h); } } guard{h};     // any failure below releases what was created so far
  CK(cudaDeviceGetAttribute(&h->sm_count, cudaDevAttrMultiProcessorCount, device));
  h->soe_itype = soe_itype; h->ncol = ncol; h->nlev = nlev; h->device = device;
  h->ncells = (size_t)ncol * (size_t)nlev;
  default_snes(h->so);
  CK(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
  CK(cudaEventCreate(&h->ev0)); CK(cudaEventCreate(&h->ev1));
  CK(h->red_out.alloc(16));
  CK(cudaMemsetAsync(h->red_out.p, 0, 16 * sizeof(double), h->stream));
  CK(h->red_scratch.alloc(REDUCE_BLOCKS * 9)); CK(h->red_counter.alloc(1));
  CK(cudaMemsetAsync(h->red_counter.p, 0, sizeof(unsigned int), h->stream));
  CK(cudaMallocHost((void **)&h->h_red, 16 * sizeof(double)));
  memset(h->h_red, 0, 16 * sizeof(double));
  CK(h->stat_its.alloc(ncol)); CK(h->stat_reason.alloc(ncol)); CK(h->stat_cuts.alloc(ncol)); CK(h->stat_nf.alloc(ncol));
  CK(cudaMemsetAsync(h->stat_its.p, 0, ncol * sizeof(int), h->stream)); CK(cudaMemsetAsync(h->stat_reason.p, 0, ncol * sizeof(int), h->stream));
  CK(cudaMemsetAsync(h->stat_cuts.p, 0, ncol * sizeof(int), h->stream)); CK(cudaMemsetAsync(h->stat_nf.p, 0, ncol * sizeof(int), h->stream));
  if (soe_itype == MPPGPU_SOE_RE_ODE) {
    const size_t N = h->ncells;
    CK(h->frac_liq.alloc(N)); CK(h->liq_sat.alloc(N)); CK(h->pressure.alloc(N)); CK(h->mass.alloc(N)); CK(h->smp.alloc(N));
    CK(h->xA.alloc(N)); CK(h->xB.alloc(N));
    h->x_committed = h->xA.p; h->x_current = h->xA.p;
    fill_kernel<<<nblk(N, 256), 256, 0, h->stream>>>(h->frac_liq.p, 1.0, (long long)N);   // SystemOfEquationsVSFMAuxType.F90:66
    CK(cudaMemsetAsync(h->liq_sat.p, 0, N * 8, h->stream)); CK(cudaMemsetAsync(h->pressure.p, 0, N * 8, h->stream));
    CK(cudaMemsetAsync(h->mass.p, 0, N * 8, h->stream)); CK(cudaMemsetAsync(h->smp.p, 0, N * 8, h->stream));
    CK(cudaMemsetAsync(h->xA.p, 0, N * 8, h->stream)); CK(cudaMemsetAsync(h->xB.p, 0, N * 8, h->stream));
    CK(h->col_mass.alloc(ncol)); CK(h->col_err.alloc(ncol)); CK(h->col_src.alloc(ncol));
    CK(cudaMemsetAsync(h->col_mass.p, 0, ncol * 8, h->stream)); CK(cudaMemsetAsync(h->col_err.p, 0, ncol * 8, h->stream));
    CK(cudaMemsetAsync(h->col_src.p, 0, ncol * 8, h->stream));
  } else if (soe_itype == MPPGPU_SOE_THERMAL_TBASED) {
    h->thermal = new ThermalState();
    if (thermal_create(h->thermal, ncol, nlev, h->stream)) return fail("thermal_create failed: %s", cudaGetErrorString(cudaGetLastError()));
  } else {
    h->th = new THState();
    if (th_create(h->th, ncol, nlev, h->stream)) return fail("th_create failed: %s", cudaGetErrorString(cudaGetLastError()));
  }
  CK(cudaStreamSynchronize(h->stream));
  guard.h = nullptr;
  *out = h;
  return 0;
}

extern "C" int mppgpu_destroy(mppgpu_handle h)
{
  if (!h) return 0;
  DeviceScope device_scope_(h->device);
  if (h->stream) cudaStreamSynchronize(h->stream);
  for (auto *c : h->bcs) delete c;
  for (auto *c : h->sss) delete c;
  if (h->thermal) { thermal_destroy(h->thermal); delete h->thermal; }
  if (h->th) { th_destroy(h->th); delete h->th; }
  if (h->elm) elm_destroy(h->elm);
  if (h->comm) comm_destroy(h->comm);
  if (h->h_red) cudaFreeHost(h->h_red);
  if (h->ev0) cudaEventDestroy(h->ev0);
  if (h->ev1) cudaEventDestroy(h->ev1);
  for (auto e : h->ev_in) cudaEventDestroy(e);
  for (auto e : h->ev_comp) cudaEventDestroy(e);
  if (h->ev_out_done) cudaEventDestroy(h->ev_out_done);
  if (h->ev_start) cudaEventDestroy(h->ev_start);
  if (h->copy_in) cudaStreamDestroy(h->copy_in);
  if (h->copy_out) cudaStreamDestroy(h->copy_out);
  if (h->own_stream && h->stream) cudaStreamDestroy(h->stream);
  delete h;
  return 0;
}

extern "C" int mppgpu_set_stream(mppgpu_handle h, void *cuda_stream)
{
  CHECK_H(h);
  CK(cudaStreamSynchronize(h->stream));
  if (cuda_stream) {
    if (h->own_stream) cudaStreamDestroy(h->stream);
    h->stream = (cudaStream_t)cuda_stream; h->own_stream = false;
  } else if (!h->own_stream) {
    CK(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking)); h->own_stream = true;
  }
  if (h->thermal) h->thermal->stream = h->stream;
  if (h->th) h->th->stream = h->stream;
  return 0;
}

extern "C" int mppgpu_synchronize(mppgpu_handle h) { CHECK_H(h); CK(cudaStreamSynchronize(h->stream)); return 0; }

// ---- setup ----------------------------------------------------------------------------------------------------
extern "C" int mppgpu_set_mesh(mppgpu_handle h, int orientation, const double *dz, const double *area, const int *col_active)
{
  CHECK_H(h);
  if (!dz || !area) return fail("mppgpu_set_mesh: dz and area are required");
  if (orientation != MPPGPU_MESH_ALONG_GRAVITY && orientation != MPPGPU_MESH_AGAINST_GRAVITY && orientation != MPPGPU_MESH_HORIZONTAL)
    return fail("mppgpu_set_mesh: unknown orientation %d", orientation);
  if (!h->bcs.empty() || !h->sss.empty()) return fail("mppgpu_set_mesh: the mesh must be set before conditions are added");
  h->orientation = orientation;
  DevBuf<double> tmp;
  if (upload_table(h, dz, tmp)) return 1;
  CK(h->dz.alloc(h->ncells));
  transpose_to_cells_kernel<<<nblk(h->ncells, 256), 256, 0, h->stream>>>(tmp.p, h->dz.p, h->ncol, h->nlev);
  CK(cudaGetLastError());
  CK(h->area.alloc(h->ncol));
  CK(cudaMemcpyAsync(h->area.p, area, h->ncol * sizeof(double), cudaMemcpyHostToDevice, h->stream));
  h->has_active = (col_active != nullptr);
  if (col_active) {
    CK(h->active.alloc(h->ncol));
    CK(cudaMemcpyAsync(h->active.p, col_active, h->ncol * sizeof(int), cudaMemcpyHostToDevice, h->stream));
  }
  CK(cudaStreamSynchronize(h->stream));
  h->mesh_set = true;
  if (h->thermal) return thermal_set_mesh(h->thermal, orientation, h->dz.p, h->area.p) ? fail("thermal_set_mesh failed") : 0;
  if (h->th) return th_set_mesh(h->th, orientation, h->dz.p, h->area.p) ? fail("th_set_mesh failed") : 0;
  return 0;
}

// (ncol, nlev-1) Fortran-order connection table -> per-cell array (cell j carries connection j -> j+1)
__global__ void conn_table_to_cells_kernel(const double *__restrict__ t, double *__restrict__ out, int ncol, int nlev)
{
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long n = (long long)ncol * nlev;
  if (i < n) { const int c = (int)(i / nlev), j = (int)(i % nlev); out[i] = (j < nlev - 1) ? t[(size_t)j * ncol + c] : 0.0; }
}

extern "C" int mppgpu_set_connection_distances(mppgpu_handle h, const double *dist_up, const double *dist_dn)
{
  CHECK_H(h);
  if (!h->thermal) return fail("mppgpu_set_connection_distances: only the thermal SoE takes explicit connection distances");
  if (!h->mesh_set) return fail("mppgpu_set_connection_distances: set the mesh first");
  if (!dist_up || !dist_dn) return fail("mppgpu_set_connection_distances: null table");
  if (h->nlev < 2) return 0;
  ThermalState *t = h->thermal;
  const size_t nconn = (size_t)h->ncol * (h->nlev - 1);
  const double *src[2] = {dist_up, dist_dn}; double **dst[2] = {&t->dist_up, &t->dist_dn};
  for (int i = 0; i < 2; ++i) {
    DevBuf<double> tmp; CK(tmp.alloc(nconn));
    CK(cudaMemcpyAsync(tmp.p, src[i], nconn * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    if (!*dst[i]) CK(mpp_dmalloc((void **)dst[i], h->ncells * sizeof(double)));
    conn_table_to_cells_kernel<<<nblk(h->ncells, 256), 256, 0, h->stream>>>(tmp.p, *dst[i], h->ncol, h->nlev);
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(h->stream));
  }
  t->custom_dist = true;
  // ELM's vertical grid is the same in every column: if the table says so, the kernels read the distances per layer from
  // their constant bank and skip 16 B per cell of HBM traffic per step
  t->dist_uniform = false;
  if (h->nlev <= 32) {
    bool same = true;
    for (int j = 0; j < h->nlev - 1 && same; ++j) {
      const double *u = dist_up + (size_t)j * h->ncol, *d = dist_dn + (size_t)j * h->ncol;
      for (int c = 1; c < h->ncol; ++c) if (u[c] != u[0] || d[c] != d[0]) { same = false; break; }
    }
    if (same) {
      for (int j = 0; j < 32; ++j) { t->lay_du[j] = 0.0; t->lay_dd[j] = 0.0; }
      for (int j = 0; j < h->nlev - 1; ++j) { t->lay_du[j] = dist_up[(size_t)j * h->ncol]; t->lay_dd[j] = dist_dn[(size_t)j * h->ncol]; }
      t->dist_uniform = true;
    }
  }
  return 0;
}

extern "C" int mppgpu_add_condition(mppgpu_handle h, int ieqn, int ss_or_bc, int cond_type, int region, int *cond_id)
{
  CHECK_H(h);
  if (!h->mesh_set) return fail("mppgpu_add_condition: set the mesh first");
  if (ss_or_bc != COND_BC && ss_or_bc != COND_SS) return fail("mppgpu_add_condition: ss_or_bc must be COND_BC (501) or COND_SS (502)");
  if (region != REGION_TOP && region != REGION_BOTTOM && region != REGION_CELLS)
    return fail("mppgpu_add_condition: unsupported region %d (SOIL_TOP_CELLS 401, SOIL_BOTTOM_CELLS 402, SOIL_CELLS 403)", region);
  if (h->soe_itype == MPPGPU_SOE_RE_ODE) {
    if (ieqn != 1) return fail("mppgpu_add_condition: the VSFM SoE has one governing equation (ieqn = 1)");
    if (ss_or_bc == COND_BC) {
      // RichardsFlux accepts only these for Darcy boundary connections (RichardsMod.F90:262-273)
      if (cond_type != COND_DIRICHLET && cond_type != COND_SEEPAGE_BC)
        return fail("mppgpu_add_condition: VSFM boundary condition type %d unsupported (COND_DIRICHLET 505, COND_SEEPAGE_BC 509)", cond_type);
      if (region == REGION_CELLS) return fail("mppgpu_add_condition: boundary conditions live on SOIL_TOP_CELLS / SOIL_BOTTOM_CELLS");
      for (auto *c : h->bcs) if (c->region == region) return fail("mppgpu_add_condition: one boundary condition per region is supported");
      if ((int)h->bcs.size() >= MAX_BC) return fail("mppgpu_add_condition: at most %d boundary conditions", MAX_BC);
    } else {
      const bool downreg = (cond_type == COND_DOWNREG_MASS_RATE_CAMPBELL || cond_type == COND_DOWNREG_MASS_RATE_FETCH2);
      if (cond_type != COND_MASS_RATE && !downreg)
        return fail("mppgpu_add_condition: VSFM source/sink type %d unsupported (COND_MASS_RATE 503, COND_DOWNREG_MASS_RATE_CAMPBELL 512, _FETCH2 513)", cond_type);
      if (downreg) for (auto *c : h->sss) if (c->itype != COND_MASS_RATE) return fail("mppgpu_add_condition: one down-regulated sink per system of equations is supported");
      if ((int)h->sss.size() >= MAX_SS) return fail("mppgpu_add_condition: at most %d source/sink conditions", MAX_SS);
    }
  } else if (h->soe_itype == MPPGPU_SOE_THERMAL_TBASED) {
    if (h->thermal && h->thermal->snow_mode)
      return fail("mppgpu_add_condition: the snow + standing-water + soil configuration already holds ELM's conditions (mppgpu_thermal_add_snow_ssw)");
    if (ieqn != 1) return fail("mppgpu_add_condition: the soil thermal SoE has one governing equation here (ieqn = 1)");
    if (ss_or_bc == COND_BC && cond_type != COND_HEAT_FLUX && cond_type != COND_DIRICHLET)
      return fail("mppgpu_add_condition: thermal boundary condition type %d unsupported (COND_HEAT_FLUX 507, COND_DIRICHLET 505)", cond_type);
    if (ss_or_bc == COND_BC && region == REGION_CELLS) return fail("mppgpu_add_condition: boundary conditions live on SOIL_TOP_CELLS / SOIL_BOTTOM_CELLS");
    if (ss_or_bc == COND_SS && cond_type != COND_HEAT_RATE)
      return fail("mppgpu_add_condition: thermal source type %d unsupported (COND_HEAT_RATE 511)", cond_type);
  } else {
    if (ieqn != 1 && ieqn != 2) return fail("mppgpu_add_condition: TH has ieqn 1 (mass) and 2 (energy)");
  }
  HostCond *c = new HostCond();
  struct CondGuard { HostCond *c; ~CondGuard() { delete c; } } cguard{c};       // released below once the handle owns the condition
  c->ieqn = ieqn; c->ss_or_bc = ss_or_bc; c->itype = cond_type; c->region = region;
  c->n = (region == REGION_CELLS) ? h->ncells : (size_t)h->ncol;
  CK(c->value.alloc(c->n)); CK(cudaMemsetAsync(c->value.p, 0, c->n * 8, h->stream));
  if (ss_or_bc == COND_SS && (cond_type == COND_DOWNREG_MASS_RATE_CAMPBELL || cond_type == COND_DOWNREG_MASS_RATE_FETCH2)) {
    CK(c->pot_pressure.alloc(c->n)); CK(c->pot_exponent.alloc(c->n));          // aux_vars_ss%pot_mass_sink_{pressure,exponent}
    fill_kernel<<<nblk(c->n, 256), 256, 0, h->stream>>>(c->pot_pressure.p, -1.0, (long long)c->n);
    CK(cudaMemsetAsync(c->pot_exponent.p, 0, c->n * 8, h->stream));
  }
  if (ss_or_bc == COND_BC) {
    CK(c->flux.alloc(c->n)); CK(cudaMemsetAsync(c->flux.p, 0, c->n * 8, h->stream));
    CK(c->mass_exc.alloc(c->n)); CK(cudaMemsetAsync(c->mass_exc.p, 0, c->n * 8, h->stream));
    if (h->soe_itype == MPPGPU_SOE_THERMAL_TBASED) {
      CK(c->dhsdT.alloc(c->n)); CK(cudaMemsetAsync(c->dhsdT.p, 0, c->n * 8, h->stream));
      CK(c->frac.alloc(c->n));  CK(cudaMemsetAsync(c->frac.p, 0, c->n * 8, h->stream));     // ThermalKSPTemperatureBaseAuxType.F90:60
    }
    h->bcs.push_back(c); cguard.c = nullptr;
    if (cond_id) *cond_id = (int)h->bcs.size();
  } else {
    h->sss.push_back(c); cguard.c = nullptr;
    if (cond_id) *cond_id = (int)h->sss.size();
  }
  return 0;
}

extern "C" int mppgpu_vsfm_set_soils(mppgpu_handle h, const double *watsat, const double *hksat, const double *bsw,
                                     const double *sucsat, const double *residual_sat, int satfunc_type, int density_type)
{
  CHECK_H(h);
  if (h->soe_itype != MPPGPU_SOE_RE_ODE) return fail("mppgpu_vsfm_set_soils: handle is not a VSFM SoE");
  if (!watsat || !hksat || !bsw || !sucsat || !residual_sat) return fail("mppgpu_vsfm_set_soils: null table");
  if (satfunc_type < 0 || satfunc_type > 3) return fail("ERROR:: Unknown vsfm_satfunc_type = %d", satfunc_type);   // MultiPhysicsProbVSFM.F90:415
  if (density_type < DENSITY_CONSTANT || density_type > DENSITY_IFC67) return fail("Unknown value for VAR_DENSITY_TYPE %d", density_type);
  const size_t N = h->ncells;
  DevBuf<double> t[5]; const double *src[5] = {watsat, hksat, bsw, sucsat, residual_sat};
  for (int i = 0; i < 5; ++i) if (upload_table(h, src[i], t[i])) return 1;
  CK(h->por.alloc(N)); CK(h->perm.alloc(N)); CK(h->sat_res.alloc(N)); CK(h->alpha.alloc(N)); CK(h->lam.alloc(N));
  h->vgn.release(); h->pu.release(); h->ps.release(); h->b2.release(); h->b3.release();
  if (satfunc_type == MPPGPU_SATFUNC_VAN_GENUCHTEN) CK(h->vgn.alloc(N));
  if (satfunc_type >= MPPGPU_SATFUNC_SBC_BZ2) { CK(h->pu.alloc(N)); CK(h->ps.alloc(N)); CK(h->b2.alloc(N)); CK(h->b3.alloc(N)); }
  DevBuf<int> bad; CK(bad.alloc(1)); CK(cudaMemsetAsync(bad.p, 0, sizeof(int), h->stream));
  convert_soils_kernel<<<nblk(N, 128), 128, 0, h->stream>>>(satfunc_type, t[0].p, t[1].p, t[2].p, t[3].p, t[4].p, h->ncol, h->nlev,
      h->por.p, h->perm.p, h->sat_res.p, h->alpha.p, h->lam.p, h->vgn.p, h->pu.p, h->ps.p, h->b2.p, h->b3.p, bad.p);
  CK(cudaGetLastError());
  int hbad = 0;
  CK(cudaMemcpyAsync(&hbad, bad.p, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  if (hbad) return fail("SatFunc_Set_*: bad param (SaturationFunction.F90:141-146,177-182,285-291,343-349)");
  h->satfunc_name = satfunc_type; h->density_type = density_type; h->soils_set = true;
  return 0;
}

extern "C" int mppgpu_set_tolerances(mppgpu_handle h, double atol, double rtol, double stol, int max_it, int max_funcs)
{
  CHECK_H(h);
  h->so.atol = atol; h->so.rtol = rtol; h->so.stol = stol; h->so.max_it = max_it; h->so.max_funcs = max_funcs;   // (step_budget is kept)
  if (h->th) h->th->so = h->so;
  return 0;
}

extern "C" int mppgpu_set_step_budget(mppgpu_handle h, int max_residual_evaluations)
{
  CHECK_H(h);
  if (max_residual_evaluations < 0) return fail("mppgpu_set_step_budget: the budget must be >= 0 (0 = unlimited, the reference's behaviour)");
  h->so.step_budget = max_residual_evaluations;
  if (h->th) h->th->so = h->so;
  return 0;
}

extern "C" int mppgpu_thermal_set_bulk_copy(mppgpu_handle h, int mode)
{
  CHECK_H(h);
  if (!h->thermal) return fail("mppgpu_thermal_set_bulk_copy: handle is not a thermal SoE");
  if (mode != 0 && mode != 1) return fail("mppgpu_thermal_set_bulk_copy: mode must be 0 (register loads) or 1 (bulk-async copies where the shape fits)");
  h->thermal->bulk_copy = (mode != 0);
  return 0;
}

extern "C" int mppgpu_set_column_ordering(mppgpu_handle h, int mode)
{
  CHECK_H(h);
  if (mode != 0 && mode != 1) return fail("mppgpu_set_column_ordering: mode must be 0 (batch order) or 1 (by the previous step's cost)");
  h->ordering = mode; h->order_valid = false;
  return 0;
}

extern "C" int mppgpu_restart(mppgpu_handle h, const double *x, int n)
{
  CHECK_H(h);
  if (!x) return fail("mppgpu_restart: null data");
  if (h->soe_itype == MPPGPU_SOE_RE_ODE) {
    if ((size_t)n != h->ncells) return fail("VSFMMPPRestart: size(data_1d) /= ncells_local (%d vs %zu)", n, h->ncells);
    // soln, soln_prev and soln_prev_clm all take the restart vector (MultiPhysicsProbVSFM.F90:675-686)
    CK(cudaMemcpyAsync(h->xA.p, x, h->ncells * 8, cudaMemcpyHostToDevice, h->stream));
    h->x_committed = h->xA.p; h->x_current = h->xA.p; h->order_valid = false;
    if (h->mesh_set && h->soils_set) {
      VsfmArgs A;
      vsfm_fill_args(h, A, 1.0);
      A.x_in = h->xA.p;
      const int sf = (h->satfunc_name == MPPGPU_SATFUNC_VAN_GENUCHTEN) ? SATFUNC_VG : (h->satfunc_name == MPPGPU_SATFUNC_BROOKS_COREY ? SATFUNC_BC : SATFUNC_SBC);
      if (h->nlev <= 16)      vsfm_restart_mailbox_kernel<16><<<nblk((long long)h->ncol * 16, 128), 128, 0, h->stream>>>(sf, A);
      else if (h->nlev <= 32) vsfm_restart_mailbox_kernel<32><<<nblk((long long)h->ncol * 32, 128), 128, 0, h->stream>>>(sf, A);
      else                    vsfm_restart_mailbox_kernel<0><<<nblk(h->ncol, 128), 128, 0, h->stream>>>(sf, A);
      CK(cudaGetLastError());
      h->launches += 1;
    }
    CK(cudaStreamSynchronize(h->stream));
    return 0;
  }
  if (h->thermal) {
    const size_t want = h->thermal->snow_mode ? h->thermal->nall : h->ncells;
    if ((size_t)n != want) return fail("mppgpu_restart: thermal expects %zu temperatures", want);
    return thermal_set_temperature(h->thermal, x, true) ? fail("thermal restart failed") : 0;
  }
  if ((size_t)n != 2 * h->ncells) return fail("mppgpu_restart: TH expects 2*ncells values [P | T]");
  return th_restart(h->th, x) ? fail("th restart failed") : 0;
}

// ---- data exchange ----------------------------------------------------------------------------------------------
static HostCond *find_cond(mppgpu_soe *h, int auxvar_type, int cond_id)
{
  std::vector<HostCond *> &v = (auxvar_type == AUXVAR_BC) ? h->bcs : h->sss;
  if (cond_id < 1 || cond_id > (int)v.size()) return nullptr;
  return v[cond_id - 1];
}

static int vsfm_field(mppgpu_soe *h, int auxvar_type, int var_type, int cond_id, bool for_set, double **p, size_t *cap)
{
  if (auxvar_type == AUXVAR_INTERNAL) {
    *cap = h->ncells;
    switch (var_type) {
    case VAR_FRAC_LIQ_SAT: *p = h->frac_liq.p; return 0;
    case VAR_TEMPERATURE:
      // stored in the SoE mailbox only; never reaches the Richards aux vars (GoveqnRichardsODEPressureType.F90:573-575)
      if (!h->temperature.p) { if (h->temperature.alloc(h->ncells) != cudaSuccess) return fail("alloc temperature");
        fill_kernel<<<nblk(h->ncells, 256), 256, 0, h->stream>>>(h->temperature.p, 298.15, (long long)h->ncells); }
      *p = h->temperature.p; return 0;
    case VAR_PRESSURE: *p = h->pressure.p; return 0;
    case VAR_LIQ_SAT: *p = h->liq_sat.p; return 0;
    case VAR_MASS: *p = h->mass.p; return 0;
    case VAR_SOIL_MATRIX_POT: *p = h->smp.p; return 0;
    }
    return fail("In VSFMSOEAuxVar%sValue: unknown var_type %d", for_set ? "Set" : "Get", var_type);
  }
  if (auxvar_type == AUXVAR_CONN_INTERNAL) {
    // SystemOfEquationsVSFMType.F90:824: one aux var per internal connection, mass_flux [kg/s] of the committed state
    if (for_set || var_type != VAR_MASS_FLUX) return fail("In VSFMSOEAuxVar%sValue: unknown var_type %d", for_set ? "Set" : "Get", var_type);
    if (!h->mesh_set || !h->soils_set) return fail("VSFMSOEGetDataForCLM: mesh and soils must be set first");
    const size_t nconn = (size_t)h->ncol * (size_t)(h->nlev > 1 ? h->nlev - 1 : 0);
    *cap = nconn;
    if (!nconn) { *p = nullptr; return 0; }
    if (h->conn_flux.n < nconn && h->conn_flux.alloc(nconn) != cudaSuccess) return fail("alloc conn_flux");
    VsfmArgs A;
    vsfm_fill_args(h, A, 1.0);
    const int sf = (h->satfunc_name == MPPGPU_SATFUNC_VAN_GENUCHTEN) ? SATFUNC_VG : (h->satfunc_name == MPPGPU_SATFUNC_BROOKS_COREY ? SATFUNC_BC : SATFUNC_SBC);
    vsfm_conn_flux_kernel<<<nblk((long long)nconn, 256), 256, 0, h->stream>>>(A, sf, h->pressure.p, h->conn_flux.p);
    CK(cudaGetLastError());
    h->launches += 1;
    *p = h->conn_flux.p; return 0;
  }
  if (auxvar_type != AUXVAR_BC && auxvar_type != AUXVAR_SS) return fail("VSFMSOE%sData: Unknown soe_auxvar_type %d", for_set ? "Set" : "Get", auxvar_type);
  HostCond *c = find_cond(h, auxvar_type, cond_id);
  if (!c) return fail("VSFMSOE%sData: condition id %d out of range", for_set ? "Set" : "Get", cond_id);
  *cap = c->n;
  if (var_type == VAR_BC_SS_CONDITION) { *p = c->value.p; return 0; }
  // VSFMMPPSetSourceSinkAuxVarRealValue (MultiPhysicsProbVSFM.F90:1437-1520)
  if (for_set && auxvar_type == AUXVAR_SS && c->pot_pressure.p && var_type == VAR_POT_MASS_SINK_PRESSURE) { *p = c->pot_pressure.p; return 0; }
  if (for_set && auxvar_type == AUXVAR_SS && c->pot_exponent.p && var_type == VAR_POT_MASS_SINK_EXPONENT) { *p = c->pot_exponent.p; return 0; }
  if (!for_set && auxvar_type == AUXVAR_BC && var_type == VAR_MASS_FLUX) { *p = c->flux.p; return 0; }
  if (!for_set && auxvar_type == AUXVAR_BC && var_type == VAR_BC_MASS_EXCHANGED) { *p = c->mass_exc.p; return 0; }
  if (!for_set && auxvar_type == AUXVAR_SS && var_type == VAR_MASS_FLUX) { *p = c->value.p; return 0; }   // ss_flux = value (GoveqnRichards...:1873)
  return fail("In VSFMSOEAuxVar%sValue: unknown var_type %d", for_set ? "Set" : "Get", var_type);
}

static int xfer(mppgpu_soe *h, int ieqn, int auxvar_type, int var_type, int cond_id, const double *src, double *dst, int n, bool set, bool device_ptr)
{
  if (n < 0) return fail("negative size");
  double *p = nullptr; size_t cap = 0;
  if (h->soe_itype == MPPGPU_SOE_RE_ODE) {
    if (vsfm_field(h, auxvar_type, var_type, cond_id, set, &p, &cap)) return 1;
  } else if (h->thermal) {
    if (thermal_field(h, h->thermal, auxvar_type, var_type, cond_id, set, &p, &cap)) return 1;
  } else {
    if (th_field(h, h->th, ieqn, auxvar_type, var_type, cond_id, set, &p, &cap)) return 1;
  }
  if ((size_t)n > cap) return fail("size(data_1d) > nauxvar (%d > %zu)", n, cap);     // SystemOfEquationsVSFMType.F90:711-716
  const cudaMemcpyKind kind = device_ptr ? cudaMemcpyDeviceToDevice : (set ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToHost);
  if (set) CK(cudaMemcpyAsync(p, src, (size_t)n * 8, kind, h->stream));
  else     CK(cudaMemcpyAsync(dst, p, (size_t)n * 8, kind, h->stream));
  if (!device_ptr) CK(cudaStreamSynchronize(h->stream));     // caller may free / read the host array right after the call
  return 0;
}

extern "C" int mppgpu_set_data(mppgpu_handle h, int ieqn, int auxvar_type, int var_type, int cond_id, const double *data, int n)
{ CHECK_H(h); if (!data) return fail("mppgpu_set_data: null data"); return xfer(h, ieqn, auxvar_type, var_type, cond_id, data, nullptr, n, true, false); }
extern "C" int mppgpu_get_data(mppgpu_handle h, int ieqn, int auxvar_type, int var_type, int cond_id, double *data, int n)
{ CHECK_H(h); if (!data) return fail("mppgpu_get_data: null data"); return xfer(h, ieqn, auxvar_type, var_type, cond_id, nullptr, data, n, false, false); }
extern "C" int mppgpu_set_data_device(mppgpu_handle h, int ieqn, int auxvar_type, int var_type, int cond_id, const double *d, int n)
{ CHECK_H(h); if (!d) return fail("mppgpu_set_data_device: null data"); return xfer(h, ieqn, auxvar_type, var_type, cond_id, d, nullptr, n, true, true); }
extern "C" int mppgpu_get_data_device(mppgpu_handle h, int ieqn, int auxvar_type, int var_type, int cond_id, double *d, int n)
{ CHECK_H(h); if (!d) return fail("mppgpu_get_data_device: null data"); return xfer(h, ieqn, auxvar_type, var_type, cond_id, nullptr, d, n, false, true); }

extern "C" int mppgpu_set_idata(mppgpu_handle h, int ieqn, int auxvar_type, int var_type, int cond_id, const int *data, int n)
{
  CHECK_H(h);
  if (!h->thermal) return fail("mppgpu_set_idata: only the thermal SoE has integer/logical data (SetIDataFromCLM/SetBDataFromCLM)");
  if (!data) return fail("mppgpu_set_idata: null data");
  return thermal_set_idata(h, h->thermal, auxvar_type, var_type, cond_id, data, n);
}

// ---- time stepping ----------------------------------------------------------------------------------------------
extern "C" int mppgpu_pre_step_dt(mppgpu_handle h)
{
  CHECK_H(h);
  if (h->soe_itype == MPPGPU_SOE_RE_ODE) {
    // VSFMSPreStepDT (SystemOfEquationsVSFMType.F90:892-923): soln = soln_prev = soln_prev_clm; reset boundary mass exchanged
    h->x_current = h->x_committed;
    for (auto *c : h->bcs) CK(cudaMemsetAsync(c->mass_exc.p, 0, c->n * 8, h->stream));
    return 0;
  }
  if (h->thermal) return thermal_pre_step_dt(h->thermal) ? fail("thermal_pre_step_dt failed") : 0;
  return th_pre_step_dt(h->th) ? fail("th_pre_step_dt failed") : 0;
}

extern "C" int mppgpu_post_step_dt(mppgpu_handle h)
{
  CHECK_H(h);
  if (h->soe_itype == MPPGPU_SOE_RE_ODE) {
    // VSFMSPostStepDT (:926-940): soln_prev_clm = soln_prev
    h->x_committed = h->x_current;
    return 0;
  }
  if (h->thermal) return thermal_post_step_dt(h->thermal) ? fail("thermal_post_step_dt failed") : 0;
  return th_post_step_dt(h->th) ? fail("th_post_step_dt failed") : 0;
}

static int vsfm_fill_args(mppgpu_soe *h, VsfmArgs &A, double dt)
{
  memset(&A, 0, sizeof(A));
  A.ncol = h->ncol; A.nlev = h->nlev;
  A.uz = (h->orientation == MPPGPU_MESH_ALONG_GRAVITY) ? -1.0 : (h->orientation == MPPGPU_MESH_AGAINST_GRAVITY ? 1.0 : 0.0);
  A.top_is_first = (h->orientation != MPPGPU_MESH_AGAINST_GRAVITY);
  A.dtab = make_density_table(h->density_type, 273.15 + 25.0);       // RichardsODEPressureAuxType.F90:92
  A.por = h->por.p; A.perm = h->perm.p; A.sat_res = h->sat_res.p; A.alpha = h->alpha.p; A.lam = h->lam.p; A.vgn = h->vgn.p;
  A.pu = h->pu.p; A.ps = h->ps.p; A.b2 = h->b2.p; A.b3 = h->b3.p; A.dz = h->dz.p; A.area = h->area.p;
  A.active = h->has_active ? h->active.p : nullptr;
  A.frac_liq = h->frac_liq.p;
  A.nss = (int)h->sss.size(); A.nbc = (int)h->bcs.size();
  for (int k = 0; k < A.nss; ++k) {
    A.ss[k].value = h->sss[k]->value.p; A.ss[k].itype = h->sss[k]->itype; A.ss[k].region = h->sss[k]->region; A.ss[k].flux = nullptr; A.ss[k].mass_exc = nullptr;
    if (h->sss[k]->itype != COND_MASS_RATE) {
      A.dr_type = h->sss[k]->itype; A.dr_region = h->sss[k]->region;
      A.dr_value = h->sss[k]->value.p; A.dr_pc = h->sss[k]->pot_pressure.p; A.dr_n = h->sss[k]->pot_exponent.p;
    }
  }
  for (int k = 0; k < A.nbc; ++k) { A.bc[k].value = h->bcs[k]->value.p; A.bc[k].itype = h->bcs[k]->itype; A.bc[k].region = h->bcs[k]->region; A.bc[k].flux = h->bcs[k]->flux.p; A.bc[k].mass_exc = h->bcs[k]->mass_exc.p; }
  A.liq_sat = h->liq_sat.p; A.pressure = h->pressure.p; A.mass = h->mass.p; A.smp = h->smp.p;
  A.stat_its = h->stat_its.p; A.stat_reason = h->stat_reason.p; A.stat_cuts = h->stat_cuts.p; A.stat_nf = h->stat_nf.p;
  A.col_mass = h->col_mass.p; A.col_err = h->col_err.p; A.col_src = h->col_src.p;
  A.dt = dt; A.so = h->so;
  return 0;
}

template <int LPC>
static void launch_vsfm2(mppgpu_soe *h, const VsfmArgs &A0, int nblocks)
{
  VsfmArgs A = A0;
  vsfm_compact_sources(A);
  const int sf = (h->satfunc_name == MPPGPU_SATFUNC_VAN_GENUCHTEN) ? SATFUNC_VG : (h->satfunc_name == MPPGPU_SATFUNC_BROOKS_COREY ? SATFUNC_BC : SATFUNC_SBC);
  // the specialisation that carries boundary conditions, the down-regulated sink and the IFC-67 density polynomial
  const bool bc = A.nbc > 0 || A.dr_type != 0 || A.dtab.type == DENSITY_IFC67;
  const int variant = A.eval_x ? 3 : (A.retry_mask ? 2 : (bc ? 1 : 0));
  // the 18 instances live in vsfm_step2_inst.cu, one translation unit per (LPC, saturation function)
  if (LPC == 8) { if (sf == SATFUNC_VG) vsfm2_launch_8_0(A, variant, nblocks, h->stream); else if (sf == SATFUNC_BC) vsfm2_launch_8_1(A, variant, nblocks, h->stream); else vsfm2_launch_8_2(A, variant, nblocks, h->stream); }
  else          { if (sf == SATFUNC_VG) vsfm2_launch_16_0(A, variant, nblocks, h->stream); else if (sf == SATFUNC_BC) vsfm2_launch_16_1(A, variant, nblocks, h->stream); else vsfm2_launch_16_2(A, variant, nblocks, h->stream); }
}

#ifdef VSFM2_PROFILE
static long long *g_prof = nullptr;
extern "C" int mppgpu_dbg_profile(long long *out7 /* 12 entries */)
{
  if (!g_prof) return 1;
  cudaDeviceSynchronize();
  cudaMemcpy(out7, g_prof, 12 * sizeof(long long), cudaMemcpyDeviceToHost);
  cudaMemset(g_prof, 0, 12 * sizeof(long long));
  return 0;
}
#endif

// shift every per-cell / per-column pointer of A to the sub-batch [col0, col0 + n)
static void vsfm_offset_args(VsfmArgs &A, int nlev, long long col0, int n, long long block0)
{
  const long long c = col0 * nlev;
  const double **cellp[] = {&A.por, &A.perm, &A.sat_res, &A.alpha, &A.lam, &A.vgn, &A.pu, &A.ps, &A.b2, &A.b3, &A.dz, &A.frac_liq, &A.x_in};
  for (auto pp : cellp) if (*pp) *pp += c;
  double **cellw[] = {&A.x_out, &A.liq_sat, &A.pressure, &A.mass, &A.smp};
  for (auto pp : cellw) if (*pp) *pp += c;
  A.area += col0; if (A.active) A.active += col0;
  for (int k = 0; k < A.nss; ++k) A.ss[k].value += (A.ss[k].region == REGION_CELLS) ? c : col0;
  if (A.dr_type) { const long long o = (A.dr_region == REGION_CELLS) ? c : col0; A.dr_value += o; A.dr_pc += o; A.dr_n += o; }
  for (int k = 0; k < A.nbc; ++k) { A.bc[k].value += col0; A.bc[k].flux += col0; A.bc[k].mass_exc += col0; }
  A.stat_its += col0; A.stat_reason += col0; A.stat_cuts += col0; A.stat_nf += col0;
  A.col_mass += col0; A.col_err += col0; A.col_src += col0;
  A.block_partials += block0 * 9;
  if (A.t_done) A.t_done += col0;
  if (A.order) A.order += col0;
  if (A.eval_x) { A.eval_x += c; A.eval_f += c; A.eval_ja += c; A.eval_jb += c; A.eval_jc += c; }
  if (A.retry_mask) { A.retry_mask += col0; A.dt_col += col0; A.rtol_col += col0; A.stol_col += col0; A.x_redo += c; }
  A.ncol = n;
}

static int vsfm_blocks_for(mppgpu_soe *h, long long ncol)
{
  const int nlev = h->nlev;
  if (nlev <= 32) return nblk(ncol * ((nlev <= 16) ? 8 : 16), VSFM2_THREADS);
  return nblk(ncol, VSFM_GENERIC_WARPS);
}

// launch the step kernel on columns [col0, col0 + n) (A holds whole-batch pointers); partials go to blocks [block0, ...)
static int vsfm_launch_range(mppgpu_soe *h, const VsfmArgs &A0, long long col0, int n, long long block0, cudaStream_t s)
{
  VsfmArgs A = A0;
  vsfm_offset_args(A, h->nlev, col0, n, block0);
#ifdef VSFM2_PROFILE
  if (!g_prof) { mpp_dmalloc((void **)&g_prof, 12 * sizeof(long long)); cudaMemset(g_prof, 0, 12 * sizeof(long long)); }
  A.prof = g_prof;
#endif
  const int nlev = h->nlev, nblocks = vsfm_blocks_for(h, n);
  cudaStream_t keep = h->stream; h->stream = s;
  if (nlev <= 16)      launch_vsfm2<8>(h, A, nblocks);
  else if (nlev <= 32) launch_vsfm2<16>(h, A, nblocks);
  else {
    const size_t smem = vsfm_generic_smem_bytes(nlev);
    if (smem > 200 * 1024) { h->stream = keep; return fail("mppgpu_step_dt: nlev = %d exceeds the generic kernel's shared-memory budget", nlev); }
    CK(cudaFuncSetAttribute(vsfm_step_generic_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    vsfm_step_generic_kernel<<<nblocks, 32 * VSFM_GENERIC_WARPS, smem, s>>>(A, h->satfunc_name == 0 ? SATFUNC_VG : (h->satfunc_name == 1 ? SATFUNC_BC : SATFUNC_SBC));
  }
  h->stream = keep;
  CK(cudaGetLastError());
  h->launches += 1;
  return 0;
}

static int vsfm_prepare_step(mppgpu_soe *h, double dt, VsfmArgs &A)
{
  if (!h->mesh_set || !h->soils_set) return fail("mppgpu_step_dt: mesh and soils must be set first");
  if (!(dt > 0.0)) return fail("mppgpu_step_dt: dt must be positive");
  vsfm_fill_args(h, A, dt);
  // soln == soln_prev at entry; write to the spare buffer if the current one is the committed (soln_prev_clm) copy
  A.x_in = h->x_current;
  double *spare = (h->x_committed == h->xA.p) ? h->xB.p : h->xA.p;
  A.x_out = (h->x_current == h->x_committed) ? spare : h->x_current;
  return 0;
}

// Build the launch order of columns [col0, col0 + n) for the NEXT StepDT from the cost (residual evaluations) of the one that has just
// been queued on `s`: three small kernels (count, scan, stable scatter), range-local indices at order[col0 ...].
static int vsfm_build_order(mppgpu_soe *h, long long col0, int n, cudaStream_t s)
{
  if (!h->ordering || h->nlev > 32) return 0;
  if (h->order.n < (size_t)h->ncol) CK(h->order.alloc(h->ncol));
  const int nb = nblk(n, ORDER_BLOCK);
  if (col0 % ORDER_BLOCK) return fail("vsfm_build_order: ranges must start at a multiple of %d columns", ORDER_BLOCK);
  const size_t c0 = (size_t)(col0 / ORDER_BLOCK) * ORDER_BUCKETS;                      // disjoint scratch per range
  const size_t need = (size_t)(nblk(h->ncol, ORDER_BLOCK) + 1) * ORDER_BUCKETS;
  if (h->order_counts.n < need) { CK(cudaStreamSynchronize(h->stream)); CK(h->order_counts.alloc(need)); }
  int *counts = h->order_counts.p + c0;
  if (c0 + (size_t)nb * ORDER_BUCKETS > h->order_counts.n) return fail("vsfm_build_order: scratch overflow");
  order_count_kernel<<<nb, ORDER_BLOCK, 0, s>>>(h->stat_nf.p + col0, n, counts);
  order_scan_kernel<<<1, 1024, 0, s>>>(counts, nb * ORDER_BUCKETS);
  order_scatter_kernel<<<nb, ORDER_BLOCK, 0, s>>>(h->stat_nf.p + col0, n, counts, h->order.p + col0);
  CK(cudaGetLastError());
  h->launches += 3;
  return 0;
}

static int vsfm_finish_step(mppgpu_soe *h, const VsfmArgs &A, int nblocks)
{
  reduce_partials_kernel<<<nblocks < REDUCE_BLOCKS ? 1 : REDUCE_BLOCKS, 256, 0, h->stream>>>(h->block_partials.p, nblocks, h->red_scratch.p, h->red_counter.p, h->red_out.p);
  CK(cudaGetLastError());
  h->launches += 1;
  h->x_current = A.x_out;
  h->nblocks_last = nblocks;
  CK(cudaMemcpyAsync(h->h_red, h->red_out.p, 9 * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  h->result_pending = true;
  return 0;
}

static int vsfm_step(mppgpu_soe *h, double dt)
{
  VsfmArgs A;
  if (vsfm_prepare_step(h, dt, A)) return 1;
  const int nblocks = vsfm_blocks_for(h, h->ncol);
  if (h->block_partials.n < (size_t)nblocks * 9) CK(h->block_partials.alloc((size_t)nblocks * 9));
  A.block_partials = h->block_partials.p;
  CK(cudaEventRecord(h->ev0, h->stream));
  A.order = (h->order_valid && h->order_chunks == 1) ? h->order.p : nullptr;
  if (vsfm_launch_range(h, A, 0, h->ncol, 0, h->stream)) return 1;
  if (vsfm_build_order(h, 0, h->ncol, h->stream)) return 1;
  h->order_valid = (h->ordering != 0 && h->nlev <= 32); h->order_chunks = 1; h->order_per = h->ncol;
  if (vsfm_finish_step(h, A, nblocks)) return 1;
  CK(cudaEventRecord(h->ev1, h->stream));
  return 0;
}

extern "C" int mppgpu_step_dt_async(mppgpu_handle h, double dt, int nstep)
{
  CHECK_H(h);
  (void)nstep;
  if (h->soe_itype == MPPGPU_SOE_RE_ODE) return vsfm_step(h, dt);
  if (h->thermal) return thermal_step(h, h->thermal, dt);
  return th_step(h, h->th, dt);
}

// device time between the events around the last StepDT; events that were never recorded (no step yet, or a step that bailed out early)
// make cudaEventElapsedTime fail, and a failure must not linger as the thread's "last error" for the next launch check
static void step_elapsed(mppgpu_soe *h)
{
  float ms = 0.0f;
  if (cudaEventElapsedTime(&ms, h->ev0, h->ev1) == cudaSuccess) h->last_ms = ms;
  else (void)cudaGetLastError();
}

extern "C" int mppgpu_step_result(mppgpu_handle h, int *converged, int *converged_reason)
{
  CHECK_H(h);
  CK(cudaStreamSynchronize(h->stream));
  step_elapsed(h);
  h->result_pending = false;
  if (h->thermal) { if (converged) *converged = 1; if (converged_reason) *converged_reason = 0; return 0; }   // KSP path: tridiagonal solve cannot diverge
  const double *r = h->h_red;
  if (converged) *converged = (r[6] == 0.0) ? 1 : 0;
  if (converged_reason) *converged_reason = (int)r[8];
  return 0;
}

extern "C" int mppgpu_step_dt(mppgpu_handle h, double dt, int nstep, int *converged, int *converged_reason)
{
  if (mppgpu_step_dt_async(h, dt, nstep)) return 1;
  return mppgpu_step_result(h, converged, converged_reason);
}

// ---- ELM coupling step, pipelined over column chunks -------------------------------------------------------------------
struct CoupledField { double *dev; size_t per_col; double *host; };
static int vsfm_coupled_pipeline(mppgpu_soe *h, double dt, const std::vector<CoupledField> &fin, const std::vector<CoupledField> &fout, int nchunks,
                                 int *converged, int *converged_reason)
{
  typedef CoupledField Field;
  // VSFMSPreStepDT (SystemOfEquationsVSFMType.F90:892-923)
  h->x_current = h->x_committed;
  for (auto *c : h->bcs) CK(cudaMemsetAsync(c->mass_exc.p, 0, c->n * 8, h->stream));
  VsfmArgs A;
  if (vsfm_prepare_step(h, dt, A)) return 1;
  // chunking: whole blocks of the step kernel (16 columns per 128-thread block at nlev <= 16), 16 chunks by default
  if (nchunks <= 0) nchunks = 16;
  const long long align = 1024;
  long long per = ((long long)h->ncol + nchunks - 1) / nchunks;
  per = ((per + align - 1) / align) * align;
  nchunks = (int)(((long long)h->ncol + per - 1) / per);
  long long total_blocks = 0;
  for (int k = 0; k < nchunks; ++k) total_blocks += vsfm_blocks_for(h, std::min<long long>(per, h->ncol - k * per));
  if (h->block_partials.n < (size_t)total_blocks * 9) CK(h->block_partials.alloc((size_t)total_blocks * 9));
  A.block_partials = h->block_partials.p;
  if (!h->copy_in)  CK(cudaStreamCreateWithFlags(&h->copy_in, cudaStreamNonBlocking));
  if (!h->copy_out) CK(cudaStreamCreateWithFlags(&h->copy_out, cudaStreamNonBlocking));
  if (!h->ev_out_done) { CK(cudaEventCreateWithFlags(&h->ev_out_done, cudaEventDisableTiming)); CK(cudaEventCreateWithFlags(&h->ev_start, cudaEventDisableTiming)); }
  while ((int)h->ev_in.size() < nchunks) {
    cudaEvent_t a, b; CK(cudaEventCreateWithFlags(&a, cudaEventDisableTiming)); CK(cudaEventCreateWithFlags(&b, cudaEventDisableTiming));
    h->ev_in.push_back(a); h->ev_comp.push_back(b);
  }
  // the copy streams start after whatever is already queued on the compute stream (earlier steps, PreStepDT memsets)
  CK(cudaEventRecord(h->ev_start, h->stream));
  CK(cudaStreamWaitEvent(h->copy_in, h->ev_start, 0));
  CK(cudaStreamWaitEvent(h->copy_out, h->ev_start, 0));
  CK(cudaEventRecord(h->ev0, h->stream));
  long long block0 = 0;
  for (int k = 0; k < nchunks; ++k) {
    const long long col0 = k * per; const int n = (int)std::min<long long>(per, h->ncol - col0);
    for (const Field &f : fin)
      CK(cudaMemcpyAsync(f.dev + col0 * f.per_col, f.host + col0 * f.per_col, (size_t)n * f.per_col * 8, cudaMemcpyHostToDevice, h->copy_in));
    CK(cudaEventRecord(h->ev_in[k], h->copy_in));
    CK(cudaStreamWaitEvent(h->stream, h->ev_in[k], 0));
    A.order = (h->order_valid && h->order_per == per && h->order_chunks == nchunks) ? h->order.p : nullptr;
    if (vsfm_launch_range(h, A, col0, n, block0, h->stream)) return 1;
    CK(cudaEventRecord(h->ev_comp[k], h->stream));
    if (vsfm_build_order(h, col0, n, h->stream)) return 1;
    CK(cudaStreamWaitEvent(h->copy_out, h->ev_comp[k], 0));
    for (const Field &f : fout)
      CK(cudaMemcpyAsync(f.host + col0 * f.per_col, f.dev + col0 * f.per_col, (size_t)n * f.per_col * 8, cudaMemcpyDeviceToHost, h->copy_out));
    block0 += vsfm_blocks_for(h, n);
  }
  CK(cudaEventRecord(h->ev_out_done, h->copy_out));
  h->order_valid = (h->ordering != 0 && h->nlev <= 32); h->order_chunks = nchunks; h->order_per = per;
  if (vsfm_finish_step(h, A, (int)total_blocks)) return 1;
  CK(cudaEventRecord(h->ev1, h->stream));
  CK(cudaStreamWaitEvent(h->stream, h->ev_out_done, 0));          // later work on the handle's stream sees the host arrays complete
  CK(cudaStreamSynchronize(h->copy_out));                         // the caller may read its arrays as soon as this returns
  return mppgpu_step_result(h, converged, converged_reason);
}

extern "C" int mppgpu_vsfm_coupled_step(mppgpu_handle h, double dt, int nstep, int nin, const mppgpu_xfer *in, int nout, const mppgpu_xfer *out,
                                        int nchunks, int *converged, int *converged_reason)
{
  CHECK_H(h);
  (void)nstep;
  if (h->soe_itype != MPPGPU_SOE_RE_ODE) return fail("mppgpu_vsfm_coupled_step: handle is not a VSFM SoE");
  if ((nin > 0 && !in) || (nout > 0 && !out) || nin < 0 || nout < 0) return fail("mppgpu_vsfm_coupled_step: bad transfer lists");
  typedef CoupledField Field;
  std::vector<Field> fin(nin), fout(nout);
  for (int i = 0; i < nin + nout; ++i) {
    const bool is_in = i < nin;
    const mppgpu_xfer &x = is_in ? in[i] : out[i - nin];
    if (!x.host) return fail("mppgpu_vsfm_coupled_step: null host array");
    double *p = nullptr; size_t cap = 0;
    if (vsfm_field(h, x.auxvar_type, x.var_type, x.cond_id, is_in, &p, &cap)) return 1;
    Field f{p, cap / (size_t)h->ncol, x.host};
    if (is_in) fin[i] = f; else fout[i - nin] = f;
  }
  // On any failure inside the pipeline, copies to and from the CALLER's host arrays may still be queued on the copy streams: drain all three
  // streams before returning (the caller may free or reuse its buffers), and put the handle back where it was (soln, launch order).
  double *const x_before = h->x_current;
  if (vsfm_coupled_pipeline(h, dt, fin, fout, nchunks, converged, converged_reason)) {
    const std::string msg = g_err;
    if (h->copy_in) cudaStreamSynchronize(h->copy_in);
    if (h->copy_out) cudaStreamSynchronize(h->copy_out);
    cudaStreamSynchronize(h->stream);
    (void)cudaGetLastError();
    h->x_current = x_before; h->order_valid = false; h->result_pending = false;
    g_err = msg;
    return 1;
  }
  return 0;
}

// ---- diagnostics --------------------------------------------------------------------------------------------------
extern "C" int mppgpu_get_column_stats(mppgpu_handle h, int *newton_its, int *reasons, int *dt_cuts, int *nfuncs)
{
  CHECK_H(h);
  const size_t nb = (size_t)h->ncol * sizeof(int);
  if (newton_its) CK(cudaMemcpyAsync(newton_its, h->stat_its.p, nb, cudaMemcpyDeviceToHost, h->stream));
  if (reasons)    CK(cudaMemcpyAsync(reasons, h->stat_reason.p, nb, cudaMemcpyDeviceToHost, h->stream));
  if (dt_cuts)    CK(cudaMemcpyAsync(dt_cuts, h->stat_cuts.p, nb, cudaMemcpyDeviceToHost, h->stream));
  if (nfuncs)     CK(cudaMemcpyAsync(nfuncs, h->stat_nf.p, nb, cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  return 0;
}

extern "C" int mppgpu_vsfm_mass_balance(mppgpu_handle h, double dt, double sums[4], double maxs[4])
{
  CHECK_H(h);
  (void)dt;
  if (h->soe_itype != MPPGPU_SOE_RE_ODE && !h->th) return fail("mppgpu_vsfm_mass_balance: not a flow SoE");
  CK(cudaStreamSynchronize(h->stream));
  for (int k = 0; k < 4; ++k) { if (sums) sums[k] = h->h_red[k]; if (maxs) maxs[k] = h->h_red[4 + k]; }
  return 0;
}

extern "C" int mppgpu_reduction_buffer_device(mppgpu_handle h, double **d_buf)
{ CHECK_H(h); if (!d_buf) return fail("null"); *d_buf = h->red_out.p; return 0; }

extern "C" int mppgpu_host_register(void *ptr, long long nbytes)
{
  if (!ptr || nbytes <= 0) return fail("mppgpu_host_register: null pointer or non-positive size");
  const cudaError_t e = cudaHostRegister(ptr, (size_t)nbytes, cudaHostRegisterPortable);
  if (e != cudaSuccess) { (void)cudaGetLastError(); return fail("mppgpu_host_register: %s", cudaGetErrorString(e)); }   // not sticky: clear it
  return 0;
}
extern "C" int mppgpu_host_unregister(void *ptr)
{
  if (!ptr) return fail("mppgpu_host_unregister: null pointer");
  const cudaError_t e = cudaHostUnregister(ptr);
  if (e != cudaSuccess) { (void)cudaGetLastError(); return fail("mppgpu_host_unregister: %s", cudaGetErrorString(e)); }
  return 0;
}
extern "C" int mppgpu_launch_count(mppgpu_handle h, long long *n) { CHECK_H(h); if (n) *n = h->launches; return 0; }
extern "C" int mppgpu_last_step_ms(mppgpu_handle h, float *ms)
{
  CHECK_H(h);
  CK(cudaStreamSynchronize(h->stream));
  step_elapsed(h);
  if (ms) *ms = h->last_ms;
  return 0;
}

// VSFMSOEResidual + VSFMJacobian (SystemOfEquationsVSFMType.F90:94-403) at x, with the accumulation of the start of the step taken at
// x_prev: the EVAL instance of the fused step kernel dumps the residual and the three Jacobian bands it assembles (N values each, cell order)
static int vsfm_eval(mppgpu_soe *h, double dt, const double *x_prev, const double *x, double *f, double *ja, double *jb, double *jc)
{
  if (!h->mesh_set || !h->soils_set) return fail("mppgpu_eval: mesh and soils must be set first");
  if (!x_prev || !x || !f || !ja || !jb || !jc) return fail("mppgpu_eval: null argument");
  if (h->nlev > 32) return fail("mppgpu_eval: the VSFM probe covers the fused kernel (nlev <= 32); nlev = %d runs on the generic kernel", h->nlev);
  if (!(dt > 0.0)) return fail("mppgpu_eval: dt must be positive");
  const size_t N = h->ncells;
  DevBuf<double> dxp, dx, df, da, db, dc;
  CK(dxp.alloc(N)); CK(dx.alloc(N)); CK(df.alloc(N)); CK(da.alloc(N)); CK(db.alloc(N)); CK(dc.alloc(N));
  CK(cudaMemcpyAsync(dxp.p, x_prev, N * 8, cudaMemcpyHostToDevice, h->stream));
  CK(cudaMemcpyAsync(dx.p, x, N * 8, cudaMemcpyHostToDevice, h->stream));
  VsfmArgs A;
  vsfm_fill_args(h, A, dt);
  A.x_in = dxp.p; A.x_out = dxp.p;
  A.eval_x = dx.p; A.eval_f = df.p; A.eval_ja = da.p; A.eval_jb = db.p; A.eval_jc = dc.p;
  const int nblocks = vsfm_blocks_for(h, h->ncol);
  if (h->block_partials.n < (size_t)nblocks * 9) CK(h->block_partials.alloc((size_t)nblocks * 9));
  A.block_partials = h->block_partials.p;
  if (vsfm_launch_range(h, A, 0, h->ncol, 0, h->stream)) return 1;
  CK(cudaMemcpyAsync(f, df.p, N * 8, cudaMemcpyDeviceToHost, h->stream));
  CK(cudaMemcpyAsync(ja, da.p, N * 8, cudaMemcpyDeviceToHost, h->stream));
  CK(cudaMemcpyAsync(jb, db.p, N * 8, cudaMemcpyDeviceToHost, h->stream));
  CK(cudaMemcpyAsync(jc, dc.p, N * 8, cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  return 0;
}

extern "C" int mppgpu_eval(mppgpu_handle h, double dt, const double *x_prev, const double *x, double *f, double *ja, double *jb, double *jc)
{
  CHECK_H(h);
  if (h->soe_itype == MPPGPU_SOE_RE_ODE) return vsfm_eval(h, dt, x_prev, x, f, ja, jb, jc);
  if (h->th) return th_eval(h, h->th, dt, x_prev, x, f, ja, jb, jc);
  return fail("mppgpu_eval: not available for the thermal SoE");
}

// thermal / TH set-soils entry points live next to their state
extern "C" int mppgpu_thermal_set_soils(mppgpu_handle h, const double *watsat, const double *csol, const double *tkmg,
                                        const double *tkdry, const int *lun_type, int nlevsoi, int istsoil)
{
  CHECK_H(h);
  if (!h->thermal) return fail("mppgpu_thermal_set_soils: handle is not a thermal SoE");
  if (!h->mesh_set) return fail("mppgpu_thermal_set_soils: set the mesh first");
  return thermal_set_soils(h, h->thermal, watsat, csol, tkmg, tkdry, lun_type, nlevsoi, istsoil);
}
extern "C" int mppgpu_thermal_set_cnfac(mppgpu_handle h, double cnfac)
{
  CHECK_H(h);
  if (!h->thermal) return fail("mppgpu_thermal_set_cnfac: handle is not a thermal SoE");
  h->thermal->cnfac = cnfac; return 0;
}
extern "C" int mppgpu_thermal_add_snow_ssw(mppgpu_handle h, int nlevsno, const double *soil_top_dist_dn)
{
  CHECK_H(h);
  if (!h->thermal) return fail("mppgpu_thermal_add_snow_ssw: handle is not a thermal SoE");
  return thermal_add_snow_ssw(h, h->thermal, nlevsno, soil_top_dist_dn);
}
extern "C" int mppgpu_thermal_elm_solve(mppgpu_handle h, double dtime, int nstep, const mppgpu_elm_thermal_columns *cols, double capr)
{
  CHECK_H(h);
  (void)nstep;
  if (!h->thermal) return fail("mppgpu_thermal_elm_solve: handle is not a thermal SoE");
  if (!(dtime > 0.0)) return fail("mppgpu_thermal_elm_solve: dtime must be positive");
  return thermal_elm_solve(h, h->thermal, dtime, cols, capr);
}
extern "C" int mppgpu_elm_set_pipeline(mppgpu_handle h, int nchunks, int static_soil_geometry)
{
  CHECK_H(h);
  if (nchunks < 0 || nchunks > 1024) return fail("mppgpu_elm_set_pipeline: nchunks must be within 0..1024 (0: default)");
  if (static_soil_geometry != 0 && static_soil_geometry != 1) return fail("mppgpu_elm_set_pipeline: static_soil_geometry must be 0 or 1");
  if (h->soe_itype == MPPGPU_SOE_RE_ODE) {
    if (static_soil_geometry) return fail("mppgpu_elm_set_pipeline: static_soil_geometry applies to the thermal SoE (the VSFM geometry is set once by mppgpu_vsfm_elm_set_geometry)");
    if (elm_need(h)) return 1;
    elm_set_chunks(h->elm, nchunks);
    return 0;
  }
  if (!h->thermal) return fail("mppgpu_elm_set_pipeline: handle has no ELM solve entry point");
  h->thermal->elm_chunks = nchunks;
  h->thermal->elm_static_soil = (static_soil_geometry != 0);
  h->thermal->elm_soil_loaded = false;
  return 0;
}
extern "C" int mppgpu_th_set_soils(mppgpu_handle h, const double *watsat, const double *hksat, const double *bsw,
                                   const double *sucsat, const double *residual_sat, const double *csol, const double *tkdry,
                                   int satfunc_type, int density_type, int int_energy_enthalpy_type)
{
  CHECK_H(h);
  if (!h->th) return fail("mppgpu_th_set_soils: handle is not a TH SoE");
  if (!h->mesh_set) return fail("mppgpu_th_set_soils: set the mesh first");
  return th_set_soils(h, h->th, watsat, hksat, bsw, sucsat, residual_sat, csol, tkdry, satfunc_type, density_type, int_energy_enthalpy_type);
}

static int th_set_energy_permeability(mppgpu_soe *h, THState *t, const double *perm);
extern "C" int mppgpu_th_set_energy_permeability(mppgpu_handle h, const double *perm, int n)
{
  CHECK_H(h);
  if (!h->th) return fail("mppgpu_th_set_energy_permeability: handle is not a TH SoE");
  if (!perm || n != h->ncells) return fail("No. of values for soil permeability is not equal to no. of grid cells.");     // GoveqnThermalEnthalpySoilType.F90:2471-2475
  return th_set_energy_permeability(h, h->th, perm);
}

#include "comm_host.inl"
#include "thermal_host.inl"
#include "vsfm_elm_host.inl"
#include "th_host.inl"
