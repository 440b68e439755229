// vsfm_elm_host.inl -- host side of mppgpu_vsfm_elm_solve: MPPVSFMALM_Solve (src/driver/alm/MPPVSFMALM_Driver.F90:204-923) with ELM's raw
// column arrays on both sides.  Included at the end of mppgpu.cu.

struct ElmState {
  bool geometry_set = false;
  int nlevsoi = 0, cond_ids[6] = {0, 0, 0, 0, 0, 0};
  int nchunks = 0;                          // column chunks of the solve pipeline (0: default)
  double watmin = 0.01;
  size_t npft_cap = 0;
  DevBuf<double> zi, dz, stage;                                                        // static geometry; staging for Fortran-order arrays
  DevBuf<double> rootr, qtran, qinfl, dews, dewg, subs, fh2osfc, qdrain, zwt, liq, ice, snowlyr, negsnow, perched;   // inputs
  DevBuf<int> snl;
  DevBuf<int> pfti, npfts, pactive; DevBuf<double> wtcol, rootr_pft, qtran_pft;         // optional patch level
  DevBuf<double> frac_ice, mass_beg, tot_flux, dt_rem, rtol, stol, t_done, smp_l, soilp, qcharge, abs_err;
  DevBuf<int> iter_count, diverged, mask, status, pending, retry_list;
};

static void elm_destroy(ElmState *e) { delete e; }
static void elm_set_chunks(ElmState *e, int nchunks) { e->nchunks = nchunks; }

static int elm_need(mppgpu_soe *h)
{
  if (h->soe_itype != MPPGPU_SOE_RE_ODE) return fail("mppgpu_vsfm_elm_*: handle is not a VSFM SoE");
  if (!h->elm) h->elm = new ElmState();
  return 0;
}

static int elm_set_geometry(mppgpu_handle h, const double *zi, const double *dz, int nlevsoi, double watmin, const int *cond_ids, bool fortran_order);
extern "C" int mppgpu_vsfm_elm_set_geometry(mppgpu_handle h, const double *zi, const double *dz, int nlevsoi, double watmin, const int *cond_ids)
{ return elm_set_geometry(h, zi, dz, nlevsoi, watmin, cond_ids, false); }
extern "C" int mppgpu_vsfm_elm_set_geometry_f(mppgpu_handle h, const double *zi, const double *dz, int nlevsoi, double watmin, const int *cond_ids)
{ return elm_set_geometry(h, zi, dz, nlevsoi, watmin, cond_ids, true); }

static int elm_set_geometry(mppgpu_handle h, const double *zi, const double *dz, int nlevsoi, double watmin, const int *cond_ids, bool fortran_order)
{
  CHECK_H(h);
  if (elm_need(h)) return 1;
  if (!zi || !dz || !cond_ids) return fail("mppgpu_vsfm_elm_set_geometry: null argument");
  if (nlevsoi < 0 || nlevsoi > h->nlev) return fail("mppgpu_vsfm_elm_set_geometry: nlevsoi must be within 0..nlev");
  ElmState *e = h->elm;
  // the six mass-rate conditions of MPPVSFMALM_Initialize.F90:836-858: infiltration, ET, dew, drainage, snow, sublimation
  static const int want_region[6] = {REGION_TOP, REGION_CELLS, REGION_TOP, REGION_CELLS, REGION_TOP, REGION_TOP};
  for (int k = 0; k < 6; ++k) {
    HostCond *c = find_cond(h, AUXVAR_SS, cond_ids[k]);
    if (!c || c->itype != COND_MASS_RATE || c->region != want_region[k])
      return fail("mppgpu_vsfm_elm_set_geometry: condition %d (slot %d) must be a COND_MASS_RATE source on %s", cond_ids[k], k,
                  want_region[k] == REGION_TOP ? "SOIL_TOP_CELLS" : "SOIL_CELLS");
    e->cond_ids[k] = cond_ids[k];
  }
  const size_t ncol = h->ncol, N = h->ncells;
  CK(e->zi.alloc(ncol * (h->nlev + 1))); CK(e->dz.alloc(N));
  if (!fortran_order) {
    CK(cudaMemcpyAsync(e->zi.p, zi, ncol * (h->nlev + 1) * 8, cudaMemcpyHostToDevice, h->stream));
    CK(cudaMemcpyAsync(e->dz.p, dz, N * 8, cudaMemcpyHostToDevice, h->stream));
  } else {
    DevBuf<double> tmp;
    CK(tmp.alloc(ncol * (h->nlev + 1)));
    CK(cudaMemcpyAsync(tmp.p, zi, ncol * (h->nlev + 1) * 8, cudaMemcpyHostToDevice, h->stream));
    transpose_to_cells_kernel<<<nblk(ncol * (h->nlev + 1), 256), 256, 0, h->stream>>>(tmp.p, e->zi.p, h->ncol, h->nlev + 1);
    CK(cudaStreamSynchronize(h->stream));
    CK(cudaMemcpyAsync(tmp.p, dz, N * 8, cudaMemcpyHostToDevice, h->stream));
    transpose_to_cells_kernel<<<nblk(N, 256), 256, 0, h->stream>>>(tmp.p, e->dz.p, h->ncol, h->nlev);
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(h->stream));
  }
  DevBuf<double> *cellb[] = {&e->rootr, &e->liq, &e->ice, &e->perched, &e->frac_ice, &e->smp_l, &e->soilp};
  for (auto b : cellb) { CK(b->alloc(N)); CK(cudaMemsetAsync(b->p, 0, N * 8, h->stream)); }    // outputs of filtered-out columns read 0
  DevBuf<double> *colb[] = {&e->qtran, &e->qinfl, &e->dews, &e->dewg, &e->subs, &e->fh2osfc, &e->qdrain, &e->zwt, &e->snowlyr, &e->negsnow,
                            &e->mass_beg, &e->tot_flux, &e->dt_rem, &e->rtol, &e->stol, &e->t_done, &e->qcharge, &e->abs_err};
  for (auto b : colb) { CK(b->alloc(ncol)); CK(cudaMemsetAsync(b->p, 0, ncol * 8, h->stream)); }
  DevBuf<int> *coli[] = {&e->snl, &e->iter_count, &e->diverged, &e->mask, &e->status, &e->retry_list};
  for (auto b : coli) { CK(b->alloc(ncol)); CK(cudaMemsetAsync(b->p, 0, ncol * sizeof(int), h->stream)); }
  CK(e->pending.alloc(2));                                               // [0] columns that need another StepDT, [1] failed columns
  CK(cudaStreamSynchronize(h->stream));
  e->nlevsoi = nlevsoi; e->watmin = watmin; e->geometry_set = true;
  return 0;
}

// column-range variants of the two layout kernels: columns [c0, c0 + n) of a (ncol, nlev) Fortran-order table <-> cell order
__global__ void transpose_to_cells_range_kernel(const double *__restrict__ t, double *__restrict__ cells, int ncol, int nlev, int c0, int n)
{
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < (long long)n * nlev) { const int c = c0 + (int)(i / nlev), j = (int)(i % nlev); cells[(size_t)c * nlev + j] = t[(size_t)j * ncol + c]; }
}
__global__ void transpose_from_cells_range_kernel(const double *__restrict__ cells, double *__restrict__ t, int ncol, int nlev, int c0, int n)
{
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;        // i runs along the rows of the table: coalesced writes
  if (i < (long long)n * nlev) { const int j = (int)(i / n), c = c0 + (int)(i % n); t[(size_t)j * ncol + c] = cells[(size_t)c * nlev + j]; }
}

// What can still be in flight when the pipeline below bails out: copies to and from the CALLER's arrays on three streams.
static int elm_solve_pipeline(mppgpu_soe *h, ElmState *e, double dtime, mppgpu_elm_columns *cols, int *nfailed, int *nattempts);

extern "C" int mppgpu_vsfm_elm_solve(mppgpu_handle h, double dtime, int nstep, mppgpu_elm_columns *cols, int *nfailed, int *nattempts)
{
  CHECK_H(h);
  (void)nstep;
  if (elm_need(h)) return 1;
  ElmState *e = h->elm;
  if (!e->geometry_set) return fail("mppgpu_vsfm_elm_solve: call mppgpu_vsfm_elm_set_geometry first");
  if (!h->mesh_set || !h->soils_set) return fail("mppgpu_vsfm_elm_solve: mesh and soils must be set first");
  if (!cols) return fail("mppgpu_vsfm_elm_solve: null column arrays");
  if (!(dtime > 0.0)) return fail("mppgpu_vsfm_elm_solve: dtime must be positive");
  if (h->nlev > 32) return fail("mppgpu_vsfm_elm_solve: columns taller than 32 layers are not supported on this path");
  if (!h->bcs.empty()) return fail("mppgpu_vsfm_elm_solve: the lateral / seepage branches of MPPVSFMALM_Solve are not part of the 1-D path (no boundary conditions)");
  const bool patches = cols->npft > 0;
  if (patches && (!cols->col_pfti || !cols->col_npfts || !cols->pft_active || !cols->pft_wtcol || !cols->rootr_pft || !cols->qflx_tran_veg_pft))
    return fail("mppgpu_vsfm_elm_solve: npft > 0 needs all patch-level arrays");
  const void *req[] = {cols->rootr_col, cols->qflx_tran_veg_col, cols->qflx_infl, cols->qflx_dew_snow, cols->qflx_dew_grnd, cols->qflx_sub_snow,
                       cols->frac_h2osfc, cols->snl, cols->qflx_drain, cols->zwt, cols->h2osoi_liq, cols->h2osoi_ice, cols->mflx_snowlyr_col,
                       cols->mflx_neg_snow_col, cols->mflx_drain_perched, cols->smp_l, cols->soilp_col, cols->qcharge};
  for (const void *q : req) if (!q) return fail("mppgpu_vsfm_elm_solve: null column array");
  double *const x_before = h->x_current, *const x_committed_before = h->x_committed;
  if (elm_solve_pipeline(h, e, dtime, cols, nfailed, nattempts)) {
    // drain every stream that may still touch the caller's arrays, and put the handle back where it was
    const std::string msg = g_err;
    if (h->copy_in) cudaStreamSynchronize(h->copy_in);
    if (h->copy_out) cudaStreamSynchronize(h->copy_out);
    cudaStreamSynchronize(h->stream);
    (void)cudaGetLastError();
    h->x_current = x_before; h->x_committed = x_committed_before; h->order_valid = false; h->result_pending = false;
    g_err = msg;
    return 1;
  }
  return 0;
}

// The solve proper, software-pipelined over column chunks like mppgpu_vsfm_coupled_step: chunk k's arrays go up on `copy_in`, are packed,
// stepped once and judged on the handle's stream, and come down on `copy_out` while chunk k+1 computes and chunk k+2 uploads.  Columns that
// the first StepDT did not settle (diverged, or mass error >= 1e-5 kg) are rare; they are finished by the whole-batch retry loop after the
// pipeline, and only then are the result arrays downloaded a second time.
static int elm_solve_pipeline(mppgpu_soe *h, ElmState *e, double dtime, mppgpu_elm_columns *cols, int *nfailed, int *nattempts)
{
  const bool patches = cols->npft > 0;
  const size_t ncol = h->ncol, N = h->ncells;
  const int nlev = h->nlev;
  cudaStream_t s = h->stream;
  const bool fo = cols->fortran_order != 0;
  if (!h->copy_in)  CK(cudaStreamCreateWithFlags(&h->copy_in, cudaStreamNonBlocking));
  if (!h->copy_out) CK(cudaStreamCreateWithFlags(&h->copy_out, cudaStreamNonBlocking));
  if (!h->ev_out_done) { CK(cudaEventCreateWithFlags(&h->ev_out_done, cudaEventDisableTiming)); CK(cudaEventCreateWithFlags(&h->ev_start, cudaEventDisableTiming)); }
  // chunks of whole launch-order blocks; small batches run as one chunk
  int nchunks = e->nchunks > 0 ? e->nchunks : 8;
  const long long align = 1024;
  long long per = ((long long)ncol + nchunks - 1) / nchunks;
  per = ((per + align - 1) / align) * align;
  if (e->nchunks <= 0 && per < 32768) per = 32768;                      // default: no chunk smaller than a few waves of the step kernel
  nchunks = (int)(((long long)ncol + per - 1) / per);
  while ((int)h->ev_in.size() < nchunks) {
    cudaEvent_t a, b; CK(cudaEventCreateWithFlags(&a, cudaEventDisableTiming)); CK(cudaEventCreateWithFlags(&b, cudaEventDisableTiming));
    h->ev_in.push_back(a); h->ev_comp.push_back(b);
  }
  // Fortran-order arrays pass through device staging tables of the caller's own layout: 3 going up, 5 coming down
  if (fo && e->stage.n < 8 * N) CK(e->stage.alloc(8 * N));
  double *const st_in[3] = {e->stage.p, e->stage.p + N, e->stage.p + 2 * N};
  double *const st_out[5] = {e->stage.p + 3 * N, e->stage.p + 4 * N, e->stage.p + 5 * N, e->stage.p + 6 * N, e->stage.p + 7 * N};
  long long total_blocks = 0;
  for (int k = 0; k < nchunks; ++k) total_blocks += vsfm_blocks_for(h, std::min<long long>(per, (long long)ncol - k * per));
  const int cb = nblk(ncol, 256);
  const size_t need_partials = (size_t)std::max<long long>(total_blocks, cb) * 9;
  if (h->block_partials.n < need_partials) CK(h->block_partials.alloc(need_partials));

  CK(cudaEventRecord(h->ev_start, s));
  CK(cudaStreamWaitEvent(h->copy_in, h->ev_start, 0));
  CK(cudaStreamWaitEvent(h->copy_out, h->ev_start, 0));
  if (patches) {
    const size_t np = cols->npft;
    if (e->npft_cap < np) {
      CK(e->pfti.alloc(ncol)); CK(e->npfts.alloc(ncol)); CK(e->pactive.alloc(np)); CK(e->wtcol.alloc(np)); CK(e->rootr_pft.alloc(np * nlev)); CK(e->qtran_pft.alloc(np));
      e->npft_cap = np;
    }
  }
  ElmArgs E;
  memset(&E, 0, sizeof(E));
  E.ncol = h->ncol; E.nlev = nlev; E.nlevsoi = e->nlevsoi; E.max_patch_per_col = cols->max_patch_per_col;
  E.col0 = 0; E.col_end = h->ncol;
  E.dtime = dtime; E.watmin = e->watmin; E.rtol0 = h->so.rtol; E.stol0 = h->so.stol;
  E.active = h->has_active ? h->active.p : nullptr;
  if (patches) { E.col_pfti = e->pfti.p; E.col_npfts = e->npfts.p; E.pft_active = e->pactive.p; E.pft_wtcol = e->wtcol.p; E.rootr_pft = e->rootr_pft.p; E.qflx_tran_veg_pft = e->qtran_pft.p; }
  E.rootr_col = e->rootr.p; E.qflx_tran_veg_col = e->qtran.p; E.qflx_infl = e->qinfl.p; E.qflx_dew_snow = e->dews.p; E.qflx_dew_grnd = e->dewg.p;
  E.qflx_sub_snow = e->subs.p; E.frac_h2osfc = e->fh2osfc.p; E.snl = e->snl.p; E.qflx_drain = e->qdrain.p; E.zwt = e->zwt.p; E.zi = e->zi.p; E.dz = e->dz.p;
  E.h2osoi_liq = e->liq.p; E.h2osoi_ice = e->ice.p; E.mflx_snowlyr_col = e->snowlyr.p; E.mflx_neg_snow = e->negsnow.p; E.mflx_drain_perched = e->perched.p;
  double **cbp[6] = {&E.c_infl, &E.c_et, &E.c_dew, &E.c_drain, &E.c_snow, &E.c_sub};
  for (int k = 0; k < 6; ++k) *cbp[k] = find_cond(h, AUXVAR_SS, e->cond_ids[k])->value.p;
  E.frac_liq = h->frac_liq.p; E.soe_mass = h->mass.p; E.soe_smp = h->smp.p; E.soe_pressure = h->pressure.p;
  E.frac_ice = e->frac_ice.p; E.mass_beg = e->mass_beg.p; E.tot_flux = e->tot_flux.p; E.dt_rem = e->dt_rem.p; E.rtol = e->rtol.p; E.stol = e->stol.p;
  E.t_done = e->t_done.p; E.iter_count = e->iter_count.p; E.diverged = e->diverged.p; E.mask = e->mask.p; E.status = e->status.p;
  E.stat_reason = h->stat_reason.p; E.pending = e->pending.p; E.retry_list = e->retry_list.p;
  E.smp_l = e->smp_l.p; E.soilp = e->soilp.p; E.qcharge = e->qcharge.p; E.abs_err = e->abs_err.p;

  // per-cell arrays: { device cell-ordered buffer, host array, staging table (Fortran order only) }
  struct CellXfer { double *dev; double *host; double *stage; };
  const CellXfer up_cells[3] = {{e->rootr.p, cols->rootr_col, st_in[0]}, {e->liq.p, cols->h2osoi_liq, st_in[1]}, {e->ice.p, cols->h2osoi_ice, st_in[2]}};
  const CellXfer down_cells[5] = {{e->liq.p, cols->h2osoi_liq, st_out[0]}, {e->ice.p, cols->h2osoi_ice, st_out[1]}, {e->smp_l.p, cols->smp_l, st_out[2]},
                                  {e->soilp.p, cols->soilp_col, st_out[3]}, {e->rootr.p, cols->rootr_col, st_out[4]}};
  const int ndown = patches ? 5 : 4;
  struct ColXfer { void *dev; const void *host; size_t elem; };
  const ColXfer up_col[12] = {{e->qtran.p, cols->qflx_tran_veg_col, 8}, {e->qinfl.p, cols->qflx_infl, 8}, {e->dews.p, cols->qflx_dew_snow, 8}, {e->dewg.p, cols->qflx_dew_grnd, 8},
                              {e->subs.p, cols->qflx_sub_snow, 8}, {e->fh2osfc.p, cols->frac_h2osfc, 8}, {e->snl.p, cols->snl, sizeof(int)}, {e->qdrain.p, cols->qflx_drain, 8},
                              {e->zwt.p, cols->zwt, 8}, {e->snowlyr.p, cols->mflx_snowlyr_col, 8}, {e->negsnow.p, cols->mflx_neg_snow_col, 8}, {nullptr, nullptr, 0}};
  const ColXfer down_col[7] = {{e->qdrain.p, cols->qflx_drain, 8}, {e->zwt.p, cols->zwt, 8}, {e->snowlyr.p, cols->mflx_snowlyr_col, 8}, {e->qcharge.p, cols->qcharge, 8},
                               {e->abs_err.p, cols->abs_mass_error, 8}, {e->iter_count.p, cols->iter_count, sizeof(int)}, {e->status.p, cols->status, sizeof(int)}};
  // device -> host of columns [c0, c0 + n) on `q` (the layout kernels, if any, have already run on the compute stream)
  auto download = [&](long long c0, int n, cudaStream_t q) -> int {
    for (int a = 0; a < ndown; ++a) {
      const CellXfer &x = down_cells[a];
      cudaError_t err;
      if (!fo) err = cudaMemcpyAsync(x.host + c0 * nlev, x.dev + c0 * nlev, (size_t)n * nlev * 8, cudaMemcpyDeviceToHost, q);
      else     err = cudaMemcpy2DAsync(x.host + c0, ncol * 8, x.stage + c0, ncol * 8, (size_t)n * 8, nlev, cudaMemcpyDeviceToHost, q);
      if (err != cudaSuccess) return 1;
    }
    for (const ColXfer &x : down_col) if (x.host)
      if (cudaMemcpyAsync((char *)x.host + c0 * x.elem, (const char *)x.dev + c0 * x.elem, (size_t)n * x.elem, cudaMemcpyDeviceToHost, q) != cudaSuccess) return 1;
    return 0;
  };
  auto to_host_layout = [&](long long c0, int n) -> int {
    if (!fo) return 0;
    for (int a = 0; a < ndown; ++a)
      transpose_from_cells_range_kernel<<<nblk((long long)n * nlev, 256), 256, 0, s>>>(down_cells[a].dev, down_cells[a].stage, h->ncol, nlev, (int)c0, n);
    h->launches += ndown;
    return cudaGetLastError() != cudaSuccess;
  };

  CK(cudaEventRecord(h->ev0, s));
  // ---- PreStepDT (:603) ----
  h->x_current = h->x_committed;
  VsfmArgs A;
  vsfm_fill_args(h, A, dtime);
  A.block_partials = h->block_partials.p;
  A.t_done = e->t_done.p;
  A.x_in = h->x_current;                                                // first StepDT: every column, the handle's tolerances, the common specialisation
  A.x_out = (h->x_committed == h->xA.p) ? h->xB.p : h->xA.p;
  A.order = (h->order_valid && h->order_per == per && h->order_chunks == nchunks) ? h->order.p : nullptr;
  CK(cudaMemsetAsync(e->pending.p, 0, sizeof(int), s));
  long long block0 = 0;
  for (int k = 0; k < nchunks; ++k) {
    const long long c0 = k * per; const int n = (int)std::min<long long>(per, (long long)ncol - c0);
    // ---- host -> device: ELM's raw arrays ----
    for (const CellXfer &x : up_cells) {
      if (!fo) CK(cudaMemcpyAsync(x.dev + c0 * nlev, x.host + c0 * nlev, (size_t)n * nlev * 8, cudaMemcpyHostToDevice, h->copy_in));
      else     CK(cudaMemcpy2DAsync(x.stage + c0, ncol * 8, x.host + c0, ncol * 8, (size_t)n * 8, nlev, cudaMemcpyHostToDevice, h->copy_in));
    }
    CK(cudaMemcpyAsync(e->perched.p + c0 * nlev, cols->mflx_drain_perched + c0 * nlev, (size_t)n * nlev * 8, cudaMemcpyHostToDevice, h->copy_in));
    if (patches) {
      // the patch tables are indexed through col%pfti: this chunk needs the patches [lo, hi) its columns point at (ELM orders patches by
      // column, so consecutive chunks take consecutive slices; any other order only makes the slices overlap)
      long long lo = cols->npft, hi = 0;
      for (long long c = c0; c < c0 + n; ++c) {
        const long long a = cols->col_pfti[c], b = a + cols->col_npfts[c];
        if (cols->col_npfts[c] > 0) { lo = std::min(lo, a); hi = std::max(hi, b); }
      }
      if (hi > lo && (lo < 0 || hi > (long long)cols->npft)) return fail("mppgpu_vsfm_elm_solve: col_pfti / col_npfts point outside the %d patches", cols->npft);
      CK(cudaMemcpyAsync(e->pfti.p + c0, cols->col_pfti + c0, (size_t)n * sizeof(int), cudaMemcpyHostToDevice, h->copy_in));
      CK(cudaMemcpyAsync(e->npfts.p + c0, cols->col_npfts + c0, (size_t)n * sizeof(int), cudaMemcpyHostToDevice, h->copy_in));
      if (hi > lo) {
        const size_t m = (size_t)(hi - lo);
        CK(cudaMemcpyAsync(e->pactive.p + lo, cols->pft_active + lo, m * sizeof(int), cudaMemcpyHostToDevice, h->copy_in));
        CK(cudaMemcpyAsync(e->wtcol.p + lo, cols->pft_wtcol + lo, m * 8, cudaMemcpyHostToDevice, h->copy_in));
        CK(cudaMemcpyAsync(e->qtran_pft.p + lo, cols->qflx_tran_veg_pft + lo, m * 8, cudaMemcpyHostToDevice, h->copy_in));
        CK(cudaMemcpyAsync(e->rootr_pft.p + lo * nlev, cols->rootr_pft + lo * nlev, m * nlev * 8, cudaMemcpyHostToDevice, h->copy_in));
      }
    }
    for (const ColXfer &x : up_col) if (x.host)
      CK(cudaMemcpyAsync((char *)x.dev + c0 * x.elem, (const char *)x.host + c0 * x.elem, (size_t)n * x.elem, cudaMemcpyHostToDevice, h->copy_in));
    CK(cudaEventRecord(h->ev_in[k], h->copy_in));
    CK(cudaStreamWaitEvent(s, h->ev_in[k], 0));
    if (fo) {
      for (const CellXfer &x : up_cells) transpose_to_cells_range_kernel<<<nblk((long long)n * nlev, 256), 256, 0, s>>>(x.stage, x.dev, h->ncol, nlev, (int)c0, n);
      CK(cudaGetLastError());
      h->launches += 3;
    }
    E.col0 = (int)c0; E.col_end = (int)c0 + n;
    if (nlev <= 16) elm_pack_kernel<16><<<nblk((long long)n * 16, 128), 128, 0, s>>>(E); else elm_pack_kernel<32><<<nblk((long long)n * 32, 128), 128, 0, s>>>(E);
    CK(cudaGetLastError());
    if (vsfm_launch_range(h, A, c0, n, block0, s)) return 1;
    if (vsfm_build_order(h, c0, n, s)) return 1;
    if (nlev <= 16) elm_decide_kernel<16><<<nblk((long long)n * 16, 128), 128, 0, s>>>(E); else elm_decide_kernel<32><<<nblk((long long)n * 32, 128), 128, 0, s>>>(E);
    CK(cudaGetLastError());
    h->launches += 2;
    if (to_host_layout(c0, n)) return fail("mppgpu_vsfm_elm_solve: layout kernel failed");
    CK(cudaEventRecord(h->ev_comp[k], s));
    // ---- device -> host: this chunk's results as the first StepDT left them ----
    CK(cudaStreamWaitEvent(h->copy_out, h->ev_comp[k], 0));
    if (download(c0, n, h->copy_out)) return fail("mppgpu_vsfm_elm_solve: download failed");
    block0 += vsfm_blocks_for(h, n);
  }
  CK(cudaEventRecord(h->ev_out_done, h->copy_out));
  h->order_valid = (h->ordering != 0 && nlev <= 32); h->order_chunks = nchunks; h->order_per = per;
  h->x_current = A.x_out;
  int attempts = 1, pending = 0;
  CK(cudaMemcpyAsync(&pending, e->pending.p, sizeof(int), cudaMemcpyDeviceToHost, s));
  CK(cudaStreamSynchronize(s));
  // ---- the retry loop (:628-923) for the columns the decision kernel marked ----
  const bool retried = pending > 0;
  E.col0 = 0; E.col_end = h->ncol;
  while (pending > 0 && attempts < 10) {
    VsfmArgs R;
    vsfm_fill_args(h, R, dtime);
    R.block_partials = h->block_partials.p;
    R.t_done = e->t_done.p;
    // only the marked columns, each with its own remaining time / tolerances / start vector, in place
    R.retry_mask = e->mask.p; R.dt_col = e->dt_rem.p; R.rtol_col = e->rtol.p; R.stol_col = e->stol.p; R.x_redo = h->x_committed;
    R.x_in = h->x_current; R.x_out = h->x_current;
    R.retry_list = e->retry_list.p; R.nretry = pending;
    // sized by the columns that need it: a handful of warps, not a pass over the whole batch
    if (nlev <= 16) launch_vsfm2<8>(h, R, nblk((long long)pending * 8, VSFM2_THREADS));
    else            launch_vsfm2<16>(h, R, nblk((long long)pending * 16, VSFM2_THREADS));
    CK(cudaGetLastError());
    attempts++;
    CK(cudaMemsetAsync(e->pending.p, 0, sizeof(int), s));
    if (nlev <= 16) elm_decide_kernel<16><<<nblk(ncol * 16, 128), 128, 0, s>>>(E); else elm_decide_kernel<32><<<nblk(ncol * 32, 128), 128, 0, s>>>(E);
    CK(cudaGetLastError());
    h->launches += 2;
    CK(cudaMemcpyAsync(&pending, e->pending.p, sizeof(int), cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
  }
  // ---- PostStepDT (:935): soln_prev_clm = soln_prev ----
  h->x_committed = h->x_current;
  {
    // the handle's reductions for the whole solve (mppgpu_vsfm_mass_balance, mppgpu_reduction_buffer_device)
    elm_column_partials_kernel<<<cb, 256, 0, s>>>(h->ncol, dtime, h->has_active ? h->active.p : nullptr, e->mass_beg.p, h->col_mass.p, e->tot_flux.p,
                                                 e->abs_err.p, e->status.p, h->stat_its.p, h->stat_reason.p, h->stat_cuts.p, h->block_partials.p);
    CK(cudaGetLastError());
    h->launches += 1;
    VsfmArgs R; memset(&R, 0, sizeof(R)); R.x_out = h->x_current;
    if (vsfm_finish_step(h, R, cb)) return 1;
  }
  CK(cudaStreamWaitEvent(s, h->ev_out_done, 0));                        // the pipelined downloads are complete before anything below touches the host arrays
  if (retried) {
    // some columns changed after their chunk had gone down: send every result array again, in stream order after the first copies
    if (to_host_layout(0, h->ncol) || download(0, h->ncol, s)) return fail("mppgpu_vsfm_elm_solve: download failed");
  }
  CK(cudaMemsetAsync(e->pending.p + 1, 0, sizeof(int), s));
  elm_count_failed_kernel<<<cb, 256, 0, s>>>(h->ncol, h->has_active ? h->active.p : nullptr, e->status.p, e->pending.p + 1);
  CK(cudaGetLastError());
  h->launches += 1;
  CK(cudaEventRecord(h->ev1, s));
  int nf = 0;
  CK(cudaMemcpyAsync(&nf, e->pending.p + 1, sizeof(int), cudaMemcpyDeviceToHost, s));
  CK(cudaStreamSynchronize(s));
  { float ms_ = 0.0f; if (cudaEventElapsedTime(&ms_, h->ev0, h->ev1) == cudaSuccess) h->last_ms = ms_; else (void)cudaGetLastError(); }
  h->result_pending = false;
  if (nfailed) *nfailed = nf;
  if (nattempts) *nattempts = attempts;
  return 0;
}
