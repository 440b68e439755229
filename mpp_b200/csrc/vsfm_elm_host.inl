// vsfm_elm_host.inl -- host side of mppgpu_vsfm_elm_solve: MPPVSFMALM_Solve (src/driver/alm/MPPVSFMALM_Driver.F90:204-923) with ELM's raw
// column arrays on both sides.  Included at the end of mppgpu.cu.

struct ElmState {
  bool geometry_set = false;
  int nlevsoi = 0, cond_ids[6] = {0, 0, 0, 0, 0, 0};
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

static int elm_need(mppgpu_soe *h)
{
  if (h->soe_itype != MPPGPU_SOE_RE_ODE) return fail("mppgpu_vsfm_elm_*: handle is not a VSFM SoE");
  if (!h->elm) h->elm = new ElmState();
  return 0;
}

__global__ void transpose_from_cells_kernel(const double *__restrict__ cells, double *__restrict__ t, int ncol, int nlev)
{
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;        // i runs over the Fortran-order table: coalesced writes
  const long long n = (long long)ncol * nlev;
  if (i < n) { const int j = (int)(i / ncol), c = (int)(i % ncol); t[i] = cells[(size_t)c * nlev + j]; }
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
  CK(e->stage.alloc(N));
  DevBuf<double> *cellb[] = {&e->rootr, &e->liq, &e->ice, &e->perched, &e->frac_ice, &e->smp_l, &e->soilp};
  for (auto b : cellb) { CK(b->alloc(N)); CK(cudaMemsetAsync(b->p, 0, N * 8, h->stream)); }    // outputs of filtered-out columns read 0
  DevBuf<double> *colb[] = {&e->qtran, &e->qinfl, &e->dews, &e->dewg, &e->subs, &e->fh2osfc, &e->qdrain, &e->zwt, &e->snowlyr, &e->negsnow,
                            &e->mass_beg, &e->tot_flux, &e->dt_rem, &e->rtol, &e->stol, &e->t_done, &e->qcharge, &e->abs_err};
  for (auto b : colb) { CK(b->alloc(ncol)); CK(cudaMemsetAsync(b->p, 0, ncol * 8, h->stream)); }
  DevBuf<int> *coli[] = {&e->snl, &e->iter_count, &e->diverged, &e->mask, &e->status, &e->retry_list};
  for (auto b : coli) { CK(b->alloc(ncol)); CK(cudaMemsetAsync(b->p, 0, ncol * sizeof(int), h->stream)); }
  CK(e->pending.alloc(1));
  CK(cudaStreamSynchronize(h->stream));
  e->nlevsoi = nlevsoi; e->watmin = watmin; e->geometry_set = true;
  return 0;
}

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
  const size_t ncol = h->ncol, N = h->ncells;
  cudaStream_t s = h->stream;
  // ---- host -> device: ELM's raw arrays ----
#define UP(buf, src, cnt) CK(cudaMemcpyAsync((buf).p, (src), (cnt) * sizeof(*(buf).p), cudaMemcpyHostToDevice, s))
  const bool fo = cols->fortran_order != 0;
  // a per-cell array in ELM's (c, j) order goes through the staging buffer and a transpose into the cell order the kernels stream
  auto up_cells = [&](DevBuf<double> &buf, const double *src) -> int {
    if (!fo) return cudaMemcpyAsync(buf.p, src, N * 8, cudaMemcpyHostToDevice, s) != cudaSuccess;
    if (cudaMemcpyAsync(e->stage.p, src, N * 8, cudaMemcpyHostToDevice, s) != cudaSuccess) return 1;
    transpose_to_cells_kernel<<<nblk(N, 256), 256, 0, s>>>(e->stage.p, buf.p, h->ncol, h->nlev);
    return cudaGetLastError() != cudaSuccess;
  };
  auto down_cells = [&](double *dst, DevBuf<double> &buf) -> int {
    if (!fo) return cudaMemcpyAsync(dst, buf.p, N * 8, cudaMemcpyDeviceToHost, s) != cudaSuccess;
    transpose_from_cells_kernel<<<nblk(N, 256), 256, 0, s>>>(buf.p, e->stage.p, h->ncol, h->nlev);
    if (cudaGetLastError() != cudaSuccess) return 1;
    return cudaMemcpyAsync(dst, e->stage.p, N * 8, cudaMemcpyDeviceToHost, s) != cudaSuccess;
  };
  if (patches) {
    const size_t np = cols->npft;
    if (e->npft_cap < np) {
      CK(e->pfti.alloc(ncol)); CK(e->npfts.alloc(ncol)); CK(e->pactive.alloc(np)); CK(e->wtcol.alloc(np)); CK(e->rootr_pft.alloc(np * h->nlev)); CK(e->qtran_pft.alloc(np));
      e->npft_cap = np;
    }
    UP(e->pfti, cols->col_pfti, ncol); UP(e->npfts, cols->col_npfts, ncol); UP(e->pactive, cols->pft_active, np); UP(e->wtcol, cols->pft_wtcol, np);
    UP(e->rootr_pft, cols->rootr_pft, np * h->nlev); UP(e->qtran_pft, cols->qflx_tran_veg_pft, np);
  }
  if (up_cells(e->rootr, cols->rootr_col)) return fail("mppgpu_vsfm_elm_solve: upload failed");   // with patches only layers 1..nlevsoi are recomputed (:206-240); the rest keeps ELM's values
  UP(e->qtran, cols->qflx_tran_veg_col, ncol); UP(e->qinfl, cols->qflx_infl, ncol); UP(e->dews, cols->qflx_dew_snow, ncol); UP(e->dewg, cols->qflx_dew_grnd, ncol);
  UP(e->subs, cols->qflx_sub_snow, ncol); UP(e->fh2osfc, cols->frac_h2osfc, ncol); UP(e->snl, cols->snl, ncol); UP(e->qdrain, cols->qflx_drain, ncol);
  UP(e->zwt, cols->zwt, ncol); UP(e->snowlyr, cols->mflx_snowlyr_col, ncol);
  if (up_cells(e->liq, cols->h2osoi_liq) || up_cells(e->ice, cols->h2osoi_ice)) return fail("mppgpu_vsfm_elm_solve: upload failed");
  UP(e->negsnow, cols->mflx_neg_snow_col, ncol); UP(e->perched, cols->mflx_drain_perched, N);
#undef UP
  ElmArgs E;
  memset(&E, 0, sizeof(E));
  E.ncol = h->ncol; E.nlev = h->nlev; E.nlevsoi = e->nlevsoi; E.max_patch_per_col = cols->max_patch_per_col;
  E.dtime = dtime; E.watmin = e->watmin; E.rtol0 = h->so.rtol; E.stol0 = h->so.stol;
  E.active = h->has_active ? h->active.p : nullptr;
  if (patches) { E.col_pfti = e->pfti.p; E.col_npfts = e->npfts.p; E.pft_active = e->pactive.p; E.pft_wtcol = e->wtcol.p; E.rootr_pft = e->rootr_pft.p; E.qflx_tran_veg_pft = e->qtran_pft.p; }
  E.rootr_col = e->rootr.p; E.qflx_tran_veg_col = e->qtran.p; E.qflx_infl = e->qinfl.p; E.qflx_dew_snow = e->dews.p; E.qflx_dew_grnd = e->dewg.p;
  E.qflx_sub_snow = e->subs.p; E.frac_h2osfc = e->fh2osfc.p; E.snl = e->snl.p; E.qflx_drain = e->qdrain.p; E.zwt = e->zwt.p; E.zi = e->zi.p; E.dz = e->dz.p;
  E.h2osoi_liq = e->liq.p; E.h2osoi_ice = e->ice.p; E.mflx_snowlyr_col = e->snowlyr.p; E.mflx_neg_snow = e->negsnow.p; E.mflx_drain_perched = e->perched.p;
  double **cb[6] = {&E.c_infl, &E.c_et, &E.c_dew, &E.c_drain, &E.c_snow, &E.c_sub};
  for (int k = 0; k < 6; ++k) *cb[k] = find_cond(h, AUXVAR_SS, e->cond_ids[k])->value.p;
  E.frac_liq = h->frac_liq.p; E.soe_mass = h->mass.p; E.soe_smp = h->smp.p; E.soe_pressure = h->pressure.p;
  E.frac_ice = e->frac_ice.p; E.mass_beg = e->mass_beg.p; E.tot_flux = e->tot_flux.p; E.dt_rem = e->dt_rem.p; E.rtol = e->rtol.p; E.stol = e->stol.p;
  E.t_done = e->t_done.p; E.iter_count = e->iter_count.p; E.diverged = e->diverged.p; E.mask = e->mask.p; E.status = e->status.p;
  E.stat_reason = h->stat_reason.p; E.pending = e->pending.p; E.retry_list = e->retry_list.p;
  E.smp_l = e->smp_l.p; E.soilp = e->soilp.p; E.qcharge = e->qcharge.p; E.abs_err = e->abs_err.p;

  CK(cudaEventRecord(h->ev0, s));
  if (h->nlev <= 16) elm_pack_kernel<16><<<nblk(ncol * 16, 128), 128, 0, s>>>(E); else elm_pack_kernel<32><<<nblk(ncol * 32, 128), 128, 0, s>>>(E);
  CK(cudaGetLastError());
  h->launches += 1;
  // ---- PreStepDT (:603) ----
  h->x_current = h->x_committed;
  const int nblocks = vsfm_blocks_for(h, h->ncol);
  if (h->block_partials.n < (size_t)nblocks * 9) CK(h->block_partials.alloc((size_t)nblocks * 9));
  int attempts = 0, pending = 0;
  for (;;) {
    VsfmArgs A;
    vsfm_fill_args(h, A, dtime);
    A.block_partials = h->block_partials.p;
    A.t_done = e->t_done.p;
    if (attempts == 0) {
      // first StepDT: every column, the handle's tolerances, the common specialisation of the step kernel
      A.x_in = h->x_current;
      A.x_out = (h->x_committed == h->xA.p) ? h->xB.p : h->xA.p;
    } else {
      // retries: only the columns the decision kernel marked, each with its own remaining time / tolerances / start vector, in place
      A.retry_mask = e->mask.p; A.dt_col = e->dt_rem.p; A.rtol_col = e->rtol.p; A.stol_col = e->stol.p; A.x_redo = h->x_committed;
      A.x_in = h->x_current; A.x_out = h->x_current;
      A.retry_list = e->retry_list.p; A.nretry = pending;
    }
    if (attempts == 0) { if (vsfm_launch_range(h, A, 0, h->ncol, 0, s)) return 1; }
    else {
      // sized by the columns that need it: a handful of warps, not a pass over the whole batch
      if (h->nlev <= 16) launch_vsfm2<8>(h, A, nblk((long long)pending * 8, VSFM2_THREADS));
      else               launch_vsfm2<16>(h, A, nblk((long long)pending * 16, VSFM2_THREADS));
      CK(cudaGetLastError());
      h->launches += 1;
    }
    h->x_current = A.x_out;
    attempts++;
    CK(cudaMemsetAsync(e->pending.p, 0, sizeof(int), s));
    if (h->nlev <= 16) elm_decide_kernel<16><<<nblk(ncol * 16, 128), 128, 0, s>>>(E); else elm_decide_kernel<32><<<nblk(ncol * 32, 128), 128, 0, s>>>(E);
    CK(cudaGetLastError());
    h->launches += 1;
    CK(cudaMemcpyAsync(&pending, e->pending.p, sizeof(int), cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    if (pending == 0 || attempts >= 10) break;
  }
  // ---- PostStepDT (:935): soln_prev_clm = soln_prev ----
  h->x_committed = h->x_current;
  {
    // the handle's reductions for the whole solve (mppgpu_vsfm_mass_balance, mppgpu_reduction_buffer_device)
    const int cb = nblk(ncol, 256);
    elm_column_partials_kernel<<<cb, 256, 0, s>>>(h->ncol, dtime, h->has_active ? h->active.p : nullptr, e->mass_beg.p, h->col_mass.p, e->tot_flux.p,
                                                 e->abs_err.p, e->status.p, h->stat_its.p, h->stat_reason.p, h->stat_cuts.p, h->block_partials.p);
    CK(cudaGetLastError());
    h->launches += 1;
    VsfmArgs R; memset(&R, 0, sizeof(R)); R.x_out = h->x_current;
    if (vsfm_finish_step(h, R, cb)) return 1;
  }
  CK(cudaEventRecord(h->ev1, s));
  // ---- device -> host: ELM's raw arrays ----
#define DOWN(dst, buf, cnt) CK(cudaMemcpyAsync((dst), (buf).p, (cnt) * sizeof(*(buf).p), cudaMemcpyDeviceToHost, s))
  // (the staging buffer is reused array by array: stream order keeps transpose -> copy -> next transpose apart)
  if (patches && down_cells(cols->rootr_col, e->rootr)) return fail("mppgpu_vsfm_elm_solve: download failed");
  if (down_cells(cols->h2osoi_liq, e->liq) || down_cells(cols->h2osoi_ice, e->ice) || down_cells(cols->smp_l, e->smp_l) || down_cells(cols->soilp_col, e->soilp))
    return fail("mppgpu_vsfm_elm_solve: download failed");
  DOWN(cols->qflx_drain, e->qdrain, ncol); DOWN(cols->zwt, e->zwt, ncol);
  DOWN(cols->mflx_snowlyr_col, e->snowlyr, ncol); DOWN(cols->qcharge, e->qcharge, ncol);
  if (cols->abs_mass_error) DOWN(cols->abs_mass_error, e->abs_err, ncol);
  if (cols->iter_count) DOWN(cols->iter_count, e->iter_count, ncol);
  std::vector<int> status(ncol);
  CK(cudaMemcpyAsync(status.data(), e->status.p, ncol * sizeof(int), cudaMemcpyDeviceToHost, s));
#undef DOWN
  CK(cudaStreamSynchronize(s));
  { float ms_ = 0.0f; if (cudaEventElapsedTime(&ms_, h->ev0, h->ev1) == cudaSuccess) h->last_ms = ms_; else (void)cudaGetLastError(); }
  int nf = 0;
  for (size_t c = 0; c < ncol; ++c) {
    if (!status[c]) nf++;
    if (cols->status) cols->status[c] = status[c];
  }
  if (h->has_active) {   // filtered-out columns are not failures
    std::vector<int> act(ncol);
    CK(cudaMemcpy(act.data(), h->active.p, ncol * sizeof(int), cudaMemcpyDeviceToHost));
    nf = 0;
    for (size_t c = 0; c < ncol; ++c) if (act[c] && !status[c]) nf++;
  }
  if (nfailed) *nfailed = nf;
  if (nattempts) *nattempts = attempts;
  return 0;
}
