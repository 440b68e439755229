// thermal_host.inl -- host side of the soil thermal SoE (sysofeqns_thermal_type, soil governing equation only):
// device-resident mailbox, set/get routing, StepDT launch.  Included at the end of mppgpu.cu.

static int th_alloc_d(double **p, size_t n, double fill, cudaStream_t s)
{
  if (mpp_dmalloc((void **)p, n * sizeof(double)) != cudaSuccess) return 1;
  if (fill == 0.0) cudaMemsetAsync(*p, 0, n * sizeof(double), s);
  else fill_kernel<<<nblk(n, 256), 256, 0, s>>>(*p, fill, (long long)n);
  return 0;
}
__global__ void fill_int_kernel(int *p, int v, long long n)
{
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = v;
}

static int thermal_create(ThermalState *t, int ncol, int nlev, cudaStream_t s)
{
  t->stream = s; t->ncol = ncol; t->nlev = nlev;
  const size_t N = (size_t)ncol * nlev;
  int rc = 0;
  rc |= th_alloc_d(&t->T_clm, N, 273.15, s); rc |= th_alloc_d(&t->T_work, N, 273.15, s);
  rc |= th_alloc_d(&t->liq, N, 0.0, s); rc |= th_alloc_d(&t->ice, N, 0.0, s); rc |= th_alloc_d(&t->snow_water, N, 0.0, s);
  rc |= th_alloc_d(&t->tuning, N, 1.0, s);                      // ThermalKSPTemperatureSoilAuxType.F90:56
  if (mpp_dmalloc((void **)&t->nsnow, N * sizeof(int)) != cudaSuccess) rc = 1;
  if (mpp_dmalloc((void **)&t->active, N * sizeof(int)) != cudaSuccess) rc = 1;
  if (!rc) { cudaMemsetAsync(t->nsnow, 0, N * sizeof(int), s); cudaMemsetAsync(t->active, 0, N * sizeof(int), s); }
  t->T_cur = t->T_clm;
  return rc;
}

static void thermal_destroy(ThermalState *t)
{
  double *d[] = {t->por, t->tkmg, t->tkdry, t->csol, t->dist_up, t->dist_dn, t->T_clm, t->T_work, t->liq, t->ice, t->snow_water,
                 t->tuning, t->frac, t->aux_dz, t->aux_dist_up, t->aux_dist_dn, t->therm_cond, t->heat_cap, t->work};
  for (double *p : d) if (p) cudaFree(p);
  if (t->lun_type) cudaFree(t->lun_type);
  if (t->nsnow) cudaFree(t->nsnow);
  if (t->active) cudaFree(t->active);
  double *e[] = {t->soil_top_dist_dn, t->hs[0], t->hs[1], t->hs[2], t->dhs[0], t->dhs[1], t->dhs[2], t->frac_soil, t->sabg_snow, t->sabg_soil};
  for (double *p : e) if (p) cudaFree(p);
  if (t->snow_top_id) cudaFree(t->snow_top_id);
  if (t->elm_stage) cudaFree(t->elm_stage);
  if (t->elm_snl) cudaFree(t->elm_snl);
}

// MPPThermalTBasedALM_Initialize.F90:150-813 in one call: snow mesh (nlevsno layers) + standing-water mesh (1 cell) next to the soil
// mesh, the three governing equations, their five conditions (heat flux at the top of snow / standing water / soil; absorbed
// solar radiation on ALL_CELLS of snow and soil), the two coupling conditions snow<->soil and ssw<->soil with their coupling
// variables, and the dist_dn = z(c,1) - zi(c,0) poked into the soil's coupling conditions (:630-639).
static int thermal_add_snow_ssw(mppgpu_soe *h, ThermalState *t, int nlevsno, const double *soil_top_dist_dn)
{
  if (!h->mesh_set) return fail("mppgpu_thermal_add_snow_ssw: set the soil mesh first");
  if (t->snow_mode) return fail("mppgpu_thermal_add_snow_ssw: already added");
  if (!h->bcs.empty() || !h->sss.empty()) return fail("mppgpu_thermal_add_snow_ssw: the ELM configuration brings its own conditions; add none before");
  if (h->orientation == MPPGPU_MESH_AGAINST_GRAVITY) return fail("mppgpu_thermal_add_snow_ssw: the ELM thermal meshes are MESH_ALONG_GRAVITY");
  if (nlevsno < 0 || (nlevsno + 1) / 2 + (h->nlev + 1) / 2 > 16)
    return fail("mppgpu_thermal_add_snow_ssw: %d snow + %d soil layers per column; at most 32 rows (each block rounded up to even) are supported", nlevsno, h->nlev);
  if (!soil_top_dist_dn) return fail("mppgpu_thermal_add_snow_ssw: null soil_top_dist_dn");
  const size_t ncol = h->ncol, NA = ncol * (size_t)(nlevsno + 1 + h->nlev), NS = ncol * (size_t)nlevsno, NG = h->ncells;
  cudaStream_t s = h->stream;
  // the internal mailbox grows from soil-only to [snow | ssw | soil]
  double **grow[] = {&t->T_clm, &t->T_work, &t->liq, &t->ice, &t->snow_water, &t->tuning, &t->frac, &t->aux_dz, &t->aux_dist_up, &t->aux_dist_dn};
  const double fillv[] = {273.15, 273.15, 0.0, 0.0, 0.0, 1.0, 0.0, 0.0, 0.0, 0.0};
  for (int i = 0; i < 10; ++i) { if (*grow[i]) cudaFree(*grow[i]); *grow[i] = nullptr; if (th_alloc_d(grow[i], NA, fillv[i], s)) return fail("out of device memory"); }
  cudaFree(t->nsnow); cudaFree(t->active); t->nsnow = t->active = nullptr;
  CK(mpp_dmalloc((void **)&t->nsnow, NA * sizeof(int))); CK(mpp_dmalloc((void **)&t->active, NA * sizeof(int)));
  CK(cudaMemsetAsync(t->nsnow, 0, NA * sizeof(int), s)); CK(cudaMemsetAsync(t->active, 0, NA * sizeof(int), s));
  // soil cells of active columns start active (MPPThermalSetSoils); snow / ssw cells wait for VAR_ACTIVE from the host model
  if (t->soils_set) fill_int_kernel<<<nblk(NG, 256), 256, 0, s>>>(t->active + (NA - NG), 1, (long long)NG);
  for (int k = 0; k < 3; ++k) { if (th_alloc_d(&t->hs[k], ncol, 0.0, s) || th_alloc_d(&t->dhs[k], ncol, 0.0, s)) return fail("out of device memory"); }
  if (th_alloc_d(&t->frac_soil, ncol, 0.0, s) || th_alloc_d(&t->sabg_snow, NS ? NS : 1, 0.0, s) || th_alloc_d(&t->sabg_soil, NG, 0.0, s) ||
      th_alloc_d(&t->soil_top_dist_dn, ncol, 0.0, s)) return fail("out of device memory");
  CK(mpp_dmalloc((void **)&t->snow_top_id, ncol * sizeof(int))); CK(cudaMemsetAsync(t->snow_top_id, 0, ncol * sizeof(int), s));
  CK(cudaMemcpyAsync(t->soil_top_dist_dn, soil_top_dist_dn, ncol * 8, cudaMemcpyHostToDevice, s));
  CK(cudaStreamSynchronize(s));
  t->T_cur = t->T_clm; t->snow_mode = true; t->nsno = nlevsno; t->nall = NA;
  return 0;
}

static int thermal_set_mesh(ThermalState *t, int orientation, const double *d_dz, const double *d_area)
{
  t->orientation = orientation; t->d_dz = d_dz; t->d_area = d_area;
  // the stale `area` of the reference's Dirichlet branch = area of the mesh's last internal connection = last column's
  double a = 1.0;
  if (cudaMemcpy(&a, d_area + (t->ncol - 1), sizeof(double), cudaMemcpyDeviceToHost) != cudaSuccess) return 1;
  t->stale_area = a;
  return 0;
}

static int thermal_set_temperature(ThermalState *t, const double *T, bool)
{
  // ThermalSOESetSolnPrevCLM (SystemOfEquationsThermalType.F90:171-199)
  const size_t n = t->snow_mode ? t->nall : (size_t)t->ncol * t->nlev;
  if (cudaMemcpyAsync(t->T_clm, T, n * sizeof(double), cudaMemcpyHostToDevice, t->stream) != cudaSuccess) return 1;
  return cudaStreamSynchronize(t->stream) != cudaSuccess;
}

static int thermal_set_soils(mppgpu_soe *h, ThermalState *t, const double *watsat, const double *csol, const double *tkmg,
                             const double *tkdry, const int *lun_type, int nlevsoi, int istsoil)
{
  if (!watsat || !csol || !tkmg || !tkdry || !lun_type) return fail("MPPThermalSetSoils: null table");
  const size_t N = h->ncells;
  const double *src[4] = {watsat, tkmg, tkdry, csol};
  double **dst[4] = {&t->por, &t->tkmg, &t->tkdry, &t->csol};
  for (int i = 0; i < 4; ++i) {
    DevBuf<double> tmp;
    if (upload_table(h, src[i], tmp)) return 1;
    if (!*dst[i]) CK(mpp_dmalloc((void **)dst[i], N * sizeof(double)));
    transpose_to_cells_kernel<<<nblk(N, 256), 256, 0, h->stream>>>(tmp.p, *dst[i], h->ncol, h->nlev);
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(h->stream));
  }
  if (!t->lun_type) CK(mpp_dmalloc((void **)&t->lun_type, h->ncol * sizeof(int)));
  CK(cudaMemcpyAsync(t->lun_type, lun_type, h->ncol * sizeof(int), cudaMemcpyHostToDevice, h->stream));
  // every column active (filter_thermal = 1): aux_vars_in%is_active = .true. (MultiPhysicsProbThermal.F90:165-170)
  fill_int_kernel<<<nblk(N, 256), 256, 0, h->stream>>>(t->active + (t->snow_mode ? t->nall - N : 0), 1, (long long)N);
  CK(cudaStreamSynchronize(h->stream));
  t->nlevsoi = nlevsoi; t->istsoil = istsoil; t->soils_set = true; h->soils_set = true;
  return 0;
}

static int thermal_lazy(double **p, size_t n, double fill, cudaStream_t s)
{
  if (*p) return 0;
  return th_alloc_d(p, n, fill, s) ? fail("out of device memory") : 0;
}

static int thermal_field(mppgpu_soe *h, ThermalState *t, int auxvar_type, int var_type, int cond_id, bool for_set, double **p, size_t *cap)
{
  const size_t N = t->snow_mode ? t->nall : h->ncells;
  if (t->snow_mode && auxvar_type != AUXVAR_INTERNAL) {
    // SoE-level condition ids of the ELM configuration (MPPThermalTBasedALM_Driver.F90:395-441):
    //   AUXVAR_BC 1 top of snow, 2 top of standing water, 3 top of soil;  AUXVAR_SS 1 snow cells, 2 soil cells
    if (auxvar_type == AUXVAR_BC && cond_id >= 1 && cond_id <= 3) {
      *cap = h->ncol;
      if (var_type == VAR_BC_SS_CONDITION) { *p = t->hs[cond_id - 1]; return 0; }
      if (var_type == VAR_DHS_DT) { *p = t->dhs[cond_id - 1]; return 0; }
      if (var_type == VAR_FRAC && cond_id == 3) { *p = t->frac_soil; return 0; }
      return fail("SOEThermalAux%sRData: unknown var_type %d for boundary condition %d", for_set ? "Set" : "Get", var_type, cond_id);
    }
    if (auxvar_type == AUXVAR_SS && (cond_id == 1 || cond_id == 2) && var_type == VAR_BC_SS_CONDITION) {
      *cap = (cond_id == 1) ? (size_t)h->ncol * t->nsno : h->ncells; *p = (cond_id == 1) ? t->sabg_snow : t->sabg_soil; return 0;
    }
    return fail("ThermalSOE%sRDataFromCLM: no condition %d of auxvar type %d / var_type %d in the snow + standing-water + soil configuration",
                for_set ? "Set" : "Get", cond_id, auxvar_type, var_type);
  }
  if (auxvar_type == AUXVAR_INTERNAL) {
    *cap = N;
    switch (var_type) {
    case VAR_TEMPERATURE:   *p = for_set ? t->T_clm : t->T_cur; return 0;      // SetSolnPrevCLM / GetSoln
    case VAR_LIQ_AREAL_DEN: *p = t->liq; return 0;
    case VAR_ICE_AREAL_DEN: *p = t->ice; return 0;
    case VAR_SNOW_WATER:    *p = t->snow_water; return 0;
    case VAR_TUNING_FACTOR: *p = t->tuning; return 0;
    // stored for the SoE mailbox; read by the snow / standing-water equations and their coupling conditions
    case VAR_FRAC:          if (thermal_lazy(&t->frac, N, 0.0, h->stream)) return 1; *p = t->frac; return 0;
    case VAR_DZ:            if (thermal_lazy(&t->aux_dz, N, 0.0, h->stream)) return 1; *p = t->aux_dz; return 0;
    case VAR_DIST_UP:       if (thermal_lazy(&t->aux_dist_up, N, 0.0, h->stream)) return 1; *p = t->aux_dist_up; return 0;
    case VAR_DIST_DN:       if (thermal_lazy(&t->aux_dist_dn, N, 0.0, h->stream)) return 1; *p = t->aux_dist_dn; return 0;
    case VAR_THERMAL_COND:
    case VAR_HEAT_CAP:
      if (for_set) break;
      if (thermal_lazy(&t->therm_cond, N, 0.0, h->stream) || thermal_lazy(&t->heat_cap, N, 0.0, h->stream)) return 1;
      if (!t->diagnostics) return fail("thermal conductivity / heat capacity diagnostics are filled by the next StepDT (request them once before stepping)");
      *p = (var_type == VAR_THERMAL_COND) ? t->therm_cond : t->heat_cap; return 0;
    }
    return fail("SOEThermalAux%sRData: unknown var_type %d", for_set ? "Set" : "Get", var_type);
  }
  if (auxvar_type != AUXVAR_BC && auxvar_type != AUXVAR_SS) return fail("ThermalSOE%sRDataFromCLM: Unknown soe_auxvar_type %d", for_set ? "Set" : "Get", auxvar_type);
  HostCond *c = find_cond(h, auxvar_type, cond_id);
  if (!c) return fail("ThermalSOE%sRDataFromCLM: condition id %d out of range", for_set ? "Set" : "Get", cond_id);
  *cap = c->n;
  if (var_type == VAR_BC_SS_CONDITION) { *p = c->value.p; return 0; }
  if (auxvar_type == AUXVAR_BC && var_type == VAR_DHS_DT) { *p = c->dhsdT.p; return 0; }
  if (auxvar_type == AUXVAR_BC && var_type == VAR_FRAC) { *p = c->frac.p; return 0; }
  if (auxvar_type == AUXVAR_BC && var_type == VAR_ACTIVE && for_set) {
    // real-valued VAR_ACTIVE on boundary aux vars (ThermKSPTempSoilAuxVarSetRValues, thermal_mms_problem.F90:633): stored as 0/1 doubles
    if (!c->mass_exc.p) return fail("internal: boundary condition without an active-flag buffer");
    *p = c->mass_exc.p; return 0;
  }
  return fail("SOEThermalAux%sRData: unknown var_type %d", for_set ? "Set" : "Get", var_type);
}

static int thermal_set_idata(mppgpu_soe *h, ThermalState *t, int auxvar_type, int var_type, int cond_id, const int *data, int n)
{
  (void)cond_id;
  if (auxvar_type != AUXVAR_INTERNAL) return fail("ThermalSOESetIDataFromCLM: only AUXVAR_INTERNAL is supported");
  const size_t N = t->snow_mode ? t->nall : h->ncells;
  if ((size_t)n > N) return fail("size(data_1d) > nauxvar (%d > %zu)", n, N);
  int *dst = nullptr;
  if (var_type == VAR_NUM_SNOW_LYR) dst = t->nsnow;
  else if (var_type == VAR_ACTIVE) dst = t->active;
  else return fail("SOEThermalAuxSetIData: unknown var_type %d", var_type);
  CK(cudaMemcpyAsync(dst, data, (size_t)n * sizeof(int), cudaMemcpyHostToDevice, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  return 0;
}

static int thermal_pre_step_dt(ThermalState *t) { t->T_cur = t->T_clm; return 0; }     // ThermalSOEPreStepDT :393-408
static int thermal_post_step_dt(ThermalState *) { return 0; }

// queue the coupled snow / standing-water / soil step kernel on columns [c0, c0 + n) of the batch (mailbox pointers are whole-batch)
static int thermal_snow_launch(mppgpu_soe *h, ThermalState *t, double dt, long long c0, int n, double *T_out)
{
  ThermalSnowArgs A;
  memset(&A, 0, sizeof(A));
  ThermalArgs &S = A.S;
  S.ncol = h->ncol; S.nlev = h->nlev; S.nlevsoi = t->nlevsoi;
  S.istsoil = t->istsoil; S.istcrop = t->istcrop; S.istice = t->istice; S.istice_mec = t->istice_mec; S.istwet = t->istwet;
  S.dt = dt; S.cnfac = t->cnfac;
  S.por = t->por; S.tkmg = t->tkmg; S.tkdry = t->tkdry; S.csol = t->csol; S.dz = h->dz.p; S.area = h->area.p;
  S.dist_up = t->custom_dist ? t->dist_up : nullptr; S.dist_dn = t->custom_dist ? t->dist_dn : nullptr;
  S.dist_uniform = (t->custom_dist && t->dist_uniform) ? 1 : 0;
  if (S.dist_uniform) { memcpy(S.lay_du, t->lay_du, sizeof(S.lay_du)); memcpy(S.lay_dd, t->lay_dd, sizeof(S.lay_dd)); }
  S.lun_type = t->lun_type; S.stale_area = t->stale_area; S.top_is_first = 1;
  A.nsno = t->nsno;
  A.T_in = t->T_cur; A.liq = t->liq; A.ice = t->ice; A.snow_water = t->snow_water; A.mdz = t->aux_dz; A.dist_up = t->aux_dist_up;
  A.dist_dn = t->aux_dist_dn; A.tuning = t->tuning; A.frac = t->frac; A.nsnow = t->nsnow; A.active = t->active;
  for (int k = 0; k < 3; ++k) { A.hs[k] = t->hs[k]; A.dhsdT[k] = t->dhs[k]; }
  A.frac_soil = t->frac_soil; A.sabg_snow = t->sabg_snow; A.sabg_soil = t->sabg_soil; A.soil_top_dist_dn = t->soil_top_dist_dn;
  A.snow_top_id = t->snow_top_id;
  A.T_out = T_out;
  A.col0 = (int)c0; A.col_end = (int)c0 + n;
  // ELM's 5 + 15 layout (and anything else that fits 8 lanes of three rows): four columns per warp; otherwise 16 lanes of two rows
  if ((t->nsno + 2) / 3 + (h->nlev + 2) / 3 <= 8 && !t->force_two_rows)
    thermal_snow_step3_kernel<8><<<nblk((long long)n * 8, TH_TILE), TH_TILE, 0, h->stream>>>(A);
  else
    thermal_snow_step_kernel<16><<<nblk((long long)n * 16, TH_TILE), TH_TILE, 0, h->stream>>>(A);
  CK(cudaGetLastError());
  h->launches += 1;
  return 0;
}

static int thermal_snow_step(mppgpu_soe *h, ThermalState *t, double dt)
{
  double *T_out = (t->T_cur == t->T_clm) ? t->T_work : t->T_cur;
  CK(cudaEventRecord(h->ev0, h->stream));
  if (thermal_snow_launch(h, t, dt, 0, h->ncol, T_out)) return 1;
  CK(cudaEventRecord(h->ev1, h->stream));
  t->T_cur = T_out;
  return 0;
}

// which arrays a tile of the bulk-async kernel carries, in the order the kernel lays out its stage; false: shape not covered
static bool thermal_tma_plan(const ThermalArgs &A, ThermalTmaPlan &P)
{
  memset(&P, 0, sizeof(P));
  if (A.dist_up && !A.dist_uniform) return false;
  for (int k = 0; k < 2; ++k) if (A.bc_type[k] != 0 && A.bc_type[k] != 507) return false;
  int ncell_ss = 0, ncol_ss = 0;
  for (int k = 0; k < A.nss; ++k) { if (A.ss_region[k] == 403) ncell_ss++; else ncol_ss++; }
  if (ncell_ss > TMA_MAX_SS_CELL || ncol_ss > TMA_MAX_SS_COL) return false;
  const void *cd[] = {A.T_in, A.dz, A.tuning, A.liq, A.ice, A.snow_water, A.por, A.tkmg, A.tkdry, A.csol};
  for (const void *q : cd) P.cell_d[P.n_cell_d++] = q;
  for (int k = 0; k < A.nss; ++k) if (A.ss_region[k] == 403) P.cell_d[P.n_cell_d++] = A.ss_value[k];
  P.cell_i[P.n_cell_i++] = A.active; P.cell_i[P.n_cell_i++] = A.nsnow;
  P.col_d[P.n_col_d++] = A.area;
  for (int k = 0; k < 2; ++k) if (A.bc_type[k] == 507) { P.col_d[P.n_col_d++] = A.bc_value[k]; P.col_d[P.n_col_d++] = A.bc_dhsdT[k]; P.col_d[P.n_col_d++] = A.bc_frac[k]; }
  for (int k = 0; k < A.nss; ++k) if (A.ss_region[k] != 403) P.col_d[P.n_col_d++] = A.ss_value[k];
  P.col_i[P.n_col_i++] = A.lun_type;
  for (int i = 0; i < P.n_cell_d; ++i) if (!P.cell_d[i] || ((uintptr_t)P.cell_d[i] & 15)) return false;
  for (int i = 0; i < P.n_col_d; ++i) if (!P.col_d[i] || ((uintptr_t)P.col_d[i] & 15)) return false;
  if (!A.active || !A.nsnow || !A.lun_type || ((uintptr_t)A.active & 15) || ((uintptr_t)A.nsnow & 15) || ((uintptr_t)A.lun_type & 15)) return false;
  const int cells = TMA_TILE_COLS * A.nlev;
  P.stage_bytes = P.n_cell_d * cells * 8 + P.n_cell_i * cells * 4 + P.n_col_d * TMA_TILE_COLS * 8 + P.n_col_i * TMA_TILE_COLS * 4;
  P.ntiles = (A.ncol + TMA_TILE_COLS - 1) / TMA_TILE_COLS;
  if ((size_t)TMA_STAGES * P.stage_bytes > 100 * 1024) return false;
  // the last tile may hang over the end of the batch by up to TMA_TILE_COLS - 1 columns: covered by the allocation slack
  if ((size_t)(TMA_TILE_COLS - 1) * A.nlev * 8 > MPP_ALLOC_SLACK) return false;
  return true;
}

static int thermal_step(mppgpu_soe *h, ThermalState *t, double dt)
{
  if (!h->mesh_set || !t->soils_set) return fail("mppgpu_step_dt: mesh and soils must be set first");
  if (!(dt > 0.0)) return fail("mppgpu_step_dt: dt must be positive");
  if (t->snow_mode) return thermal_snow_step(h, t, dt);
  ThermalArgs A;
  memset(&A, 0, sizeof(A));
  A.ncol = h->ncol; A.nlev = h->nlev; A.nlevsoi = t->nlevsoi;
  A.istsoil = t->istsoil; A.istcrop = t->istcrop; A.istice = t->istice; A.istice_mec = t->istice_mec; A.istwet = t->istwet;
  A.dt = dt; A.cnfac = t->cnfac;
  A.por = t->por; A.tkmg = t->tkmg; A.tkdry = t->tkdry; A.csol = t->csol; A.dz = h->dz.p; A.area = h->area.p;
  A.dist_up = t->custom_dist ? t->dist_up : nullptr; A.dist_dn = t->custom_dist ? t->dist_dn : nullptr;
  A.dist_uniform = (t->custom_dist && t->dist_uniform && h->nlev <= 32) ? 1 : 0;
  if (A.dist_uniform) { memcpy(A.lay_du, t->lay_du, sizeof(A.lay_du)); memcpy(A.lay_dd, t->lay_dd, sizeof(A.lay_dd)); }
  A.lun_type = t->lun_type;
  A.T_in = t->T_cur; A.liq = t->liq; A.ice = t->ice; A.snow_water = t->snow_water; A.tuning = t->tuning;
  A.nsnow = t->nsnow; A.active = t->active;
  A.top_is_first = (h->orientation != MPPGPU_MESH_AGAINST_GRAVITY);
  A.stale_area = t->stale_area;
  for (auto *c : h->bcs) {
    const int k = (c->region == REGION_TOP) ? 0 : 1;
    if (A.bc_type[k]) return fail("mppgpu_step_dt: one thermal boundary condition per region is supported");
    A.bc_type[k] = c->itype; A.bc_value[k] = c->value.p; A.bc_dhsdT[k] = c->dhsdT.p; A.bc_frac[k] = c->frac.p;
    A.bc_active[k] = c->mass_exc.p;                  // boundary aux var is_active flags (0/1 doubles)
  }
  for (auto *c : h->sss) {
    if (A.nss >= TH_MAX_SS) return fail("mppgpu_step_dt: at most %d thermal source conditions", TH_MAX_SS);
    A.ss_value[A.nss] = c->value.p; A.ss_region[A.nss] = c->region; A.nss++;
  }
  A.T_out = (t->T_cur == t->T_clm) ? t->T_work : t->T_cur;
  if (t->therm_cond && t->heat_cap) { A.therm_cond = t->therm_cond; A.heat_cap = t->heat_cap; t->diagnostics = true; }
  CK(cudaEventRecord(h->ev0, h->stream));
  const int nlev = h->nlev;
  // bulk-async (1-D TMA) persistent variant for the shapes it covers and batches that fill the machine (thermal_kernels.cuh)
  ThermalTmaPlan P;
  const bool tma = t->bulk_copy && nlev <= 16 && thermal_tma_plan(A, P) && P.ntiles >= 2 * h->sm_count;
  if (tma) {
    const size_t smem = (size_t)TMA_STAGES * P.stage_bytes;
    if (!t->tma_attr_set) { CK(cudaFuncSetAttribute(thermal_step2_tma_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024)); t->tma_attr_set = true; }
    const int grid = std::min(P.ntiles, h->sm_count * TMA_BLOCKS_PER_SM);
    thermal_step2_tma_kernel<8><<<grid, TH_TILE, smem, h->stream>>>(A, P);
  } else if (nlev <= 32) {
    if (nlev <= 16) thermal_step2_kernel<8><<<nblk((long long)h->ncol * 8, TH_TILE), TH_TILE, 0, h->stream>>>(A);
    else            thermal_step_kernel<32><<<nblk((long long)h->ncol * 32, TH_TILE), TH_TILE, 0, h->stream>>>(A);
  } else {
    if (!t->work) CK(mpp_dmalloc((void **)&t->work, 4 * h->ncells * sizeof(double)));
    thermal_step_generic_kernel<<<nblk(h->ncol, 64), 64, 0, h->stream>>>(A, t->work);
  }
  CK(cudaGetLastError());
  CK(cudaEventRecord(h->ev1, h->stream));
  h->launches += 1;
  t->T_cur = A.T_out;                              // PostSolve: soln -> soln_prev (SOEBasePostSolve :650-668)
  return 0;
}


// MPPThermalTBasedALM_Solve (src/driver/alm/MPPThermalTBasedALM_Driver.F90:150-452): ELM's raw column arrays in, tvector out.
// Software-pipelined over column chunks: chunk k's rows go up on `copy_in` (one strided copy per array: the arrays are layer-major), are
// packed, stepped and unpacked on the handle's stream, and its tvector rows come down on `copy_out` while chunk k+1 computes.
static int thermal_elm_pipeline(mppgpu_soe *h, ThermalState *t, double dtime, const mppgpu_elm_thermal_columns *cols, double capr);
static int thermal_elm_solve(mppgpu_soe *h, ThermalState *t, double dtime, const mppgpu_elm_thermal_columns *cols, double capr)
{
  if (!t->snow_mode) return fail("mppgpu_thermal_elm_solve: call mppgpu_thermal_add_snow_ssw first (the ELM configuration)");
  if (!t->soils_set) return fail("mppgpu_thermal_elm_solve: soils must be set first");
  if (h->nlev < 2) return fail("mppgpu_thermal_elm_solve: the surface tuning factor reads z(c,2): at least two soil layers are needed");
  if (!cols) return fail("mppgpu_thermal_elm_solve: null column arrays");
  const void *req[] = {cols->snl, cols->z, cols->dz, cols->zi, cols->t_soisno, cols->h2osoi_liq, cols->h2osoi_ice, cols->frac_sno_eff, cols->h2osno,
                       cols->h2osfc, cols->frac_h2osfc, cols->t_h2osfc, cols->sabg_lyr, cols->dhsdT, cols->hs_soil, cols->hs_top_snow, cols->hs_h2osfc, cols->tvector};
  for (const void *q : req) if (!q) return fail("mppgpu_thermal_elm_solve: null column array");
  double *const T_before = t->T_cur;
  if (thermal_elm_pipeline(h, t, dtime, cols, capr)) {
    // copies to and from the caller's arrays may still be queued: drain the three streams before handing the arrays back
    const std::string msg = g_err;
    if (h->copy_in) cudaStreamSynchronize(h->copy_in);
    if (h->copy_out) cudaStreamSynchronize(h->copy_out);
    cudaStreamSynchronize(h->stream);
    (void)cudaGetLastError();
    t->T_cur = T_before; t->elm_soil_loaded = false;
    g_err = msg;
    return 1;
  }
  return 0;
}

static int thermal_elm_pipeline(mppgpu_soe *h, ThermalState *t, double dtime, const mppgpu_elm_thermal_columns *cols, double capr)
{
  const size_t ncol = h->ncol, nsno = t->nsno, nlev = h->nlev, nl = nsno + nlev, nrow = nl + 1;
  cudaStream_t s = h->stream;
  if (!h->copy_in)  CK(cudaStreamCreateWithFlags(&h->copy_in, cudaStreamNonBlocking));
  if (!h->copy_out) CK(cudaStreamCreateWithFlags(&h->copy_out, cudaStreamNonBlocking));
  if (!h->ev_out_done) { CK(cudaEventCreateWithFlags(&h->ev_out_done, cudaEventDisableTiming)); CK(cudaEventCreateWithFlags(&h->ev_start, cudaEventDisableTiming)); }
  int nchunks = t->elm_chunks > 0 ? t->elm_chunks : 8;
  const long long align = 1024;
  long long per = ((long long)ncol + nchunks - 1) / nchunks;
  per = ((per + align - 1) / align) * align;
  if (t->elm_chunks <= 0 && per < 32768) per = 32768;
  nchunks = (int)(((long long)ncol + per - 1) / per);
  while ((int)h->ev_in.size() < nchunks) {
    cudaEvent_t a, b; CK(cudaEventCreateWithFlags(&a, cudaEventDisableTiming)); CK(cudaEventCreateWithFlags(&b, cudaEventDisableTiming));
    h->ev_in.push_back(a); h->ev_comp.push_back(b);
  }
  // staging tables in the caller's own layer-major layout: z, dz, t, liq, ice (nl rows of ncol) | zi (nl+1) | sabg (nsno+1) | 9 per-column rows | tvector (nrow)
  const size_t total = 5 * ncol * nl + ncol * (nl + 1) + ncol * (nsno + 1) + 9 * ncol + ncol * nrow;
  if (!t->elm_stage) { CK(mpp_dmalloc((void **)&t->elm_stage, total * 8)); CK(mpp_dmalloc((void **)&t->elm_snl, ncol * sizeof(int))); t->elm_soil_loaded = false; }
  // the soil rows of z / dz / zi are ELM's fixed vertical grid: with mppgpu_elm_set_pipeline(h, ., 1) they go up with the first solve only
  const bool snow_rows_only = t->elm_static_soil && t->elm_soil_loaded;
  struct Table { double *dev; const double *host; size_t rows; };
  double *p = t->elm_stage;
  auto table = [&](const double *src, size_t rows_total, size_t rows_up) -> Table { Table x{p, src, rows_up}; p += rows_total * ncol; return x; };
  Table up[16];
  int nup = 0;
  ElmThermalArgs E;
  memset(&E, 0, sizeof(E));
  E.ncol = h->ncol; E.nlev = h->nlev; E.nsno = t->nsno; E.capr = capr;
  E.active_col = h->has_active ? h->active.p : nullptr;
  up[nup] = table(cols->z, nl, snow_rows_only ? nsno : nl); E.z = up[nup++].dev;
  up[nup] = table(cols->dz, nl, snow_rows_only ? nsno : nl); E.dz = up[nup++].dev;
  up[nup] = table(cols->t_soisno, nl, nl); E.t_soisno = up[nup++].dev;
  up[nup] = table(cols->h2osoi_liq, nl, nl); E.h2osoi_liq = up[nup++].dev;
  up[nup] = table(cols->h2osoi_ice, nl, nl); E.h2osoi_ice = up[nup++].dev;
  up[nup] = table(cols->zi, nl + 1, snow_rows_only ? nsno + 1 : nl + 1); E.zi = up[nup++].dev;
  up[nup] = table(cols->sabg_lyr, nsno + 1, nsno + 1); E.sabg_lyr = up[nup++].dev;
  up[nup] = table(cols->frac_sno_eff, 1, 1); E.frac_sno_eff = up[nup++].dev;
  up[nup] = table(cols->h2osno, 1, 1); E.h2osno = up[nup++].dev;
  up[nup] = table(cols->h2osfc, 1, 1); E.h2osfc = up[nup++].dev;
  up[nup] = table(cols->frac_h2osfc, 1, 1); E.frac_h2osfc = up[nup++].dev;
  up[nup] = table(cols->t_h2osfc, 1, 1); E.t_h2osfc = up[nup++].dev;
  up[nup] = table(cols->dhsdT, 1, 1); E.dhsdT = up[nup++].dev;
  up[nup] = table(cols->hs_soil, 1, 1); E.hs_soil = up[nup++].dev;
  up[nup] = table(cols->hs_top_snow, 1, 1); E.hs_top_snow = up[nup++].dev;
  up[nup] = table(cols->hs_h2osfc, 1, 1); E.hs_h2osfc = up[nup++].dev;
  E.tvector = p;                                                     // entries the driver does not assign keep the caller's values
  E.snl = t->elm_snl;
  // SetSolnPrevCLM + Set{R,I,B}DataFromCLM + PreStepDT (:332-441): written straight into the mailbox
  E.T = t->T_clm; E.liq = t->liq; E.ice = t->ice; E.snow_water = t->snow_water; E.mdz = t->aux_dz; E.dist_up = t->aux_dist_up; E.dist_dn = t->aux_dist_dn;
  E.tuning = t->tuning; E.frac = t->frac; E.nsnow = t->nsnow; E.active = t->active;
  for (int k = 0; k < 3; ++k) { E.hs[k] = t->hs[k]; E.dhs[k] = t->dhs[k]; }
  E.frac_soil = t->frac_soil; E.sabg_snow = t->sabg_snow; E.sabg_soil = t->sabg_soil;
  t->T_cur = t->T_clm;                                               // PreStepDT
  double *const T_out = t->T_work;
  E.T_out = T_out;                                                   // GetSoln

  CK(cudaEventRecord(h->ev_start, s));
  CK(cudaStreamWaitEvent(h->copy_in, h->ev_start, 0));
  CK(cudaStreamWaitEvent(h->copy_out, h->ev_start, 0));
  CK(cudaEventRecord(h->ev0, s));
  const size_t pitch = ncol * 8;
  for (int k = 0; k < nchunks; ++k) {
    const long long c0 = k * per; const int n = (int)std::min<long long>(per, (long long)ncol - c0);
    for (int a = 0; a < nup; ++a)
      CK(cudaMemcpy2DAsync(up[a].dev + c0, pitch, up[a].host + c0, pitch, (size_t)n * 8, up[a].rows, cudaMemcpyHostToDevice, h->copy_in));
    CK(cudaMemcpy2DAsync(E.tvector + c0, pitch, cols->tvector + c0, pitch, (size_t)n * 8, nrow, cudaMemcpyHostToDevice, h->copy_in));
    CK(cudaMemcpyAsync(t->elm_snl + c0, cols->snl + c0, (size_t)n * sizeof(int), cudaMemcpyHostToDevice, h->copy_in));
    CK(cudaEventRecord(h->ev_in[k], h->copy_in));
    CK(cudaStreamWaitEvent(s, h->ev_in[k], 0));
    E.col0 = (int)c0; E.ncols = n;
    elm_thermal_pack_kernel<<<nblk((long long)n * nrow, 256), 256, 0, s>>>(E);
    CK(cudaGetLastError());
    if (thermal_snow_launch(h, t, dtime, c0, n, T_out)) return 1;     // StepDT
    elm_thermal_unpack_kernel<<<nblk((long long)n * nrow, 256), 256, 0, s>>>(E);
    CK(cudaGetLastError());
    h->launches += 2;
    CK(cudaEventRecord(h->ev_comp[k], s));
    CK(cudaStreamWaitEvent(h->copy_out, h->ev_comp[k], 0));
    CK(cudaMemcpy2DAsync(cols->tvector + c0, pitch, E.tvector + c0, pitch, (size_t)n * 8, nrow, cudaMemcpyDeviceToHost, h->copy_out));
  }
  CK(cudaEventRecord(h->ev_out_done, h->copy_out));
  CK(cudaStreamWaitEvent(s, h->ev_out_done, 0));
  CK(cudaEventRecord(h->ev1, s));
  t->T_cur = T_out;
  CK(cudaStreamSynchronize(s));
  t->elm_soil_loaded = true;
  return 0;
}
