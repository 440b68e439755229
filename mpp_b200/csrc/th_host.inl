// th_host.inl -- host side of the coupled thermal-hydrology SoE (sysofeqns_th_type: Richards ieqn 1 + enthalpy ieqn 2):
// device-resident soil tables, cell-interleaved (P,T) solution, set/get routing, StepDT launch.
// Included at the end of mppgpu.cu.

// [P(0..N-1) | T(0..N-1)] -> interleaved (P,T) per cell, and back
__global__ void th_interleave_kernel(const double *__restrict__ pt, double *__restrict__ x, long long n)
{
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) { x[2 * i] = pt[i]; x[2 * i + 1] = pt[n + i]; }
}
__global__ void th_deinterleave_kernel(const double *__restrict__ x, double *__restrict__ P, double *__restrict__ T, long long n)
{
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) { P[i] = x[2 * i]; T[i] = x[2 * i + 1]; }
}

static int th_create(THState *t, int ncol, int nlev, cudaStream_t s)
{
  t->stream = s; t->ncol = ncol; t->nlev = nlev;
  const size_t N = (size_t)ncol * nlev;
  int rc = 0;
  rc |= th_alloc_d(&t->x, 2 * N, 0.0, s);
  rc |= th_alloc_d(&t->Pout, N, 0.0, s); rc |= th_alloc_d(&t->Tout, N, 0.0, s);
  rc |= th_alloc_d(&t->liq_sat, N, 0.0, s); rc |= th_alloc_d(&t->mass, N, 0.0, s);
  return rc;
}

static void th_destroy(THState *t)
{
  double *d[] = {t->por, t->perm, t->sat_res, t->alpha, t->lam, t->vgn, t->pu, t->ps, t->b2, t->b3, t->tkdry, t->csol, t->perm_e,
                 t->x, t->Pout, t->Tout, t->liq_sat, t->mass};
  for (double *p : d) if (p) cudaFree(p);
}

static int th_set_mesh(THState *t, int orientation, const double *d_dz, const double *d_area)
{
  t->orientation = orientation; t->d_dz = d_dz; t->d_area = d_area;
  return 0;
}

static int th_restart(THState *t, const double *x)
{
  const size_t N = (size_t)t->ncol * t->nlev;
  double *tmp = nullptr;
  if (mpp_dmalloc((void **)&tmp, 2 * N * sizeof(double)) != cudaSuccess) return 1;
  int rc = 0;
  if (cudaMemcpyAsync(tmp, x, 2 * N * sizeof(double), cudaMemcpyHostToDevice, t->stream) != cudaSuccess) rc = 1;
  if (!rc) {
    th_interleave_kernel<<<nblk(N, 256), 256, 0, t->stream>>>(tmp, t->x, (long long)N);
    if (cudaGetLastError() != cudaSuccess || cudaStreamSynchronize(t->stream) != cudaSuccess) rc = 1;
  }
  cudaFree(tmp);
  t->views_stale = true;
  return rc;
}

// MPPTHSetSoils (MultiPhysicsProbTH.F90:75-560): the five hydraulic tables feed both governing equations' saturation
// functions; the energy equation's permeability stays at its aux-var default (the setter is commented out, :293)
static int th_set_soils(mppgpu_soe *h, THState *t, const double *watsat, const double *hksat, const double *bsw, const double *sucsat,
                        const double *residual_sat, const double *csol, const double *tkdry, int satfunc_type, int density_type, int iee_type)
{
  if (!watsat || !hksat || !bsw || !sucsat || !residual_sat || !csol || !tkdry) return fail("MPPTHSetSoils: null table");
  if (satfunc_type < 0 || satfunc_type > 3) return fail("ERROR:: Unknown satfunc_type = %d", satfunc_type);
  if (density_type < DENSITY_CONSTANT || density_type > DENSITY_IFC67) return fail("Unknown value for VAR_DENSITY_TYPE %d", density_type);
  if (iee_type != INT_ENERGY_ENTHALPY_CONSTANT && iee_type != INT_ENERGY_ENTHALPY_IFC67) return fail("Unknown int_energy_enthalpy_type %d", iee_type);
  const size_t N = h->ncells;
  DevBuf<double> tb[7]; const double *src[7] = {watsat, hksat, bsw, sucsat, residual_sat, csol, tkdry};
  for (int i = 0; i < 7; ++i) if (upload_table(h, src[i], tb[i])) return 1;
  double **need[] = {&t->por, &t->perm, &t->sat_res, &t->alpha, &t->lam, &t->vgn, &t->pu, &t->ps, &t->b2, &t->b3, &t->tkdry, &t->csol};
  for (double **p : need) if (!*p) CK(mpp_dmalloc((void **)p, N * sizeof(double)));
  DevBuf<int> bad; CK(bad.alloc(1)); CK(cudaMemsetAsync(bad.p, 0, sizeof(int), h->stream));
  convert_soils_kernel<<<nblk(N, 128), 128, 0, h->stream>>>(satfunc_type, tb[0].p, tb[1].p, tb[2].p, tb[3].p, tb[4].p, h->ncol, h->nlev,
      t->por, t->perm, t->sat_res, t->alpha, t->lam, t->vgn, t->pu, t->ps, t->b2, t->b3, bad.p);
  CK(cudaGetLastError());
  transpose_to_cells_kernel<<<nblk(N, 256), 256, 0, h->stream>>>(tb[5].p, t->csol, h->ncol, h->nlev);
  transpose_to_cells_kernel<<<nblk(N, 256), 256, 0, h->stream>>>(tb[6].p, t->tkdry, h->ncol, h->nlev);
  CK(cudaGetLastError());
  int hbad = 0;
  CK(cudaMemcpyAsync(&hbad, bad.p, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  if (hbad) return fail("SatFunc_Set_*: bad param (SaturationFunction.F90:141-146,177-182,285-291,343-349)");
  t->satfunc_name = satfunc_type; t->density_type = density_type; t->iee_type = iee_type;
  t->soils_set = true; h->soils_set = true;
  return 0;
}

// goveq_enthalpy%SetSoilPermeability (GoveqnThermalEnthalpySoilType.F90:2454-2480): per-cell permeability of the ENERGY equation's aux vars
static int th_set_energy_permeability(mppgpu_soe *h, THState *t, const double *perm)
{
  const size_t N = (size_t)h->ncells;
  if (!t->perm_e && mpp_dmalloc((void **)&t->perm_e, N * sizeof(double)) != cudaSuccess) return fail("cudaMalloc failed (energy permeability)");
  CK(cudaMemcpyAsync(t->perm_e, perm, N * sizeof(double), cudaMemcpyHostToDevice, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  return 0;
}

static int th_refresh_views(mppgpu_soe *h, THState *t)
{
  if (!t->views_stale) return 0;
  th_deinterleave_kernel<<<nblk(h->ncells, 256), 256, 0, h->stream>>>(t->x, t->Pout, t->Tout, (long long)h->ncells);
  CK(cudaGetLastError());
  t->views_stale = false;
  return 0;
}

static int th_field(mppgpu_soe *h, THState *t, int ieqn, int auxvar_type, int var_type, int cond_id, bool for_set, double **p, size_t *cap)
{
  if (auxvar_type == AUXVAR_INTERNAL) {
    *cap = h->ncells;
    if (for_set) return fail("SOETHSetDataFromCLM: internal aux vars are set through Restart (P, T) only");
    switch (var_type) {
    case VAR_PRESSURE:    if (th_refresh_views(h, t)) return 1; *p = t->Pout; return 0;
    case VAR_TEMPERATURE: if (th_refresh_views(h, t)) return 1; *p = t->Tout; return 0;
    case VAR_LIQ_SAT:     *p = t->liq_sat; return 0;
    case VAR_MASS:        *p = t->mass; return 0;
    }
    return fail("SOETHGetDataForCLM: unknown var_type %d", var_type);
  }
  if (auxvar_type != AUXVAR_BC && auxvar_type != AUXVAR_SS) return fail("SOETH%sData: Unknown soe_auxvar_type %d", for_set ? "Set" : "Get", auxvar_type);
  HostCond *c = find_cond(h, auxvar_type, cond_id);
  if (!c) return fail("SOETH%sData: condition id %d out of range", for_set ? "Set" : "Get", cond_id);
  if (c->ieqn != ieqn) return fail("SOETH%sData: condition %d belongs to governing equation %d, not %d", for_set ? "Set" : "Get", cond_id, c->ieqn, ieqn);
  *cap = c->n;
  if (var_type == VAR_BC_SS_CONDITION) { *p = c->value.p; return 0; }
  if (auxvar_type == AUXVAR_BC && var_type == VAR_PRESSURE && c->ieqn == 2) {
    // the reference's drivers poke aux_vars_bc(:)%pressure of the energy equation directly (mass_and_heat_model_problem.F90:616-621);
    // the flux buffer of a boundary condition is unused by the TH SoE and carries that pressure (default 0, RichardsODEPressureAuxType.F90:90)
    *p = c->flux.p; return 0;
  }
  return fail("SOETH%sData: unknown var_type %d", for_set ? "Set" : "Get", var_type);
}

static int th_pre_step_dt(THState *) { return 0; }
static int th_post_step_dt(THState *) { return 0; }

static int th_fill_args(mppgpu_soe *h, THState *t, THArgs &A, double dt)
{
  memset(&A, 0, sizeof(A));
  A.ncol = h->ncol; A.nlev = h->nlev;
  A.uz = (h->orientation == MPPGPU_MESH_ALONG_GRAVITY) ? -1.0 : (h->orientation == MPPGPU_MESH_AGAINST_GRAVITY ? 1.0 : 0.0);
  A.top_is_first = (h->orientation != MPPGPU_MESH_AGAINST_GRAVITY);
  A.satfunc = (t->satfunc_name == MPPGPU_SATFUNC_VAN_GENUCHTEN) ? SATFUNC_VG : (t->satfunc_name == MPPGPU_SATFUNC_BROOKS_COREY ? SATFUNC_BC : SATFUNC_SBC);
  A.density_type = t->density_type; A.iee_type = t->iee_type;
  A.por = t->por; A.perm = t->perm; A.sat_res = t->sat_res; A.alpha = t->alpha; A.lam = t->lam;
  A.vgn = (A.satfunc == SATFUNC_VG) ? t->vgn : nullptr;
  const bool sbc = (A.satfunc == SATFUNC_SBC);
  A.pu = sbc ? t->pu : nullptr; A.ps = sbc ? t->ps : nullptr; A.b2 = sbc ? t->b2 : nullptr; A.b3 = sbc ? t->b3 : nullptr;
  A.dz = h->dz.p; A.area = h->area.p; A.tkdry = t->tkdry; A.csol = t->csol; A.perm_e = t->perm_e;
  A.x_in = t->x; A.x_out = t->x;
  for (auto *c : h->bcs) {
    if (A.nbc >= 4) return fail("mppgpu_step_dt: at most 4 TH boundary conditions");
    if (c->region == REGION_CELLS) return fail("mppgpu_step_dt: TH boundary conditions live on SOIL_TOP_CELLS / SOIL_BOTTOM_CELLS");
    if (c->itype != COND_DIRICHLET) return fail("mppgpu_step_dt: TH boundary condition type %d unsupported (COND_DIRICHLET 505)", c->itype);
    THCondDev &d = A.bc[A.nbc++];
    d.value = c->value.p; d.bc_pressure = c->flux.p; d.ieqn = c->ieqn; d.itype = c->itype; d.region = c->region;
  }
  for (auto *c : h->sss) {
    if (A.nss >= 4) return fail("mppgpu_step_dt: at most 4 TH source/sink conditions");
    if (!((c->ieqn == 1 && c->itype == COND_MASS_RATE) || (c->ieqn == 2 && c->itype == COND_HEAT_RATE)))
      return fail("mppgpu_step_dt: TH source type %d on equation %d unsupported (COND_MASS_RATE on 1, COND_HEAT_RATE on 2)", c->itype, c->ieqn);
    THCondDev &d = A.ss[A.nss++];
    d.value = c->value.p; d.bc_pressure = nullptr; d.ieqn = c->ieqn; d.itype = c->itype; d.region = c->region;
  }
  // one Dirichlet temperature at the top of columns of <= 15 layers: the boundary connection rides on the kernel's padding lane
  A.bc_on_pad_lane = (h->nlev <= 15 && A.top_is_first && A.nbc == 1 && A.bc[0].ieqn == 2 && A.bc[0].region == REGION_TOP && A.bc[0].itype == COND_DIRICHLET) ? 1 : 0;
  A.liq_sat = t->liq_sat; A.mass = t->mass;
  A.stat_its = h->stat_its.p; A.stat_reason = h->stat_reason.p; A.stat_cuts = h->stat_cuts.p; A.stat_nf = h->stat_nf.p;
  A.dt = dt; A.so = h->so;
  return 0;
}

// nlev <= 16: lane-per-cell register kernel (th_kernels2.cuh); taller columns: one warp per column out of shared memory
static int th_launch(mppgpu_soe *h, THState *, THArgs &A, int *nblocks_out)
{
  const bool fast = h->nlev <= 16;
  const int nblocks = fast ? nblk((long long)h->ncol * 16, TH2_THREADS) : h->ncol;
  if (!A.eval_x) {
    if (h->block_partials.n < (size_t)nblocks * 9) CK(h->block_partials.alloc((size_t)nblocks * 9));
    A.block_partials = h->block_partials.p;
  }
  if (fast) {
    // the model combinations the reference's drivers use (and ELM's default curve) get compile-time specialisations in th_step2_inst.cu;
    // anything else dispatches at run time
    const bool tanaka_const = A.density_type == DENSITY_TGDPB01 && A.iee_type == INT_ENERGY_ENTHALPY_CONSTANT;
    const bool padbc = A.bc_on_pad_lane != 0 && !A.eval_x;
    if (A.satfunc == SATFUNC_VG && tanaka_const) { if (padbc) th2_launch_0p(A, nblocks, h->stream); else th2_launch_0(A, nblocks, h->stream); }
    else if (A.satfunc == SATFUNC_VG && A.density_type == DENSITY_IFC67 && A.iee_type == INT_ENERGY_ENTHALPY_IFC67) th2_launch_1(A, nblocks, h->stream);
    else if (A.satfunc == SATFUNC_SBC && tanaka_const) { if (padbc) th2_launch_2p(A, nblocks, h->stream); else th2_launch_2(A, nblocks, h->stream); }   // ELM's default curve (mpp_varctl.F90:17)
    else th2_launch_3(A, nblocks, h->stream);
  } else {
    const size_t smem = (size_t)TH_NARR * h->nlev * sizeof(double);
    if (smem > 200 * 1024) return fail("mppgpu_step_dt: nlev = %d exceeds the TH kernel's shared-memory budget", h->nlev);
    CK(cudaFuncSetAttribute(th_step_generic_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    th_step_generic_kernel<<<nblocks, 32, smem, h->stream>>>(A);
  }
  CK(cudaGetLastError());
  *nblocks_out = nblocks;
  return 0;
}

static int th_step(mppgpu_soe *h, THState *t, double dt)
{
  if (!h->mesh_set || !t->soils_set) return fail("mppgpu_step_dt: mesh and soils must be set first");
  if (!(dt > 0.0)) return fail("mppgpu_step_dt: dt must be positive");
  THArgs A;
  if (th_fill_args(h, t, A, dt)) return 1;
  int nblocks = 0;
  CK(cudaEventRecord(h->ev0, h->stream));
  const bool ordered = h->ordering != 0 && h->nlev <= 16;               // the register kernel; columns are independent: results do not depend on it
  A.order = (ordered && h->order_valid && h->order_chunks == 1 && h->order_per == h->ncol) ? h->order.p : nullptr;
  if (th_launch(h, t, A, &nblocks)) return 1;
  if (ordered) { if (vsfm_build_order(h, 0, h->ncol, h->stream)) return 1; }
  h->order_valid = ordered; h->order_chunks = 1; h->order_per = h->ncol;
  reduce_partials_kernel<<<nblocks < REDUCE_BLOCKS ? 1 : REDUCE_BLOCKS, 256, 0, h->stream>>>(h->block_partials.p, nblocks, h->red_scratch.p, h->red_counter.p, h->red_out.p);
  CK(cudaGetLastError());
  CK(cudaEventRecord(h->ev1, h->stream));
  h->launches += 2;
  h->nblocks_last = nblocks;
  t->views_stale = true;
  CK(cudaMemcpyAsync(h->h_red, h->red_out.p, 9 * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  h->result_pending = true;
  return 0;
}

// residual + Jacobian blocks at x (cell-interleaved (P,T)), accumulation at x_prev: kernel unit-test probe, no state change
static int th_eval(mppgpu_soe *h, THState *t, double dt, const double *x_prev, const double *x, double *f, double *ja, double *jb, double *jc)
{
  if (!h->mesh_set || !t->soils_set) return fail("mppgpu_eval: mesh and soils must be set first");
  if (!x_prev || !x || !f || !ja || !jb || !jc) return fail("mppgpu_eval: null argument");
  const size_t N = h->ncells;
  DevBuf<double> dxp, dx, df, da, db, dc;
  CK(dxp.alloc(2 * N)); CK(dx.alloc(2 * N)); CK(df.alloc(2 * N)); CK(da.alloc(4 * N)); CK(db.alloc(4 * N)); CK(dc.alloc(4 * N));
  CK(cudaMemcpyAsync(dxp.p, x_prev, 2 * N * 8, cudaMemcpyHostToDevice, h->stream));
  CK(cudaMemcpyAsync(dx.p, x, 2 * N * 8, cudaMemcpyHostToDevice, h->stream));
  THArgs A;
  if (th_fill_args(h, t, A, dt)) return 1;
  A.x_in = dxp.p; A.x_out = nullptr;
  A.eval_x = dx.p; A.eval_f = df.p; A.eval_ja = da.p; A.eval_jb = db.p; A.eval_jc = dc.p;
  int nb = 0;
  if (th_launch(h, t, A, &nb)) return 1;
  h->launches += 1;
  CK(cudaMemcpyAsync(f, df.p, 2 * N * 8, cudaMemcpyDeviceToHost, h->stream));
  CK(cudaMemcpyAsync(ja, da.p, 4 * N * 8, cudaMemcpyDeviceToHost, h->stream));
  CK(cudaMemcpyAsync(jb, db.p, 4 * N * 8, cudaMemcpyDeviceToHost, h->stream));
  CK(cudaMemcpyAsync(jc, dc.p, 4 * N * 8, cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  return 0;
}
