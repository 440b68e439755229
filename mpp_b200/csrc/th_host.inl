// th_host.inl -- placeholder (filled in by the TH milestone)
static int th_create(THState *t, int, int, cudaStream_t s) { t->stream = s; return 0; }
static void th_destroy(THState *) {}
static int th_set_mesh(THState *, int, const double *, const double *) { return 0; }
static int th_restart(THState *, const double *) { return 1; }
static int th_field(mppgpu_soe *, THState *, int, int, int, int, bool, double **, size_t *) { return fail("TH SoE not implemented yet"); }
static int th_pre_step_dt(THState *) { return 0; }
static int th_post_step_dt(THState *) { return 0; }
static int th_step(mppgpu_soe *, THState *, double) { return fail("TH SoE not implemented yet"); }
static int th_eval(mppgpu_soe *, THState *, double, const double *, const double *, double *, double *, double *, double *) { return fail("TH SoE not implemented yet"); }
static int th_set_soils(mppgpu_soe *, THState *, const double *, const double *, const double *, const double *, const double *, const double *, const double *, int, int, int) { return fail("TH SoE not implemented yet"); }
