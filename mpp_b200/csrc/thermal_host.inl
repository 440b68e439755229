// thermal_host.inl -- placeholder (filled in by the thermal milestone)
static int thermal_create(ThermalState *t, int, int, cudaStream_t s) { t->stream = s; return 0; }
static void thermal_destroy(ThermalState *) {}
static int thermal_set_mesh(ThermalState *, int, const double *, const double *) { return 0; }
static int thermal_set_temperature(ThermalState *, const double *, bool) { return 1; }
static int thermal_field(mppgpu_soe *, ThermalState *, int, int, int, bool, double **, size_t *) { return fail("thermal SoE not implemented yet"); }
static int thermal_set_idata(mppgpu_soe *, ThermalState *, int, int, int, const int *, int) { return fail("thermal SoE not implemented yet"); }
static int thermal_pre_step_dt(ThermalState *) { return 0; }
static int thermal_post_step_dt(ThermalState *) { return 0; }
static int thermal_step(mppgpu_soe *, ThermalState *, double) { return fail("thermal SoE not implemented yet"); }
static int thermal_set_soils(mppgpu_soe *, ThermalState *, const double *, const double *, const double *, const double *, const int *, int, int) { return fail("thermal SoE not implemented yet"); }
