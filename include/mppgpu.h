/*
 * mppgpu.h -- C ABI of libmppgpu.so: the B200-native (sm_100a, fp64) replacement for the
 * per-soil-column implicit solve of MPP-LSM/MPP (reference: Fortran 2003 + PETSc).
 *
 * Seam: the type-bound procedures of class(sysofeqns_base_type) that every reference driver
 * calls.  A Fortran ISO_C_BINDING shim forwards them to the entry points
 * below; the MPP problem/governing-equation setup API stays in Fortran.
 *
 *   reference interface (file:line under the reference root)                    entry point here
 *   ---------------------------------------------------------------------------------------------
 *   MPPSetupProblem / SOE creation   MultiPhysicsProbBaseType.F90:1058-1213     mppgpu_create
 *   MeshCreate, CreateFromCLMCols    MeshType.F90:173-269, 293-645              mppgpu_set_mesh
 *   mpp%CreateAndAddConnectionSet    MPPThermalTBasedALM_Initialize.F90:379-381,
 *                                    466-468                                     mppgpu_set_connection_distances
 *   soe%AddConditionInGovEqn         SystemOfEquationsBaseType.F90:995+         mppgpu_add_condition
 *   VSFMMPPSetSoils                  MultiPhysicsProbVSFM.F90:211-475           mppgpu_vsfm_set_soils
 *   MPPThermalSetSoils               MultiPhysicsProbThermal.F90:76-208         mppgpu_thermal_set_soils
 *   MPPTHSetSoils                    MultiPhysicsProbTH.F90:75-401              mppgpu_th_set_soils
 *   goveq_enthalpy%SetSoilPermeability GoveqnThermalEnthalpySoilType.F90:2454  mppgpu_th_set_energy_permeability
 *   SNESSetTolerances                MultiPhysicsProbBaseType.F90:1110-1196,
 *                                    MPPVSFMALM_Driver.F90:632-640              mppgpu_set_tolerances
 *   mpp%Restart(data_1d)             MultiPhysicsProbVSFM.F90:603-707           mppgpu_restart
 *   soe%SetDataFromCLM               SystemOfEquationsVSFMType.F90:663-724      mppgpu_set_data
 *   soe%SetSolnPrevCLM/SetRDataFromCLM/SetIDataFromCLM/SetBDataFromCLM
 *                                    SystemOfEquationsThermalType.F90:171-330   mppgpu_set_data / mppgpu_set_idata
 *   soe%GetDataForCLM(AUXVAR_CONN_INTERNAL, VAR_MASS_FLUX): SystemOfEquationsVSFMType.F90:824 -> mppgpu_get_data(.., 704, 644, ..),
 *                                    ncol*(nlev-1) internal-connection mass fluxes [kg/s], connection j->j+1 of column c at c*(nlev-1)+j
 *   soe%GetDataForCLM, GetSoln       SystemOfEquationsVSFMType.F90:781-845,
 *                                    SystemOfEquationsThermalType.F90:336       mppgpu_get_data
 *   soe%PreStepDT                    SystemOfEquationsVSFMType.F90:892-923      mppgpu_pre_step_dt
 *   soe%StepDT                       SystemOfEquationsBaseType.F90:334-647      mppgpu_step_dt
 *   soe%PostStepDT                   SystemOfEquationsVSFMType.F90:926-940      mppgpu_post_step_dt
 *   per-column mass-balance check    MPPVSFMALM_Driver.F90:556-601, 845-863     mppgpu_vsfm_mass_balance
 *   ELM coupling step (set x7, step, get x4)  MPPVSFMALM_Driver.F90:379-705   mppgpu_vsfm_coupled_step
 *   MPPVSFMALM_Solve (raw ELM arrays, packing, retry loop, unpacking)  MPPVSFMALM_Driver.F90:204-923   mppgpu_vsfm_elm_solve
 *
 * Conventions
 *   - plain pointers and sizes only; every array argument is a HOST pointer owned by the caller
 *     and may be freed right after the call (the reference copies eagerly too).  The *_device
 *     variants take DEVICE pointers on the handle's device for callers that already live on the GPU.
 *   - 1-D vectors are cell-ordered, icell = c*nlev + j (layer fastest; MultiPhysicsProbVSFM.F90:364).
 *   - soil tables are Fortran (ncol,nlev) column-major, t[j*ncol + c] (the reference's (c,j) arrays).
 *   - integer codes are the reference's own (MultiPhysicsProbConstants.F90:17-196).
 *   - return value: 0 = ok (PetscErrorCode-like); non-zero = configuration error, message via
 *     mppgpu_last_error().  Solver failure is NOT an error: *converged = 0 and *converged_reason
 *     carries the PETSc SNESConvergedReason code (worst column).
 *   - one handle = one system of equations on one GPU; not re-entrant (neither is the reference).
 *   - there is no CPU fallback: without a CUDA device mppgpu_create fails.
 */
#ifndef MPPGPU_H
#define MPPGPU_H

#ifdef __cplusplus
extern "C" {
#endif

typedef struct mppgpu_soe *mppgpu_handle;

/* soe_itype (MultiPhysicsProbConstants.F90:33-36) */
#define MPPGPU_SOE_RE_ODE          101
#define MPPGPU_SOE_THERMAL_TBASED  102
#define MPPGPU_SOE_TH              104
/* mesh orientation (:65-66); 313 is ours for CONN_IN_X_DIR chains (no gravity component) */
#define MPPGPU_MESH_ALONG_GRAVITY   311
#define MPPGPU_MESH_AGAINST_GRAVITY 312
#define MPPGPU_MESH_HORIZONTAL      313
/* saturation function names accepted by VSFMMPPSetSoils (MultiPhysicsProbVSFM.F90:391-417) */
#define MPPGPU_SATFUNC_VAN_GENUCHTEN 0
#define MPPGPU_SATFUNC_BROOKS_COREY  1
#define MPPGPU_SATFUNC_SBC_BZ2       2
#define MPPGPU_SATFUNC_SBC_BZ3       3

const char *mppgpu_last_error(void);
int  mppgpu_version(void);
int  mppgpu_device_count(void);

/* ---- life cycle -------------------------------------------------------------------------------
 * A handle lives on the CUDA device it was created on.  Every entry point selects that device for the duration of the call and hands the
 * caller's current device back on return. */
int  mppgpu_create(int soe_itype, int ncol, int nlev, int device, mppgpu_handle *out);
int  mppgpu_destroy(mppgpu_handle h);
/* run all work of this handle on a caller-provided cudaStream_t (NULL = the handle's own stream) */
int  mppgpu_set_stream(mppgpu_handle h, void *cuda_stream);
int  mppgpu_synchronize(mppgpu_handle h);

/* ---- setup ------------------------------------------------------------------------------------ */
/* dz: (ncol,nlev) Fortran order [m]; area: ncol [m^2]; col_active: ncol ints or NULL (all active) */
int  mppgpu_set_mesh(mppgpu_handle h, int orientation, const double *dz, const double *area, const int *col_active);
/* centroid-to-face distances of the internal (vertical) connections j -> j+1, (ncol, nlev-1) Fortran order;
 * default (never called) = dz/2 of the two cells.  Thermal SoE only (ELM's soil thermal mesh). */
int  mppgpu_set_connection_distances(mppgpu_handle h, const double *dist_up, const double *dist_dn);
/* ieqn: 1-based governing-equation rank in the SoE (TH: 1 = mass, 2 = energy); ss_or_bc: COND_SS/COND_BC;
 * returns the 1-based condition id in *cond_id (separate numbering for BCs and SSs, = soe_auxvar_id) */
int  mppgpu_add_condition(mppgpu_handle h, int ieqn, int ss_or_bc, int cond_type, int region, int *cond_id);
int  mppgpu_vsfm_set_soils(mppgpu_handle h, const double *watsat, const double *hksat, const double *bsw,
                           const double *sucsat, const double *residual_sat, int satfunc_type, int density_type);
int  mppgpu_thermal_set_soils(mppgpu_handle h, const double *watsat, const double *csol, const double *tkmg,
                              const double *tkdry, const int *lun_type, int nlevsoi, int istsoil);
int  mppgpu_thermal_set_cnfac(mppgpu_handle h, double cnfac);
/* ELM's real thermal column: adds the snow (nlevsno layers, variable active count) and standing-surface-water
 * (one cell) governing equations next to the soil equation and couples the three into ONE linear system per column and step,
 * i.e. everything add_meshes / add_goveqns / add_conditions_to_goveqns / allocate_auxvars of
 * src/driver/alm/MPPThermalTBasedALM_Initialize.F90:150-727 set up: GE 1 snow, GE 2 standing water, GE 3 soil; boundary conditions
 * 1 = heat flux at the top of snow, 2 = at the top of standing water, 3 = at the top of soil; sources 1 = absorbed solar
 * radiation on the snow cells (ncol*nlevsno values), 2 = on the soil cells; COND_DIRICHLET_FRM_OTR_GOVEQ coupling snow<->soil and
 * ssw<->soil (GoveqnThermalKSPTemperature{Snow,SSW,Soil}Type.F90).  soil_top_dist_dn[ncol] = z(c,1) - zi(c,0), the dist_dn the
 * driver writes into the soil's coupling conditions (:630-639).  Call after mppgpu_set_mesh (soil mesh, MESH_ALONG_GRAVITY) and
 * before any data is set; mppgpu_add_condition is not used in this configuration.  From then on every AUXVAR_INTERNAL array
 * (set_data / set_idata / restart / get_data) has ncol*(nlevsno+1+nlev) entries in the reference's SoE order
 * [snow cells, column-major | standing-water cells | soil cells], as MPPThermalTBasedALM_Driver.F90:204-452 packs them;
 * inactive cells come back as 0 like the reference's identity rows.  ceil(nlevsno/2) + ceil(nlev/2) <= 16. */
int  mppgpu_thermal_add_snow_ssw(mppgpu_handle h, int nlevsno, const double *soil_top_dist_dn);
/* MPPThermalTBasedALM_Solve (src/driver/alm/MPPThermalTBasedALM_Driver.F90:150-452) with ELM's raw column arrays: the packing into the
 * SoE mailbox (:204-330: active snow layers, the standing-water cell, tuning factors, fractions, heat fluxes and absorbed radiation),
 * SetSolnPrevCLM / Set{R,I,B}DataFromCLM, PreStepDT, StepDT, GetSoln and the unpacking into tvector (:460-505) run on the device.
 * Arrays are HOST pointers in ELM's own (c, j) Fortran order (column index fastest): value of column c at layer j sits at
 * [(j - jlo)*ncol + c], with jlo = -nlevsno+1 for z, dz, t_soisno, h2osoi_liq, h2osoi_ice, sabg_lyr (up to j = 1) and jlo = -nlevsno
 * for zi and tvector.  Needs mppgpu_thermal_add_snow_ssw; columns filtered out by mppgpu_set_mesh's col_active are skipped. */
typedef struct {
  const int *snl;                                          /* ncol: minus the number of snow layers */
  const double *z, *dz, *zi;                               /* col%z, col%dz (ncol, -nlevsno+1:nlev), col%zi (ncol, -nlevsno:nlev) */
  const double *t_soisno, *h2osoi_liq, *h2osoi_ice;        /* (ncol, -nlevsno+1:nlev) */
  const double *frac_sno_eff, *h2osno, *h2osfc, *frac_h2osfc, *t_h2osfc;   /* ncol */
  const double *sabg_lyr;                                  /* (ncol, -nlevsno+1:1) */
  const double *dhsdT, *hs_soil, *hs_top_snow, *hs_h2osfc; /* ncol */
  double *tvector;                                         /* in/out (ncol, -nlevsno:nlev): only the entries the driver assigns change */
} mppgpu_elm_thermal_columns;
int  mppgpu_thermal_elm_solve(mppgpu_handle h, double dtime, int nstep, const mppgpu_elm_thermal_columns *cols, double capr);
/* The two ELM solve entry points (mppgpu_thermal_elm_solve above, mppgpu_vsfm_elm_solve below) are software-pipelined over column chunks
 * on three CUDA streams -- chunk k+1's host->device copies, chunk k's kernels and chunk k-1's device->host copies overlap -- with results
 * bit-identical to an unpipelined solve.  nchunks: number of chunks (0 = default: 8, none smaller than 32768 columns; 1 = no pipelining).
 * static_soil_geometry (thermal SoE only): 1 = the soil rows (j >= 1) of z / dz / zi are ELM's fixed vertical grid (initVertical); they are
 * uploaded by the first solve after this call and only the snow rows (and zi(c,0)) on later solves; 0 (default) = every row, every solve.
 * Host arrays should be page-locked (mppgpu_host_register) for the copies to overlap. */
int  mppgpu_elm_set_pipeline(mppgpu_handle h, int nchunks, int static_soil_geometry);
int  mppgpu_th_set_soils(mppgpu_handle h, const double *watsat, const double *hksat, const double *bsw,
                         const double *sucsat, const double *residual_sat, const double *csol, const double *tkdry,
                         int satfunc_type, int density_type, int int_energy_enthalpy_type);
/* goveq_enthalpy%SetSoilPermeability (GoveqnThermalEnthalpySoilType.F90:2454-2480, called by th_mms_problem.F90:739 and
 * mass_and_heat_model_problem.F90): per-cell permeability [m^2] of the ENERGY equation's aux vars, cell order, n = ncol*nlev.  MPPTHSetSoils
 * leaves them at the aux-var default 8.3913e-12 (MultiPhysicsProbTH.F90:293 is commented out); so does this library until this is called. */
int  mppgpu_th_set_energy_permeability(mppgpu_handle h, const double *perm, int n);
int  mppgpu_set_tolerances(mppgpu_handle h, double atol, double rtol, double stol, int max_it, int max_funcs);
/* NOT in the reference (default 0 = off = the reference's behaviour).  The reference's StepDT halves dt up to 20 times and then
 * sub-steps with the smallest dt that converged, so one pathological column can spend millions of residual evaluations in one
 * StepDT; in a batch that stalls every other column of the launch.  With a budget > 0 a column that has used this many residual
 * evaluations inside one StepDT gives up exactly like one that ran out of dt cuts: converged = 0, reason
 * SNES_DIVERGED_FUNCTION_COUNT (-2), solution left at the last converged sub-step. */
int  mppgpu_set_step_budget(mppgpu_handle h, int max_residual_evaluations);
/* NOT in the reference (scheduling only, results are identical either way; VSFM with nlev <= 32, TH with nlev <= 16).  mode 1 (default):
 * the step kernel visits the columns grouped by the number of residual evaluations their previous StepDT needed, most expensive first, so
 * that the columns a warp advances (four, TH: two) finish together and the longest jobs start first; mode 0: batch order. */
int  mppgpu_set_column_ordering(mppgpu_handle h, int mode);
/* VSFM/thermal: x has ncells entries (pressure or temperature); TH: 2*ncells, [P(0..N-1) | T(0..N-1)] */
/* NOT in the reference (scheduling only, results are bit-identical either way; soil thermal SoE, nlev <= 16).  mode 0 (default): every
 * warp loads its own columns into registers; mode 1: batches that fill the GPU run on the persistent kernel whose inputs arrive by
 * bulk-async copies (cp.async.bulk, the 1-D TMA path) into shared-memory stages -- measured 23 % slower on B200 (DESIGN.md section 4.2:
 * the step is bound by the latency of its fp64 chains, not by its loads) and kept as an option for shapes where that may differ. */
int  mppgpu_thermal_set_bulk_copy(mppgpu_handle h, int mode);
int  mppgpu_restart(mppgpu_handle h, const double *x, int n);

/* ---- per-step data exchange ------------------------------------------------------------------- */
int  mppgpu_set_data(mppgpu_handle h, int ieqn, int auxvar_type, int var_type, int cond_id, const double *data, int n);
int  mppgpu_set_idata(mppgpu_handle h, int ieqn, int auxvar_type, int var_type, int cond_id, const int *data, int n);
int  mppgpu_get_data(mppgpu_handle h, int ieqn, int auxvar_type, int var_type, int cond_id, double *data, int n);
int  mppgpu_set_data_device(mppgpu_handle h, int ieqn, int auxvar_type, int var_type, int cond_id, const double *d_data, int n);
int  mppgpu_get_data_device(mppgpu_handle h, int ieqn, int auxvar_type, int var_type, int cond_id, double *d_data, int n);

/* ---- time stepping ---------------------------------------------------------------------------- */
int  mppgpu_pre_step_dt(mppgpu_handle h);
int  mppgpu_step_dt(mppgpu_handle h, double dt, int nstep, int *converged, int *converged_reason);
/* same, without the blocking read-back of converged/converged_reason (query them later) */
int  mppgpu_step_dt_async(mppgpu_handle h, double dt, int nstep);
int  mppgpu_step_result(mppgpu_handle h, int *converged, int *converged_reason);
int  mppgpu_post_step_dt(mppgpu_handle h);

/* One ELM coupling step of the VSFM SoE with HOST buffers, software-pipelined over column chunks:
 *   for every in[i]:  SetDataFromCLM(in[i]);   PreStepDT;  StepDT(dt, nstep);   for every out[i]: GetDataForCLM(out[i])
 * (MPPVSFMALM_Driver.F90:379-463, 603, 642, 674-705) -- same results as those separate calls, but the host->device copies of
 * chunk k+1, the Newton kernel of chunk k and the device->host copies of chunk k-1 overlap on three CUDA streams.
 * Host arrays should be page-locked (cudaHostRegister / pinned allocation) for the copies to be asynchronous; pageable
 * memory works but serialises.  Every array covers the whole handle (ncells or ncol entries, by the condition's region).
 * nchunks <= 0 picks a default.  The caller still calls PostStepDT to commit. */
typedef struct { int ieqn, auxvar_type, var_type, cond_id; double *host; } mppgpu_xfer;
int  mppgpu_vsfm_coupled_step(mppgpu_handle h, double dt, int nstep, int nin, const mppgpu_xfer *in, int nout, const mppgpu_xfer *out,
                              int nchunks, int *converged, int *converged_reason);

/* ---- MPPVSFMALM_Solve with ELM's raw column arrays (src/driver/alm/MPPVSFMALM_Driver.F90:204-923) -----------------
 * Everything the host model's driver does around StepDT runs on the device: root-fraction weighting of transpiration over the patches
 * of a column (:204-240, optional), source/sink packing incl. the drainage distribution below the water table (:325-404),
 * frac_liq_sat from the ice fraction (:435-450), the retry loop (:628-923; <= 10 StepDT calls: a diverged step continues with the
 * remaining time and stol = 1e-10, a second divergence sets frac_liq_sat = 1; a converged step whose mass-balance error is >= 1e-5 kg
 * is redone from soln_prev_clm with rtol or stol tightened tenfold), and unpacking (h2osoi_liq / h2osoi_ice, smp_l [mm], soil
 * pressure, water-table depth, qcharge = 0), then PostStepDT.  The reference takes the retry decisions per MPI rank; here per column.
 * Lateral-flux and seepage branches are not part of the 1-D path (no boundary conditions allowed).  All arrays are HOST pointers;
 * per-cell arrays are cell-ordered (c*nlev + j); `zi` has nlev+1 interfaces per column, zi(c,0) first.  nlev <= 32.  Columns filtered out by
 * mppgpu_set_mesh's col_active keep their in/out arrays and read 0 in the pure outputs (smp_l, soilp_col, qcharge, status). */
typedef struct {
  /* layout of the per-cell arrays marked (cells) below: 0 = cell-ordered 1-D vectors (c*nlev + j), 1 = ELM's own (c, j) Fortran arrays
   * (column index fastest: value of column c at layer j = 1..nlev at [(j-1)*ncol + c]; pass the address of element (begc, 1)) */
  int fortran_order;
  /* patch level, optional (npft = 0: rootr_col is an input) -- col%pfti (0-based), col%npfts, pft%active, pft%wtcol, rootr_patch(p,j), qflx_tran_veg_patch */
  int npft, max_patch_per_col;
  const int *col_pfti, *col_npfts, *pft_active; const double *pft_wtcol, *rootr_pft, *qflx_tran_veg_pft;
  /* column level */
  double *rootr_col;                     /* (cells); input, or output when patches are given */
  const double *qflx_tran_veg_col, *qflx_infl, *qflx_dew_snow, *qflx_dew_grnd, *qflx_sub_snow, *frac_h2osfc;   /* ncol, [mm/s] */
  const int *snl;                        /* ncol, minus the number of snow layers */
  double *qflx_drain, *zwt;              /* ncol, in/out */
  double *h2osoi_liq, *h2osoi_ice;       /* (cells), in/out [kg/m^2] */
  double *mflx_snowlyr_col;              /* ncol, in/out (zeroed) */
  const double *mflx_neg_snow_col;       /* ncol */
  const double *mflx_drain_perched;      /* ncells, always cell-ordered (ELM keeps it as mflx_drain_perched_col_1d) */
  /* outputs */
  double *smp_l, *soilp_col;             /* (cells): matric potential [mm], soil water pressure [Pa] */
  double *qcharge;                       /* ncol */
  double *abs_mass_error;                /* ncol or NULL */
  int *iter_count, *status;              /* ncol or NULL: StepDT calls used; 1 = accepted, 0 = failed all retries (the reference would endrun) */
} mppgpu_elm_columns;
/* static part: interface depths zi (ncol*(nlev+1)), thicknesses dz (ncells, cell-ordered), nlevsoi, clm_varcon's watmin (0.01 mm), and the
 * ids of the six COND_MASS_RATE sources in the order infiltration, ET, dew, drainage, snow, sublimation (MPPVSFMALM_Initialize.F90:836-858) */
int  mppgpu_vsfm_elm_set_geometry(mppgpu_handle h, const double *zi, const double *dz, int nlevsoi, double watmin, const int *cond_ids);
/* the same with ELM's (c, j) arrays: zi = address of col%zi(begc, 0) (nlev+1 layers), dz = address of col%dz(begc, 1) */
int  mppgpu_vsfm_elm_set_geometry_f(mppgpu_handle h, const double *zi, const double *dz, int nlevsoi, double watmin, const int *cond_ids);
/* After the call mppgpu_vsfm_mass_balance / mppgpu_reduction_buffer_device describe the whole solve (retries included). */
int  mppgpu_vsfm_elm_solve(mppgpu_handle h, double dtime, int nstep, mppgpu_elm_columns *cols, int *nfailed, int *nattempts);

/* ---- host arrays ------------------------------------------------------------------------------ */
/* Page-lock a host array the caller owns for the rest of the run (ELM's column arrays are allocated once:
 * clm_instMod / ColumnDataType), so that every later copy of it by this library (mppgpu_set_data, mppgpu_get_data,
 * mppgpu_vsfm_coupled_step, mppgpu_*_elm_solve) is a direct PCIe DMA instead of a staged pageable copy.  A host model
 * written in Fortran needs no CUDA runtime binding for this.  Unregister before the array is freed.  Optional: every
 * entry point accepts pageable memory.  Registering a range twice (or two small arrays that share a page) is refused by
 * the CUDA driver and reported as an error; the library stays usable. */
int  mppgpu_host_register(void *ptr, long long nbytes);
int  mppgpu_host_unregister(void *ptr);

/* ---- diagnostics ------------------------------------------------------------------------------ */
/* per-column Newton iterations, SNES reason, dt cuts, residual evaluations of the last StepDT (any may be NULL) */
int  mppgpu_get_column_stats(mppgpu_handle h, int *newton_its, int *reasons, int *dt_cuts, int *nfuncs);
/* Rank-local reduction of the ELM driver's mass-balance bookkeeping (MPPVSFMALM_Driver.F90:556-601,845-863):
 * sums[0] = sum mass at the start of the last StepDT, sums[1] = at its end, sums[2] = sum of all mass-rate
 * sources * dt, sums[3] = sum boundary mass exchanged; maxs[0] = max per-column |m_beg - m_end + q dt| [kg],
 * maxs[1] = max Newton its, maxs[2] = any column diverged (0/1), maxs[3] = max dt cuts.
 * The same 8 doubles live in a device buffer (mppgpu_reduction_buffer_device: sums[4] then maxs[4]) so a
 * multi-GPU driver can all-reduce them with NCCL without a host round trip. */
int  mppgpu_vsfm_mass_balance(mppgpu_handle h, double dt, double sums[4], double maxs[4]);
int  mppgpu_reduction_buffer_device(mppgpu_handle h, double **d_buf);

/* ---- global reductions over the ranks of a column-sharded batch ------------------
 * The reference keeps the mass bookkeeping of MPPVSFMALM_Driver.F90:124-133, 556-601, 845-898 per MPI rank; with the batch
 * sharded by column over the GPUs of one box these calls give the global figures: ONE ncclAllGather of the 9 doubles above
 * per rank per step on the handle's stream, folded on the device in rank order.  NCCL is loaded with dlopen by
 * mppgpu_comm_unique_id / mppgpu_comm_init (no link-time dependency).  Usage from an MPI host model: rank 0 calls
 * mppgpu_comm_unique_id, MPI_Bcast of the MPPGPU_COMM_ID_BYTES bytes, every rank calls mppgpu_comm_init(h, nranks, rank, id)
 * once per handle; then after every StepDT every rank calls mppgpu_global_mass_balance (collective; blocks until the result
 * is on the host) or mppgpu_global_reduce_async (collective; only queues the work, a later mppgpu_global_mass_balance returns
 * it).  nranks = 1 needs no id and no NCCL.  worst_reason is the minimum SNESConvergedReason over all columns of all ranks. */
#define MPPGPU_COMM_ID_BYTES 128
int  mppgpu_comm_unique_id(void *id128);
int  mppgpu_comm_init(mppgpu_handle h, int nranks, int rank, const void *id128);
int  mppgpu_global_reduce_async(mppgpu_handle h);
int  mppgpu_global_mass_balance(mppgpu_handle h, double sums[4], double maxs[4], int *worst_reason);
/* launches made by the library since creation, and device-time of the last StepDT kernel(s) in ms */
int  mppgpu_launch_count(mppgpu_handle h, long long *n);
int  mppgpu_last_step_ms(mppgpu_handle h, float *ms);
/* One residual + Jacobian evaluation at x with the accumulation term of the start of the step taken at x_prev: the finer seam of the
 * reference, SOEResidual / SOEJacobian (SystemOfEquationsBasePointerType.F90:40-109 -> VSFMSOEResidual / VSFMJacobian
 * SystemOfEquationsVSFMType.F90:94-403, SOETHResidual / SOETHJacobian SystemOfEquationsTHType.F90:736-1004), as a debugging cross-check.
 * VSFM (nlev <= 32): f and the sub-, main and super-diagonal ja, jb, jc of the tridiagonal Jacobian, ncells values each, cell order,
 * written by the EVAL instance of the fused step kernel (the same assembly code the time step runs).  TH: x interleaved (P, T) per
 * cell, f of 2N, 2x2 row-major blocks ja, jb, jc of 4N each.  Not available for the (linear) thermal SoE. */
int  mppgpu_eval(mppgpu_handle h, double dt, const double *x_prev, const double *x, double *f, double *ja, double *jb, double *jc);

#ifdef __cplusplus
}
#endif
#endif
