/* cabi_smoke.c -- the C ABI driven from plain C99 (no C++, no Python): proves include/mppgpu.h is a valid C header and that
 * create / set / step / get work for a caller that only has the shared library.
 *   gcc -std=c99 -Wall -Wextra -pedantic -Iinclude tests/cabi_smoke.c -o cabi_smoke -Lmpp_b200 -lmppgpu -lm
 * With no argument the program only checks that the library loads and reports its version and device count (CPU suite);
 * with "run" it steps 4 ELM-like columns x 15 layers on device 0 and checks the mass balance (GPU suite).
 * The sequence is the one the Fortran shim makes: MPPSetupProblem -> VSFMMPPSetSoils -> Restart -> SetDataFromCLM -> PreStepDT
 * -> StepDT -> GetDataForCLM (vsfm_celia1990_problem.F90:106-137, 383-394; MPPVSFMALM_Driver.F90:379-463, 603, 642, 674-705). */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "mppgpu.h"

#define NCOL 4
#define NLEV 15
#define CHECK(call) do { if ((call) != 0) { fprintf(stderr, "FAILED %s: %s\n", #call, mppgpu_last_error()); return 1; } } while (0)

int main(int argc, char **argv)
{
  printf("mppgpu version %d, %d CUDA device(s)\n", mppgpu_version(), mppgpu_device_count());
  if (mppgpu_version() < 100) return 1;
  if (argc < 2 || strcmp(argv[1], "run") != 0) return 0;

  mppgpu_handle h = NULL;
  CHECK(mppgpu_create(MPPGPU_SOE_RE_ODE, NCOL, NLEV, 0, &h));
  /* ELM node depths z_j = 0.025 (exp(0.5 (j - 0.5)) - 1); tables are Fortran (c, j): t[j * ncol + c] */
  double z[NLEV], zi[NLEV + 1], dz[NLEV * NCOL], area[NCOL], tab[5][NLEV * NCOL], x0[NLEV * NCOL];
  int j, c;
  for (j = 0; j < NLEV; ++j) z[j] = 0.025 * (exp(0.5 * (j + 0.5)) - 1.0);
  zi[0] = 0.0;
  for (j = 1; j < NLEV; ++j) zi[j] = 0.5 * (z[j - 1] + z[j]);
  zi[NLEV] = z[NLEV - 1] + 0.5 * (z[NLEV - 1] - z[NLEV - 2]);
  for (c = 0; c < NCOL; ++c) {
    area[c] = 1.0;
    for (j = 0; j < NLEV; ++j) {
      const int t = j * NCOL + c;
      dz[t] = zi[j + 1] - zi[j];
      tab[0][t] = 0.40 + 0.02 * c;      /* watsat */
      tab[1][t] = 0.005 * (1 + c);      /* hksat [mm/s] */
      tab[2][t] = 4.0 + c;              /* bsw */
      tab[3][t] = 100.0 + 50.0 * c;     /* sucsat [mm] */
      tab[4][t] = 0.0;                  /* residual saturation */
      /* hydrostatic start, water table at 3 m (MPPVSFMALM_Initialize.F90:1058-1060); vectors are cell ordered: c * nlev + j */
      x0[c * NLEV + j] = 101325.0 + 997.16 * 9.80665 * (0.5 * (zi[j] + zi[j + 1]) - 3.0);
    }
  }
  CHECK(mppgpu_set_mesh(h, MPPGPU_MESH_ALONG_GRAVITY, dz, area, NULL));
  int infil = 0;
  CHECK(mppgpu_add_condition(h, 1, 502 /* COND_SS */, 503 /* COND_MASS_RATE */, 401 /* SOIL_TOP_CELLS */, &infil));
  CHECK(mppgpu_vsfm_set_soils(h, tab[0], tab[1], tab[2], tab[3], tab[4], MPPGPU_SATFUNC_VAN_GENUCHTEN, 2 /* DENSITY_TGDPB01 */));
  CHECK(mppgpu_restart(h, x0, NCOL * NLEV));
  CHECK(mppgpu_comm_init(h, 1, 0, NULL));
  double rate[NCOL] = {1e-5, 2e-5, 3e-5, 4e-5}, mass0[NCOL * NLEV], mass1[NCOL * NLEV], P[NCOL * NLEV], sat[NCOL * NLEV];
  CHECK(mppgpu_get_data(h, 1, 701 /* AUXVAR_INTERNAL */, 610 /* VAR_MASS */, 1, mass0, NCOL * NLEV));
  CHECK(mppgpu_set_data(h, 1, 703 /* AUXVAR_SS */, 607 /* VAR_BC_SS_CONDITION */, infil, rate, NCOL));
  int converged = 0, reason = 0;
  CHECK(mppgpu_pre_step_dt(h));
  CHECK(mppgpu_step_dt(h, 1800.0, 1, &converged, &reason));
  CHECK(mppgpu_get_data(h, 1, 701, 604 /* VAR_PRESSURE */, 1, P, NCOL * NLEV));
  CHECK(mppgpu_get_data(h, 1, 701, 608 /* VAR_LIQ_SAT */, 1, sat, NCOL * NLEV));
  CHECK(mppgpu_get_data(h, 1, 701, 610, 1, mass1, NCOL * NLEV));
  CHECK(mppgpu_post_step_dt(h));
  double sums[4], maxs[4]; int worst = 0;
  CHECK(mppgpu_global_mass_balance(h, sums, maxs, &worst));
  printf("converged %d reason %d worst %d max |mass error| %.3e kg\n", converged, reason, worst, maxs[0]);
  if (!converged || reason < 2 || reason > 4 || worst != reason) return 2;
  for (c = 0; c < NCOL; ++c) {
    double m0 = 0.0, m1 = 0.0;
    for (j = 0; j < NLEV; ++j) { m0 += mass0[c * NLEV + j]; m1 += mass1[c * NLEV + j]; if (!(sat[c * NLEV + j] > 0.0 && sat[c * NLEV + j] <= 1.0)) return 3; }
    const double err = fabs(m0 - m1 + rate[c] * 1800.0);
    printf("column %d: mass %.6f -> %.6f kg, balance error %.3e kg, P(top) %.3f Pa\n", c, m0, m1, err, P[c * NLEV]);
    if (!(err < 1e-5)) return 4;                     /* the reference's gate, MPPVSFMALM_Driver.F90:140 */
  }
  if (!(maxs[0] < 1e-5)) return 5;
  int its[NCOL], rs[NCOL], cuts[NCOL], nf[NCOL];
  CHECK(mppgpu_get_column_stats(h, its, rs, cuts, nf));
  for (c = 0; c < NCOL; ++c) if (its[c] < 1 || its[c] > 50 || cuts[c] != 0) return 6;
  CHECK(mppgpu_destroy(h));
  printf("ok\n");
  return 0;
}
