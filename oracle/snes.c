/*
 * snes.c -- oracle restatement of the solver arithmetic the reference delegates
 * to PETSc (an un-vendored dependency; pinned v3.16.2, README.md:35):
 *   SNESSolve_NEWTONLS        (src/snes/impls/ls/ls.c in PETSc 3.16)
 *   SNESLineSearchApply_BT    (src/snes/linesearch/impls/bt/linesearchbt.c)
 *   SNESConvergedDefault      (src/snes/interface/snesut.c)
 *   KSPSolve GMRES + PCILU(0) on a (block-)tridiagonal AIJ matrix: ILU(0) in
 *   natural ordering of a tridiagonal matrix is its exact LU, so GMRES
 *   converges in one iteration to the Thomas-algorithm solution (SURVEY.md
 *   section 8a3).  For the TH 2x2-block system the reference's segregated
 *   ordering makes ILU(0) inexact; the oracle solves the Newton system exactly
 *   (SURVEY.md section 7 "hard parts" (c), Appendix C).
 * Reference call sites: SNESSetTolerances MultiPhysicsProbBaseType.F90:1110-1114,1196
 * (atol 1e-50, rtol 1e-8, stol 1e-10, max_it 50, max_funcs 10000); SNESSolve
 * SystemOfEquationsBaseType.F90:478.
 *
 * TEST INFRASTRUCTURE ONLY (see mpp_oracle.h).
 *
 * Parity-unpinned PETSc behaviours (no reference golden exercises them):
 * lambda <= minlambda abort, "stol*xnorm > ynorm" early exit, maxstep clipping,
 * NaN/Inf back-off, max_funcs exhaustion.
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "mpp_oracle.h"

/* ORC_SNES_MONITOR=1 in the environment prints one line per Newton iteration (debugging aid for hard columns) */
static int snes_monitor(void)
{
  static int on = -1;
  if (on < 0) { const char *e = getenv("ORC_SNES_MONITOR"); on = (e && e[0] == '1') ? 1 : 0; }
  return on;
}

void orc_snes_default_opts(orc_snes_opts *o)
{
  o->atol = 1.e-50; o->rtol = 1.e-8; o->stol = 1.e-10; o->divtol = 1.e4;   /* MultiPhysicsProbBaseType.F90:1110-1114 + PETSc defaults */
  o->max_it = 50; o->max_funcs = 10000;
  o->ls_alpha = 1.e-4; o->ls_minlambda = 1.e-12; o->ls_maxstep = 1.e8; o->ls_max_its = 40;  /* SNESLineSearch defaults */
}

/* Thomas algorithm == LU without pivoting in natural ordering == PETSc ILU(0)+GMRES on a tridiagonal AIJ */
void orc_tridiag_solve(int n, const double *a, const double *b, const double *c, const double *d, double *x)
{
  double *cp = (double *)malloc(sizeof(double) * (size_t)n * 2), *dp = cp + n;
  int i;
  double m;
  cp[0] = c[0] / b[0];
  dp[0] = d[0] / b[0];
  for (i = 1; i < n; i++) {
    m     = b[i] - a[i] * cp[i - 1];
    cp[i] = c[i] / m;
    dp[i] = (d[i] - a[i] * dp[i - 1]) / m;
  }
  x[n - 1] = dp[n - 1];
  for (i = n - 2; i >= 0; i--) x[i] = dp[i] - cp[i] * x[i + 1];
  free(cp);
}

/* 2x2 helpers, row-major blocks [m00 m01 m10 m11] */
static void inv2(const double *m, double *r)
{
  double det = m[0] * m[3] - m[1] * m[2];
  r[0] = m[3] / det; r[1] = -m[1] / det; r[2] = -m[2] / det; r[3] = m[0] / det;
}
static void mul22(const double *x, const double *y, double *r)
{
  r[0] = x[0] * y[0] + x[1] * y[2]; r[1] = x[0] * y[1] + x[1] * y[3];
  r[2] = x[2] * y[0] + x[3] * y[2]; r[3] = x[2] * y[1] + x[3] * y[3];
}
static void mul21(const double *x, const double *v, double *r)
{
  r[0] = x[0] * v[0] + x[1] * v[1]; r[1] = x[2] * v[0] + x[3] * v[1];
}

/* block Thomas for 2x2 blocks; a,b,c are ncell blocks of 4 doubles, d and x are 2*ncell */
void orc_blocktridiag2_solve(int ncell, const double *a, const double *b, const double *c, const double *d, double *x)
{
  double *cp = (double *)malloc(sizeof(double) * (size_t)ncell * 6), *dp = cp + 4 * (size_t)ncell;
  double binv[4], m[4], t[4], v[2];
  int i;
  inv2(b, binv);
  mul22(binv, c, cp);
  mul21(binv, d, dp);
  for (i = 1; i < ncell; i++) {
    mul22(a + 4 * i, cp + 4 * (i - 1), t);
    m[0] = b[4 * i] - t[0]; m[1] = b[4 * i + 1] - t[1]; m[2] = b[4 * i + 2] - t[2]; m[3] = b[4 * i + 3] - t[3];
    inv2(m, binv);
    mul22(binv, c + 4 * i, cp + 4 * i);
    mul21(a + 4 * i, dp + 2 * (i - 1), v);
    v[0] = d[2 * i] - v[0]; v[1] = d[2 * i + 1] - v[1];
    mul21(binv, v, dp + 2 * i);
  }
  x[2 * (ncell - 1)] = dp[2 * (ncell - 1)]; x[2 * (ncell - 1) + 1] = dp[2 * (ncell - 1) + 1];
  for (i = ncell - 2; i >= 0; i--) {
    mul21(cp + 4 * i, x + 2 * (i + 1), v);
    x[2 * i] = dp[2 * i] - v[0]; x[2 * i + 1] = dp[2 * i + 1] - v[1];
  }
  free(cp);
}

static double norm2(int n, const double *v)
{
  double s = 0.0; int i;
  for (i = 0; i < n; i++) s += v[i] * v[i];
  return sqrt(s);
}

/* w = J y for the block-tridiagonal J (MatMult in SNESLineSearchApply_BT) */
static void jac_mult(const orc_system *sys, const double *a, const double *b, const double *c, const double *y, double *w)
{
  int ch, i, bs = sys->bs;
  for (ch = 0; ch < sys->nchain; ch++) {
    int lo = sys->col_start[ch], hi = sys->col_start[ch + 1];
    for (i = lo; i < hi; i++) {
      if (bs == 1) {
        double s = b[i] * y[i];
        if (i > lo)     s += a[i] * y[i - 1];
        if (i < hi - 1) s += c[i] * y[i + 1];
        w[i] = s;
      } else {
        double s0 = b[4 * i] * y[2 * i] + b[4 * i + 1] * y[2 * i + 1];
        double s1 = b[4 * i + 2] * y[2 * i] + b[4 * i + 3] * y[2 * i + 1];
        if (i > lo) {
          s0 += a[4 * i] * y[2 * i - 2] + a[4 * i + 1] * y[2 * i - 1];
          s1 += a[4 * i + 2] * y[2 * i - 2] + a[4 * i + 3] * y[2 * i - 1];
        }
        if (i < hi - 1) {
          s0 += c[4 * i] * y[2 * i + 2] + c[4 * i + 1] * y[2 * i + 3];
          s1 += c[4 * i + 2] * y[2 * i + 2] + c[4 * i + 3] * y[2 * i + 3];
        }
        w[2 * i] = s0; w[2 * i + 1] = s1;
      }
    }
  }
}

static void lin_solve(const orc_system *sys, const double *a, const double *b, const double *c, const double *f, double *y)
{
  int ch, bs = sys->bs;
  for (ch = 0; ch < sys->nchain; ch++) {
    int lo = sys->col_start[ch], n = sys->col_start[ch + 1] - lo;
    if (bs == 1) orc_tridiag_solve(n, a + lo, b + lo, c + lo, f + lo, y + lo);
    else         orc_blocktridiag2_solve(n, a + 4 * lo, b + 4 * lo, c + 4 * lo, f + 2 * lo, y + 2 * lo);
  }
}

/* SNESConvergedDefault (PETSc 3.16 snesut.c) */
static int converged_default(const orc_snes_opts *o, int it, double xnorm, double snorm, double fnorm,
                             double *ttol, double *rnorm0, int nfuncs)
{
  int reason = SNES_CONVERGED_ITERATING;
  if (!it) { *ttol = fnorm * o->rtol; *rnorm0 = fnorm; }
  if (isnan(fnorm) || isinf(fnorm))                          reason = SNES_DIVERGED_FNORM_NAN;
  else if (fnorm < o->atol)                                  reason = SNES_CONVERGED_FNORM_ABS;   /* (it || !forceiteration), forceiteration = false */
  else if (nfuncs >= o->max_funcs && o->max_funcs >= 0)      reason = SNES_DIVERGED_FUNCTION_COUNT;
  if (it && !reason) {
    if (fnorm <= *ttol)                                      reason = SNES_CONVERGED_FNORM_RELATIVE;
    else if (snorm < o->stol * xnorm)                        reason = SNES_CONVERGED_SNORM_RELATIVE;
    else if (o->divtol > 0 && fnorm > o->divtol * (*rnorm0)) reason = SNES_DIVERGED_DTOL;
  }
  return reason;
}

enum { LS_SUCCEEDED = 0, LS_FAILED_NANORINF = 1, LS_FAILED_REDUCT = 3, LS_FAILED_FUNCTION = 5 };

/*
 * SNESSolve_NEWTONLS with SNESLineSearchApply_BT (cubic order, the default).
 * x is updated in place.  Work vectors are allocated here.
 */
void orc_snes_solve(const orc_system *sys, const orc_snes_opts *o, double *X, orc_snes_result *res)
{
  int n = sys->n, bs = sys->bs, ncell = sys->ncell, i, it, its = 0, reason = 0, nfuncs = 0;
  size_t nb = (size_t)ncell * (size_t)(bs * bs);
  double *F = (double *)malloc(sizeof(double) * (size_t)n * 4);
  double *Y = F + n, *W = Y + n, *G = W + n;
  double *ja = (double *)calloc(nb * 3, sizeof(double)), *jb = ja + nb, *jc = jb + nb;
  double fnorm, xnorm = 0.0, ynorm = 0.0, gnorm = 0.0, ttol = 0.0, rnorm0 = 0.0, lambda = 1.0;

  sys->residual(sys->ctx, X, F); nfuncs++;
  fnorm = norm2(n, F);
  res->fnorm0 = fnorm;
  if (isnan(fnorm) || isinf(fnorm)) { reason = SNES_DIVERGED_FNORM_NAN; goto done; }
  reason = converged_default(o, 0, 0.0, 0.0, fnorm, &ttol, &rnorm0, nfuncs);
  if (reason) goto done;

  for (it = 0; it < o->max_it; it++) {
    int ls_reason = LS_SUCCEEDED, count;
    double f, g, gprev = 0.0, initslope, lambdatemp, lambdaprev = 0.0, t1, t2, a, b, d;

    sys->jacobian(sys->ctx, X, ja, jb, jc);
    lin_solve(sys, ja, jb, jc, F, Y);                    /* J Y = F */

    /* ---- SNESLineSearchApply_BT ---- */
    lambda = 1.0;                                         /* damping */
    ynorm = norm2(n, Y);
    xnorm = norm2(n, X);
    if (ynorm == 0.0) {
      memcpy(W, X, sizeof(double) * (size_t)n); memcpy(G, F, sizeof(double) * (size_t)n);
      gnorm = fnorm; ls_reason = LS_FAILED_REDUCT; goto ls_done_nocopy;
    }
    if (ynorm > o->ls_maxstep) {
      double s = o->ls_maxstep / ynorm;
      for (i = 0; i < n; i++) Y[i] *= s;
      ynorm = o->ls_maxstep;
    }
    f = fnorm * fnorm;
    jac_mult(sys, ja, jb, jc, Y, W);
    initslope = 0.0;
    for (i = 0; i < n; i++) initslope += F[i] * W[i];
    if (initslope > 0.0)  initslope = -initslope;
    if (initslope == 0.0) initslope = -1.0;

    for (;;) {
      for (i = 0; i < n; i++) W[i] = X[i] - lambda * Y[i];
      if (nfuncs >= o->max_funcs && o->max_funcs >= 0) { reason = SNES_DIVERGED_FUNCTION_COUNT; ls_reason = LS_FAILED_FUNCTION; goto ls_done_nocopy; }
      sys->residual(sys->ctx, W, G); nfuncs++;
      gnorm = norm2(n, G);
      g = gnorm * gnorm;
      if (!(isnan(g) || isinf(g))) break;
      if (lambda <= o->ls_minlambda) { reason = SNES_DIVERGED_FNORM_NAN; ls_reason = LS_FAILED_NANORINF; goto ls_done_nocopy; }
      lambda = .5 * lambda;
    }

    if (.5 * g <= .5 * f + lambda * o->ls_alpha * initslope) {
      /* sufficient reduction with the full step */
    } else {
      if (o->stol * xnorm > ynorm) { ls_reason = LS_FAILED_REDUCT; goto ls_done_nocopy; }
      /* quadratic fit */
      lambdatemp = -initslope / (g - f - 2.0 * lambda * initslope);
      lambdaprev = lambda;
      gprev      = g;
      if (lambdatemp > .5 * lambda)  lambdatemp = .5 * lambda;
      if (lambdatemp <= .1 * lambda) lambda = .1 * lambda;
      else                           lambda = lambdatemp;

      for (i = 0; i < n; i++) W[i] = X[i] - lambda * Y[i];
      if (nfuncs >= o->max_funcs && o->max_funcs >= 0) { reason = SNES_DIVERGED_FUNCTION_COUNT; ls_reason = LS_FAILED_FUNCTION; goto ls_done_nocopy; }
      sys->residual(sys->ctx, W, G); nfuncs++;
      gnorm = norm2(n, G);
      g = gnorm * gnorm;
      if (isnan(g) || isinf(g)) { ls_reason = LS_FAILED_NANORINF; goto ls_done_nocopy; }
      if (.5 * g < .5 * f + lambda * o->ls_alpha * initslope) {
        /* quadratically determined step accepted */
      } else {
        /* cubic fits */
        for (count = 0; count < o->ls_max_its; count++) {
          if (lambda <= o->ls_minlambda) { ls_reason = LS_FAILED_REDUCT; goto ls_done_nocopy; }
          t1 = .5 * (g - f) - lambda * initslope;
          t2 = .5 * (gprev - f) - lambdaprev * initslope;
          a  = (t1 / (lambda * lambda) - t2 / (lambdaprev * lambdaprev)) / (lambda - lambdaprev);
          b  = (-lambdaprev * t1 / (lambda * lambda) + lambda * t2 / (lambdaprev * lambdaprev)) / (lambda - lambdaprev);
          d  = b * b - 3 * a * initslope;
          if (d < 0.0) d = 0.0;
          if (a == 0.0) lambdatemp = -initslope / (2.0 * b);
          else          lambdatemp = (-b + sqrt(d)) / (3.0 * a);
          lambdaprev = lambda;
          gprev      = g;
          if (lambdatemp > .5 * lambda)  lambdatemp = .5 * lambda;
          if (lambdatemp <= .1 * lambda) lambda = .1 * lambda;
          else                           lambda = lambdatemp;
          for (i = 0; i < n; i++) W[i] = X[i] - lambda * Y[i];
          if (nfuncs >= o->max_funcs && o->max_funcs >= 0) { reason = SNES_DIVERGED_FUNCTION_COUNT; ls_reason = LS_FAILED_FUNCTION; goto ls_done_nocopy; }
          sys->residual(sys->ctx, W, G); nfuncs++;
          gnorm = norm2(n, G);
          g = gnorm * gnorm;
          if (isnan(g) || isinf(g)) { ls_reason = LS_FAILED_NANORINF; goto ls_done_nocopy; }
          if (.5 * g < .5 * f + lambda * o->ls_alpha * initslope) break;
        }
        /* PETSc falls out of the loop after max_its fits and accepts the last trial point */
      }
    }
    /* success: copy the solution over */
    memcpy(X, W, sizeof(double) * (size_t)n);
    memcpy(F, G, sizeof(double) * (size_t)n);
    xnorm = norm2(n, X);
    fnorm = gnorm;
    goto ls_done;

ls_done_nocopy:
    /* failure paths return before "copy the solution over": X and F keep their old
     * values; the norms handed back are (xnorm, fnorm, ynorm) as set so far */
    ;
ls_done:
    /* ---- back in SNESSolve_NEWTONLS ---- */
    if (reason) break;                            /* set inside the line search (function count / NaN) */
    if (isnan(fnorm) || isinf(fnorm)) { reason = SNES_DIVERGED_FNORM_NAN; break; }
    if (ls_reason) {
      if (o->stol * xnorm > ynorm) { reason = SNES_CONVERGED_SNORM_RELATIVE; break; }
      reason = SNES_DIVERGED_LINE_SEARCH;         /* maxFailures = 1 */
      break;
    }
    its = it + 1;                                 /* snes->iter */
    reason = converged_default(o, its, xnorm, ynorm, fnorm, &ttol, &rnorm0, nfuncs);
    if (snes_monitor()) {
      int im = 0, iy = 0;
      for (i = 1; i < n; i++) { if (fabs(F[i]) > fabs(F[im])) im = i; if (fabs(Y[i]) > fabs(Y[iy])) iy = i; }
      fprintf(stderr, "  snes it %d fnorm %.6e ynorm %.6e xnorm %.6e lambda %.3e nfuncs %d reason %d | max|F| at %d: F %.3e X %.9e | max|Y| at %d: Y %.3e X %.9e\n",
              its, fnorm, ynorm, xnorm, lambda, nfuncs, reason, im, F[im], X[im], iy, Y[iy], X[iy]);
    }
    if (reason) break;
  }
  if (!reason) reason = SNES_DIVERGED_MAX_IT;

done:
  if (snes_monitor()) fprintf(stderr, "snes done: reason %d its %d nfuncs %d fnorm0 %.6e fnorm %.6e ynorm %.6e lambda %.3e\n", reason, its, nfuncs, res->fnorm0, fnorm, ynorm, lambda);
  res->reason = reason; res->its = its; res->nfuncs = nfuncs;
  res->fnorm = fnorm; res->xnorm = xnorm; res->ynorm = ynorm; res->last_lambda = lambda;
  free(F); free(ja);
}
