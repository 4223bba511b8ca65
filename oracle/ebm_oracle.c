/*
 * ebm_oracle.c -- CPU oracle: literal restatement of the reference's time-stepping path.
 *
 * TEST INFRASTRUCTURE ONLY (see ebm_oracle.h).  PARITY UNPINNED (no Julia, fixture absent).
 *
 * Build: gcc -O2 -ffp-contract=off -fno-fast-math -fopenmp -fPIC -shared   (see Makefile)
 * Julia performs no FMA contraction and evaluates `@.` expressions left to right, so every
 * expression below is written in the reference's association order and the file must be
 * compiled with -ffp-contract=off.  Bool*Float in Julia is a "strong zero" (NaN*false == 0),
 * so every mask is a select, never a multiply.
 *
 * Citations (file:line) are relative to /root/reference.
 */
#include "ebm_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

static const double JL_PI = 3.141592653589793; /* Float64(pi) */

int ebm_oracle_max_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

/* ---------------------------------------------------------------- helpers */

/* Julia min/max propagate NaN (unlike C fmin/fmax) */
static inline double jl_min(double a, double b) {
  if (a != a || b != b) return NAN;
  if (a == b) return signbit(a) ? a : b;
  return a < b ? a : b;
}
/* Julia clamp(x, lo, hi) = ifelse(x > hi, hi, ifelse(x < lo, lo, x)); NaN passes through */
static inline double jl_clamp(double x, double lo, double hi) { return x > hi ? hi : (x < lo ? lo : x); }

/* Base.mapreduce_impl pairwise sum with blksize 1024 (what Statistics.mean uses) */
static double jl_pairwise_sum(const double* a, long stride, int ifirst, int ilast) {
  if (ifirst == ilast) return a[(long)ifirst * stride];
  if (ilast - ifirst < 1024) {
    double v = a[(long)ifirst * stride] + a[(long)(ifirst + 1) * stride];
    for (int i = ifirst + 2; i <= ilast; ++i) v = v + a[(long)i * stride];
    return v;
  }
  int imid = ifirst + ((ilast - ifirst) >> 1);
  double v1 = jl_pairwise_sum(a, stride, ifirst, imid);
  double v2 = jl_pairwise_sum(a, stride, imid + 1, ilast);
  return v1 + v2;
}

/* Forcing call, src/infrastructure.jl:294-307.  A constant forcing has all breakpoints 0 and
 * base == peak == cool, so the generic branch chain returns `cool` == base for every T >= 0. */
double ebm_oracle_forcing(const double* fr, double T) {
  if (T < fr[OF_d2]) return fr[OF_base];
  else if (T < fr[OF_d3]) return fr[OF_base] + fr[OF_rup] * (T - fr[OF_d2]);
  else if (T < fr[OF_d4]) return fr[OF_peak];
  else if (T < fr[OF_d5]) return fr[OF_peak] + fr[OF_rdown] * (T - fr[OF_d4]);
  else return fr[OF_cool];
}

/* hemispheric_mean, src/utilities.jl:397-403 */
double ebm_oracle_hemispheric_mean(const double* v, const double* x, int nx) {
  double acc = 0.0;
  for (int i = 0; i < nx - 1; ++i) acc += (v[i] + v[i + 1]) * (x[i + 1] - x[i]) / 2.0;
  return acc;
}

void ebm_oracle_diag(const double* T, const double* E, const double* phi, const double* x, int nx, double* out4) {
  double* ind = (double*)malloc(sizeof(double) * (size_t)nx);
  double edge = 1.0;
  int found = 0;
  for (int j = 0; j < nx; ++j) {
    int ice = phi ? (phi[j] > 0.0) : (E[j] < 0.0);
    ind[j] = phi ? phi[j] : (E[j] < 0.0 ? 1.0 : 0.0);
    if (ice && !found) { edge = x[j]; found = 1; }
  }
  out4[0] = ebm_oracle_hemispheric_mean(T, x, nx);
  out4[1] = ebm_oracle_hemispheric_mean(E, x, nx);
  out4[2] = 2.0 * JL_PI * ebm_oracle_hemispheric_mean(ind, x, nx);
  out4[3] = edge;
  free(ind);
}

/* time of global step tinx (1-based): SpaceTime.T = dt/2 : dt : dur - dt/2
 * (src/infrastructure.jl:130), a TwicePrecision range == correctly rounded (2*tinx-1)/(2*nt) */
static inline double global_time(long tinx, int nt) { return (double)(2 * tinx - 1) / (double)(2L * nt); }

static inline long jl_mod1(long a, long n) { long r = a % n; return r == 0 ? n : r; }

/* savesol!, src/infrastructure.jl:549-591.  cur[nvar][nx] = this step's stored variables. */
typedef struct {
  int nx, nt, dur, nvar, winter_inx, summer_inx, lastonly;
  double* annual;   /* [nt][nvar][nx] scratch year buffer (annusol.raw) */
  double* raw;      /* NULL or [nraw][nvar][nx] */
  double* seasonal; /* NULL or [dur][3][nvar][nx] */
} sampler_t;

static void sampler_store(sampler_t* s, const double* cur, long tinx) {
  const int nx = s->nx, nt = s->nt, nvar = s->nvar;
  const size_t fsz = (size_t)nvar * nx;
  int year = (int)ceil(global_time(tinx, nt));   /* :553 */
  int ti = (int)jl_mod1(tinx, nt);               /* :554 */
  memcpy(s->annual + (size_t)(ti - 1) * fsz, cur, fsz * sizeof(double)); /* :556-559 */
  if (s->raw) {
    if (!s->lastonly) memcpy(s->raw + (size_t)(tinx - 1) * fsz, cur, fsz * sizeof(double));    /* :561-565 */
    else if (tinx > (long)nt * s->dur - nt)
      memcpy(s->raw + (size_t)(ti - 1) * fsz, cur, fsz * sizeof(double));                       /* :566-570 */
  }
  if (!s->seasonal) return;
  double* yr = s->seasonal + (size_t)(year - 1) * 3 * fsz;
  if (ti == s->winter_inx) memcpy(yr + 0 * fsz, cur, fsz * sizeof(double));        /* :573-577 */
  else if (ti == s->summer_inx) memcpy(yr + 1 * fsz, cur, fsz * sizeof(double));   /* :578-582 */
  else if (ti == nt) {                                                             /* :583-588 */
    /* annual_mean -> crossmean (src/utilities.jl:390-395): per-cell Statistics.mean over the nt steps */
    for (int v = 0; v < nvar; ++v)
      for (int j = 0; j < nx; ++j)
        yr[2 * fsz + (size_t)v * nx + j] =
            jl_pairwise_sum(s->annual + (size_t)v * nx + j, (long)fsz, 0, nt - 1) / (double)nt;
  }
}

/* get_diffop, src/infrastructure.jl:480-489: lambda[j] couples cell j and j+1 (0-based j = 0..nx-2) */
static void diffop_lambda(int nx, double* lambda) {
  double dx = 1.0 / nx;
  for (int j = 1; j <= nx - 1; ++j) {
    double xb = (double)j / (double)nx; /* range dx:dx:1-dx (TwicePrecision) == correctly rounded j/nx */
    lambda[j - 1] = (1 - xb * xb) / (dx * dx);
  }
}

/* ---------------------------------------------------------------- classic */

typedef struct {
  int nx, nt;
  double dt, cg_tau, dt_tau, dc, M, kLf;
  double* aw;    /* [nx] */
  double* S;     /* [nt+1][nx] (column i of the reference's S is row i-1 here) */
  double* koff;  /* [nx-1] kappa super-diagonal: row j, column j+1 */
  double* ksub;  /* [nx-1] kappa sub-diagonal: row j+1, column j (== koff with get_diffop; differs with the generic stencil) */
  double* kdiag; /* [nx] */
} classic_statics;

/* get_statics, src/classic.jl:16-33 */
static void classic_statics_init(classic_statics* st, int nx, int nt, const double* x, const double* t,
                                 const double* p, int stencil) {
  st->nx = nx; st->nt = nt;
  st->dt = 1.0 / nt;                         /* infrastructure.jl:128 */
  st->cg_tau = p[OC_cg] / p[OC_tau];         /* :18 */
  st->dt_tau = st->dt / p[OC_tau];           /* :19 */
  st->dc = st->dt_tau * st->cg_tau;          /* :20 */
  st->M = p[OC_B] + st->cg_tau;              /* :27 */
  st->kLf = p[OC_k] * p[OC_Lf];              /* :29 */
  st->aw = (double*)malloc(sizeof(double) * nx);
  st->S = (double*)malloc(sizeof(double) * (size_t)nx * (nt + 1));
  st->koff = (double*)malloc(sizeof(double) * nx);
  st->ksub = (double*)malloc(sizeof(double) * nx);
  st->kdiag = (double*)malloc(sizeof(double) * nx);
  double* lambda = (double*)malloc(sizeof(double) * nx);
  diffop_lambda(nx, lambda);
  /* kappa = (1+dt_tau)*I - dt*D*diffop/cg  (:21), evaluated as ((dt*D)*diffop)/cg */
  double dtD = st->dt * p[OC_D];
  /* stencil == 0: get_diffop(nx) whatever the grid, as the reference does (classic.jl:21).
   * stencil == 1 (extension, SURVEY 8f-4: classic on non-uniform grids): the coefficients of the generic flux-form
   * stencil (infrastructure.jl:510-524) -- lower lo_j = m-_j / (dx_{j-1/2} (x_{j+1/2} - x_{j-1/2})), upper
   * hi_j = m+_j / (dx_{j+1/2} (x_{j+1/2} - x_{j-1/2})), zero flux at both ends; the matrix is no longer symmetric. */
  double* lo = (double*)calloc((size_t)nx, sizeof(double));
  double* hi = (double*)calloc((size_t)nx, sizeof(double));
  if (stencil) {
    double* xe = (double*)malloc(sizeof(double) * (nx + 2));
    xe[0] = -x[0];
    for (int j = 0; j < nx; ++j) xe[j + 1] = x[j];
    xe[nx + 1] = 2 - x[nx - 1];
    for (int j = 0; j < nx; ++j) {
      int i = j + 1;
      double xxph = (xe[i + 1] + xe[i]) / 2.0, xxmh = (xe[i] + xe[i - 1]) / 2.0;
      double phmmh = xxph - xxmh;
      if (j > 0) lo[j] = (1.0 - xxmh * xxmh) / ((xe[i] - xe[i - 1]) * phmmh);
      if (j < nx - 1) hi[j] = (1.0 - xxph * xxph) / ((xe[i + 1] - xe[i]) * phmmh);
    }
    free(xe);
  } else {
    for (int j = 0; j < nx; ++j) { lo[j] = j > 0 ? lambda[j - 1] : 0.0; hi[j] = j < nx - 1 ? lambda[j] : 0.0; }
  }
  for (int j = 0; j < nx; ++j) {
    double lm = lo[j], lp = hi[j];
    /* diffop diagonal = -l3 with l3 = -l1 - l2, l1[j] = -lm (0.0 at j=0), l2[j] = -lp (0.0 at the end) */
    double l1 = j > 0 ? -lm : 0.0, l2 = j < nx - 1 ? -lp : 0.0;
    double l3 = -l1 - l2;
    double dop_diag = -l3;
    st->kdiag[j] = (1 + st->dt_tau) - (dtD * dop_diag) / p[OC_cg];
    if (j < nx - 1) st->koff[j] = -((dtD * lp) / p[OC_cg]);
    if (j > 0) st->ksub[j - 1] = -((dtD * lm) / p[OC_cg]);
  }
  free(lo); free(hi);
  for (int j = 0; j < nx; ++j) st->aw[j] = p[OC_a0] - p[OC_a2] * (x[j] * x[j]);   /* :28 */
  /* S = (S0 - S2*x^2) - (S1*cos(2*pi*t)) * x   (:23-24), column nt+1 := column 1 (:25) */
  for (int i = 0; i < nt; ++i) {
    double sc = p[OC_S1] * cos(2.0 * JL_PI * t[i]);
    for (int j = 0; j < nx; ++j)
      st->S[(size_t)i * nx + j] = (p[OC_S0] - p[OC_S2] * (x[j] * x[j])) - sc * x[j];
  }
  memcpy(st->S + (size_t)nt * nx, st->S, sizeof(double) * nx);
  free(lambda);
}
static void classic_statics_free(classic_statics* st) { free(st->aw); free(st->S); free(st->koff); free(st->ksub); free(st->kdiag); }

/* tridiagonal solve in LU order (what dense LU without row swaps reduces to on a tridiagonal
 * matrix): l = a/w; w' = d - l*c; y' = r - l*y; back x = (y - c*x')/w.  sub[j] couples j and j-1. */
static void solve_tridiag(int n, const double* sub, const double* off, const double* diag, const double* rhs, double* xout,
                          double* w, double* y) {
  /* sub[j-1]: row j, column j-1; off[j]: row j, column j+1 */
  w[0] = diag[0]; y[0] = rhs[0];
  for (int j = 1; j < n; ++j) {
    double l = sub[j - 1] / w[j - 1];
    w[j] = diag[j] - l * off[j - 1];
    y[j] = rhs[j] - l * y[j - 1];
  }
  xout[n - 1] = y[n - 1] / w[n - 1];
  for (int j = n - 2; j >= 0; --j) xout[j] = (y[j] - off[j] * xout[j + 1]) / w[j];
}

/* dense LU with partial pivoting (what the reference's `\` on a dense Matrix does, classic.jl:55-63;
 * unblocked right-looking getrf + getrs).  A is n*n row-major, destroyed. */
static void solve_dense_lu(int n, double* A, double* b) {
  int* piv = (int*)malloc(sizeof(int) * n);
  for (int k = 0; k < n; ++k) {
    int p = k; double best = fabs(A[(size_t)k * n + k]);
    for (int i = k + 1; i < n; ++i) { double v = fabs(A[(size_t)i * n + k]); if (v > best) { best = v; p = i; } }
    piv[k] = p;
    if (p != k) {
      for (int c = 0; c < n; ++c) { double tmp = A[(size_t)k * n + c]; A[(size_t)k * n + c] = A[(size_t)p * n + c]; A[(size_t)p * n + c] = tmp; }
      double tb = b[k]; b[k] = b[p]; b[p] = tb;
    }
    double rp = 1.0 / A[(size_t)k * n + k];
    for (int i = k + 1; i < n; ++i) {
      double l = A[(size_t)i * n + k] * rp;
      A[(size_t)i * n + k] = l;
      if (l != 0.0) {
        for (int c = k + 1; c < n; ++c) A[(size_t)i * n + c] -= l * A[(size_t)k * n + c];
        b[i] -= l * b[k];
      }
    }
  }
  for (int i = n - 1; i >= 0; --i) {
    double s = b[i];
    for (int c = i + 1; c < n; ++c) s -= A[(size_t)i * n + c] * b[c];
    b[i] = s / A[(size_t)i * n + i];
  }
  free(piv);
}

/* step!(::Val{:Classic}), src/classic.jl:37-71.  i1 = 1-based index of t in the year (:45).
 * E, Tg updated in place; T, h written. */
/* debug menu (ODBG_*): the locals of step! a `debug` expression of the reference usually names (classic.jl:67-69 evaluates
 * an arbitrary expression in this scope; the menu offers the per-cell ones): dbg = NULL or [nx] */
static void classic_step_dbg(const classic_statics* st, const double* p, int i1, double f,
                             double* E, double* Tg, double* T, double* h, int solver, double* work, int which, double* dbg);
static void classic_step(const classic_statics* st, const double* p, int i1, double f,
                         double* E, double* Tg, double* T, double* h, int solver, double* work) {
  classic_step_dbg(st, p, i1, f, E, Tg, T, h, solver, work, 0, NULL);
}
static void classic_step_dbg(const classic_statics* st, const double* p, int i1, double f,
                             double* E, double* Tg, double* T, double* h, int solver, double* work, int which, double* dbg) {
  const int nx = st->nx;
  const double* Si = st->S + (size_t)(i1 - 1) * nx;
  const double* Sn = st->S + (size_t)i1 * nx; /* column i+1 */
  double* diag = work; double* rhs = work + nx; double* w = work + 2 * nx; double* y = work + 3 * nx;
  for (int j = 0; j < nx; ++j) {
    double Ej = E[j];
    double alpha = Ej > 0.0 ? st->aw[j] : (Ej < 0.0 ? p[OC_ai] : 0.0);                       /* :47 */
    double C = alpha * Si[j] + st->cg_tau * Tg[j] - p[OC_A] + f;                              /* :48 */
    double T0 = C / (st->M - st->kLf / Ej);                                                   /* :50 */
    double Tj = (Ej >= 0.0 ? Ej / p[OC_cw] : 0.0) + ((Ej < 0.0 && T0 < 0.0) ? T0 : 0.0);      /* :51 */
    T[j] = Tj;
    Ej = Ej + st->dt * (C - st->M * Tj + p[OC_Fb]);                                           /* :53 */
    E[j] = Ej;
    int m = (T0 < 0.0) && (Ej < 0.0);          /* T0 from the OLD E, E already updated (:56,61) */
    if (dbg) dbg[j] = which == ODBG_ALPHA ? alpha : which == ODBG_C ? C : which == ODBG_T0 ? T0 : which == ODBG_S ? Si[j]
                    : which == ODBG_MASK ? (double)m : NAN;
    double g = st->M - st->kLf / Ej;
    diag[j] = st->kdiag[j] - (m ? st->dc / g : 0.0);                                          /* :56 */
    rhs[j] = Tg[j] + (st->dt_tau * ((Ej >= 0.0 ? Ej / p[OC_cw] : 0.0) +
                                    (m ? (p[OC_ai] * Sn[j] - p[OC_A] + f) / g : 0.0)));       /* :58-62 */
    h[j] = Ej < 0.0 ? -Ej / p[OC_Lf] : 0.0;                                                   /* :65 */
  }
  if (solver == OSOLVE_DENSE_LU) {
    double* A = (double*)calloc((size_t)nx * nx, sizeof(double));
    for (int j = 0; j < nx; ++j) {
      A[(size_t)j * nx + j] = diag[j];
      if (j < nx - 1) { A[(size_t)j * nx + j + 1] = st->koff[j]; A[(size_t)(j + 1) * nx + j] = st->ksub[j]; }
    }
    solve_dense_lu(nx, A, rhs);
    memcpy(Tg, rhs, sizeof(double) * nx);
    free(A);
  } else {
    solve_tridiag(nx, st->ksub, st->koff, diag, rhs, Tg, w, y);
  }
}

int ebm_oracle_classic_step(int nx, int nt, const double* x, const double* t, const double* par15, int i1, double f,
                            double* E, double* Tg, double* T, double* h, int which, double* dbg) {
  if (nx < 2 || nt < 1 || i1 < 1 || i1 > nt) return -1;
  classic_statics st;
  classic_statics_init(&st, nx, nt, x, t, par15, 0);
  double* work = (double*)malloc(sizeof(double) * 4 * (size_t)nx);
  classic_step_dbg(&st, par15, i1, f, E, Tg, T, h, OSOLVE_TRIDIAG, work, which, dbg);
  free(work);
  classic_statics_free(&st);
  return 0;
}

int ebm_oracle_classic_run(int nx, int nt, int dur, const double* x, const double* t,
                           int winter_inx, int summer_inx, int nmem,
                           const double* par, const double* forc,
                           double* E, double* Tg, int solver, int lastonly,
                           double* raw, double* seasonal, int nthreads) {
  if (nx < 2 || nt < 1 || dur < 1 || nmem < 0) return -1;
  const int stencil = (solver >> 8) & 1;   /* bit 8 of `solver`: generic flux-form stencil in kappa (extension) */
  solver &= 0xff;
  const size_t fsz = (size_t)OCV_NVAR * nx;
  const size_t nraw = lastonly ? (size_t)nt : (size_t)nt * dur;
  const double dt = 1.0 / nt;
#ifdef _OPENMP
  if (nthreads > 0) omp_set_num_threads(nthreads);
#else
  (void)nthreads;
#endif
#pragma omp parallel for schedule(dynamic, 1)
  for (int m = 0; m < nmem; ++m) {
    const double* p = par + (size_t)m * OC_NPAR;
    const double* fr = forc + (size_t)m * OF_NF;
    classic_statics st;
    classic_statics_init(&st, nx, nt, x, t, p, stencil);
    double* cur = (double*)malloc(sizeof(double) * fsz);
    double* work = (double*)malloc(sizeof(double) * 4 * nx);
    sampler_t s = {nx, nt, dur, OCV_NVAR, winter_inx, summer_inx, lastonly, NULL, NULL, NULL};
    int sampling = (raw != NULL) || (seasonal != NULL);
    if (sampling) s.annual = (double*)malloc(sizeof(double) * fsz * nt);
    if (raw) s.raw = raw + (size_t)m * nraw * fsz;
    if (seasonal) {
      s.seasonal = seasonal + (size_t)m * dur * 3 * fsz;
      for (size_t q = 0; q < (size_t)dur * 3 * fsz; ++q) s.seasonal[q] = NAN;
    }
    double* Em = E + (size_t)m * nx; double* Tgm = Tg + (size_t)m * nx;
    for (long tinx = 1; tinx <= (long)nt * dur; ++tinx) {            /* infrastructure.jl:630 */
      int ti = (int)jl_mod1(tinx, nt);
      double tt = t[ti - 1];
      /* classic.jl:45: i = round(Int, mod1((t + dt/2)*nt, nt)) */
      double v = (tt + dt / 2.0) * nt;
      double r = fmod(v, (double)nt); if (r == 0.0) r = nt;
      int i1 = (int)nearbyint(r);
      double f = ebm_oracle_forcing(fr, global_time(tinx, nt));      /* infrastructure.jl:631 */
      classic_step(&st, p, i1, f, Em, Tgm, cur + (size_t)OCV_T * nx, cur + (size_t)OCV_h * nx, solver, work);
      if (sampling) {
        memcpy(cur + (size_t)OCV_E * nx, Em, sizeof(double) * nx);
        sampler_store(&s, cur, tinx);                                /* infrastructure.jl:632 */
      }
    }
    free(cur); free(work); free(s.annual);
    classic_statics_free(&st);
  }
  return 0;
}

/* ---------------------------------------------------------------- MIZ */

typedef struct {
  int nx, kind;
  /* generic stencil caches, src/infrastructure.jl:509-519 */
  double *diffx /* [nx+1] */, *mxxph, *mxxmh, *phmmh; /* [nx] */
  /* identity grid: D * get_diffop (src/infrastructure.jl:497): lower/diag/upper of (D*diffop) */
  double *ml, *md, *mu;
} miz_diff;

static void miz_diff_init(miz_diff* d, int nx, int kind, const double* x, double D) {
  d->nx = nx; d->kind = kind;
  d->diffx = (double*)calloc(nx + 1, sizeof(double));
  d->mxxph = (double*)calloc(nx, sizeof(double)); d->mxxmh = (double*)calloc(nx, sizeof(double));
  d->phmmh = (double*)calloc(nx, sizeof(double));
  d->ml = (double*)calloc(nx, sizeof(double)); d->md = (double*)calloc(nx, sizeof(double)); d->mu = (double*)calloc(nx, sizeof(double));
  if (kind == 1) {
    double* xe = (double*)malloc(sizeof(double) * (nx + 2));
    xe[0] = -x[0]; memcpy(xe + 1, x, sizeof(double) * nx); xe[nx + 1] = 2 - x[nx - 1];   /* :510 */
    for (int q = 0; q < nx + 1; ++q) d->diffx[q] = xe[q + 1] - xe[q];                     /* :511 */
    for (int j = 0; j < nx; ++j) {
      int i = j + 1; /* 0-based index into xe of cell j */
      double xxph = (xe[i + 1] + xe[i]) / 2.0;   /* :514 */
      double xxmh = (xe[i] + xe[i - 1]) / 2.0;   /* :515 */
      d->mxxph[j] = 1.0 - xxph * xxph;           /* :516 */
      d->mxxmh[j] = 1.0 - xxmh * xxmh;           /* :517 */
      d->phmmh[j] = xxph - xxmh;                 /* :518 */
    }
    free(xe);
  } else {
    double* lambda = (double*)malloc(sizeof(double) * nx);
    diffop_lambda(nx, lambda);
    for (int j = 0; j < nx; ++j) {
      double lm = j > 0 ? lambda[j - 1] : 0.0, lp = j < nx - 1 ? lambda[j] : 0.0;
      double l1 = j > 0 ? -lm : 0.0, l2 = j < nx - 1 ? -lp : 0.0;
      double l3 = -l1 - l2;
      d->ml[j] = D * lm; d->md[j] = D * (-l3); d->mu[j] = D * lp;   /* par.D * get_diffop(nx) */
    }
    free(lambda);
  }
}
static void miz_diff_free(miz_diff* d) {
  free(d->diffx); free(d->mxxph); free(d->mxxmh); free(d->phmmh); free(d->ml); free(d->md); free(d->mu);
}

/* out[j] = diffusion term added to a zero (or given) base: returns the increment only.
 * generic: :521-524; identity: sparse mat-vec in CSC column order (:497). */
static void miz_diffusion(const miz_diff* d, double D, const double* temp, double* out) {
  const int nx = d->nx;
  if (d->kind == 1) {
    for (int j = 0; j < nx; ++j) {
      double dTp = j < nx - 1 ? temp[j + 1] - temp[j] : 0.0; /* diffT[i]   (:522-523) */
      double dTm = j > 0 ? temp[j] - temp[j - 1] : 0.0;      /* diffT[i-1] */
      out[j] = D * (d->mxxph[j] * dTp / d->diffx[j + 1] - d->mxxmh[j] * dTm / d->diffx[j]) / d->phmmh[j]; /* :524 */
    }
  } else {
    for (int j = 0; j < nx; ++j) {
      double acc = 0.0;
      if (j > 0) acc += d->ml[j] * temp[j - 1];
      acc += d->md[j] * temp[j];
      if (j < nx - 1) acc += d->mu[j] * temp[j + 1];
      out[j] = acc;
    }
  }
}

/* tridiagonal coefficients of the linear operator L (increment = lo*T[j-1] + di*T[j] + up*T[j+1]),
 * used only for the semi-smooth Newton Jacobian of the closure */
static void miz_diff_coeffs(const miz_diff* d, double D, double* lo, double* di, double* up) {
  const int nx = d->nx;
  for (int j = 0; j < nx; ++j) {
    if (d->kind == 1) {
      double cu = j < nx - 1 ? D * d->mxxph[j] / d->diffx[j + 1] / d->phmmh[j] : 0.0;
      double cl = j > 0 ? D * d->mxxmh[j] / d->diffx[j] / d->phmmh[j] : 0.0;
      lo[j] = cl; up[j] = cu; di[j] = -(cl + cu);
    } else {
      lo[j] = j > 0 ? d->ml[j] : 0.0; up[j] = j < nx - 1 ? d->mu[j] : 0.0; di[j] = d->md[j];
    }
  }
}

typedef struct {
  int nx; double dt;
  const double* x; const double* p; const miz_diff* df;
  double pi_cos;     /* cos(2*pi*t) for the current step */
  double *lo, *di, *up;       /* Jacobian stencil */
  double *buf;                /* scratch, 12*nx */
} miz_ctx;

/* S0 - S1*x*cos(2*pi*t) - S2*x^2 as written in solar!, src/miz.jl:8-13 */
static inline double miz_insol(const double* p, double x, double c) {
  return p[OM_S0] - p[OM_S1] * x * c - p[OM_S2] * (x * x);
}

/* T0eq residual, src/miz.jl:33-45 */
static void miz_T0eq(const miz_ctx* c, const double* T0, const double* hp, const double* Tw, const double* phi,
                     double f, double* res, double* tb, double* dif) {
  const double* p = c->p; const int nx = c->nx;
  for (int j = 0; j < nx; ++j) {
    double ti = jl_min(T0[j], p[OM_Tm]);                /* ice_temp, :31 */
    tb[j] = ti * phi[j] + (1 - phi[j]) * Tw[j];         /* Tbar!, :21-25 */
  }
  miz_diffusion(c->df, p[OM_D], tb, dif);
  for (int j = 0; j < nx; ++j) {
    double v = p[OM_k] * (p[OM_Tm] - T0[j]) / hp[j];                    /* :39 SCM */
    v = v + p[OM_ai] * miz_insol(p, c->x[j], c->pi_cos);                /* :40 solar on ice */
    v = v + ((-p[OM_A]) - p[OM_B] * (T0[j] - p[OM_Tm]));                /* :41 OLR */
    v = v + dif[j];                                                     /* :42 diffusion */
    v = v + f;                                                          /* :43 forcing */
    res[j] = v;
  }
}

/* solveTi, src/miz.jl:47-68.  The reference calls NonlinearSolve.TrustRegion (abstol=1e-8, reltol=1e-6;
 * un-vendored dependency, compat "4.12.0").  The residual is piecewise linear with a tridiagonal
 * generalised Jacobian, so a semi-smooth Newton iteration stopped at max|res| <= tol converges to the
 * same (unique) root the reference's solver is asked for.  Returns iterations; *fail set if not converged. */
static int miz_solveTi(miz_ctx* c, const double* h, const double* Tw, const double* phi, double f,
                       double* T0, double* Ti, double tol, int* fail) {
  const double* p = c->p; const int nx = c->nx;
  double* hp = c->buf; double* res = hp + nx; double* tb = res + nx; double* dif = tb + nx;
  double* jd = dif + nx; double* jl = jd + nx; double* ju = jl + nx; double* w = ju + nx; double* y = w + nx;
  for (int j = 0; j < nx; ++j) hp[j] = (h[j] == 0.0) ? p[OM_hmin] : h[j];   /* :51 */
  int it = 0; *fail = 0;
  for (;;) {
    miz_T0eq(c, T0, hp, Tw, phi, f, res, tb, dif);
    double rmax = 0.0; int bad = 0;
    for (int j = 0; j < nx; ++j) { double a = fabs(res[j]); if (!(a <= rmax)) { if (a != a) bad = 1; else rmax = a; } }
    if (!bad && rmax <= tol) break;
    if (bad || it >= 100) { *fail = 1; break; }
    /* J = -diag(k/hp + B) + L*diag(phi*[T0<Tm]) */
    for (int j = 0; j < nx; ++j) {
      double gj = (T0[j] < p[OM_Tm]) ? phi[j] : 0.0;
      double gm = (j > 0 && T0[j - 1] < p[OM_Tm]) ? phi[j - 1] : 0.0;
      double gp = (j < nx - 1 && T0[j + 1] < p[OM_Tm]) ? phi[j + 1] : 0.0;
      jd[j] = -(p[OM_k] / hp[j] + p[OM_B]) + c->di[j] * gj;
      jl[j] = c->lo[j] * gm; ju[j] = c->up[j] * gp;
    }
    /* solve J*delta = -res (Thomas, general tridiagonal) and update */
    w[0] = jd[0]; y[0] = -res[0];
    for (int j = 1; j < nx; ++j) {
      double l = jl[j] / w[j - 1];
      w[j] = jd[j] - l * ju[j - 1];
      y[j] = -res[j] - l * y[j - 1];
    }
    double dl = y[nx - 1] / w[nx - 1];
    T0[nx - 1] += dl;
    for (int j = nx - 2; j >= 0; --j) { dl = (y[j] - ju[j] * dl) / w[j]; T0[j] += dl; }
    ++it;
  }
  for (int j = 0; j < nx; ++j) {
    double ti = jl_min(T0[j], p[OM_Tm]);     /* :65 */
    Ti[j] = (h[j] == 0.0) ? 0.0 : ti;        /* :66 zeroref!(Ti, h) */
  }
  return it;
}

/* step!(::Val{:MIZ}), src/miz.jl:150-196.  State Ei,Ew,h,D,phi updated in place; stored-only outputs
 * Tw,Ti,n,E,T written with the reference's NaN masks applied (:193-194). */
static int miz_step(miz_ctx* c, double f, double* Ei, double* Ew, double* h, double* D, double* phi, double* T0,
                    double* oTw, double* oTi, double* on, double* oE, double* oT, double tol, int* fail) {
  const double* p = c->p; const int nx = c->nx; const double dt = c->dt;
  double* Tw = oTw; double* Ti = oTi; double* n = on;
  double* tb = c->buf + 9 * c->nx; double* dif = tb + nx; /* miz_solveTi uses buf[0..9nx) */
  for (int j = 0; j < nx; ++j) {
    double v = p[OM_Tm] + Ew[j] / ((1 - phi[j]) * p[OM_cw]);   /* water_temp :30, :156 */
    Tw[j] = (v != v) ? 0.0 : v;                                 /* :157 */
  }
  int iters = miz_solveTi(c, h, Tw, phi, f, T0, Ti, tol, fail); /* :158 */
  for (int j = 0; j < nx; ++j) {
    double v = phi[j] / (p[OM_alpha] * (D[j] * D[j]));          /* num :84 */
    n[j] = (D[j] == 0.0) ? 0.0 : v;                             /* :85 */
    tb[j] = Ti[j] * phi[j] + (1 - phi[j]) * Tw[j];              /* Tbar(Ti,Tw,phi) */
  }
  miz_diffusion(c->df, p[OM_D], tb, dif);
  const double Tm_m2 = pow(p[OM_Tm], p[OM_m2]);                 /* wlat :71 -- `Tm^m2` binds to Tm (sic) */
  const double denom_dn = p[OM_Lf] * p[OM_alpha] * (p[OM_Dmin] * p[OM_Dmin]) * p[OM_hmin]; /* psinplus :127 */
  for (int j = 0; j < nx; ++j) {
    double xj = c->x[j];
    double ins = miz_insol(p, xj, c->pi_cos);
    double Lolr = p[OM_A] + p[OM_B] * (tb[j] - p[OM_Tm]);                                   /* :99 */
    double sol_i = 0.0 + p[OM_ai] * ins;                                                    /* solar(ice) :11,16-18 */
    double sol_w = 0.0 + (p[OM_a0] - p[OM_a2] * (xj * xj)) * ins;                           /* solar(water) :14 */
    double difz = 0.0 + dif[j];
    double Fvi = sol_i - Lolr + difz + p[OM_Fb] + f;                                        /* :100 */
    double Fvw = sol_w - Lolr + difz + p[OM_Fb] + f;
    double wl = p[OM_m1] * (Tw[j] - Tm_m2);                                                 /* :71 */
    double Flat = phi[j] * h[j] * p[OM_Lf] * wl * JL_PI / (p[OM_alpha] * D[j]);             /* :104 */
    if (D[j] == 0.0) Flat = 0.0;                                                            /* :105 */
    double rEi = Ei[j] + (phi[j] * Fvi + Flat) * dt;                                        /* :137,148,166 */
    double rEw = Ew[j] + ((1 - phi[j]) * Fvw - Flat) * dt;                                  /* :138,148,167 */
    /* redistributeE :109-117 */
    double cEi = jl_clamp(rEi, -INFINITY, 0.0), cEw = jl_clamp(rEw, 0.0, INFINITY);
    double psiEidt = rEi - cEi, psiEwdt = rEw - cEw;
    double Ei_n = cEi + psiEwdt, Ew_n = cEw + psiEidt;
    /* area_lead :90-93 (uses n from the start of the step) */
    double d2rl = D[j] + 2.0 * p[OM_rl];
    double ring = p[OM_alpha] * n[j] * (d2rl * d2rl - D[j] * D[j]);
    double Al = jl_min(ring, 1.0 - phi[j]);
    /* split_psiEw :120-125 on psiEwdt/dt (:173) */
    double psiEw = psiEwdt / dt;
    double Ql = Al / (1 - phi[j]) * psiEw;
    if (phi[j] == 1.0) Ql = 0.0;
    double Qp = psiEw - Ql;
    double dn = dt * (-Qp / denom_dn);                                                      /* :127,174 */
    /* D_t :140-146 */
    double lat_melt = -JL_PI / 2.0 * p[OM_alpha] * wl;                                      /* :141 (sic) */
    double lat_grow = -D[j] / (2 * p[OM_Lf] * h[j] * phi[j]) * Ql;                          /* :142 */
    double D3 = D[j] * D[j] * D[j];
    double weld = p[OM_kappa] * p[OM_alpha] / 4 * phi[j] * D3;                              /* :143 */
    if (h[j] == 0.0) lat_grow = 0.0;                                                        /* :144 */
    double Dt = lat_melt + lat_grow + weld;                                                 /* :145 */
    double rD = D[j] + Dt * dt;                                                             /* :175 */
    /* average :129-134 */
    double total = n[j] + dn;
    double Dn = (n[j] * rD + dn * p[OM_Dmin]) / total;
    if (total == 0.0) Dn = 0.0;
    Dn = jl_clamp(Dn, p[OM_Dmin], p[OM_Dmax]);                                              /* :177 */
    if (Ei_n == 0.0) Dn = 0.0;                                                              /* :178 */
    double rh = h[j] + (-1 / p[OM_Lf] * Fvi) * dt;                                          /* :139,179 */
    rh = jl_clamp(rh, 0.0, INFINITY);                                                       /* :180 */
    double hn = (n[j] * rh + dn * p[OM_hmin]) / total;                                      /* :181 */
    if (total == 0.0) hn = 0.0;
    /* concentration :74-80 */
    double ph = -Ei_n / (p[OM_Lf] * hn);
    if (hn == 0.0) ph = 0.0;
    if (ph > 1.0) ph = 1.0;
    if (hn == 0.0) Ei_n = 0.0;                                                              /* :185 */
    oE[j] = ph * Ei_n + (1 - ph) * Ew_n;                                                    /* :186 */
    oT[j] = Ti[j] * ph + (1 - ph) * Tw[j];                                                  /* :187 Tbar(Ti,Tw,phi_new) */
    Ei[j] = Ei_n; Ew[j] = Ew_n; D[j] = Dn; h[j] = hn; phi[j] = ph;
  }
  for (int j = 0; j < nx; ++j) {            /* stored-only masks :193-194 */
    if (Ei[j] == 0.0) Ti[j] = NAN;
    if (phi[j] > 0.99) Tw[j] = NAN;
  }
  return iters;
}

int ebm_oracle_miz_run(int nx, int nt, int dur, const double* x, const double* t,
                       int winter_inx, int summer_inx, int grid_kind, int nmem,
                       const double* par, const double* forc,
                       double* Ei, double* Ew, double* h, double* D, double* phi, double* T0,
                       double newton_tol, int lastonly,
                       double* raw, double* seasonal,
                       long long* newton_iters, long long* nonconv, int nthreads) {
  if (nx < 2 || nt < 1 || dur < 1 || nmem < 0) return -1;
  const size_t fsz = (size_t)OMV_NVAR * nx;
  const size_t nraw = lastonly ? (size_t)nt : (size_t)nt * dur;
#ifdef _OPENMP
  if (nthreads > 0) omp_set_num_threads(nthreads);
#else
  (void)nthreads;
#endif
#pragma omp parallel for schedule(dynamic, 1)
  for (int m = 0; m < nmem; ++m) {
    const double* p = par + (size_t)m * OM_NPAR;
    const double* fr = forc + (size_t)m * OF_NF;
    miz_diff df; miz_diff_init(&df, nx, grid_kind, x, p[OM_D]);
    miz_ctx c; c.nx = nx; c.dt = 1.0 / nt; c.x = x; c.p = p; c.df = &df;
    c.lo = (double*)malloc(sizeof(double) * nx); c.di = (double*)malloc(sizeof(double) * nx); c.up = (double*)malloc(sizeof(double) * nx);
    c.buf = (double*)malloc(sizeof(double) * 12 * nx);
    miz_diff_coeffs(&df, p[OM_D], c.lo, c.di, c.up);
    double* cur = (double*)malloc(sizeof(double) * fsz);
    sampler_t s = {nx, nt, dur, OMV_NVAR, winter_inx, summer_inx, lastonly, NULL, NULL, NULL};
    int sampling = (raw != NULL) || (seasonal != NULL);
    if (sampling) s.annual = (double*)malloc(sizeof(double) * fsz * nt);
    if (raw) s.raw = raw + (size_t)m * nraw * fsz;
    if (seasonal) {
      s.seasonal = seasonal + (size_t)m * dur * 3 * fsz;
      for (size_t q = 0; q < (size_t)dur * 3 * fsz; ++q) s.seasonal[q] = NAN;
    }
    double *Eim = Ei + (size_t)m * nx, *Ewm = Ew + (size_t)m * nx, *hm = h + (size_t)m * nx,
           *Dm = D + (size_t)m * nx, *phim = phi + (size_t)m * nx, *T0m = T0 + (size_t)m * nx;
    long long iters = 0, fails = 0;
    for (long tinx = 1; tinx <= (long)nt * dur; ++tinx) {
      int ti = (int)jl_mod1(tinx, nt);
      c.pi_cos = cos(2.0 * JL_PI * t[ti - 1]);
      double f = ebm_oracle_forcing(fr, global_time(tinx, nt));
      int fail = 0;
      iters += miz_step(&c, f, Eim, Ewm, hm, Dm, phim, T0m,
                        cur + (size_t)OMV_Tw * nx, cur + (size_t)OMV_Ti * nx, cur + (size_t)OMV_n * nx,
                        cur + (size_t)OMV_E * nx, cur + (size_t)OMV_T * nx, newton_tol, &fail);
      fails += fail;
      if (sampling) {
        memcpy(cur + (size_t)OMV_Ei * nx, Eim, sizeof(double) * nx);
        memcpy(cur + (size_t)OMV_Ew * nx, Ewm, sizeof(double) * nx);
        memcpy(cur + (size_t)OMV_h * nx, hm, sizeof(double) * nx);
        memcpy(cur + (size_t)OMV_D * nx, Dm, sizeof(double) * nx);
        memcpy(cur + (size_t)OMV_phi * nx, phim, sizeof(double) * nx);
        sampler_store(&s, cur, tinx);
      }
    }
    if (newton_iters) newton_iters[m] = iters;
    if (nonconv) nonconv[m] = fails;
    free(cur); free(s.annual); free(c.lo); free(c.di); free(c.up); free(c.buf);
    miz_diff_free(&df);
  }
  return 0;
}
