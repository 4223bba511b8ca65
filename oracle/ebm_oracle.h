/*
 * ebm_oracle.h -- CPU oracle for the EnergyBalanceModel.jl time-stepping path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product: only
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load this library, and only as the checker / CPU baseline.
 *
 * PARITY UNPINNED: the reference is Julia; Julia is not installed here and the
 * reference's single golden fixture (test/solution_1year.jld2) is absent from
 * the mount (.MISSING_LARGE_BLOBS).  This file is a literal restatement of the
 * reference's arithmetic (citations are relative to /root/reference), pinned
 * only by the docstring known-answer values (grid, forcing, parameters) and
 * cross-checked by an independent NumPy restatement under tests/.
 */
#ifndef EBM_ORACLE_H
#define EBM_ORACLE_H

#ifdef __cplusplus
extern "C" {
#endif

/* parameter vectors are plain double arrays, one row per member */
enum { /* classic: src/infrastructure.jl:442-444 (par.F is never read by step!) */
  OC_D, OC_A, OC_B, OC_cw, OC_S0, OC_S1, OC_S2, OC_a0, OC_a2, OC_ai, OC_Fb, OC_k, OC_Lf, OC_cg, OC_tau,
  OC_NPAR
};
enum { /* MIZ: src/infrastructure.jl:436-441 */
  OM_D, OM_A, OM_B, OM_cw, OM_S0, OM_S1, OM_S2, OM_a0, OM_a2, OM_ai, OM_Fb, OM_k, OM_Lf, OM_Tm, OM_m1, OM_m2,
  OM_alpha, OM_rl, OM_Dmin, OM_Dmax, OM_hmin, OM_kappa,
  OM_NPAR
};
/* forcing row: base, peak, cool, rate_up, rate_down, then 5 breakpoints (years, stored as doubles) */
enum { OF_base, OF_peak, OF_cool, OF_rup, OF_rdown, OF_d1, OF_d2, OF_d3, OF_d4, OF_d5, OF_NF };

/* classic stored variables (src/infrastructure.jl:621): E, T, h  */
enum { OCV_E, OCV_T, OCV_h, OCV_NVAR };
/* MIZ stored variables (src/infrastructure.jl:621-624) in the order of EnergyBalanceModel.jl:63 */
enum { OMV_T, OMV_Ei, OMV_Ti, OMV_D, OMV_n, OMV_h, OMV_phi, OMV_E, OMV_Ew, OMV_Tw, OMV_NVAR };

/* solver selection for the classic implicit step */
enum { OSOLVE_TRIDIAG = 0, OSOLVE_DENSE_LU = 1 };

double ebm_oracle_forcing(const double* frow, double T);

/*
 * Classic ensemble integration (integrate + step!(:Classic) + savesol!).
 *   x[nx], t[nt]            : grid and in-year times, verbatim from SpaceTime
 *   winter_inx, summer_inx  : 1-based season indices (SpaceTime.winter.inx / summer.inx)
 *   par[nmem][OC_NPAR], forc[nmem][OF_NF]
 *   E[nmem][nx], Tg[nmem][nx] : initial state in, final state out
 *   raw      : NULL or [nmem][nraw][OCV_NVAR][nx], nraw = lastonly ? nt : nt*dur
 *   seasonal : NULL or [nmem][dur][3 (winter,summer,avg)][OCV_NVAR][nx]  (NaN where never stored)
 * returns 0 on success.
 */
/* one step!(Val(:Classic), ...) of one member (src/classic.jl:37-71) with the debug menu: dbg = NULL or [nx] receives the
 * step's local `which` (classic.jl:67-69 evaluates a user expression in that scope) */
enum { ODBG_NONE = 0, ODBG_ALPHA = 1, ODBG_C = 2, ODBG_T0 = 3, ODBG_S = 4, ODBG_MASK = 5 };
int ebm_oracle_classic_step(int nx, int nt, const double* x, const double* t, const double* par15, int i1, double f,
                            double* E, double* Tg, double* T, double* h, int which, double* dbg);

int ebm_oracle_classic_run(int nx, int nt, int dur, const double* x, const double* t,
                           int winter_inx, int summer_inx, int nmem,
                           const double* par, const double* forc,
                           double* E, double* Tg, int solver, int lastonly,
                           double* raw, double* seasonal, int nthreads);

/*
 * MIZ ensemble integration (integrate + step!(:MIZ) + savesol!).
 *   grid_kind : 0 = SpaceTime{identity} (sparse diffop mat-vec), 1 = generic flux-form stencil
 *   state arrays [nmem][nx]: Ei, Ew, h, D, phi in/out;  T0 = closure warm start in/out
 *   newton_tol : residual max-norm stop (reference: abstol = 1e-8, miz.jl:137)
 *   newton_iters[nmem] (optional) : total Newton iterations taken;  nonconv[nmem] (optional)
 */
int ebm_oracle_miz_run(int nx, int nt, int dur, const double* x, const double* t,
                       int winter_inx, int summer_inx, int grid_kind, int nmem,
                       const double* par, const double* forc,
                       double* Ei, double* Ew, double* h, double* D, double* phi, double* T0,
                       double newton_tol, int lastonly,
                       double* raw, double* seasonal,
                       long long* newton_iters, long long* nonconv, int nthreads);

/* hemispheric_mean (src/utilities.jl:397-403) */
double ebm_oracle_hemispheric_mean(const double* v, const double* x, int nx);

/*
 * L0 diagnostics from one stored field set: out[4] = {hemispheric mean T, hemispheric mean E,
 * ice area, ice-edge x}.  Ice area follows plot_seasonal (src/plot.jl:173-190):
 * 2*pi*hemispheric_mean(phi) when phi is given, else 2*pi*hemispheric_mean(E<0).
 * Ice edge = x of the lowest-latitude ice cell (E<0, or phi>0), 1.0 if there is none.
 */
void ebm_oracle_diag(const double* T, const double* E, const double* phi, const double* x, int nx, double* out4);

int ebm_oracle_max_threads(void);

#ifdef __cplusplus
}
#endif
#endif
