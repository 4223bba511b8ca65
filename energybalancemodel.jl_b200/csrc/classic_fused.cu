// classic_fused.cu -- round-2 classic (Wagner-Eisenman) EBM ensemble kernel for grids up to 208 cells: one CTA
// barrier per time step.
//
// Replaces, for a whole ensemble and many years per launch, the reference's integrate loop
// (src/infrastructure.jl:630-634), step!(::Val{:Classic}) (src/classic.jl:43-65; arithmetic spec SURVEY.md
// Appendix A) and savesol! / annual_mean (src/infrastructure.jl:536-591).
//
// Mapping (measured B200 facts, profiles/r2_microbench.txt: DFMA latency 8.8 cycles, a sub-partition issues one warp
// DFMA per ~2.2 cycles whatever the number of active lanes, MUFU.RCP64H seed is good to 2^-20 so ONE cubic step gives
// 1 ulp, a 4-warp CTA barrier costs ~20 cycles when nobody is late):
//   * lane = member, warp = 2 adjacent latitude bands of K = 13 cells x 16 members, CTA = 16 members x WB bands.
//     E and Tg of the thread's cells stay in registers for the whole launch; what is touched once or twice per step
//     (carried reciprocals, annual sums, the band's pivot rows) lives in thread-private shared memory [row][thread].
//     Which band pair a warp owns rotates with the CTA index: in a partially ice-covered member only the polar
//     bands take the expensive path, and without the rotation all three resident CTAs put that warp on the same
//     SM sub-partition.
//   * threads whose cells are all open water (E > 0): folded physics (7 FP64 instructions per cell), band-local
//     elimination with precomputed pivots (tables built once per launch; parameter-uniform groups, TAB = true).
//   * threads with ice: physics, the masked diagonal (classic.jl:56) and the band's elimination are fused into one
//     pass over the cells: pivots in determinant form (one dependent DFMA per row), reciprocals independent.
//   * the upward pass reduces every row to x_i = al_i - be_i xL - ga_i xn (xL / xn: last unknown of the previous /
//     this band), so the back substitution after the interface solve has no dependent chain.
//   * interface system (WB unknowns per member): every warp solves it redundantly for its own 16 members, two lanes
//     per member sweeping from both ends, from the rows all bands posted to a double-buffered shared array -- ONE
//     __syncthreads per step and no warp waits for another warp's serial phase.
//   * sampling (savesol!): annual sums of E are updated every second step with E_old + E_new (nt even); steps that
//     store output take a second instantiation.
// TAB = false integrates the groups whose table-building parameters (D, cg, tau, S0, S2, a0, a2) differ between
// members: shared tables hold geometry only, every thread takes the fused path with its own coefficients.
#include <type_traits>

#include "ebm_internal.cuh"

namespace {

constexpr double kTwoPi = 6.283185307179586;
constexpr int MW = 16;   // members per CTA: a warp = 2 bands x 16 members

// reciprocal: MUFU.RCP64H seed (relative error <= 2^-20, measured) + one cubic step -> <= 1 ulp; operands are
// well scaled (pivots of a diagonally dominant matrix, M*E - kLf), no denormal / overflow slow path
__device__ __forceinline__ double rcp3(double w) {
  double x;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(x) : "d"(w));
  const double e = fma(-w, x, 1.0);
  const double t = fma(e, e, e);
  return fma(x, t, x);
}
// sign / zero tests on the integer pipe.  The state never holds -0.0: it is normalised when loaded, and an FMA
// result is -0.0 only if both addends are (classic.jl:47's masks are E>0 / E<0, so -0.0 would otherwise differ).
__device__ __forceinline__ bool is_neg(double v) { return __double2hiint(v) < 0; }
__device__ __forceinline__ bool is_zero(double v) { return ((__double2hiint(v) << 1) | __double2loint(v)) == 0; }
// strictly positive, normal, finite
__device__ __forceinline__ bool is_pos(double v) { return (unsigned)(__double2hiint(v) - 0x00100000) < 0x7fe00000u; }

// v with both words ANDed with m (m = 0 or -1): v or +0.0, on the integer pipe
__device__ __forceinline__ double and_mask(double v, int m) {
  return __hiloint2double(__double2hiint(v) & m, __double2loint(v) & m);
}

struct __align__(16) PhysA { double S0x, x; };    // S0 - S2 x^2, x           (TAB = false: x^2, x)
struct __align__(16) PhysB { double aw, wts; };   // a0 - a2 x^2, trapezoid w (TAB = false: x^2, w)
struct __align__(16) CoefT { double kjj, ac; };   // kappa_jj, aoff_j^2       (TAB = false: lam_lo + lam_hi, lam_lo^2)
struct __align__(16) ElimF { double iw, tq; };    // no-mask band elimination: 1/w_i, aoff_i/w_i
struct __align__(16) ElimB { double be, ga; };    // reduced rows x_i = al_i - be_i xL - ga_i xn; last row: (s, q)

template <int K, int WB>
constexpr size_t fused_smem_bytes(bool fields) {
  return (size_t)K * WB * (sizeof(PhysA) + sizeof(PhysB) + sizeof(CoefT) + sizeof(ElimF) + sizeof(ElimB) + 2 * sizeof(double)) +
         sizeof(double) * (16 + (size_t)2 * WB * 6 * MW + (size_t)(4 + (fields ? 2 : 0)) * K * WB * MW);
}

template <int K, int WB, bool TAB, bool DEPBS>
struct Fused {
  static constexpr int NXP = K * WB, NT = WB * MW, H = WB / 2;
  // shared memory
  const PhysA* pa; const PhysB* pb; const CoefT* cf; const double* aoff; const ElimF* ef; const ElimB* eb; const double* eq;
  double* iface;   // [2][WB][6][MW]
  double* rows;    // [4 (+2)][K][NT] thread-private: r, sumE, q, s (, sumT, sumH)
  // member constants
  double ai, cg_tau, M, kLf, inv_cw, dt, dt_tau, dttau_cw, Fb, A, c1, inv_nt, inv_Lf;
  double S0m, S2m, a0m, a2m, fac, fac2, one_dttau;   // TAB = false only
  long long thr_bits;                                  // bits of kLf / M (sign of M - kLf/E for E > 0)
  // identity
  int tid, h, mi, pair, j0, nv;   // nv: real cells of this band (K unless the band holds pad cells)
  bool active, sel, cta_fields, pairsum;
  long long mo, msel;
  // state
  double E[K], Tg[K], accT, accEw;
  int buf;
  bool gen;      // this step's band rows are in shared memory (fused path) rather than in the tables
  bool rvalid;   // row 0 holds 1/(M - kLf/E) of the current E (the open-water path does not maintain it)

  __device__ __forceinline__ double& row(int r, int i) { return rows[(r * K + i) * NT + tid]; }
  __device__ __forceinline__ PhysA physA(int j) const {
    PhysA p = pa[j];
    if constexpr (!TAB) p.S0x = fma(-S2m, p.S0x, S0m);
    return p;
  }
  __device__ __forceinline__ PhysB physB(int j) const {
    PhysB p = pb[j];
    if constexpr (!TAB) p.aw = fma(-a2m, p.aw, a0m);
    return p;
  }
  __device__ __forceinline__ CoefT coef(int j) const {
    CoefT c = cf[j];
    if constexpr (!TAB) { c.kjj = fma(fac, c.kjj, one_dttau); c.ac = fac2 * c.ac; }
    return c;
  }
  __device__ __forceinline__ double sub(int j) const {   // sub-diagonal of row j == super-diagonal of row j-1
    if constexpr (TAB) return aoff[j]; else return -fac * aoff[j];
  }

  // sampled output of cell i after its update: savesol! (infrastructure.jl:549-591).  SLOW instantiation only.
  __device__ __forceinline__ void sample(const ClassicKArgs& a, const int i, const double wj, const double xj, const double En,
                                         const double T, const double sEi, const int season, const int ti, const int year,
                                         double& dgT, double& dgE, double& dgA, double& dgX) {
    const int nx = a.nx, nt = a.nt;
    const int j = j0 + i;
    const double Eneg = is_neg(En) ? En : 0.0;
    if (cta_fields) { row(4, i) += T; row(5, i) += Eneg; }
    const bool rawstep = sel && a.raw != nullptr && (!a.lastonly || year == a.dur - 1);
    if (rawstep && j < nx) {
      const long long nraw = a.lastonly ? (long long)nt : (long long)nt * a.dur;
      const long long rawidx = a.lastonly ? (ti - 1) : ((long long)year * nt + ti - 1);
      double* o = a.raw + ((msel * nraw + rawidx) * 3) * (long long)nx + j;
      o[0] = En; o[nx] = T; o[2 * nx] = -Eneg * inv_Lf;            // h = -E/Lf*(E<0)  (classic.jl:65)
    }
    if (season >= 0) {
      double vT = T, vE = En, vN = Eneg;
      if (season == 2) {                                            // annual mean (infrastructure.jl:583-588)
        vE = sEi * inv_nt;
        if (cta_fields) { vT = row(4, i) * inv_nt; vN = row(5, i) * inv_nt; }
      }
      dgT = fma(wj, vT, dgT);
      dgE = fma(wj, vE, dgE);
      if (vE < 0.0 && j < nx) { dgA += wj; dgX = fmin(dgX, xj); }
      if (sel && a.seasonal != nullptr && j < nx) {
        double* o = a.seasonal + ((((msel * a.dur + year) * 3 + season) * 3) * (long long)nx) + j;
        o[0] = vE; o[nx] = vT; o[2 * nx] = -vN * inv_Lf;
      }
    }
    if (ti == nt && cta_fields) { row(4, i) = 0.0; row(5, i) = 0.0; }
  }

  // Pad cells (a band's cells beyond nx) follow the band: frozen when every real cell of the band is ice, warm when
  // every real cell is open water, untouched otherwise.  Called at launch and from the (rare) G path.
  __device__ __forceinline__ void set_pads() {
    if (nv < K && nv > 0) {
      int hand = -1, hor = 0;
#pragma unroll
      for (int i = 0; i < K; ++i) if (i < nv) { hand &= __double2hiint(E[i]); hor |= __double2hiint(E[i]); }
      const bool allice = hand < 0, allwater = hor >= 0;
#pragma unroll
      for (int i = 0; i < K; ++i) {
        if (i >= nv) {
          if (allice && !is_neg(E[i])) { E[i] = -10.0; Tg[i] = -10.0; rvalid = false; }
          if (allwater && is_neg(E[i])) { E[i] = 100.0; Tg[i] = 10.0; rvalid = false; }
        }
      }
    }
  }

  // ptxas keeps the instruction order of the source to a large extent (and a warp issues in order): code written
  // cell after cell runs the cells' dependent chains one after the other (measured: 3 issue cycles per FP64
  // instruction).  The hot paths below are therefore written statement-major over groups of GI cells -- every
  // statement for all cells of the group before the next statement -- which gives each warp GI independent chains.
  static constexpr int GI = 4;

  // band-local elimination of the rows phase 1 left behind (diagonal in row 2, right-hand side in Tg): pivots in
  // determinant form, P_i = w_0 ... w_i = d_i P_{i-1} - a_i^2 P_{i-2} (one dependent DFMA per row), so that the K
  // reciprocals 1/w_i = P_{i-1}/P_i are independent of each other; then x_i + q_i x_{i+1} + s_i xL = y_i
  __device__ __forceinline__ void eliminate(double& i_sl, double& i_ql, double& i_yl) {
    double P[K + 1];
    P[0] = 1.0;
    {
      double dg[K], ac[K];
#pragma unroll
      for (int i = 0; i < K; ++i) { dg[i] = row(2, i); ac[i] = (i == 0) ? 0.0 : coef(j0 + i).ac; }
      P[1] = dg[0];
#pragma unroll
      for (int i = 1; i < K; ++i) P[i + 1] = fma(dg[i], P[i], -(ac[i] * P[i - 1]));
    }
    // 1 / w_i = P_{i-1} / P_i, GI rows at a time
#pragma unroll
    for (int i0 = 0; i0 < K; i0 += GI) {
      double x[GI], e[GI];
#pragma unroll
      for (int g = 0; g < GI; ++g) if (i0 + g < K) asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(x[g]) : "d"(P[i0 + g + 1]));
#pragma unroll
      for (int g = 0; g < GI; ++g) if (i0 + g < K) e[g] = fma(-P[i0 + g + 1], x[g], 1.0);
#pragma unroll
      for (int g = 0; g < GI; ++g) if (i0 + g < K) e[g] = fma(e[g], e[g], e[g]);
#pragma unroll
      for (int g = 0; g < GI; ++g) if (i0 + g < K) x[g] = fma(x[g], e[g], x[g]);
#pragma unroll
      for (int g = 0; g < GI; ++g) if (i0 + g < K) P[i0 + g] = P[i0 + g] * x[g];      // P[i] now holds 1 / w_i
    }
    // forward substitution: y and the left spike s are two chains of one dependent operation per row
    double yprev = 0.0, sprev = 0.0;
    double aj = sub(j0);
#pragma unroll
    for (int i = 0; i < K; ++i) {
      const double cj = sub(j0 + i + 1);
      const double iw = P[i];
      const double tq = aj * iw;
      const double ri = Tg[i] * iw;
      const double y = (i == 0) ? ri : fma(-tq, yprev, ri);
      const double s = (i == 0) ? tq : -tq * sprev;
      const double q = cj * iw;
      row(2, i) = q; row(3, i) = s; Tg[i] = y;
      yprev = y; sprev = s; aj = cj;
      if (i == K - 1) { i_sl = s; i_ql = q; i_yl = y; }
    }
  }

  // physics of GI cells and their rows of the implicit system (classic.jl:47-62), statement-major.  MIXED = false:
  // every cell of the band is ice (alpha = ai); the remaining masks -- frozen surface (C < 0, hence T0 < 0) and
  // E' < 0 -- are bit masks on operands, exact in every case:
  //   T = T0 [C<0];  mk = (T0<0) & (E'<0);  um = mk ? dt_tau/(M - kLf/E') : 0
  //   diag = kappa_jj - cg_tau um;  rhs = Tg + dt_tau/cw E' [E'>=0] + um (ai S' - A + f)
  // MIXED = true: cells of either sign (ice edge inside the band, E == 0): alpha, T and E' are selected per cell
  // between the ice expressions above and the open-water ones of the W path; for an all-ice band the results are
  // bit-identical to MIXED = false.
  template <int I0, bool MIXED, bool SLOW, bool SUMNOW>
  __device__ __forceinline__ void cell_group(const ClassicKArgs& a, const double fmA, const double S1c0, const double S1c1,
                                             const double wold, const int season, const int ti, const int year,
                                             double& dgT, double& dgE, double& dgA, double& dgX) {
    constexpr int N = (I0 + GI <= K) ? GI : K - I0;
    PhysA p[N]; double kjj[N], wj[N], rv[N], se[N], aw[N];
#pragma unroll
    for (int g = 0; g < N; ++g) {
      p[g] = physA(j0 + I0 + g); kjj[g] = coef(j0 + I0 + g).kjj; rv[g] = row(0, I0 + g);
      if constexpr (MIXED) { const PhysB q = physB(j0 + I0 + g); wj[g] = q.wts; aw[g] = q.aw; }
      else wj[g] = pb[j0 + I0 + g].wts;
      se[g] = (SUMNOW || SLOW) ? row(1, I0 + g) : 0.0;
    }
    double S[N], C[N], T[N], En[N], r[N], um[N], G[N];
    int ice[N], tneg[N];   // 0 / -1: cell is ice; T0 < 0
#pragma unroll
    for (int g = 0; g < N; ++g) S[g] = fma(-S1c0, p[g].x, p[g].S0x);
#pragma unroll
    for (int g = 0; g < N; ++g) C[g] = fma(cg_tau, Tg[I0 + g], fmA);
    if constexpr (MIXED) {
#pragma unroll
      for (int g = 0; g < N; ++g) {
        ice[g] = __double2hiint(E[I0 + g]) >> 31;
        const double al = ice[g] ? ai : (is_zero(E[I0 + g]) ? 0.0 : aw[g]);   // alpha = aw [E>0] + ai [E<0]       :47
        C[g] = fma(al, S[g], C[g]);                                                                //                 :48
      }
    } else {
#pragma unroll
      for (int g = 0; g < N; ++g) C[g] = fma(ai, S[g], C[g]);                                  //                 :48
    }
#pragma unroll
    for (int g = 0; g < N; ++g) G[g] = fma(-S1c1, p[g].x, p[g].S0x);
#pragma unroll
    for (int g = 0; g < N; ++g) T[g] = C[g] * rv[g];                // C / (M - kLf/E), reciprocal carried        :50
#pragma unroll
    for (int g = 0; g < N; ++g) G[g] = fma(ai, G[g], fmA);          // ai S[j,i+1] - A + f                         :61
#pragma unroll
    for (int g = 0; g < N; ++g) T[g] = and_mask(T[g], __double2hiint(C[g]) >> 31);   // T0 [T0 < 0], ice cells     :51
#pragma unroll
    for (int g = 0; g < N; ++g) En[g] = fma(-M, T[g], C[g]);
#pragma unroll
    for (int g = 0; g < N; ++g) En[g] = En[g] + Fb;
#pragma unroll
    for (int g = 0; g < N; ++g) En[g] = fma(dt, En[g], E[I0 + g]);                           //                 :53
    if constexpr (MIXED) {
#pragma unroll
      for (int g = 0; g < N; ++g) {
        const double Eo = E[I0 + g];
        const bool water = is_pos(Eo) || (!ice[g] && !is_zero(Eo));   // E > 0 (denormals included)
        const double Enw = fma(dt, C[g] + Fb, c1 * Eo);               // open water: the W path's expression
        const bool cneg = __double2hiint(C[g]) < 0;
        // sign of T0 = C / (M - kLf/E) of a water cell (matters when it freezes in this step): E > 0
        const bool small = __double_as_longlong(Eo) < thr_bits;
        tneg[g] = ice[g] ? (__double2hiint(C[g]) >> 31) : ((water && (cneg != small)) ? -1 : 0);
        En[g] = water ? Enw : En[g];
        T[g] = water ? Eo * inv_cw : (ice[g] ? T[g] : 0.0);           // T = E/cw [E>=0] + T0 [E<0][T0<0]            :51
      }
    }
#pragma unroll
    for (int g = 0; g < N; ++g) {
      if constexpr (MIXED) {   // hemispheric mean of T: ice cells through accT, water cells through accEw (as in the W path)
        accT = fma(wj[g], ice[g] ? T[g] : 0.0, accT);
        accEw = fma(wj[g], ice[g] ? 0.0 : E[I0 + g], accEw);
      } else {
        accT = fma(wj[g], T[g], accT);
      }
    }
#pragma unroll
    for (int g = 0; g < N; ++g) r[g] = fma(M, En[g], -kLf);
    {
      double x[N], e[N];
#pragma unroll
      for (int g = 0; g < N; ++g) asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(x[g]) : "d"(r[g]));
#pragma unroll
      for (int g = 0; g < N; ++g) e[g] = fma(-r[g], x[g], 1.0);
#pragma unroll
      for (int g = 0; g < N; ++g) e[g] = fma(e[g], e[g], e[g]);
#pragma unroll
      for (int g = 0; g < N; ++g) x[g] = fma(x[g], e[g], x[g]);
#pragma unroll
      for (int g = 0; g < N; ++g) r[g] = En[g] * x[g];              // 1/(M - kLf/E) of the new enthalpy
    }
#pragma unroll
    for (int g = 0; g < N; ++g) um[g] = dt_tau * r[g];
#pragma unroll
    for (int g = 0; g < N; ++g) {
      const int t0n = MIXED ? tneg[g] : (__double2hiint(C[g]) >> 31);
      const int mk = t0n & (__double2hiint(En[g]) >> 31);                    // (T0<0) & (E<0), E updated        :56,61
      um[g] = and_mask(um[g], mk);
    }
#pragma unroll
    for (int g = 0; g < N; ++g) {
      const double Ep = and_mask(En[g], ~(__double2hiint(En[g]) >> 31));     // E [E >= 0]                         :59
      S[g] = fma(dttau_cw, Ep, Tg[I0 + g]);
    }
#pragma unroll
    for (int g = 0; g < N; ++g) kjj[g] = fma(-cg_tau, um[g], kjj[g]);         // kappa_jj - dc/(M - kLf/E) [masked]   :56
#pragma unroll
    for (int g = 0; g < N; ++g) Tg[I0 + g] = fma(um[g], G[g], S[g]);          // right-hand side                  :58-62
#pragma unroll
    for (int g = 0; g < N; ++g) {
      if (SUMNOW) se[g] += fma(wold, E[I0 + g], En[g]);
      if (SLOW) {
        sample(a, I0 + g, wj[g], p[g].x, En[g], T[g], se[g], season, ti, year, dgT, dgE, dgA, dgX);
        if (ti == a.nt) se[g] = 0.0;
      }
      E[I0 + g] = En[g];
    }
#pragma unroll
    for (int g = 0; g < N; ++g) {
      row(0, I0 + g) = r[g]; row(2, I0 + g) = kjj[g];
      if (SUMNOW || SLOW) row(1, I0 + g) = se[g];
    }
  }
  template <int I0, bool MIXED, bool SLOW, bool SUMNOW>
  __device__ __forceinline__ void cell_groups(const ClassicKArgs& a, const double fmA, const double S1c0, const double S1c1,
                                              const double wold, const int season, const int ti, const int year,
                                              double& dgT, double& dgE, double& dgA, double& dgX) {
    if constexpr (I0 < K) {
      cell_group<I0, MIXED, SLOW, SUMNOW>(a, fmA, S1c0, S1c1, wold, season, ti, year, dgT, dgE, dgA, dgX);
      cell_groups<I0 + GI, MIXED, SLOW, SUMNOW>(a, fmA, S1c0, S1c1, wold, season, ti, year, dgT, dgE, dgA, dgX);
    }
  }

  // ---- phases A + B of a step: physics, band-local elimination, reduction, post the interface rows.
  // Three straight-line paths; which one a thread takes depends on its state alone (never on what is sampled), so a
  // member's trajectory does not depend on the output options:
  //   W  every cell open water before and after the step          -> folded physics, table pivots
  //   I  every cell ice with a frozen surface (T0 < 0, E' < 0)     -> no selects, fused physics + elimination
  //   G  anything else (ice edge inside the band, melting, freeze-up, E == 0): literal masks as selects
  // W and I compute the new enthalpies first and commit only when their assumption held for all K cells.
  // SUMNOW: this step adds to the annual sums of E (every step when nt is odd; E_old + E_new on even steps otherwise).
  template <bool SLOW, bool SUMNOW>
  __device__ __forceinline__ void advance(const ClassicKArgs& a, const double f, const double S1c0, const double S1c1,
                                          const int ti, const int year, double (&dg)[4]) {
    const int nt = a.nt;
    const double fmA = f - A;
    const double fmAFb = fmA + Fb;
    const int season = SLOW ? ((ti == a.winter_inx) ? 0 : (ti == a.summer_inx) ? 1 : (ti == nt) ? 2 : -1) : -1;
    const double wold = pairsum ? 1.0 : 0.0;
    double dgT = 0.0, dgE = 0.0, dgA = 0.0, dgX = 2.0;
    double i_sl, i_ql, i_yl, i_al, i_be, i_ga;

    int hand = -1, hor = 0;            // AND / OR of the high words: all negative / any negative
    bool allpos = true;
#pragma unroll
    for (int i = 0; i < K; ++i) {
      const int hw = __double2hiint(E[i]);
      hand &= hw; hor |= hw;
      allpos = allpos & is_pos(E[i]);
    }
    bool done = false;
    if (TAB && allpos && !(a.dbg & 1)) {
      // ---- W: alpha = aw, T = E/cw, rows of the implicit system carry kappa alone
      double En[K];
      double acc = accEw;
      bool stay = true;
#pragma unroll
      for (int i = 0; i < K; ++i) {
        const PhysA p = pa[j0 + i];
        const PhysB q = pb[j0 + i];
        const double S = fma(-S1c0, p.x, p.S0x);                    // S[j,i]                              classic.jl:23-25
        const double u = fma(cg_tau, Tg[i], fmAFb);
        const double v = fma(q.aw, S, u);                           // C + Fb                                       :48
        En[i] = fma(dt, v, c1 * E[i]);                              // E + dt (C - M E/cw + Fb), c1 = 1 - dt M/cw   :53
        acc = fma(q.wts, E[i], acc);                                // hemispheric mean of T = E/cw, scaled later
        stay = stay & is_pos(En[i]);
      }
      if (stay) {
        done = true; gen = false; rvalid = false;
        accEw = acc;
        double yprev = 0.0;
#pragma unroll
        for (int i = 0; i < K; ++i) {
          const ElimF e = ef[j0 + i];
          const double rhs = fma(dttau_cw, En[i], Tg[i]);           // Tg + dt_tau E/cw                            :58-59
          const double y = fma(-e.tq, yprev, rhs * e.iw);
          if (SUMNOW || SLOW) {
            double se = row(1, i);
            if (SUMNOW) se += fma(wold, E[i], En[i]);
            if (SLOW) {
              sample(a, i, pb[j0 + i].wts, pa[j0 + i].x, En[i], E[i] * inv_cw, se, season, ti, year, dgT, dgE, dgA, dgX);
              if (ti == nt) se = 0.0;
            }
            row(1, i) = se;
          }
          E[i] = En[i];
          Tg[i] = y; yprev = y;
        }
        double al = Tg[K - 2];
#pragma unroll
        for (int i = K - 3; i >= 0; --i) { al = fma(-eq[j0 + i], al, Tg[i]); Tg[i] = al; }
        const ElimB l = eb[j0 + K - 1], z = eb[j0];
        i_sl = l.be; i_ql = l.ga; i_yl = Tg[K - 1]; i_al = al; i_be = z.be; i_ga = z.ga;
      }
    }
    if (!done) {
      // ---- I / G: physics with the literal masks, then the band's elimination
      gen = true;
      const bool allice = hand < 0;
      if (!allice) set_pads();
      if (!rvalid) {   // first step after open-water steps: r is a pure function of E
#pragma unroll
        for (int i = 0; i < K; ++i) row(0, i) = E[i] * rcp3(fma(M, E[i], -kLf));
        rvalid = true;
      }
      // the all-ice specialisation only when no thread of the warp needs the general code (both give the same bits
      // for an all-ice band, so the vote does not influence any result)
      if (__all_sync(__activemask(), allice) && !(a.dbg & 2))
        cell_groups<0, false, SLOW, SUMNOW>(a, fmA, S1c0, S1c1, wold, season, ti, year, dgT, dgE, dgA, dgX);
      else
        cell_groups<0, true, SLOW, SUMNOW>(a, fmA, S1c0, S1c1, wold, season, ti, year, dgT, dgE, dgA, dgX);
      eliminate(i_sl, i_ql, i_yl);
    }
    if (gen) {
      double al = Tg[K - 2], be = row(3, K - 2), ga = row(2, K - 2);
#pragma unroll
      for (int i = K - 3; i >= 0; --i) {
        const double q = row(2, i), s = row(3, i);
        al = fma(-q, al, Tg[i]);
        be = fma(-q, be, s);
        ga = -q * ga;
        if constexpr (!DEPBS) { Tg[i] = al; row(3, i) = be; row(2, i) = ga; }
      }
      i_al = al; i_be = be; i_ga = ga;
    }
    if (SLOW) {
      if (season == 2) dgT = (accT + inv_cw * accEw) * inv_nt;     // mean over the year of the hemispheric mean (linear)
      if (ti == nt) { accT = 0.0; accEw = 0.0; }
      dg[0] = dgT; dg[1] = dgE; dg[2] = dgA; dg[3] = dgX;
    }
    const int band = pair * 2 + h;
    double* f6 = iface + ((buf * WB + band) * 6) * MW + mi;
    f6[0 * MW] = i_sl; f6[1 * MW] = i_ql; f6[2 * MW] = i_yl; f6[3 * MW] = i_al; f6[4 * MW] = i_be; f6[5 * MW] = i_ga;
  }

  template <int PP>
  __device__ __forceinline__ void pick(const double (&z)[H], double& xL, double& xn) const {
    if (pair == PP) {
      // z[k] of this lane is unknown (h ? WB-1-k : k); wanted: zA = z(2PP-1), zB = z(2PP), zC = z(2PP+1)
      auto get = [&](auto IDX) -> double {
        constexpr int idx = decltype(IDX)::value;
        if constexpr (idx < 0) return 0.0;
        else {
          constexpr bool upper = idx >= H;
          constexpr int kk = upper ? WB - 1 - idx : idx;
          const double other = __shfl_xor_sync(0xffffffffu, z[kk], MW);
          return ((h != 0) == upper) ? z[kk] : other;
        }
      };
      const double zA = get(std::integral_constant<int, 2 * PP - 1>{});
      const double zB = get(std::integral_constant<int, 2 * PP>{});
      const double zC = get(std::integral_constant<int, 2 * PP + 1>{});
      xL = h ? zB : zA;
      xn = h ? zC : zB;
    } else if constexpr (PP + 1 < H) {
      pick<PP + 1>(z, xL, xn);
    }
  }

  // ---- phases C + D: interface system (every warp, for its own 16 members) and back substitution
  __device__ __forceinline__ void solve() {
    const double* base = iface + (buf * WB * 6) * MW + mi;
    // two lanes per member sweep from both ends towards the middle; pivots carried as determinants D_k so that the
    // only dependent chain is one DFMA per row; all reciprocals are independent of each other
    double a_[H], c_[H], r_[H], Dm[H + 1];
    Dm[0] = 1.0;
#pragma unroll
    for (int k = 0; k < H; ++k) {
      const int b = h ? (WB - 1 - k) : k;
      const double* g6 = base + (b * 6) * MW;
      const double* n6 = (b + 1 < WB) ? g6 + 6 * MW : g6;          // the last band has q = 0: any finite row will do
      const double sl = g6[0 * MW], ql = g6[1 * MW], yl = g6[2 * MW];
      const double dgn = fma(-ql, n6[4 * MW], 1.0), sup = -ql * n6[5 * MW];
      r_[k] = fma(-ql, n6[3 * MW], yl);
      a_[k] = h ? sup : sl;                                         // coupling to the previously eliminated row
      c_[k] = h ? sl : sup;                                         // coupling to the next row in sweep order
      Dm[k + 1] = (k == 0) ? dgn : fma(dgn, Dm[k], -(a_[k] * c_[k - 1]) * Dm[k - 1]);
    }
    double cq[H], cy[H];
#pragma unroll
    for (int k = 0; k < H; ++k) {
      const double iw = Dm[k] * rcp3(Dm[k + 1]);
      cq[k] = c_[k] * iw;
      const double g = a_[k] * iw, ri = r_[k] * iw;
      cy[k] = (k == 0) ? ri : fma(-g, cy[k - 1], ri);
    }
    // the two sweeps meet between rows H-1 and H:  x_own = cy_own - cq_own * x_other
    const double ocq = __shfl_xor_sync(0xffffffffu, cq[H - 1], MW);
    const double ocy = __shfl_xor_sync(0xffffffffu, cy[H - 1], MW);
    double z[H];
    z[H - 1] = fma(-cq[H - 1], ocy, cy[H - 1]) * rcp3(fma(-cq[H - 1], ocq, 1.0));
#pragma unroll
    for (int k = H - 2; k >= 0; --k) z[k] = fma(-cq[k], z[k + 1], cy[k]);
    double xL = 0.0, xn = 0.0;
    pick<0>(z, xL, xn);
    // back substitution with the true neighbours
    Tg[K - 1] = xn;
    if (gen) {
      if constexpr (DEPBS) {
#pragma unroll
        for (int i = K - 2; i >= 0; --i) { xn = fma(-row(2, i), xn, fma(-row(3, i), xL, Tg[i])); Tg[i] = xn; }
      } else {
#pragma unroll
        for (int i = K - 2; i >= 0; --i) Tg[i] = fma(-row(2, i), xn, fma(-row(3, i), xL, Tg[i]));
      }
    } else {
#pragma unroll
      for (int i = K - 2; i >= 0; --i) {
        const ElimB e = eb[j0 + i];
        Tg[i] = fma(-e.ga, xn, fma(-e.be, xL, Tg[i]));
      }
    }
  }
};

template <int K, int WB, int MAXR, bool TAB, bool DEPBS, bool ROT>
__global__ void __maxnreg__(MAXR) classic_fused_kernel(const ClassicKArgs a) {
  static_assert(WB % 2 == 0, "two bands per warp");
  extern __shared__ __align__(16) unsigned char smem_raw[];
  using F = Fused<K, WB, TAB, DEPBS>;
  constexpr int NXP = K * WB, NT = WB * MW, H = WB / 2;
  const int tid = threadIdx.x;
  const int lane = tid & 31, warp = tid >> 5;
  const int mi = lane & (MW - 1), h = lane >> 4;
  const int pair = ROT ? (int)((warp + blockIdx.x) % H) : warp;
  const int band = pair * 2 + h;
  const long long nmem = a.nmem;
  const long long m_first = (long long)blockIdx.x * MW;
  const long long m_raw = m_first + mi;
  const bool active = m_raw < nmem;
  const long long m = active ? m_raw : nmem - 1;
  const int nx = a.nx, nt = a.nt;

  // the TAB instance integrates the 32-member groups whose table-building parameters agree, the other instance the rest
  if (ebm_classic_group_uniform<MW>(a.par, nmem, m_first, mi) != TAB) return;
  double par[EBM_CLASSIC_NPAR];
#pragma unroll
  for (int k = 0; k < EBM_CLASSIC_NPAR; ++k) par[k] = a.par[(long long)k * nmem + m];

  // ---- shared memory carve-up
  PhysA* pa = reinterpret_cast<PhysA*>(smem_raw);                 // [NXP]
  PhysB* pb = reinterpret_cast<PhysB*>(pa + NXP);                 // [NXP]
  CoefT* cf = reinterpret_cast<CoefT*>(pb + NXP);                 // [NXP]
  ElimF* ef = reinterpret_cast<ElimF*>(cf + NXP);                 // [NXP]
  ElimB* eb = reinterpret_cast<ElimB*>(ef + NXP);                 // [NXP]
  double* eq = reinterpret_cast<double*>(eb + NXP);               // [NXP]
  double* aoff = eq + NXP;                                        // [NXP + 1] (+ pad to 16 doubles)
  double* iface = aoff + NXP + 16;                                // [2][WB][6][MW]
  double* rows = iface + 2 * WB * 6 * MW;                         // [4 (+2)][K][NT]

  const double pD = par[0], pA = par[1], pB = par[2], pcw = par[3], pS0 = par[4], pS1 = par[5], pS2 = par[6];
  const double pa0 = par[7], pa2 = par[8], pai = par[9], pFb = par[10], pk = par[11], pLf = par[12], pcg = par[13];
  const double ptau = par[14];
  const double dt = 1.0 / nt;
  const double cg_tau = pcg / ptau, dt_tau = dt / ptau;
  const double fac = dt * pD / pcg;          // kappa = (1+dt_tau) I - fac*diffop        classic.jl:21
  const double one_dttau = 1.0 + dt_tau;

  for (int j = tid; j < NXP + 1; j += NT) {
    const bool v = j < nx;
    const double ll = v ? a.g.lam_lo[j] : 0.0;
    aoff[j] = TAB ? -fac * ll : ll;
    if (j < NXP) {
      const double xj = v ? a.g.x[j] : 0.0, x2 = v ? a.g.x2[j] : 0.0;
      const double lh = v ? a.g.lam_hi[j] : 0.0;
      PhysA p; PhysB q; CoefT c;
      if constexpr (TAB) {
        // pad cells (j >= nx): decoupled rows of weight 0, bistable by construction -- as open water they absorb
        // aw S = 1000 W/m^2 and stay warm, as ice ai S ~ 0 and they stay frozen -- so that they never change the path
        // their band takes; Fused::advance flips them when the band's real cells have all changed sign
        p.S0x = v ? fma(-pS2, x2, pS0) : 1e-3; q.aw = v ? fma(-pa2, x2, pa0) : 1e6;
        c.kjj = fma(fac, ll + lh, one_dttau); c.ac = (fac * ll) * (fac * ll);
      } else {   // geometry only: Fused::physA / physB / coef apply the member's parameters
        p.S0x = x2; q.aw = x2;
        c.kjj = ll + lh; c.ac = ll * ll;
      }
      p.x = xj; q.wts = v ? a.g.wts[j] : 0.0;
      pa[j] = p; pb[j] = q; cf[j] = c;
    }
  }
  __syncthreads();
  // band-local elimination of the constant matrix kappa (rows without the ice-mask term)
  if (TAB && tid < WB) {
    const int b = tid;
    double qv[K], sv[K];
    double qprev = 0.0, sprev = 0.0;
    for (int i = 0; i < K; ++i) {
      const int j = b * K + i;
      const double w = (i == 0) ? cf[j].kjj : cf[j].kjj - aoff[j] * qprev;
      ElimF e; e.iw = 1.0 / w; e.tq = aoff[j] * e.iw;
      qv[i] = aoff[j + 1] * e.iw; sv[i] = (i == 0) ? e.tq : -e.tq * sprev;
      ef[j] = e; eq[j] = qv[i];
      qprev = qv[i]; sprev = sv[i];
    }
    // rows K-1 (interface row: s, q) and K-2 (be = s, ga = q) keep the eliminated coefficients
    double be = sv[K - 2], ga = qv[K - 2];
    for (int i = K - 1; i >= 0; --i) {
      ElimB r;
      if (i >= K - 2) { r.be = sv[i]; r.ga = qv[i]; }
      else { be = sv[i] - qv[i] * be; ga = -qv[i] * ga; r.be = be; r.ga = ga; }
      eb[b * K + i] = r;
    }
  }

  F cx;
  cx.pa = pa; cx.pb = pb; cx.cf = cf; cx.aoff = aoff; cx.ef = ef; cx.eb = eb; cx.eq = eq;
  cx.iface = iface; cx.rows = rows;
  cx.S0m = pS0; cx.S2m = pS2; cx.a0m = pa0; cx.a2m = pa2; cx.fac = fac; cx.fac2 = fac * fac; cx.one_dttau = one_dttau;
  cx.ai = pai; cx.cg_tau = cg_tau; cx.M = pB + cg_tau; cx.kLf = pk * pLf; cx.inv_cw = 1.0 / pcw; cx.dt = dt;
  cx.dt_tau = dt_tau; cx.dttau_cw = dt_tau * cx.inv_cw; cx.Fb = pFb; cx.A = pA;
  cx.c1 = fma(-dt * cx.M, cx.inv_cw, 1.0);
  cx.inv_nt = 1.0 / nt; cx.inv_Lf = 1.0 / pLf;
  cx.thr_bits = __double_as_longlong(cx.kLf / cx.M);
  cx.tid = tid; cx.h = h; cx.mi = mi; cx.pair = pair; cx.j0 = band * K;
  cx.nv = min(max(nx - band * K, 0), K);
  cx.active = active;
  cx.pairsum = (nt % 2) == 0;
  const long long mo = a.orig != nullptr ? a.orig[m] : m;   // ebm_classic_device_args_t.member_index
  cx.mo = mo;
  cx.sel = active && a.field_stride > 0 && (mo % a.field_stride) == 0;
  cx.msel = cx.sel ? mo / a.field_stride : 0;
  cx.cta_fields = __syncthreads_or(cx.sel && (a.seasonal != nullptr)) != 0;
  cx.accT = 0.0; cx.accEw = 0.0;
  cx.buf = 0; cx.gen = true; cx.rvalid = true;
#pragma unroll
  for (int i = 0; i < K; ++i) {
    const int j = cx.j0 + i;
    const bool v = j < nx;
    cx.E[i] = (v ? a.E[(long long)j * nmem + m] : 100.0) + 0.0;   // -0.0 -> +0.0; pad cells: see set_pads
    cx.Tg[i] = v ? a.Tg[(long long)j * nmem + m] : 10.0;
    cx.row(0, i) = 0.0;
    cx.row(1, i) = 0.0;
    if (cx.cta_fields) { cx.row(4, i) = 0.0; cx.row(5, i) = 0.0; }
  }
  cx.set_pads();
  cx.rvalid = false;   // the first fused step computes the carried reciprocals from E
  // Forcing{true}: base == peak == cool, all breakpoints 0 -> the call is the constant `base`
  double fr[10];
#pragma unroll
  for (int k = 0; k < 10; ++k) fr[k] = (k == 5) ? 0.0 : a.forc[(long long)k * nmem + m];
  const double fbase = fr[0];
  const bool myconst = fr[1] == fbase && fr[2] == fbase && fr[6] == 0.0 && fr[7] == 0.0 && fr[8] == 0.0 && fr[9] == 0.0;
  const bool constf = __syncthreads_and(myconst) != 0;
  const bool has_raw = __syncthreads_or(cx.sel && (a.raw != nullptr)) != 0;

  double S1c_next = pS1 * __ldg(a.g.ctab);   // S1*cos(2*pi*t_1); ctab[nt] == ctab[0] closes the year (classic.jl:25)
  for (int year = a.year0; year < a.year0 + a.nyears; ++year) {
    const bool raw_year = has_raw && (!a.lastonly || year == a.dur - 1);
    for (int ti = 1; ti <= nt; ++ti) {
      // column i+1 of this step is column i of the next: one table load per step
      const double S1c0 = S1c_next, S1c1 = pS1 * __ldg(a.g.ctab + ti);
      S1c_next = S1c1;
      double f = fbase;
      if (!constf) {
        const long long tinx = (long long)(year + a.start_year) * nt + ti;
        const double* fp = a.forc + m;
        f = ebm_forcing_eval(fbase, __ldg(fp + 1 * nmem), __ldg(fp + 2 * nmem), __ldg(fp + 3 * nmem), __ldg(fp + 4 * nmem),
                             __ldg(fp + 6 * nmem), __ldg(fp + 7 * nmem), __ldg(fp + 8 * nmem), __ldg(fp + 9 * nmem),
                             ebm_global_time(tinx, nt));
      }
      const bool season_step = ti == a.winter_inx || ti == a.summer_inx || ti == nt;
      const bool slow = cx.cta_fields || raw_year || season_step;
      double dg[4];
      const bool sum_now = !cx.pairsum || (ti & 1) == 0;
      if (slow) {
        if (sum_now) cx.template advance<true, true>(a, f, S1c0, S1c1, ti, year, dg);
        else cx.template advance<true, false>(a, f, S1c0, S1c1, ti, year, dg);
      } else if (sum_now) cx.template advance<false, true>(a, f, S1c0, S1c1, ti, year, dg);
      else cx.template advance<false, false>(a, f, S1c0, S1c1, ti, year, dg);
      __syncthreads();
      cx.solve();
      if (season_step) {
        // L0 diagnostics of this season: the WB band partials of a member are summed through the interface buffer
        // this step just consumed (the next step posts to the other one)
        __syncthreads();
        double* red = iface + (cx.buf * WB * 6) * MW;   // [WB][4][MW] fits in [WB][6][MW]
        double* r4 = red + (band * 4) * MW + mi;
        r4[0 * MW] = dg[0]; r4[1 * MW] = dg[1]; r4[2 * MW] = dg[2]; r4[3 * MW] = dg[3];
        __syncthreads();
        if (tid < MW && a.diag != nullptr && active) {
          const int season = (ti == a.winter_inx) ? 0 : (ti == a.summer_inx) ? 1 : 2;
          double t0 = 0.0, t1 = 0.0, t2 = 0.0, t3 = 2.0;
          for (int b = 0; b < WB; ++b) {
            const double* q4 = red + (b * 4) * MW + mi;
            t0 += q4[0 * MW]; t1 += q4[1 * MW]; t2 += q4[2 * MW]; t3 = fmin(t3, q4[3 * MW]);
          }
          double* o = a.diag + ((mo * a.dur + year) * 3 + season) * 4;
          o[0] = t0; o[1] = t1; o[2] = kTwoPi * t2; o[3] = (t3 > 1.5) ? 1.0 : t3;
        }
      }
      cx.buf ^= 1;
    }
  }

  bool bad = false;
#pragma unroll
  for (int i = 0; i < K; ++i) {
    const int j = cx.j0 + i;
    if (j < nx && active) {
      a.E[(long long)j * nmem + m] = cx.E[i];
      a.Tg[(long long)j * nmem + m] = cx.Tg[i];
      bad = bad || !(fabs(cx.E[i]) < 1e300) || !(fabs(cx.Tg[i]) < 1e300);
    }
  }
  if (bad && a.flags != nullptr) atomicOr(a.flags + mo, 1);
}

template <int K, int WB, int MAXR, bool TAB, bool DEPBS = false, bool ROT = true>
int launch_fused(const ClassicKArgs& a, cudaStream_t stream) {
  if (a.nx > K * WB) {
    ebm_set_error("classic_fused: nx=%d exceeds %d bands of %d cells", a.nx, WB, K);
    return EBM_ERR_UNSUPPORTED;
  }
  const bool fields = a.seasonal != nullptr && a.field_stride > 0;
  const size_t smem = fused_smem_bytes<K, WB>(fields);
  auto kern = classic_fused_kernel<K, WB, MAXR, TAB, DEPBS, ROT>;
  EBM_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  EBM_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
  const long long blocks = (a.nmem + MW - 1) / MW;
  kern<<<(unsigned)blocks, WB * MW, smem, stream>>>(a);
  EBM_CUDA_TRY(cudaGetLastError());
  ebm_count_launch();
  return EBM_OK;
}

}  // namespace

// parameter-uniform 32-member groups.  variant (env EBM_CLASSIC_VARIANT, development): alternative instantiations
int ebm_launch_classic_fused(const ClassicKArgs& a, int variant, cudaStream_t stream) {
  if (a.nx > 104) return launch_fused<13, 16, 255, true>(a, stream);          // 16 bands, 8 warps per CTA, 1 CTA per SM
  switch (variant) {
    case 21: return launch_fused<13, 8, 168, true, true, true>(a, stream);    // dependent back substitution (fewer smem stores)
    case 22: return launch_fused<13, 8, 168, true, false, false>(a, stream);  // no band-pair rotation
    case 23: return launch_fused<13, 8, 255, true>(a, stream);                // 2 CTAs per SM, no register cap
    case 24: if (a.nx <= 100) return launch_fused<10, 10, 136, true>(a, stream);   // 10 bands of 10 cells, 15 warps per SM
             return launch_fused<13, 8, 168, true>(a, stream);
    default: return launch_fused<13, 8, 168, true>(a, stream);                // EBM_CLASSIC_VARIANT = 20
  }
}

// groups whose table-building parameters differ between members (e.g. a sweep over D)
int ebm_launch_classic_fused_general(const ClassicKArgs& a, cudaStream_t stream) {
  if (a.nx > 104) return launch_fused<13, 16, 255, false>(a, stream);
  return launch_fused<13, 8, 168, false>(a, stream);
}
