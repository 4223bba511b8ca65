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
  int tid, h, mi, pair, j0;
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
    } else if (hand < 0 && !(a.dbg & 2)) {
      // ---- I: every cell ice (alpha = ai).  The remaining masks of classic.jl:51,56-61 -- frozen surface (C < 0,
      // hence T0 < 0) and E' < 0 -- are applied as bit masks on operands, exact in every case:
      //   T = C<0 ? T0 : 0;  masked = (C<0) & (E'<0);  um = masked ? dt_tau/(M - kLf/E') : 0
      //   diag = kappa_jj - cg_tau um;  rhs = Tg + dt_tau/cw max(E', 0) + um (ai S' - A + f)
      done = true; gen = true;
      if (!rvalid) {   // first step after open-water steps: r is a pure function of E
#pragma unroll
        for (int i = 0; i < K; ++i) row(0, i) = E[i] * rcp3(fma(M, E[i], -kLf));
        rvalid = true;
      }
      double Pm1 = 1.0, Pm2 = 1.0, yprev = 0.0, sprev = 0.0;
      double aj = sub(j0);
      // software pipeline: the loads of cell i+1 (tables, carried reciprocal, annual sum) are issued before the stores
      // of cell i -- the scheduler cannot prove that table loads and row stores do not alias, and would otherwise
      // run the cells strictly one after the other
      PhysA p_n = physA(j0); CoefT c_n = coef(j0);
      double cj_n = sub(j0 + 1), wj_n = pb[j0].wts, rv_n = row(0, 0), se_n = (SUMNOW || SLOW) ? row(1, 0) : 0.0;
#pragma unroll
      for (int i = 0; i < K; ++i) {
        const int j = j0 + i;
        const PhysA p = p_n; const CoefT c = c_n;
        const double cj = cj_n, wj = wj_n, rv = rv_n;
        double se = se_n;
        if (i + 1 < K) {
          p_n = physA(j + 1); c_n = coef(j + 1); cj_n = sub(j + 2); wj_n = pb[j + 1].wts; rv_n = row(0, i + 1);
          if (SUMNOW || SLOW) se_n = row(1, i + 1);
        }
        const double Eo = E[i], Tgo = Tg[i];
        const double S = fma(-S1c0, p.x, p.S0x);
        const double C = fma(ai, S, fma(cg_tau, Tgo, fmA));        //                                              :48
        const double T0 = C * rv;                                   // C / (M - kLf/E), reciprocal carried          :50
        const double T = and_mask(T0, __double2hiint(C) >> 31);     // T0 [T0 < 0]                                   :51
        const double En = fma(dt, fma(-M, T, C) + Fb, Eo);          //                                              :53
        const int mk = (__double2hiint(C) & __double2hiint(En)) >> 31;   // (T0<0) & (E<0), E updated            :56,61
        const double r = En * rcp3(fma(M, En, -kLf));               // 1/(M - kLf/E) of the new enthalpy
        const double um = and_mask(dt_tau * r, mk);
        const double Ep = and_mask(En, ~(__double2hiint(En) >> 31));   // E [E >= 0]                                 :59
        const double G = fma(ai, fma(-S1c1, p.x, p.S0x), fmA);      // ai S[j,i+1] - A + f                          :61
        const double rhs = fma(um, G, fma(dttau_cw, Ep, Tgo));      //                                           :58-62
        const double diag = fma(-cg_tau, um, c.kjj);                // kappa_jj - dc/(M - kLf/E) [masked]            :56
        // pivots in determinant form: P_i = w_0 ... w_i = d_i P_{i-1} - a_i^2 P_{i-2}; 1/w_i = P_{i-1}/P_i
        const double P = (i == 0) ? diag : fma(diag, Pm1, -(c.ac * Pm2));
        const double iw = Pm1 * rcp3(P);
        const double tq = aj * iw;
        const double y = (i == 0) ? rhs * iw : fma(-tq, yprev, rhs * iw);
        const double s = (i == 0) ? tq : -tq * sprev;
        const double q = cj * iw;
        accT = fma(wj, T, accT);
        row(0, i) = r;
        if (SUMNOW || SLOW) {
          if (SUMNOW) se += fma(wold, Eo, En);
          if (SLOW) {
            sample(a, i, wj, p.x, En, T, se, season, ti, year, dgT, dgE, dgA, dgX);
            if (ti == nt) se = 0.0;
          }
          row(1, i) = se;
        }
        row(2, i) = q; row(3, i) = s;
        E[i] = En; Tg[i] = y;
        Pm2 = Pm1; Pm1 = P; yprev = y; sprev = s; aj = cj;
        if (i == K - 1) { i_sl = s; i_ql = q; i_yl = y; }
      }
    }
    if (!done) {
      // ---- G: literal masks (classic.jl:47-63) as selects, fused with the elimination
      gen = true;
      if (!rvalid) {
#pragma unroll
        for (int i = 0; i < K; ++i) row(0, i) = E[i] * rcp3(fma(M, E[i], -kLf));
        rvalid = true;
      }
      double Pm1 = 1.0, Pm2 = 1.0, yprev = 0.0, sprev = 0.0;
      double aj = sub(j0);
#pragma unroll
      for (int i = 0; i < K; ++i) {
        const int j = j0 + i;
        const PhysA p = physA(j);
        const PhysB pq = physB(j);
        const CoefT c = coef(j);
        const double cj = sub(j + 1);
        const double S = fma(-S1c0, p.x, p.S0x);
        const double Eo = E[i], Tgo = Tg[i];
        const bool ice = is_neg(Eo), zero = is_zero(Eo);
        const double alpha = ice ? ai : (zero ? 0.0 : pq.aw);       // alpha = aw [E>0] + ai [E<0]                  :47
        const double inner = fma(cg_tau, Tgo, fmA);
        const double C = fma(alpha, S, inner);                      //                                              :48
        const double T0 = C * row(0, i);                            //                                              :50
        const bool Cneg = is_neg(C);                                // E < 0: M - kLf/E > 0, so T0 < 0 <=> C < 0
        const double Ti = Cneg ? T0 : 0.0;
        const double T = ice ? Ti : Eo * inv_cw;                    //                                              :51
        // same arithmetic as the W / I paths for water / ice cells
        const double En_w = fma(dt, (C + Fb), c1 * Eo);
        const double En_i = fma(dt, fma(-M, Ti, C) + Fb, Eo);
        const double En = (ice | zero) ? En_i : En_w;               //                                              :53
        const bool negn = is_neg(En);
        // sign of T0 of a water cell (only matters when it freezes in this step): sign(C) * sign(M - kLf/E), E > 0
        const bool small = __double_as_longlong(Eo) < thr_bits;
        const bool T0neg = ice ? Cneg : ((!zero) & (Cneg != small));
        const bool masked = T0neg & negn;                           // (T0<0) & (E<0), E updated                :56,61
        const double r = En * rcp3(fma(M, En, -kLf));
        row(0, i) = r;
        const double u = dt_tau * r;
        const double G = fma(ai, fma(-S1c1, p.x, p.S0x), fmA);
        const double rhs = masked ? fma(u, G, Tgo) : (negn ? Tgo : fma(dttau_cw, En, Tgo));               // :58-62
        const double diag = masked ? fma(-cg_tau, u, c.kjj) : c.kjj;                                       // :56
        const double P = (i == 0) ? diag : fma(diag, Pm1, -(c.ac * Pm2));
        const double iw = Pm1 * rcp3(P);
        const double tq = aj * iw;
        const double y = (i == 0) ? rhs * iw : fma(-tq, yprev, rhs * iw);
        const double s = (i == 0) ? tq : -tq * sprev;
        const double q = cj * iw;
        row(2, i) = q; row(3, i) = s;
        // hemispheric mean of T: ice cells through accT, water cells through accEw (as in the W path)
        accT = fma(pq.wts, ice ? Ti : 0.0, accT);
        accEw = fma(pq.wts, ice ? 0.0 : Eo, accEw);
        if (SUMNOW || SLOW) {
          double se = row(1, i);
          if (SUMNOW) se += fma(wold, Eo, En);
          if (SLOW) {
            sample(a, i, pq.wts, p.x, En, T, se, season, ti, year, dgT, dgE, dgA, dgX);
            if (ti == nt) se = 0.0;
          }
          row(1, i) = se;
        }
        E[i] = En; Tg[i] = y;
        Pm2 = Pm1; Pm1 = P; yprev = y; sprev = s; aj = cj;
        if (i == K - 1) { i_sl = s; i_ql = q; i_yl = y; }
      }
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
        // pad cells (j >= nx): decoupled rows that stay open water (S0x = 1000 keeps their E positive), weight 0
        p.S0x = v ? fma(-pS2, x2, pS0) : 1000.0; q.aw = v ? fma(-pa2, x2, pa0) : 1.0;
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
    cx.E[i] = (v ? a.E[(long long)j * nmem + m] : 1.0) + 0.0;     // pad cells: decoupled open-water rows; -0.0 -> +0.0
    cx.Tg[i] = v ? a.Tg[(long long)j * nmem + m] : 0.0;
    cx.row(0, i) = cx.E[i] * rcp3(fma(cx.M, cx.E[i], -cx.kLf));   // same expression as in the step: r is a pure function of E
    cx.row(1, i) = 0.0;
    if (cx.cta_fields) { cx.row(4, i) = 0.0; cx.row(5, i) = 0.0; }
  }
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
    default: return launch_fused<13, 8, 168, true>(a, stream);
  }
}

// groups whose table-building parameters differ between members (e.g. a sweep over D)
int ebm_launch_classic_fused_general(const ClassicKArgs& a, cudaStream_t stream) {
  if (a.nx > 104) return launch_fused<13, 16, 255, false>(a, stream);
  return launch_fused<13, 8, 168, false>(a, stream);
}
