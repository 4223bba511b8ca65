/*
 * ebm_cuda.h -- C ABI of libebm_cuda.so: the B200 (sm_100a) ensemble integrator for the
 * time-stepping path of EnergyBalanceModel.jl.
 *
 * This is the drop-in boundary.  The reference has no FFI of its own (it is pure Julia); the entry
 * points below are what a Julia package extension (julia/ext/EBMCUDAExt.jl, same mechanism as
 * ext/CairoExt.jl:7-12 + Project.toml:18-24) binds with `ccall` to replace, for whole ensembles,
 *
 *     Infrastructure.integrate      src/infrastructure.jl:615-636   (time loop + savesol!)
 *     Infrastructure.step!(:Classic) src/classic.jl:37-71
 *     Infrastructure.step!(:MIZ)     src/miz.jl:150-196
 *     savesol! / annual_mean         src/infrastructure.jl:536-591
 *     (forcing::Forcing)(T)          src/infrastructure.jl:294-307
 *     hemispheric_mean               src/utilities.jl:397-403       (L0 diagnostics)
 *
 * Conventions
 *   - extern "C", plain pointers and sizes; Float64 / Int32 / Int64 only.  No exceptions cross.
 *   - Every function returns an int32 status: 0 = ok, <0 = error class; ebm_last_error() gives the
 *     message (thread local).  There is NO CPU fallback: without a CUDA device every compute entry
 *     point fails with EBM_ERR_CUDA.
 *   - The caller owns every buffer it passes; the library never keeps a pointer after returning.
 *   - "host" entry points take host buffers and do H2D / D2H themselves; "device" entry points take
 *     device pointers (e.g. torch tensors' data_ptr) and only enqueue kernels on the given stream.
 *   - All 1-based indices (winter_inx, summer_inx) are exactly the values stored in Julia's SpaceTime.
 */
#ifndef EBM_CUDA_H
#define EBM_CUDA_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define EBM_OK 0
#define EBM_ERR_INVALID (-1)     /* invalid argument (Julia side: ArgumentError) */
#define EBM_ERR_CUDA (-2)        /* CUDA runtime / no device / kernel fault */
#define EBM_ERR_OOM (-3)         /* device or host allocation failed */
#define EBM_ERR_UNSUPPORTED (-4) /* e.g. `debug::Expr` has no device equivalent */

/* ---- grid: SpaceTime{F} (src/infrastructure.jl:109-137), arrays passed verbatim ------------- */
typedef struct ebm_grid {
  int32_t nx;         /* st.nx  */
  int32_t nt;         /* st.nt  */
  int32_t dur;        /* st.dur */
  int32_t grid_kind;  /* 0: SpaceTime{identity} (get_diffop, infrastructure.jl:480-497); 1: generic stencil (:500-527) */
  int32_t winter_inx; /* st.winter.inx (1-based) */
  int32_t summer_inx; /* st.summer.inx (1-based) */
  const double* x;    /* st.x [nx], host */
  const double* t;    /* st.t [nt], host */
} ebm_grid_t;

/* ---- parameters: Collection{Float64} marshalled by name -------------------------------------- */
/* classic_paramset, src/infrastructure.jl:442-444 (par.F is never read by step!) */
typedef struct ebm_classic_params {
  double D, A, B, cw, S0, S1, S2, a0, a2, ai, Fb, k, Lf, cg, tau;
} ebm_classic_params_t;
#define EBM_CLASSIC_NPAR 15

/* miz_paramset, src/infrastructure.jl:436-441 */
typedef struct ebm_miz_params {
  double D, A, B, cw, S0, S1, S2, a0, a2, ai, Fb, k, Lf, Tm, m1, m2, alpha, rl, Dmin, Dmax, hmin, kappa;
} ebm_miz_params_t;
#define EBM_MIZ_NPAR 22

/* Forcing{C}, src/infrastructure.jl:208-215.  A constant forcing has domain = {0,0,0,0,0} and
 * base == peak == cool.  domain holds integer years, stored as doubles so the struct is 10 doubles. */
typedef struct ebm_forcing {
  double base, peak, cool, rate_up, rate_down;
  double domain[5];
} ebm_forcing_t;
#define EBM_NFORCING 10

/* ---- stored variables (src/infrastructure.jl:621-624) ---------------------------------------- */
enum { EBM_CV_E = 0, EBM_CV_T = 1, EBM_CV_h = 2, EBM_CLASSIC_NVAR = 3 };
/* order of src/EnergyBalanceModel.jl:63 */
enum { EBM_MV_T = 0, EBM_MV_Ei, EBM_MV_Ti, EBM_MV_D, EBM_MV_n, EBM_MV_h, EBM_MV_phi, EBM_MV_E, EBM_MV_Ew, EBM_MV_Tw,
       EBM_MIZ_NVAR = 10 };
enum { EBM_SEASON_WINTER = 0, EBM_SEASON_SUMMER = 1, EBM_SEASON_AVG = 2, EBM_NSEASON = 3 };
/* L0 diagnostics per member-year-season */
enum { EBM_DIAG_MEAN_T = 0, EBM_DIAG_MEAN_E = 1, EBM_DIAG_ICE_AREA = 2, EBM_DIAG_ICE_EDGE = 3, EBM_NDIAG = 4 };

/* ---- options ----------------------------------------------------------------------------------- */
typedef struct ebm_options {
  int32_t device;           /* CUDA ordinal; -1 = current device */
  int32_t lastonly;         /* integrate(...; lastonly): raw holds the last year (1) or every step (0) */
  int32_t field_stride;     /* members with m % field_stride == 0 get L1/L2 field output; 0 = none */
  int32_t strict;           /* 1: literal-arithmetic kernel (IEEE div, no FMA contraction, serial LU-order
                               tridiagonal solve) -- slow, for parity debugging */
  int32_t years_per_launch; /* 0: whole run in one launch */
  int32_t newton_maxit;     /* MIZ closure iteration cap, 0 -> 100 */
  double newton_tol;        /* MIZ closure max|residual| stop, 0 -> 1e-8 (miz.jl:59 abstol) */
  int32_t step_limit;       /* MIZ only, > 0: stop after this many time steps (partial last year; outputs of the
                               steps not taken stay NaN).  Used to compare short horizons from a given state: the
                               MIZ dynamics amplify rounding differences, so long-run pointwise
                               parity is not defined even for the reference itself. */
  int32_t start_year;       /* years already simulated before this run (restart): Forcing is evaluated at
                               T + start_year; output slots stay indexed from the start of this run */
  int32_t classic_stencil;  /* classic only.  0: kappa from get_diffop(nx) whatever the grid -- what the reference does
                               (src/classic.jl:21), correct on SpaceTime{identity} only; 1 (extension): kappa
                               from the generic flux-form stencil of src/infrastructure.jl:500-527, i.e. the classic model
                               on non-uniform grids (on the identity grid both agree to rounding) */
} ebm_options_t;

/* ---- outputs (host entry points).  Any pointer may be NULL = not wanted. ------------------------
 * nsel = number of members with m % field_stride == 0 = ceil(nmem / field_stride); nraw = lastonly ? nt : nt*dur.
 * Entries the reference would leave `undef` (never assigned by savesol!) are NaN.                   */
typedef struct ebm_classic_outputs {
  double* diag;     /* [nmem][dur][EBM_NSEASON][EBM_NDIAG]                          (L0) */
  double* seasonal; /* [nsel][dur][EBM_NSEASON][EBM_CLASSIC_NVAR][nx]               (L1: Solutions.seasonal) */
  double* raw;      /* [nsel][nraw][EBM_CLASSIC_NVAR][nx]                           (L2: Solutions.raw) */
  double* E_final;  /* [nmem][nx] */
  double* Tg_final; /* [nmem][nx]  (classic Tg is not in Solutions; returned so runs can be chained) */
  int32_t* flags;   /* [nmem] bit0: NaN/Inf in final state */
} ebm_classic_outputs_t;

typedef struct ebm_miz_outputs {
  double* diag;     /* [nmem][dur][EBM_NSEASON][EBM_NDIAG] */
  double* seasonal; /* [nsel][dur][EBM_NSEASON][EBM_MIZ_NVAR][nx] */
  double* raw;      /* [nsel][nraw][EBM_MIZ_NVAR][nx] */
  double* Ei_final; double* Ew_final; double* h_final; double* D_final; double* phi_final; /* [nmem][nx] */
  double* T0_final; /* [nmem][nx] closure warm start (solveTi's persistent T0, miz.jl:47,64) */
  int64_t* newton_iters; /* [nmem] total closure iterations */
  int64_t* nonconv;      /* [nmem] steps whose closure hit newton_maxit (reference: @warn only, miz.jl:61-63) */
  int32_t* flags;        /* [nmem] bit0: NaN/Inf in final state */
} ebm_miz_outputs_t;

/* ---- device entry points: caller-owned device memory, member index fastest ---------------------- */
typedef struct ebm_classic_device_args {
  int64_t nmem;
  const double* par;  /* [EBM_CLASSIC_NPAR][nmem] */
  const double* forc; /* [EBM_NFORCING][nmem] */
  double* E;          /* [nx][nmem] in/out */
  double* Tg;         /* [nx][nmem] in/out */
  double* diag;       /* NULL or [nmem][dur][3][4]; must be pre-filled by the caller (NaN) */
  double* seasonal;   /* NULL or [nsel][dur][3][3][nx] */
  double* raw;        /* NULL or [nsel][nraw][3][nx] */
  int32_t* flags;     /* NULL or [nmem], zero-initialised by the caller */
  const int64_t* member_index; /* NULL, or [nmem] (device): original index of the member held in slot m.  Output rows
                         (diag, flags) and the field_stride selection use it, so a caller may hand the members over in
                         any order -- e.g. sorted by regime, which is what ebm_classic_run does internally: lanes of a
                         warp are members, and neighbouring members in the same regime do not diverge */
} ebm_classic_device_args_t;

typedef struct ebm_miz_device_args {
  int64_t nmem;
  const double* par;  /* [EBM_MIZ_NPAR][nmem] */
  const double* forc; /* [EBM_NFORCING][nmem] */
  double* Ei; double* Ew; double* h; double* D; double* phi; double* T0; /* [nx][nmem] in/out */
  double* diag; double* seasonal; double* raw;
  int64_t* newton_iters; int64_t* nonconv; int32_t* flags;               /* NULL or [nmem], zeroed by caller */
} ebm_miz_device_args_t;

/* ---- library ------------------------------------------------------------------------------------ */
const char* ebm_version(void);
const char* ebm_last_error(void);
int32_t ebm_device_count(void);
/* number of kernels this library has launched so far in this process (for bench accounting) */
int64_t ebm_launch_count(void);
/* release cached device tables, the workspace, and the NCCL communicators of the *_multi entry points */
int32_t ebm_shutdown(void);

/* ---- classic: integrate(:Classic, st, forcing[], par[], init[]) for nmem members ----------------
 * replaces src/infrastructure.jl:615-636 + src/classic.jl:37-71 + savesol! (:549-591).
 * par[nmem], forc[nmem]; E0/Tg0 [nmem][nx] (init.E, init.Tg).                                         */
int32_t ebm_classic_run(const ebm_grid_t* grid, int64_t nmem, const ebm_classic_params_t* par,
                        const ebm_forcing_t* forc, const double* E0, const double* Tg0,
                        const ebm_options_t* opt, ebm_classic_outputs_t* out);
/* same, inputs resident in HBM; enqueues on `stream` (cudaStream_t) and returns without waiting for the integration.
 * (nx <= 104: a 124-byte read-back at entry -- are the 15 parameters the same for every member? -- waits for work
 * queued earlier on `stream`; such ensembles take the kernel instance with the member constants in the argument block.) */
int32_t ebm_classic_run_device(const ebm_grid_t* grid, const ebm_classic_device_args_t* args,
                               const ebm_options_t* opt, void* stream);
/* one step of one member, step!(Val(:Classic), t, f, vars, st, par) (src/classic.jl:37-71); `ti` is the
 * 1-based index of t in st.t.  E, Tg in/out; T, h out; all [nx] host.  Uses the strict kernel.          */
int32_t ebm_classic_step(const ebm_grid_t* grid, const ebm_classic_params_t* par, int32_t ti, double f,
                         double* E, double* Tg, double* T, double* h);
/* the same step with a debug variable.  The reference evaluates an arbitrary `debug::Expr` in the scope of step!
 * (src/classic.jl:67-69) and stores it as vars.debug; an expression cannot run on the device, so the seam offers the
 * per-cell locals such expressions name as a fixed menu: debug_out [nx] receives the selected one
 * (alpha :47, C :48, T0 :50, S = stat.S[:,i] :48, mask = (T0<0)&(E<0) :56 as 0/1).                       */
enum { EBM_DEBUG_NONE = 0, EBM_DEBUG_ALPHA = 1, EBM_DEBUG_C = 2, EBM_DEBUG_T0 = 3, EBM_DEBUG_S = 4, EBM_DEBUG_MASK = 5 };
int32_t ebm_classic_step_debug(const ebm_grid_t* grid, const ebm_classic_params_t* par, int32_t ti, double f,
                               double* E, double* Tg, double* T, double* h, int32_t which, double* debug_out);

/* ---- MIZ: integrate(:MIZ, ...) -- src/miz.jl:150-196 + closure :33-68 ----------------------------- */
int32_t ebm_miz_run(const ebm_grid_t* grid, int64_t nmem, const ebm_miz_params_t* par,
                    const ebm_forcing_t* forc, const double* Ei0, const double* Ew0, const double* h0,
                    const double* D0, const double* phi0, const double* T0guess /* NULL = zeros */,
                    const ebm_options_t* opt, ebm_miz_outputs_t* out);
int32_t ebm_miz_run_device(const ebm_grid_t* grid, const ebm_miz_device_args_t* args,
                           const ebm_options_t* opt, void* stream);
/* one step of one member, step!(Val(:MIZ), ...); state in/out, T0 = closure warm start in/out,
 * vars_out [EBM_MIZ_NVAR][nx] = the ten stored variables after the step (with the NaN masks).          */
int32_t ebm_miz_step(const ebm_grid_t* grid, const ebm_miz_params_t* par, int32_t ti, double f,
                     double* Ei, double* Ew, double* h, double* D, double* phi, double* T0,
                     double* vars_out, int32_t* newton_iters);

/* ---- several GPUs behind one call ------------------------------------------------
 * Members are independent (src/infrastructure.jl:615-636 integrates one member; nothing couples two), so the path
 * shards with no data-path collective.  Single process, one host thread + stream per GPU; members are dealt to
 * the GPUs in packets after a stable sort by what they cost (classic: regime of the initial state).  Every output
 * row comes back at the member's original index; seasonal / raw rows belong to the members whose ORIGINAL index is a
 * multiple of field_stride, in that order, exactly as in the single-GPU call.                                    */
typedef struct ebm_multi {
  int32_t ndevices;       /* GPUs to use; 0 = every visible device */
  int32_t diag_device;    /* -1: out->diag is host memory, every GPU copies its own rows back;
                             >= 0: out->diag is DEVICE memory on that ordinal (one of `devices`), pre-filled by the
                             caller; the other GPUs send their rows there over NVLink -- ncclSend/ncclRecv on
                             communicators the library owns (created lazily, destroyed by ebm_shutdown) */
  int32_t packet;         /* members per dealt packet, 0 -> 32 */
  int32_t reserved;
  const int32_t* devices; /* NULL = ordinals 0 .. ndevices-1 */
} ebm_multi_t;
int32_t ebm_classic_run_multi(const ebm_grid_t* grid, int64_t nmem, const ebm_classic_params_t* par,
                              const ebm_forcing_t* forc, const double* E0, const double* Tg0,
                              const ebm_options_t* opt /* opt->device is ignored */, const ebm_multi_t* multi,
                              ebm_classic_outputs_t* out);
int32_t ebm_miz_run_multi(const ebm_grid_t* grid, int64_t nmem, const ebm_miz_params_t* par,
                          const ebm_forcing_t* forc, const double* Ei0, const double* Ew0, const double* h0,
                          const double* D0, const double* phi0, const double* T0guess /* NULL = zeros */,
                          const ebm_options_t* opt, const ebm_multi_t* multi, ebm_miz_outputs_t* out);

/* ---- layout helpers (device, on `stream`): [rows][cols] -> [cols][rows] ---------------------------- */
int32_t ebm_transpose_device(const double* src, double* dst, int64_t rows, int64_t cols, void* stream);

/* ---- FP64 roof: measured DFMA throughput of the device (TFLOP/s, FMA = 2 flop) --------------------- */
int32_t ebm_fp64_peak(int32_t device, double* tflops, double* sm_clock_mhz_est);

#ifdef __cplusplus
}
#endif
#endif /* EBM_CUDA_H */
