// miz_kernel.cu -- fast flavour of the MIZ ensemble kernel (see miz_kernel.cuh) and the launch dispatch.
#include "miz_kernel.cuh"

int ebm_launch_miz_fast(const MizKArgs& a, cudaStream_t stream) { return miz_launch_any(a, stream); }

int ebm_launch_miz(const MizKArgs& a, int strict, cudaStream_t stream) {
  return strict ? ebm_launch_miz_strict(a, stream) : ebm_launch_miz_fast(a, stream);
}

// step!(Val(:MIZ), ...) for one member (src/miz.jl:150-196) with the literal-order kernel: the ten stored
// variables of the step go to vars_out [EBM_MIZ_NVAR][nx]; state and the closure warm start are updated in place.
int ebm_launch_miz_single_step(const EbmGridTables& g, const double* par22, int ti, double f, double tol, int maxit,
                               double* Ei, double* Ew, double* h, double* D, double* phi, double* T0,
                               double* vars_out, long long* iters, cudaStream_t stream) {
  MizKArgs a;
  memset(&a, 0, sizeof(a));
  a.nx = g.nx; a.nt = g.nt; a.dur = 1; a.nmem = 1; a.year0 = 0; a.nyears = 1;
  a.winter_inx = -1; a.summer_inx = -1; a.lastonly = 1; a.field_stride = 1;
  a.maxit = maxit; a.tol = tol; a.single_ti = ti; a.single_f = f;
  a.g = g; a.par = par22; a.forc = nullptr;
  a.Ei = Ei; a.Ew = Ew; a.h = h; a.D = D; a.phi = phi; a.T0 = T0;
  a.raw = vars_out; a.newton_iters = iters;
  return ebm_launch_miz_strict(a, stream);
}
