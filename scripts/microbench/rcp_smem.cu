// dev microbenchmark (not part of the library): accuracy of the MUFU.RCP64H seed and of the refinement variants used
// by the classic kernel; shared-memory wavefront costs of the access shapes it uses; DFMA issue rate of a half-active
// warp; cost of a 4-warp CTA barrier.
#include <cstdio>
#include <cmath>
#include <cuda_runtime.h>

__device__ __forceinline__ double seed(double w) { double x; asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(x) : "d"(w)); return x; }

__global__ void rcp_err(double* maxerr, int n) {
  // relative errors of: seed, seed + 1 Newton, seed + 2 Newton, seed + cubic (Halley-like), seed + cubic + Newton
  double m[5] = {0, 0, 0, 0, 0};
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    // w sweeps [1, 2) finely and a few binades
    const double w = ldexp(1.0 + (double)i / n, (i % 7) * 9 - 20) * ((i & 1) ? -1.0 : 1.0);
    const double ex = 1.0 / w;
    const double x0 = seed(w);
    double e = fma(-w, x0, 1.0);
    const double x1 = fma(x0, e, x0);
    double e1 = fma(-w, x1, 1.0);
    const double x2 = fma(x1, e1, x1);
    const double t = fma(e, e, e);
    const double xc = fma(x0, t, x0);
    double ec = fma(-w, xc, 1.0);
    const double xcn = fma(xc, ec, xc);
    const double v[5] = {x0, x1, x2, xc, xcn};
    for (int k = 0; k < 5; ++k) m[k] = fmax(m[k], fabs((v[k] - ex) / ex));
  }
  for (int k = 0; k < 5; ++k) {
    for (int o = 16; o; o >>= 1) m[k] = fmax(m[k], __shfl_xor_sync(0xffffffffu, m[k], o));
    if ((threadIdx.x & 31) == 0) atomicMax((unsigned long long*)&maxerr[k], (unsigned long long)__double_as_longlong(m[k]));
  }
}

// ---- shared memory access shapes: cycles per warp-instruction with `warps` warps of one CTA hammering
template <int MODE>
__global__ void smem_rate(double* out, int iters, long long* cyc) {
  extern __shared__ __align__(16) double sm[];
  const int tid = threadIdx.x, lane = tid & 31;
  for (int i = tid; i < 8192; i += blockDim.x) sm[i] = i;
  __syncthreads();
  double acc = 0.0;
  double2 acc2 = make_double2(0, 0);
  long long t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      if (MODE == 0) acc += sm[(k * 128 + tid) & 8191];                                   // LDS.64 private [row][thread]
      if (MODE == 1) { const double2 v = *reinterpret_cast<double2*>(&sm[2 * ((k * 16 + (lane >> 4) * 13 + (tid >> 5) * 26) & 1023)]); acc2.x += v.x; acc2.y += v.y; }  // LDS.128, 2 addresses per warp
      if (MODE == 2) acc += sm[(k * 16 + (lane >> 4) * 13 + (tid >> 5) * 26) & 1023];     // LDS.64, 2 addresses per warp
      if (MODE == 3) sm[(k * 128 + tid) & 8191] = acc + k;                                  // STS.64 private
      if (MODE == 4) { const double2 v = *reinterpret_cast<double2*>(&sm[2 * ((k * 128 + tid) & 4095)]); acc2.x += v.x; acc2.y += v.y; }  // LDS.128 private
    }
  }
  long long t1 = clock64();
  if (acc + acc2.x + acc2.y == 1.2345) out[0] = acc;
  if (tid == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}

__global__ void half_warp_dfma(double* out, int iters, long long* cyc, int active_lanes) {
  double a0 = threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
  long long t0 = clock64();
  if ((threadIdx.x & 31) < active_lanes) {
#pragma unroll 1
    for (int i = 0; i < iters; ++i) {
      a0 = fma(a0, 1.0000001, 1e-9); a1 = fma(a1, 1.0000001, 1e-9); a2 = fma(a2, 1.0000001, 1e-9); a3 = fma(a3, 1.0000001, 1e-9);
      a4 = fma(a4, 1.0000001, 1e-9); a5 = fma(a5, 1.0000001, 1e-9); a6 = fma(a6, 1.0000001, 1e-9); a7 = fma(a7, 1.0000001, 1e-9);
    }
  }
  long long t1 = clock64();
  if (a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7 == 1.2345) out[0] = a0;
  if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}

__global__ void barrier_cost(double* out, int iters, long long* cyc) {
  double a = threadIdx.x;
  long long t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < iters; ++i) { a = fma(a, 1.0000001, 1e-9); __syncthreads(); }
  long long t1 = clock64();
  if (a == 1.2345) out[0] = a;
  if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}

int main() {
  double* out; long long* cyc; double* me;
  cudaMalloc(&out, 8); cudaMalloc(&cyc, 8); cudaMalloc(&me, 40); cudaMemset(me, 0, 40);
  rcp_err<<<296, 256>>>(me, 1 << 26);
  double h[5]; cudaMemcpy(h, me, 40, cudaMemcpyDeviceToHost);
  printf("rcp relative error: seed %.3e (2^%.1f)  +1 Newton %.3e  +2 Newton %.3e  cubic %.3e  cubic+Newton %.3e\n",
         h[0], log2(h[0]), h[1], h[2], h[3], h[4]);
  const char* names[5] = {"LDS.64 private", "LDS.128 2-address broadcast", "LDS.64 2-address broadcast", "STS.64 private", "LDS.128 private"};
  long long c; const int iters = 2048;
  for (int warps : {1, 4, 12}) {
    void (*k[5])(double*, int, long long*) = {smem_rate<0>, smem_rate<1>, smem_rate<2>, smem_rate<3>, smem_rate<4>};
    for (int m = 0; m < 5; ++m) {
      k[m]<<<1, 32 * warps, 65536>>>(out, iters, cyc);
      cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
      printf("%-30s warps=%2d: %.2f cycles per warp-instruction (SM-wide)\n", names[m], warps, (double)c / (iters * 8.0 * warps));
    }
  }
  for (int al : {32, 16, 8}) {
    for (int warps : {4, 8, 16}) {
      half_warp_dfma<<<1, 32 * warps>>>(out, 4096, cyc, al);
      cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
      printf("DFMA active lanes=%2d warps/SM=%2d: %.2f cycles per warp-DFMA per SMSP\n", al, warps, (double)c / (4096.0 * 8 * warps / 4));
    }
  }
  for (int warps : {4, 8}) {
    barrier_cost<<<1, 32 * warps>>>(out, 4096, cyc);
    cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
    printf("__syncthreads + 1 DFMA, %d warps: %.1f cycles per iteration\n", warps, (double)c / 4096.0);
  }
  printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
  return 0;
}
