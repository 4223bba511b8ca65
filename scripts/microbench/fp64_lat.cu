// dev microbenchmark: FP64 pipe latency / issue behaviour on B200 (not part of the library)
#include <cstdio>
#include <cuda_runtime.h>
template <int ILP>
__global__ void chain(double* out, int iters, double a, double b, long long* cyc) {
  double acc[ILP];
#pragma unroll
  for (int k = 0; k < ILP; ++k) acc[k] = threadIdx.x + k;
  long long t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int k = 0; k < ILP; ++k) acc[k] = fma(acc[k], a, b);
#pragma unroll
    for (int k = 0; k < ILP; ++k) acc[k] = fma(acc[k], a, b);
#pragma unroll
    for (int k = 0; k < ILP; ++k) acc[k] = fma(acc[k], a, b);
#pragma unroll
    for (int k = 0; k < ILP; ++k) acc[k] = fma(acc[k], a, b);
  }
  long long t1 = clock64();
  double r = 0; 
#pragma unroll
  for (int k = 0; k < ILP; ++k) r += acc[k];
  if (r == 1.2345) out[0] = r;
  if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}
__global__ void rcpchain(double* out, int iters, double a, long long* cyc) {
  double x = a + threadIdx.x;
  long long t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < iters; ++i) {
    double y; asm volatile("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x)); x = y;
    asm volatile("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x)); x = y;
    asm volatile("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x)); x = y;
    asm volatile("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x)); x = y;
  }
  long long t1 = clock64();
  if (x == 1.2345) out[0] = x;
  if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}
__global__ void ldschain(double* out, int iters, long long* cyc) {
  __shared__ int idx[1024];
  for (int i = threadIdx.x; i < 1024; i += blockDim.x) idx[i] = (i * 33 + 7) & 1023;
  __syncthreads();
  int p = threadIdx.x;
  long long t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < iters; ++i) { p = idx[p]; p = idx[p]; p = idx[p]; p = idx[p]; }
  long long t1 = clock64();
  if (p == 12345) out[0] = p;
  if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}
template <int ILP> void run(int warps, double* out, long long* cyc) {
  int iters = 4096; long long h;
  chain<ILP><<<1, 32 * warps>>>(out, iters, 1.0000001, 1e-9, cyc);
  cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
  printf("DFMA ILP=%d warps/SM=%2d: %.2f cycles per DFMA-round (=> %.2f cyc/instr/warp, SM rate %.1f thread-DFMA/cycle)\n", ILP, warps,
         (double)h / (iters * 4), (double)h / (iters * 4.0 * ILP), 32.0 * warps * ILP * iters * 4 / (double)h);
}
int main() {
  double* out; long long* cyc; cudaMalloc(&out, 8); cudaMalloc(&cyc, 8);
  for (int w : {1, 2, 4, 8, 16}) { run<1>(w, out, cyc); }
  for (int w : {1, 4, 8, 16}) { run<2>(w, out, cyc); run<4>(w, out, cyc); run<8>(w, out, cyc); }
  long long h; int iters = 4096;
  rcpchain<<<1, 32>>>(out, iters, 3.0, cyc); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
  printf("MUFU.RCP64H dependent latency: %.2f cycles\n", (double)h / (iters * 4));
  ldschain<<<1, 32>>>(out, iters, cyc); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
  printf("LDS dependent latency: %.2f cycles\n", (double)h / (iters * 4));
  printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
  return 0;
}
