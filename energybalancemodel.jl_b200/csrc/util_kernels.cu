// util_kernels.cu -- layout helpers and the FP64 roof microbenchmark.
#include "ebm_internal.cuh"
#include <algorithm>

namespace {

// [rows][cols] -> [cols][rows], 32x32 tiles through shared memory (coalesced on both sides).  The larger tile count
// goes to gridDim.x (2^31 - 1 blocks), the smaller to gridDim.y (65535): either dimension may be the member axis.
__global__ void transpose_kernel(const double* __restrict__ src, double* __restrict__ dst, long long rows, long long cols,
                                 int rows_on_x) {
  __shared__ double tile[32][33];
  const long long r0 = (long long)(rows_on_x ? blockIdx.x : blockIdx.y) * 32;
  const long long c0 = (long long)(rows_on_x ? blockIdx.y : blockIdx.x) * 32;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const long long r = r0 + i, c = c0 + threadIdx.x;
    if (r < rows && c < cols) tile[i][threadIdx.x] = src[r * cols + c];
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const long long c = c0 + i, r = r0 + threadIdx.x;
    if (r < rows && c < cols) dst[c * rows + r] = tile[threadIdx.x][i];
  }
}

// member reordering: forward  dst[c*rows + r] = src[idx[r]*cols + c];  inverse  dst[idx[r]*cols + c] = src[c*rows + r]
__global__ void gather_transpose_kernel(const double* __restrict__ src, double* __restrict__ dst, long long rows, long long cols,
                                        const long long* __restrict__ idx, int inverse) {
  const long long n = rows * cols;
  for (long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x; q < n; q += (long long)gridDim.x * blockDim.x) {
    const long long c = q / rows, r = q % rows;     // consecutive threads: consecutive members of one column
    if (inverse) dst[idx[r] * cols + c] = src[q];
    else dst[q] = src[idx[r] * cols + c];
  }
}

// dst[idx[r]][:] = src[r][:]  (rows of `rowlen` doubles)
__global__ void scatter_rows_kernel(const double* __restrict__ src, double* __restrict__ dst, long long nrows, long long rowlen,
                                    const long long* __restrict__ idx) {
  const long long n = nrows * rowlen;
  for (long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x; q < n; q += (long long)gridDim.x * blockDim.x) {
    const long long r = q / rowlen, c = q % rowlen;
    dst[idx[r] * rowlen + c] = src[q];
  }
}

__global__ void fill_kernel(double* dst, long long n, double v) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    dst[i] = v;
}

// FP64 FMA roof: 8 independent dependent-chains per thread, 8 warps per SMSP-quad worth of TLP.
__global__ void __launch_bounds__(256) fp64_peak_kernel(double* out, int iters, double a, double b, long long* cycles) {
  double acc0 = threadIdx.x, acc1 = acc0 + 1, acc2 = acc0 + 2, acc3 = acc0 + 3;
  double acc4 = acc0 + 4, acc5 = acc0 + 5, acc6 = acc0 + 6, acc7 = acc0 + 7;
  const long long t0 = clock64();
#pragma unroll 4
  for (int i = 0; i < iters; ++i) {
    acc0 = fma(acc0, a, b); acc1 = fma(acc1, a, b); acc2 = fma(acc2, a, b); acc3 = fma(acc3, a, b);
    acc4 = fma(acc4, a, b); acc5 = fma(acc5, a, b); acc6 = fma(acc6, a, b); acc7 = fma(acc7, a, b);
  }
  const long long t1 = clock64();
  const double r = ((acc0 + acc1) + (acc2 + acc3)) + ((acc4 + acc5) + (acc6 + acc7));
  if (r == 123.456) out[0] = r;  // keep the chains alive
  if (threadIdx.x == 0 && blockIdx.x == 0) cycles[0] = t1 - t0;
}

}  // namespace

int ebm_launch_transpose(const double* src, double* dst, long long rows, long long cols, cudaStream_t stream) {
  if (rows <= 0 || cols <= 0) return EBM_OK;
  const long long tr = (rows + 31) / 32, tc = (cols + 31) / 32;
  const int rows_on_x = tr >= tc;
  const long long gx = rows_on_x ? tr : tc, gy = rows_on_x ? tc : tr;
  if (gx > 0x7fffffffLL || gy > 65535) { ebm_set_error("transpose: %lld x %lld is too large", rows, cols); return EBM_ERR_INVALID; }
  transpose_kernel<<<dim3((unsigned)gx, (unsigned)gy), dim3(32, 8), 0, stream>>>(src, dst, rows, cols, rows_on_x);
  EBM_CUDA_TRY(cudaGetLastError());
  ebm_count_launch();
  return EBM_OK;
}

int ebm_launch_gather_transpose(const double* src, double* dst, long long rows, long long cols, const long long* idx,
                                int inverse, cudaStream_t stream) {
  if (rows <= 0 || cols <= 0) return EBM_OK;
  gather_transpose_kernel<<<1184, 256, 0, stream>>>(src, dst, rows, cols, idx, inverse);
  EBM_CUDA_TRY(cudaGetLastError());
  ebm_count_launch();
  return EBM_OK;
}

int ebm_launch_scatter_rows(const double* src, double* dst, long long nrows, long long rowlen, const long long* idx,
                            cudaStream_t stream) {
  if (nrows <= 0 || rowlen <= 0) return EBM_OK;
  scatter_rows_kernel<<<1184, 256, 0, stream>>>(src, dst, nrows, rowlen, idx);
  EBM_CUDA_TRY(cudaGetLastError());
  ebm_count_launch();
  return EBM_OK;
}

int ebm_launch_fill(double* dst, long long n, double v, cudaStream_t stream) {
  if (n <= 0) return EBM_OK;
  fill_kernel<<<1184, 256, 0, stream>>>(dst, n, v);
  EBM_CUDA_TRY(cudaGetLastError());
  ebm_count_launch();
  return EBM_OK;
}

// head[k] = par[k][0]; *differs |= 1 when any member's parameter k is not bit-identical to member 0's
__global__ void par_uniform_kernel(const double* __restrict__ par, int npar, long long nmem, double* __restrict__ head,
                                   int* __restrict__ differs) {
  bool diff = false;
  for (int k = 0; k < npar; ++k) {
    const double* row = par + (long long)k * nmem;
    const long long ref = __double_as_longlong(row[0]);
    if (blockIdx.x == 0 && threadIdx.x == 0) head[k] = row[0];
    for (long long m = (long long)blockIdx.x * blockDim.x + threadIdx.x; m < nmem; m += (long long)gridDim.x * blockDim.x)
      diff = diff || (__double_as_longlong(row[m]) != ref);
  }
  if (__syncthreads_or(diff) && threadIdx.x == 0) atomicOr(differs, 1);
}

int ebm_launch_par_uniform(const double* par, int npar, long long nmem, double* head, int* differs, cudaStream_t stream) {
  EBM_CUDA_TRY(cudaMemsetAsync(differs, 0, sizeof(int), stream));
  const int blocks = (int)std::min<long long>(592, (nmem + 255) / 256);
  par_uniform_kernel<<<blocks, 256, 0, stream>>>(par, npar, nmem, head, differs);
  EBM_CUDA_TRY(cudaGetLastError());
  ebm_count_launch();
  return EBM_OK;
}

int ebm_run_fp64_peak(int device, double* tflops, double* mhz) {
  if (device >= 0) EBM_CUDA_TRY(cudaSetDevice(device));
  cudaDeviceProp prop;
  int dev = 0;
  EBM_CUDA_TRY(cudaGetDevice(&dev));
  EBM_CUDA_TRY(cudaGetDeviceProperties(&prop, dev));
  double* out = nullptr; long long* cyc = nullptr;
  EBM_CUDA_TRY(cudaMalloc(&out, sizeof(double)));
  EBM_CUDA_TRY(cudaMalloc(&cyc, sizeof(long long)));
  const int blocks = prop.multiProcessorCount * 8, threads = 256, iters = 1 << 16;
  cudaEvent_t e0, e1;
  EBM_CUDA_TRY(cudaEventCreate(&e0));
  EBM_CUDA_TRY(cudaEventCreate(&e1));
  double best = 0.0, best_mhz = 0.0;
  for (int rep = 0; rep < 5; ++rep) {
    EBM_CUDA_TRY(cudaEventRecord(e0));
    fp64_peak_kernel<<<blocks, threads>>>(out, iters, 1.0000001, 1e-9, cyc);
    EBM_CUDA_TRY(cudaEventRecord(e1));
    EBM_CUDA_TRY(cudaEventSynchronize(e1));
    ebm_count_launch();
    float ms = 0.f;
    EBM_CUDA_TRY(cudaEventElapsedTime(&ms, e0, e1));
    const double flop = 2.0 * 8.0 * (double)iters * threads * (double)blocks;
    const double tf = flop / (ms * 1e-3) / 1e12;
    if (rep > 0 && tf > best) {
      best = tf;
      long long hc = 0;
      EBM_CUDA_TRY(cudaMemcpy(&hc, cyc, sizeof(hc), cudaMemcpyDeviceToHost));
      // blocks run in waves; a single block's cycle count vs its share of wall time gives the clock
      const double waves = (double)blocks / (prop.multiProcessorCount * (2048 / threads));
      best_mhz = (double)hc * (waves < 1.0 ? 1.0 : waves) / (ms * 1e-3) / 1e6;
    }
  }
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  cudaFree(out); cudaFree(cyc);
  if (tflops) *tflops = best;
  if (mhz) *mhz = best_mhz;
  return EBM_OK;
}
