// peak.cu -- FP64 pipe microbenchmark (the roofline denominator for the FP64-bound kernels;
// MEASURED_PEAKS.json carries HBM and bf16 numbers only).  Not part of the numerical path:
// this is the one place where explicit FMAs are issued on purpose.
#include "common.cuh"
#include "launch.h"

namespace bdlm {
namespace {

__global__ void __launch_bounds__(256) dfma_kernel(double *out, int iters, double a, double b) {
  double x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5,
         x6 = x0 + 6, x7 = x0 + 7;
  for (int i = 0; i < iters; ++i) {
    x0 = __fma_rn(x0, a, b); x1 = __fma_rn(x1, a, b); x2 = __fma_rn(x2, a, b);
    x3 = __fma_rn(x3, a, b); x4 = __fma_rn(x4, a, b); x5 = __fma_rn(x5, a, b);
    x6 = __fma_rn(x6, a, b); x7 = __fma_rn(x7, a, b);
  }
  const double s = ((x0 + x1) + (x2 + x3)) + ((x4 + x5) + (x6 + x7));
  if (s == 123.456) out[0] = s;  // keep the chains alive
}

}  // namespace

// Returns DFMA TFLOP/s (2 flops per DFMA); best of 3.
cudaError_t measure_fp64_peak(cudaStream_t stream, double *scratch, double *tflops) {
  int dev = 0, sms = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int blocks = sms * 8, threads = 256, iters = 1 << 14;
  cudaEvent_t e0, e1;
  cudaError_t e = cudaEventCreate(&e0);
  if (e != cudaSuccess) return e;
  e = cudaEventCreate(&e1);
  if (e != cudaSuccess) return e;
  double best = 0.0;
  for (int rep = 0; rep < 4; ++rep) {
    cudaEventRecord(e0, stream);
    dfma_kernel<<<blocks, threads, 0, stream>>>(scratch, iters, 0.999999, 1e-6);
    cudaEventRecord(e1, stream);
    e = cudaEventSynchronize(e1);
    if (e != cudaSuccess) break;
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    const double fl = 2.0 * 8.0 * iters * (double)blocks * threads;
    if (rep > 0 && ms > 0.f) best = fl / (ms * 1e-3) / 1e12 > best ? fl / (ms * 1e-3) / 1e12 : best;
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  *tflops = best;
  return e == cudaSuccess ? cudaGetLastError() : e;
}

}  // namespace bdlm
