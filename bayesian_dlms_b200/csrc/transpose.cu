// transpose.cu -- out[c][r] = in[r][c] for fp64 matrices (layout conversion between the
// series-major order of the reference's Vector[KfState] and the device-native
// time-major SoA).  Classic 32x32 shared-memory tile, +1 padding, coalesced both ways;
// pure HBM traffic: 16 bytes per element.
#include "common.cuh"
#include "launch.h"

namespace bdlm {
namespace {

__global__ void __launch_bounds__(256)
transpose_kernel(const double *__restrict__ in, double *__restrict__ out, int64_t rows,
                 int64_t cols, int64_t out_pitch) {
  __shared__ double tile[32][33];
  const int64_t c0 = (int64_t)blockIdx.x * 32, r0 = (int64_t)blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
#pragma unroll
  for (int k = 0; k < 32; k += 8) {
    const int64_t r = r0 + ty + k, c = c0 + tx;
    if (r < rows && c < cols) tile[ty + k][tx] = ld_stream(in + r * cols + c);
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < 32; k += 8) {
    const int64_t c = c0 + ty + k, r = r0 + tx;
    if (r < rows && c < cols) st_stream(out + c * out_pitch + r, tile[tx][ty + k]);
  }
}

// Per-series time grids (Data.time differs by series, Dlm.scala:94): dt into every observation,
// computed exactly as the host does for a shared grid -- prev = min(times) - 1.0
// (KalmanFilter.initialiseState, KalmanFilter.scala:112-118) or the saved state's time, then
// dt[t] = times[t] - prev, prev = times[t].  One thread per series.
__global__ void __launch_bounds__(128)
dt_kernel(const double *__restrict__ times, int64_t tsb, int64_t tsr, double *__restrict__ dt,
          int64_t dsb, int64_t dsr, int64_t B, int T, int has_init, double t_init) {
  const int64_t b = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (b >= B) return;
  const double *tp = times + b * tsb;
  double prev;
  if (has_init) {
    prev = t_init;
  } else {
    double tmin = tp[0];
    for (int t = 1; t < T; ++t) tmin = fmin(tmin, tp[(int64_t)t * tsr]);
    prev = tmin - 1.0;
  }
  double *dp = dt + b * dsb;
  for (int t = 0; t < T; ++t) {
    const double cur = tp[(int64_t)t * tsr];
    dp[(int64_t)t * dsr] = cur - prev;
    prev = cur;
  }
}

}  // namespace

cudaError_t launch_dt_from_times(const double *times, int64_t tsb, int64_t tsr, double *dt,
                                 int64_t dsb, int64_t dsr, int64_t B, int T, const double *t_init,
                                 cudaStream_t stream) {
  if (B <= 0 || T <= 0) return cudaSuccess;
  dt_kernel<<<(unsigned)((B + 127) / 128), 128, 0, stream>>>(times, tsb, tsr, dt, dsb, dsr, B, T,
                                                             t_init != nullptr, t_init ? *t_init : 0.0);
  return cudaGetLastError();
}

cudaError_t launch_transpose(const double *in, double *out, int64_t rows, int64_t cols,
                             cudaStream_t stream) {
  if (rows <= 0 || cols <= 0) return cudaSuccess;
  // grid.y is limited to 65535 blocks: loop over bands of <= 65535 * 32 rows (a band of `in`
  // rows is a contiguous block of `in` and a column range of `out`, reachable by pointer offset
  // with the full pitches kept as kernel arguments)
  const int64_t gx = (cols + 31) / 32;
  if (gx > 2147483647LL) return cudaErrorInvalidValue;
  const int64_t band = (int64_t)65535 * 32;
  for (int64_t r0 = 0; r0 < rows; r0 += band) {
    const int64_t rcnt = rows - r0 < band ? rows - r0 : band;
    transpose_kernel<<<dim3((unsigned)gx, (unsigned)((rcnt + 31) / 32)), 256, 0, stream>>>(
        in + r0 * cols, out + r0, rcnt, cols, rows);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
  }
  return cudaSuccess;
}

}  // namespace bdlm
