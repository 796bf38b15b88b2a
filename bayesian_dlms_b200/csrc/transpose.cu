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
                 int64_t cols) {
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
    if (r < rows && c < cols) st_stream(out + c * rows + r, tile[tx][ty + k]);
  }
}

}  // namespace

cudaError_t launch_transpose(const double *in, double *out, int64_t rows, int64_t cols,
                             cudaStream_t stream) {
  if (rows <= 0 || cols <= 0) return cudaSuccess;
  const int64_t gx = (cols + 31) / 32, gy = (rows + 31) / 32;
  if (gy > 65535) {  // put the long dimension on x
    // out[c][r] = in[r][c]  <=>  transpose of the (cols x rows) view is not expressible
    // by swapping arguments, so tile over y in chunks instead.
    for (int64_t y0 = 0; y0 < gy; y0 += 65535) {
      const int64_t ny = (gy - y0 < 65535) ? gy - y0 : 65535;
      const int64_t rbeg = y0 * 32, rcnt = (rows - rbeg < ny * 32) ? rows - rbeg : ny * 32;
      // rows [rbeg, rbeg + rcnt) of `in` -> columns [rbeg, ...) of `out`: needs the full
      // output pitch, so use the strided kernel through pointer offsets.
      (void)rcnt;
      return cudaErrorInvalidValue;  // not needed by any caller (rows*k <= 2M)
    }
  }
  transpose_kernel<<<dim3((unsigned)gx, (unsigned)gy), 256, 0, stream>>>(in, out, rows, cols);
  return cudaGetLastError();
}

}  // namespace bdlm
