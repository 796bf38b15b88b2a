// Microbenchmark: does the WRITE PATTERN of the fused filter+smoother kernel (every warp
// advancing 20 separate output streams, 256 contiguous bytes per stream per step, streams
// of one field-row 8*B bytes apart) cap HBM throughput below a plain streaming write?
//   pattern 0: time-major SoA  [rows][k][B]   (what kf_small.cu writes today)
//   pattern 1: tile-major      [B/32][rows][k][32]  (one contiguous 20*256 B record per warp-step)
//   pattern 2: per-field tile-major [k-field][B/32][rows][kf][32]
// nvcc -O3 -gencode arch=compute_100a,code=sm_100a stream_pattern.cu -o stream_pattern
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
__global__ void __launch_bounds__(128, 5) k(double *out, long B, int rows, int K, int pattern, int spin) {
  const long b = blockIdx.x * (long)blockDim.x + threadIdx.x;
  if (b >= B) return;
  double v = (double)b;
  for (int r = 0; r < rows; ++r) {
    for (int s = 0; s < spin; ++s) v = v * 1.0000001 + 1e-9;  // stand-in for the recursion latency
    for (int kk = 0; kk < K; ++kk) {
      long idx;
      if (pattern == 0) idx = ((long)r * K + kk) * B + b;
      else idx = (((b >> 5) * rows + r) * K + kk) * 32 + (b & 31);
      __stcs(out + idx, v + kk);
    }
  }
}
int main(int argc, char **argv) {
  const int K = 20, rows = 1000, spin = argc > 1 ? atoi(argv[1]) : 40;
  const long B = 94720L * 4;
  double *out;
  cudaMalloc(&out, sizeof(double) * B * rows * K);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int pattern = 0; pattern < 2; ++pattern) {
    float best = 1e9;
    for (int it = 0; it < 4; ++it) {
      cudaEventRecord(e0);
      k<<<(unsigned)((B + 127) / 128), 128>>>(out, B, rows, K, pattern, spin);
      cudaEventRecord(e1); cudaEventSynchronize(e1);
      float ms; cudaEventElapsedTime(&ms, e0, e1); if (it && ms < best) best = ms;
    }
    printf("pattern %d spin %d: %.2f ms  %.0f GB/s  (%s)\n", pattern, spin, best,
           (double)B * rows * K * 8 / best / 1e6, cudaGetErrorString(cudaGetLastError()));
  }
  return 0;
}
