// rng.cuh -- counter-based normals for the FFBS kernels' on-device RNG mode (z == NULL).
//
// The reference draws rand.gaussian(0, 1) inside MultivariateGaussianSvd.draw / SvdSampler.rnorm
// (MultivariateGaussianSvd.scala:20, SvdSampler.scala:94-102) from Breeze's Mersenne twister;
// bit-for-bit parity therefore needs INJECTED normals (the `z` argument).  Without them the
// kernels generate their own: Philox4x32-10 (cuRAND device API) keyed by the context seed, one
// subsequence per series / chain, the offset addressing (sweep, row, component) -- reproducible
// for a given (seed, sweep) whatever the launch geometry, kernel variant or memory layout.
#pragma once
#include <curand_kernel.h>

namespace bdlm {

struct RngKey {
  unsigned long long seed, sweep;
};

// N(0, 1) value consumed when drawing component k of theta[row] of series b.
__device__ __forceinline__ double philox_normal(const RngKey &key, long long b, int rows, int row,
                                                int n, int k) {
  curandStatePhilox4_32_10_t st;
  const unsigned long long idx = ((key.sweep * (unsigned long long)rows + (unsigned long long)row) *
                                  (unsigned long long)n + (unsigned long long)k);
  curand_init(key.seed, (unsigned long long)b, 4ULL * idx, &st);
  return curand_normal_double(&st);
}

}  // namespace bdlm
