// gibbs_draw.cu -- on-device conjugate parameter draws, so that a
// whole Gibbs sweep -- FFBS + sufficient statistics (kf_warp.cu / kf_group.cu) + the draw of
// V and W -- stays on the GPU without a host round trip per iteration.
//
//   invgamma_kernel   GibbsSampling.sampleObservationMatrix (Gibbs.scala:41-49) and
//                     sampleSystemMatrix (:72-77): per diagonal element
//                     shape = prior.shape + count / 2, rate = prior.scale + ss / 2,
//                     draw = InverseGamma(shape, rate).draw = 1 / Gamma(shape, 1 / rate).draw
//                     (InverseGamma.scala:14); one thread per matrix element of a chain.
//   invwishart_kernel GibbsWishart.sampleSystemMatrix (GibbsWishart.scala:16-35) +
//                     InverseWishart.draw (InverseWishart.scala:17-25) with the Bartlett factor of
//                     Wishart.scala:34-43; one warp per chain, matrices in shared memory.
//
// Random numbers: counter-based Philox4x32-10 (cuRAND device API), one subsequence per (chain,
// element), offset by the sweep number -- reproducible for a given seed, independent of the
// launch geometry.  The Gamma variate is Marsaglia-Tsang (what Breeze's Gamma.draw implements).
// For bit-exact parity tests the variates can be INJECTED instead (standard Gamma(shape, 1)
// values, Bartlett factors): the deterministic arithmetic then equals oracle/bdlm_oracle.c.
#include <curand_kernel.h>

#include "common.cuh"
#include "launch.h"
#include "warp_linalg.cuh"

namespace bdlm {

namespace {

// The FFBS kernels' normals (rng.cuh) use the caller's seed as the Philox key; the conjugate
// draws use a different key so the two families of variates never share a counter block.
__device__ __forceinline__ unsigned long long draw_key(unsigned long long seed) {
  return seed ^ 0x9E3779B97F4A7C15ULL;
}

// Standard Gamma(shape, 1), Marsaglia & Tsang (2000); shape < 1 through the U^(1/shape) boost.
__device__ double gamma_mt(curandStatePhilox4_32_10_t *rng, double shape) {
  double boost = 1.0;
  if (shape < 1.0) {
    boost = pow(curand_uniform_double(rng), 1.0 / shape);
    shape += 1.0;
  }
  const double d = shape - 1.0 / 3.0;
  const double c = 1.0 / sqrt(9.0 * d);
  for (int it = 0; it < 64; ++it) {
    double x, v;
    do {
      x = curand_normal_double(rng);
      v = 1.0 + c * x;
    } while (v <= 0.0);
    v = v * v * v;
    const double x2 = x * x;
    const double u = curand_uniform_double(rng);
    if (u < 1.0 - 0.0331 * (x2 * x2) || log(u) < 0.5 * x2 + d * (1.0 - v + log(v)))
      return boost * (d * v);
  }
  return boost * d;  // unreachable in practice (acceptance > 95 % per round)
}

__global__ void __launch_bounds__(256)
invgamma_kernel(const GibbsDrawArgs a) {
  const int pp = a.p * a.p, nn = a.n * a.n;
  const int per_chain = pp + (a.wishart ? 0 : nn);
  const int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (idx >= a.B * per_chain) return;
  const int64_t b = idx / per_chain;
  int e = (int)(idx - b * per_chain);
  const bool is_v = e < pp;
  if (!is_v) e -= pp;
  const int dim = is_v ? a.p : a.n;
  const int i = e % dim, j = e / dim;
  const View &out = is_v ? a.V : a.W;
  if (!out.ptr) return;
  double val = 0.0;
  if (i == j) {
    double count, ss, prior_shape, prior_scale;
    if (is_v) {
      count = a.stats.ny.ptr[b * a.stats.ny.sb + i * a.stats.ny.sk];
      ss = a.stats.ssy.ptr[b * a.stats.ssy.sb + i * a.stats.ssy.sk];
      prior_shape = a.v_shape; prior_scale = a.v_scale;
    } else {
      count = (double)a.T;  // theta.size - 1 (Gibbs.scala:72)
      ss = a.stats.ssw.ptr[b * a.stats.ssw.sb + i * a.stats.ssw.sk];
      prior_shape = a.w_shape; prior_scale = a.w_scale;
    }
    const double shape = prior_shape + count * 0.5;
    const double rate = prior_scale + ss * 0.5;
    const View &inj = is_v ? a.gv : a.gw;
    double g;
    if (inj.ptr) {
      g = inj.ptr[b * inj.sb + i * inj.sk];
    } else {
      curandStatePhilox4_32_10_t rng;
      curand_init(draw_key(a.seed),
                  (unsigned long long)((a.base + b) * (a.p + a.n) + (is_v ? i : a.p + i)),
                  a.sweep * 1024ULL, &rng);
      g = gamma_mt(&rng, shape);
    }
    val = 1.0 / ((1.0 / rate) * g);
    const View &so = is_v ? a.v_shape_rate : a.w_shape_rate;  // optional [2*dim]: shapes | rates
    if (so.ptr) {
      so.ptr[b * so.sb + i * so.sk] = shape;
      so.ptr[b * so.sb + (dim + i) * so.sk] = rate;
    }
  }
  out.ptr[b * out.sb + e * out.sk] = val;
}

// dpotrf 'L' with the strict upper triangle zeroed (oracle chol_lower).  S -> L, n x n.
__device__ __forceinline__ int w_chol_lower(int lane, int n, const double *S, double *L) {
  int st = 0;
  for (int k = lane; k < n * n; k += 32) L[k] = S[k];
  __syncwarp();
  for (int j = 0; j < n; ++j) {
    double d = L[j + j * n];
    for (int k = 0; k < j; ++k) d = d - L[j + k * n] * L[j + k * n];
    if (!(d > 0.0)) st = BDLM_ST_NOTPD;
    d = sqrt(d);
    __syncwarp();
    if (lane == 0) L[j + j * n] = d;
    for (int i = j + 1 + lane; i < n; i += 32) {  // rows strided over the lanes (n up to 48)
      double v = L[i + j * n];
      for (int k = 0; k < j; ++k) v = v - L[i + k * n] * L[j + k * n];
      L[i + j * n] = v / d;
    }
    for (int i = lane; i < j; i += 32) L[i + j * n] = 0.0;
    __syncwarp();
  }
  return st;
}

// out = inv(M) as dgesv against the identity (oracle inv_lu); M destroyed.
__device__ __forceinline__ int w_inv(int lane, int n, double *M, double *out) {
  for (int k = lane; k < n * n; k += 32) out[k] = (k % (n + 1) == 0) ? 1.0 : 0.0;
  __syncwarp();
  return w_lu_solve(lane, n, M, n, out);
}

__global__ void __launch_bounds__(128)
invwishart_kernel(const GibbsDrawArgs a) {
  extern __shared__ double smem[];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int64_t b = blockIdx.x * (int64_t)(blockDim.x >> 5) + wib;
  if (b >= a.B) return;  // whole warp exits together
  const int n = a.n, nn = n * n;
  double *sc = smem + (size_t)wib * 6 * nn, *isc = sc + nn, *l = isc + nn, *il = l + nn,
         *A = il + nn, *ia = A + nn;
  // scale = psi + squaredSum (GibbsWishart.scala:31-32)
  for (int k = lane; k < nn; k += 32)
    sc[k] = a.psi[k] + a.stats.scatter.ptr[b * a.stats.scatter.sb + k * a.stats.scatter.sk];
  // Bartlett factor of Wishart(dof, .) (Wishart.scala:34-43), dof = nu + T
  const double dof = a.w_nu + (double)a.T;
  for (int k = lane; k < nn; k += 32) {
    double v;
    if (a.bart.ptr) {
      v = a.bart.ptr[b * a.bart.sb + k * a.bart.sk];
    } else {
      const int i = k % n, j = k / n;
      v = 0.0;
      if (i >= j) {
        curandStatePhilox4_32_10_t rng;
        curand_init(draw_key(a.seed) ^ 0x5DEECE66DULL, (unsigned long long)((a.base + b) * nn + k),
                    a.sweep * 1024ULL, &rng);
        // ChiSquared(k) = Gamma(k / 2, 2)
        v = (i == j) ? sqrt(2.0 * gamma_mt(&rng, 0.5 * (dof - i))) : curand_normal_double(&rng);
      }
    }
    A[k] = v;
  }
  __syncwarp();
  int st = 0;
  st |= w_inv(lane, n, sc, isc);        // inv(scale)            (sc destroyed)
  st |= w_chol_lower(lane, n, isc, l);  // l = cholesky(inv(psi)) (InverseWishart.scala:17)
  st |= w_inv(lane, n, l, il);          // invl                   (l destroyed)
  st |= w_inv(lane, n, A, ia);          // inva                   (A destroyed)
  // invl.t * inva.t * inva * invl, left to right (InverseWishart.scala:24)
  w_mm(lane, n, n, n, il, n, true, ia, n, true, sc, n);
  w_mm(lane, n, n, n, sc, n, false, ia, n, false, isc, n);
  w_mm(lane, n, n, n, isc, n, false, il, n, false, l, n);
  for (int k = lane; k < nn; k += 32) a.W.ptr[b * a.W.sb + k * a.W.sk] = l[k];
  if (a.status && st && lane == 0) atomicOr(a.status + b, st);
}

}  // namespace

cudaError_t launch_gibbs_draw(const GibbsDrawArgs &a, cudaStream_t stream, int64_t *launches) {
  if (a.B == 0) return cudaSuccess;
  const int per_chain = a.p * a.p + (a.wishart ? 0 : a.n * a.n);
  const int64_t total = a.B * per_chain;
  invgamma_kernel<<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(a);
  ++*launches;
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess || !a.wishart) return e;
  const int warps = a.n <= 16 ? 4 : 1;
  const size_t smem = sizeof(double) * 6 * a.n * a.n * warps;
  e = cudaFuncSetAttribute(invwishart_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  invwishart_kernel<<<(unsigned)((a.B + warps - 1) / warps), warps * 32, smem, stream>>>(a);
  ++*launches;
  return cudaGetLastError();
}

}  // namespace bdlm
