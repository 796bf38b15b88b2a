// small_steps.cuh -- register-level building blocks shared by the thread-per-series kernels
// (kf_small.cu) and the parallel-in-time scan kernels (scan.cu): unrolled small matrix
// products, the dgesv restatement and the Kalman / RTS steps for p = 1.
//
// Arithmetic mirrors oracle/bdlm_oracle.c operation for operation (see common.cuh);
// reference citations: KalmanFilter.scala:64-107,273-321 and Smoothing.scala:31-64.
#pragma once
#include "common.cuh"

namespace bdlm {
namespace small {

// out(AR x BC) = A(AR x AC) * B(AC x BC), column-major, TA/TB = operand stored
// transposed.  Products summed in increasing inner index, first product initialises.
template <int AR, int AC, int BC, bool TA, bool TB>
__host__ __device__ __forceinline__ void smm(const double *A, const double *B, double *out) {
#pragma unroll
  for (int j = 0; j < BC; ++j)
#pragma unroll
    for (int i = 0; i < AR; ++i) {
      double acc = 0.0;
#pragma unroll
      for (int k = 0; k < AC; ++k) {
        const double a = TA ? A[k + i * AC] : A[i + k * AR];
        const double b = TB ? B[j + k * BC] : B[k + j * AC];
        const double prod = a * b;
        acc = (k == 0) ? prod : acc + prod;
      }
      out[i + j * AR] = acc;
    }
}

// dgesv restatement (see oracle lu_solve): A is N x N (destroyed), Bm is N x NR.
template <int N, int NR>
__host__ __device__ __forceinline__ int lu_solve(double (&A)[N * N], double (&Bm)[N * NR]) {
  int st = 0;
#pragma unroll
  for (int j = 0; j < N; ++j) {
    int jp = j;
    double best = fabs(A[j + j * N]);
#pragma unroll
    for (int i = j + 1; i < N; ++i) {
      const double v = fabs(A[i + j * N]);
      if (v > best) { best = v; jp = i; }
    }
    double pv = A[j + j * N];
#pragma unroll
    for (int i = j + 1; i < N; ++i)
      if (jp == i) pv = A[i + j * N];
    if (pv != 0.0) {
#pragma unroll
      for (int i = j + 1; i < N; ++i)
        if (jp == i) {
#pragma unroll
          for (int c = 0; c < N; ++c) {
            const double t = A[j + c * N]; A[j + c * N] = A[i + c * N]; A[i + c * N] = t;
          }
#pragma unroll
          for (int c = 0; c < NR; ++c) {
            const double t = Bm[j + c * N]; Bm[j + c * N] = Bm[i + c * N]; Bm[i + c * N] = t;
          }
        }
      const double r = 1.0 / A[j + j * N];
#pragma unroll
      for (int i = j + 1; i < N; ++i) A[i + j * N] = A[i + j * N] * r;
    } else {
      st = BDLM_ST_SINGULAR;
    }
#pragma unroll
    for (int c = j + 1; c < N; ++c)
#pragma unroll
      for (int i = j + 1; i < N; ++i)
        A[i + c * N] = A[i + c * N] - A[i + j * N] * A[j + c * N];
  }
#pragma unroll
  for (int c = 0; c < NR; ++c) {
#pragma unroll
    for (int k = 0; k < N; ++k)
#pragma unroll
      for (int i = k + 1; i < N; ++i)
        Bm[i + c * N] = Bm[i + c * N] - Bm[k + c * N] * A[i + k * N];
#pragma unroll
    for (int k = N - 1; k >= 0; --k) {
      Bm[k + c * N] = Bm[k + c * N] / A[k + k * N];
#pragma unroll
      for (int i = 0; i < k; ++i)
        Bm[i + c * N] = Bm[i + c * N] - Bm[k + c * N] * A[i + k * N];
    }
  }
  return st;
}

// KalmanFilter.advState (KalmanFilter.scala:273-286)
template <int N, bool REG>
__host__ __device__ __forceinline__ void advance(const double *G, const double (&W)[N * N],
                                        double dt, const double (&m)[N],
                                        const double (&C)[N * N], double (&a)[N],
                                        double (&R)[N * N]) {
  if (!REG && dt == 0.0) {
#pragma unroll
    for (int i = 0; i < N; ++i) a[i] = m[i];
#pragma unroll
    for (int k = 0; k < N * N; ++k) R[k] = C[k];
    return;
  }
  double t1[N * N];
  smm<N, N, 1, false, false>(G, m, a);
  smm<N, N, N, false, false>(G, C, t1);
  smm<N, N, N, false, true>(t1, G, R);
#pragma unroll
  for (int k = 0; k < N * N; ++k) R[k] = R[k] + (REG ? W[k] : W[k] * dt);  // W * 1.0 == W
}

// oneStepPrediction (:311-321) + updateState (:64-94) for p = 1.
// RECIP (parallel-in-time scan only, 1e-9 contract): one reciprocal of Q instead of N divisions.
template <int N, bool RECIP = false>
__host__ __device__ __forceinline__ void update(const double *F, double V, double y,
                                       const double (&a)[N], const double (&R)[N * N],
                                       double &f, double &Q, double (&m)[N],
                                       double (&C)[N * N], int &st) {
  double fr[N];
  smm<1, N, 1, true, false>(F, a, &f);
  smm<1, N, N, true, false>(F, R, fr);
  smm<1, N, 1, false, false>(fr, F, &Q);
  Q = Q + V;
  if (isnan(y)) {  // all missing (:74-75)
#pragma unroll
    for (int i = 0; i < N; ++i) m[i] = a[i];
#pragma unroll
    for (int k = 0; k < N * N; ++k) C[k] = R[k];
    return;
  }
  // oneStepMissing (:44-53) on the full F, V repeats the same operations: reuse f, Q.
  const double e = y - f;
  double rhs[N], K[N], D[N * N], t1[N * N], t2[N], C2[N * N];
  smm<1, N, N, true, true>(F, R, rhs);  // F^T R^T
  if (Q == 0.0) st |= BDLM_ST_SINGULAR;
  if (RECIP) {
    const double rq = 1.0 / Q;
#pragma unroll
    for (int i = 0; i < N; ++i) K[i] = rhs[i] * rq;
  } else {
#pragma unroll
    for (int i = 0; i < N; ++i) K[i] = rhs[i] / Q;  // (Q^T \ (F^T R^T))^T  (:83)
  }
#pragma unroll
  for (int i = 0; i < N; ++i) m[i] = a[i] + K[i] * e;
  smm<N, 1, N, false, true>(K, F, D);  // K F^T
#pragma unroll
  for (int j = 0; j < N; ++j)
#pragma unroll
    for (int i = 0; i < N; ++i) D[i + j * N] = ((i == j) ? 1.0 : 0.0) - D[i + j * N];
  smm<N, N, N, false, false>(D, R, t1);
  smm<N, N, N, false, true>(t1, D, C);
#pragma unroll
  for (int i = 0; i < N; ++i) t2[i] = K[i] * V;
  smm<N, 1, N, false, true>(t2, K, C2);
#pragma unroll
  for (int k = 0; k < N * N; ++k) C[k] = C[k] + C2[k];
}

// Smoothing.smoothStep (Smoothing.scala:31-47)
template <int N>
__host__ __device__ __forceinline__ void rts_step(const double *G, const double (&m)[N],
                                         const double (&C)[N * N], const double (&a1)[N],
                                         const double (&R1)[N * N], bool textbook,
                                         double (&s)[N], double (&S)[N * N], int &st) {
  double rhs[N * N], At[N * N], Bg[N * N], d[N], t[N], Dm[N * N], t1[N * N], t2[N * N];
  smm<N, N, N, false, true>(G, C, rhs);  // G C^T
#pragma unroll
  for (int j = 0; j < N; ++j)
#pragma unroll
    for (int i = 0; i < N; ++i) At[i + j * N] = R1[j + i * N];
  st |= lu_solve<N, N>(At, rhs);
#pragma unroll
  for (int j = 0; j < N; ++j)
#pragma unroll
    for (int i = 0; i < N; ++i) Bg[i + j * N] = rhs[j + i * N];
#pragma unroll
  for (int i = 0; i < N; ++i) d[i] = s[i] - a1[i];
  smm<N, N, 1, false, false>(Bg, d, t);
#pragma unroll
  for (int k = 0; k < N * N; ++k) Dm[k] = R1[k] - S[k];
  smm<N, N, N, false, false>(Bg, Dm, t1);
  if (textbook) smm<N, N, N, false, true>(t1, Bg, t2);
  else smm<N, N, N, false, false>(t1, Bg, t2);  // Smoothing.scala:44 (no transpose)
#pragma unroll
  for (int i = 0; i < N; ++i) s[i] = m[i] + t[i];
#pragma unroll
  for (int k = 0; k < N * N; ++k) S[k] = C[k] - t2[k];
}


}  // namespace small
}  // namespace bdlm
