// pair_steps.cuh -- Kalman / RTS steps for n = 4, p = 1 with the state of ONE series split over
// TWO lanes (kf_pair.cu).
//
// Why: the thread-per-series kernel (kf_small.cu) needs 255 registers at n = 4, i.e. 8 warps per
// SM, and runs at about half of HBM with neither the FP64 pipe nor the issue slots full: it is
// latency bound.  Here lane h of a pair owns COLUMNS 2h, 2h+1 of every 4 x 4 matrix (8 doubles
// instead of 16), vectors are held in full by both lanes, and the few products that need a whole
// matrix gather it from the partner.  Every output element is still produced by ONE lane with
// the oracle's operation order (products summed in increasing inner index, no FMA), so results
// stay bit-identical to oracle/bdlm_oracle.c and to kf_small.cu -- pinned on the CPU by
// tests/test_pair_steps_cpu.py through a host build of these very functions
// (bdlm_debug_pair_filter_smooth_host, two host threads standing in for the two lanes).
//
// PX is the pair context: PX::h (0 | 1), gather_mat<SITE>(local 8 -> full 16, column-major),
// gather_vec<SITE>(own 2 -> full 4), or_int.  Local element (i, jj) of a matrix is global element
// (i, 2h + jj): loc[i + 4 * jj] == full[i + 4 * (2h + jj)].
//
// Reference citations: KalmanFilter.scala:64-107,273-321 and Smoothing.scala:31-64 (as small_steps.cuh).
#pragma once
#include "common.cuh"
#include "small_steps.cuh"

namespace bdlm {
namespace pairk {

constexpr int N = 4;    // state dimension served
constexpr int NL = 8;   // matrix elements per lane (two columns)

// gather sites (compile-time: the shared-memory transport rotates its buffers by site)
enum { kSiteAdvT1 = 0, kSiteFr = 1, kSiteR = 2, kSiteUpdT1 = 3, kSiteR1 = 4, kSiteX = 5, kSiteT = 6,
       kSiteRtsT1 = 7 };

// rows 2h, 2h+1 of a full column-major 4 x 4 matrix: out[jj + 2 * k] = M[(2h + jj) + 4 * k]
#pragma nv_exec_check_disable
template <class PX>
__host__ __device__ __forceinline__ void own_rows(const PX &px, const double *M, double (&out)[NL]) {
#pragma unroll
  for (int k = 0; k < N; ++k)
#pragma unroll
    for (int jj = 0; jj < 2; ++jj) out[jj + 2 * k] = px.h ? M[2 + jj + 4 * k] : M[jj + 4 * k];
}

// KalmanFilter.advState (KalmanFilter.scala:273-286): a = G m, R = G C G^T + W dt.
// G: full; Gr: its rows 2h, 2h+1 (own_rows); W, C, R: own columns.
#pragma nv_exec_check_disable
template <bool REG, class PX>
__host__ __device__ __forceinline__ void advance(PX &px, const double *G, const double (&Gr)[NL],
                                                 const double (&Wl)[NL], double dt,
                                                 const double (&m)[N], const double (&Cl)[NL],
                                                 double (&a)[N], double (&Rl)[NL]) {
  if (!REG && dt == 0.0) {
#pragma unroll
    for (int i = 0; i < N; ++i) a[i] = m[i];
#pragma unroll
    for (int k = 0; k < NL; ++k) Rl[k] = Cl[k];
    return;
  }
  double t1l[NL], t1[N * N];
  small::smm<N, N, 1, false, false>(G, m, a);
  small::smm<N, N, 2, false, false>(G, Cl, t1l);          // columns of G C
  px.template gather_mat<kSiteAdvT1>(t1l, t1);
  // R[i, j] = sum_k t1[i, k] G[j, k] for the own columns j = 2h + jj
#pragma unroll
  for (int jj = 0; jj < 2; ++jj)
#pragma unroll
    for (int i = 0; i < N; ++i) {
      double acc = 0.0;
#pragma unroll
      for (int k = 0; k < N; ++k) {
        const double prod = t1[i + 4 * k] * Gr[jj + 2 * k];
        acc = (k == 0) ? prod : acc + prod;
      }
      Rl[i + 4 * jj] = acc;
    }
#pragma unroll
  for (int k = 0; k < NL; ++k) Rl[k] = Rl[k] + (REG ? Wl[k] : Wl[k] * dt);
}

// oneStepPrediction (:311-321) + updateState (:64-94), p = 1.  f, Q, m: both lanes; C, R: own columns.
#pragma nv_exec_check_disable
template <class PX>
__host__ __device__ __forceinline__ void update(PX &px, const double *F, double V, double y,
                                                const double (&a)[N], const double (&Rl)[NL],
                                                double &f, double &Q, double (&m)[N],
                                                double (&Cl)[NL], int &st) {
  double frl[2], fr[N];
  small::smm<1, N, 1, true, false>(F, a, &f);
  small::smm<1, N, 2, true, false>(F, Rl, frl);            // (F^T R)[j], own columns
  px.template gather_vec<kSiteFr>(frl[0], frl[1], fr);
  small::smm<1, N, 1, false, false>(fr, F, &Q);
  Q = Q + V;
  if (isnan(y)) {  // all missing (:74-75)
#pragma unroll
    for (int i = 0; i < N; ++i) m[i] = a[i];
#pragma unroll
    for (int k = 0; k < NL; ++k) Cl[k] = Rl[k];
    return;
  }
  const double e = y - f;
  double R[N * N], rhs[N], K[N], D[N * N], Dr[NL], t1l[NL], t1[N * N];
  px.template gather_mat<kSiteR>(Rl, R);
  small::smm<1, N, N, true, true>(F, R, rhs);              // F^T R^T (needs rows of R)
  if (Q == 0.0) st |= BDLM_ST_SINGULAR;
#pragma unroll
  for (int i = 0; i < N; ++i) K[i] = rhs[i] / Q;           // (:83)
#pragma unroll
  for (int i = 0; i < N; ++i) m[i] = a[i] + K[i] * e;
  small::smm<N, 1, N, false, true>(K, F, D);               // K F^T
#pragma unroll
  for (int j = 0; j < N; ++j)
#pragma unroll
    for (int i = 0; i < N; ++i) D[i + j * N] = ((i == j) ? 1.0 : 0.0) - D[i + j * N];
  small::smm<N, N, 2, false, false>(D, Rl, t1l);           // columns of D R
  px.template gather_mat<kSiteUpdT1>(t1l, t1);
  own_rows(px, D, Dr);
  const double K0 = px.h ? K[2] : K[0], K1 = px.h ? K[3] : K[1];
#pragma unroll
  for (int jj = 0; jj < 2; ++jj)
#pragma unroll
    for (int i = 0; i < N; ++i) {
      double acc = 0.0;
#pragma unroll
      for (int k = 0; k < N; ++k) {                        // (t1 D^T)[i, j]
        const double prod = t1[i + 4 * k] * Dr[jj + 2 * k];
        acc = (k == 0) ? prod : acc + prod;
      }
      const double t2 = K[i] * V;                          // (K V) K^T
      const double c2 = t2 * (jj ? K1 : K0);
      Cl[i + 4 * jj] = acc + c2;
    }
}

// Smoothing.smoothStep (Smoothing.scala:31-47).  m, a1, s: both lanes; Cl, R1l, Sl: own columns;
// Cr: rows 2h, 2h+1 of C_t (Cr[jj + 2 * k] = C[(2h + jj), k]).
#pragma nv_exec_check_disable
template <class PX>
__host__ __device__ __forceinline__ void rts_step(PX &px, const double *G, const double (&m)[N],
                                                  const double (&Cl)[NL], const double (&Cr)[NL],
                                                  const double (&a1)[N], const double (&R1l)[NL],
                                                  bool textbook, double (&s)[N], double (&Sl)[NL],
                                                  int &st) {
  double X[NL], R1[N * N], At[N * N], d[N], tl[2], t[N], Dml[NL], Xf[N * N], t1l[NL], t1[N * N];
  // rhs = G C^T, own columns: rhs[i, j] = sum_k G[i, k] C[j, k]
#pragma unroll
  for (int jj = 0; jj < 2; ++jj)
#pragma unroll
    for (int i = 0; i < N; ++i) {
      double acc = 0.0;
#pragma unroll
      for (int k = 0; k < N; ++k) {
        const double prod = G[i + 4 * k] * Cr[jj + 2 * k];
        acc = (k == 0) ? prod : acc + prod;
      }
      X[i + 4 * jj] = acc;
    }
  px.template gather_mat<kSiteR1>(R1l, R1);
#pragma unroll
  for (int j = 0; j < N; ++j)
#pragma unroll
    for (int i = 0; i < N; ++i) At[i + j * N] = R1[j + i * N];
  st |= small::lu_solve<N, 2>(At, X);                      // factorisation on both lanes, own right-hand sides
  // Bg = X^T: the own COLUMNS of X are the own ROWS of Bg
#pragma unroll
  for (int i = 0; i < N; ++i) d[i] = s[i] - a1[i];
#pragma unroll
  for (int jj = 0; jj < 2; ++jj) {
    double acc = 0.0;
#pragma unroll
    for (int k = 0; k < N; ++k) {                          // (Bg d)[2h + jj]
      const double prod = X[k + 4 * jj] * d[k];
      acc = (k == 0) ? prod : acc + prod;
    }
    tl[jj] = acc;
  }
  px.template gather_vec<kSiteT>(tl[0], tl[1], t);
#pragma unroll
  for (int k = 0; k < NL; ++k) Dml[k] = R1l[k] - Sl[k];
  px.template gather_mat<kSiteX>(X, Xf);
  // t1 = Bg Dm, own columns: t1[i, j] = sum_k Bg[i, k] Dm[k, j] = sum_k X[k, i] Dm[k, j]
#pragma unroll
  for (int jj = 0; jj < 2; ++jj)
#pragma unroll
    for (int i = 0; i < N; ++i) {
      double acc = 0.0;
#pragma unroll
      for (int k = 0; k < N; ++k) {
        const double prod = Xf[k + 4 * i] * Dml[k + 4 * jj];
        acc = (k == 0) ? prod : acc + prod;
      }
      t1l[i + 4 * jj] = acc;
    }
  px.template gather_mat<kSiteRtsT1>(t1l, t1);
  double Xr[NL];
  own_rows(px, Xf, Xr);                                    // rows 2h, 2h+1 of X = columns of Bg
#pragma unroll
  for (int jj = 0; jj < 2; ++jj)
#pragma unroll
    for (int i = 0; i < N; ++i) {
      double acc = 0.0;
#pragma unroll
      for (int k = 0; k < N; ++k) {
        // textbook: (t1 Bg^T)[i, j] = sum_k t1[i, k] X[k, j]; Smoothing.scala:44 (no transpose):
        // (t1 Bg)[i, j] = sum_k t1[i, k] X[j, k]
        const double b = textbook ? X[k + 4 * jj] : Xr[jj + 2 * k];
        const double prod = t1[i + 4 * k] * b;
        acc = (k == 0) ? prod : acc + prod;
      }
      Sl[i + 4 * jj] = Cl[i + 4 * jj] - acc;
    }
#pragma unroll
  for (int i = 0; i < N; ++i) s[i] = m[i] + t[i];
}

}  // namespace pairk
}  // namespace bdlm
