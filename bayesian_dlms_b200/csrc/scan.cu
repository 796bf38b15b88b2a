// scan.cu -- parallel-in-time Kalman filter / RTS smoother for ONE long series
// (BASELINE.json config 5; not in the reference, whose time recursion is strictly
// sequential -- Filter.scala:41-62, and whose FilterTs.scanRight is literally `???`).
//
// Formulation: Sarkka & Garcia-Fernandez, "Temporal Parallelization of Bayesian Smoothers"
// (IEEE TAC 2021) mapped to DLM notation (H = F^T, Q = W dt, R = V).
// Two-level scan, organised for HBM:
//   level 1  every thread owns kSub consecutive time points and composes their filtering
//            elements (A, b, C, eta, J) sequentially, reading only y (8 B / step);
//   level 2  the per-thread aggregates (T / kSub of them, a few MB) are scanned with the
//            associative operator (block-wide Hillis-Steele in shared memory + one
//            recursion over block totals);
//   apply    every thread starts from its exclusive prefix -- whose (b, C) IS the filtered
//            (m, C) at its first time point -- and runs the ordinary sequential Kalman
//            recursion (the same register code as kf_small.cu) over its kSub steps,
//            writing the KfState outputs.  The smoother repeats this backwards with the
//            elements (E, g, L); its apply phase is the sequential RTS step in textbook mode.
// Elements are never materialised per time point: HBM traffic is y twice + outputs once.
//
// Multi-GPU: a rank's chunk aggregate (3n^2+2n doubles forward, 2n^2+n backward) is all
// that crosses the fabric (one all-gather per pass); the carry is folded on the host.
//
// Parity: equals the sequential kernel in textbook-smoother mode to 1e-9 relative (for n = 1
// that is the reference itself); the apply phases reuse the sequential step code, so the
// only difference is the rounding of the scanned start states.
#include <vector>

#include "common.cuh"
#include "launch.h"
#include "small_steps.cuh"

namespace bdlm {

using namespace small;

// ------------------------------------------------------------------ elements & operators

template <int N>
struct FElem {  // filtering element
  double A[N * N], b[N], C[N * N], eta[N], J[N * N];
};
template <int N>
struct SElem {  // smoothing element
  double E[N * N], g[N], L[N * N];
};

template <int N>
__host__ __device__ __forceinline__ void f_identity(FElem<N> &e) {
#pragma unroll
  for (int k = 0; k < N * N; ++k) { e.A[k] = (k % (N + 1) == 0) ? 1.0 : 0.0; e.C[k] = 0.0; e.J[k] = 0.0; }
#pragma unroll
  for (int k = 0; k < N; ++k) { e.b[k] = 0.0; e.eta[k] = 0.0; }
}

template <int N>
__host__ __device__ __forceinline__ void f_state(FElem<N> &e, const double *m, const double *C) {
#pragma unroll
  for (int k = 0; k < N * N; ++k) { e.A[k] = 0.0; e.C[k] = C[k]; e.J[k] = 0.0; }
#pragma unroll
  for (int k = 0; k < N; ++k) { e.b[k] = m[k]; e.eta[k] = 0.0; }
}

// out = ei (x) ej, ei earlier in time.  (Appendix C)
template <int N>
__host__ __device__ __forceinline__ void f_combine(const FElem<N> &ei, const FElem<N> &ej,
                                                   FElem<N> &out) {
  double CJ[N * N], Mt[N * N], X[N * N], JC[N * N], Nt[N * N], Y[N * N], t1[N * N], t2[N * N],
      v1[N], v2[N];
  // X = A_j (I + C_i J_j)^-1 :  X^T = (I + C_i J_j)^-T A_j^T
  smm<N, N, N, false, false>(ei.C, ej.J, CJ);
#pragma unroll
  for (int j = 0; j < N; ++j)
#pragma unroll
    for (int i = 0; i < N; ++i) {
      Mt[i + j * N] = ((i == j) ? 1.0 : 0.0) + CJ[j + i * N];  // (I + C_i J_j)^T
      X[i + j * N] = ej.A[j + i * N];                            // A_j^T
    }
  lu_solve<N, N>(Mt, X);  // X := (I + C_i J_j)^-T A_j^T = (A_j M)^T
  // Y^T = (I + J_j C_i)^-T A_i  ->  Y = A_i^T (I + J_j C_i)^-1
  smm<N, N, N, false, false>(ej.J, ei.C, JC);
#pragma unroll
  for (int j = 0; j < N; ++j)
#pragma unroll
    for (int i = 0; i < N; ++i) {
      Nt[i + j * N] = ((i == j) ? 1.0 : 0.0) + JC[j + i * N];
      Y[i + j * N] = ei.A[i + j * N];
    }
  lu_solve<N, N>(Nt, Y);  // Y := (I + J_j C_i)^-T A_i = (A_i^T N)^T
  // A = (A_j M) A_i
  smm<N, N, N, true, false>(X, ei.A, out.A);
  // b = (A_j M)(b_i + C_i eta_j) + b_j
  smm<N, N, 1, false, false>(ei.C, ej.eta, v1);
#pragma unroll
  for (int k = 0; k < N; ++k) v1[k] = ei.b[k] + v1[k];
  smm<N, N, 1, true, false>(X, v1, v2);
#pragma unroll
  for (int k = 0; k < N; ++k) out.b[k] = v2[k] + ej.b[k];
  // C = (A_j M) C_i A_j^T + C_j
  smm<N, N, N, true, false>(X, ei.C, t1);
  smm<N, N, N, false, true>(t1, ej.A, t2);
#pragma unroll
  for (int k = 0; k < N * N; ++k) out.C[k] = t2[k] + ej.C[k];
  // eta = (A_i^T N)(eta_j - J_j b_i) + eta_i
  smm<N, N, 1, false, false>(ej.J, ei.b, v1);
#pragma unroll
  for (int k = 0; k < N; ++k) v1[k] = ej.eta[k] - v1[k];
  smm<N, N, 1, true, false>(Y, v1, v2);
#pragma unroll
  for (int k = 0; k < N; ++k) out.eta[k] = v2[k] + ei.eta[k];
  // J = (A_i^T N) J_j A_i + J_i
  smm<N, N, N, true, false>(Y, ej.J, t1);
  smm<N, N, N, false, false>(t1, ei.A, t2);
#pragma unroll
  for (int k = 0; k < N * N; ++k) out.J[k] = t2[k] + ei.J[k];
}

// Filtering element of one observation (p = 1): Q = W dt.
template <int N>
__host__ __device__ __forceinline__ void f_element(const double *G, const double *F,
                                                   const double (&Q)[N * N], double V, double y,
                                                   FElem<N> &e) {
  if (isnan(y)) {  // missing: pure prediction
#pragma unroll
    for (int k = 0; k < N * N; ++k) { e.A[k] = G[k]; e.C[k] = Q[k]; e.J[k] = 0.0; }
#pragma unroll
    for (int k = 0; k < N; ++k) { e.b[k] = 0.0; e.eta[k] = 0.0; }
    return;
  }
  double QF[N], GtF[N], K[N], IKH[N * N], S;
  smm<N, N, 1, false, false>(Q, F, QF);
  smm<1, N, 1, true, false>(F, QF, &S);
  S = S + V;
  smm<N, N, 1, true, false>(G, F, GtF);  // G^T F
#pragma unroll
  for (int i = 0; i < N; ++i) K[i] = QF[i] / S;
#pragma unroll
  for (int j = 0; j < N; ++j)
#pragma unroll
    for (int i = 0; i < N; ++i) IKH[i + j * N] = ((i == j) ? 1.0 : 0.0) - K[i] * F[j];
  smm<N, N, N, false, false>(IKH, G, e.A);
  smm<N, N, N, false, false>(IKH, Q, e.C);
#pragma unroll
  for (int i = 0; i < N; ++i) { e.b[i] = K[i] * y; e.eta[i] = GtF[i] * y / S; }
#pragma unroll
  for (int j = 0; j < N; ++j)
#pragma unroll
    for (int i = 0; i < N; ++i) e.J[i + j * N] = GtF[i] * GtF[j] / S;
}

template <int N>
__host__ __device__ __forceinline__ void s_identity(SElem<N> &e) {
#pragma unroll
  for (int k = 0; k < N * N; ++k) { e.E[k] = (k % (N + 1) == 0) ? 1.0 : 0.0; e.L[k] = 0.0; }
#pragma unroll
  for (int k = 0; k < N; ++k) e.g[k] = 0.0;
}

template <int N>
__host__ __device__ __forceinline__ void s_state(SElem<N> &e, const double *s, const double *S) {
#pragma unroll
  for (int k = 0; k < N * N; ++k) { e.E[k] = 0.0; e.L[k] = S[k]; }
#pragma unroll
  for (int k = 0; k < N; ++k) e.g[k] = s[k];
}

// out = ei (x) ej, ei earlier:  E = E_i E_j ; g = E_i g_j + g_i ; L = E_i L_j E_i^T + L_i
template <int N>
__host__ __device__ __forceinline__ void s_combine(const SElem<N> &ei, const SElem<N> &ej,
                                                   SElem<N> &out) {
  double t1[N * N], t2[N * N], v[N];
  smm<N, N, N, false, false>(ei.E, ej.E, out.E);
  smm<N, N, 1, false, false>(ei.E, ej.g, v);
#pragma unroll
  for (int k = 0; k < N; ++k) out.g[k] = v[k] + ei.g[k];
  smm<N, N, N, false, false>(ei.E, ej.L, t1);
  smm<N, N, N, false, true>(t1, ei.E, t2);
#pragma unroll
  for (int k = 0; k < N * N; ++k) out.L[k] = t2[k] + ei.L[k];
}

// Smoothing element of a filtered row that has a successor: E = C G^T R1^-1 (the RTS gain),
// g = m - E a1, L = C - E R1 E^T, with (a1, R1) the prediction from (m, C).
template <int N>
__host__ __device__ __forceinline__ void s_element(const double *G, const double (&W)[N * N],
                                                   const double (&m)[N], const double (&C)[N * N],
                                                   SElem<N> &e) {
  double a1[N], R1[N * N], rhs[N * N], At[N * N], t1[N * N], t2[N * N], v[N];
  advance<N, true>(G, W, 1.0, m, C, a1, R1);
  smm<N, N, N, false, true>(G, C, rhs);
#pragma unroll
  for (int j = 0; j < N; ++j)
#pragma unroll
    for (int i = 0; i < N; ++i) At[i + j * N] = R1[j + i * N];
  lu_solve<N, N>(At, rhs);
#pragma unroll
  for (int j = 0; j < N; ++j)
#pragma unroll
    for (int i = 0; i < N; ++i) e.E[i + j * N] = rhs[j + i * N];
  smm<N, N, 1, false, false>(e.E, a1, v);
#pragma unroll
  for (int k = 0; k < N; ++k) e.g[k] = m[k] - v[k];
  smm<N, N, N, false, false>(e.E, R1, t1);
  smm<N, N, N, false, true>(t1, e.E, t2);
#pragma unroll
  for (int k = 0; k < N * N; ++k) e.L[k] = C[k] - t2[k];
}

namespace {

constexpr int kSub = 64;       // time points per thread (level 1)
constexpr int kScanBlock = 256;

template <int N>
struct ScanModel {
  double G[N * N], F[N], W[N * N], V;
};

// ---- level 1, forward: per-thread aggregate over observations [c*kSub, (c+1)*kSub)
template <int N>
__global__ void __launch_bounds__(128)
fwd_reduce_kernel(const ScanModel<N> md, const double *__restrict__ y, int64_t T, int64_t M,
                  FElem<N> *agg /* [M + 1], slot 0 reserved for the start element */) {
  const int64_t c = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (c >= M) return;
  const int64_t t0 = c * kSub, t1 = (t0 + kSub < T) ? t0 + kSub : T;
  FElem<N> acc, e, tmp;
  f_element<N>(md.G, md.F, md.W, md.V, ld_stream(y + t0), acc);
  for (int64_t t = t0 + 1; t < t1; ++t) {
    f_element<N>(md.G, md.F, md.W, md.V, ld_stream(y + t), e);
    f_combine<N>(acc, e, tmp);
    acc = tmp;
  }
  agg[c + 1] = acc;
}

// ---- level 2: inclusive scan of an element array with an associative operator.
// IDXREV: scan position p maps to array index M-1-p (suffix scan).  OPREV: an element later
// in SCAN order is EARLIER in time (so it is the left operand).  The top level of a suffix
// scan uses (IDXREV, OPREV) = (true, true); its block totals are stored in scan order and
// are therefore scanned with (false, true).
template <class E>
struct Op;
template <int N>
struct Op<FElem<N>> {
  __device__ static void apply(const FElem<N> &earlier, const FElem<N> &later, FElem<N> &o) {
    f_combine<N>(earlier, later, o);
  }
};
template <int N>
struct Op<SElem<N>> {
  __device__ static void apply(const SElem<N> &earlier, const SElem<N> &later, SElem<N> &o) {
    s_combine<N>(earlier, later, o);
  }
};

template <class E, bool IDXREV, bool OPREV>
__global__ void __launch_bounds__(kScanBlock)
block_scan_kernel(E *x, int64_t M, E *totals) {
  extern __shared__ unsigned char raw[];
  E *buf0 = reinterpret_cast<E *>(raw), *buf1 = buf0 + kScanBlock;
  const int tid = threadIdx.x;
  const int64_t base = (int64_t)blockIdx.x * kScanBlock;
  const int64_t p = base + tid;
  const int64_t idx = IDXREV ? (M - 1 - p) : p;
  const bool act = p < M;
  if (act) buf0[tid] = x[idx];
  __syncthreads();
  E *src = buf0, *dst = buf1;
  for (int off = 1; off < kScanBlock; off <<= 1) {
    if (act) {
      if (tid >= off) {
        E o;
        if (OPREV) Op<E>::apply(src[tid], src[tid - off], o);
        else Op<E>::apply(src[tid - off], src[tid], o);
        dst[tid] = o;
      } else {
        dst[tid] = src[tid];
      }
    }
    __syncthreads();
    E *t = src; src = dst; dst = t;
  }
  if (act) x[idx] = src[tid];
  const int64_t last = (base + kScanBlock <= M) ? kScanBlock - 1 : (M - 1 - base);
  if (totals && tid == last) totals[blockIdx.x] = src[tid];
}

template <class E, bool IDXREV, bool OPREV>
__global__ void __launch_bounds__(kScanBlock)
add_prefix_kernel(E *x, int64_t M, const E *totals /* scanned, scan order */) {
  if (blockIdx.x == 0) return;
  const int64_t p = (int64_t)blockIdx.x * kScanBlock + threadIdx.x;
  if (p >= M) return;
  const int64_t idx = IDXREV ? (M - 1 - p) : p;
  const E pre = totals[blockIdx.x - 1], cur = x[idx];
  E o;
  if (OPREV) Op<E>::apply(cur, pre, o);
  else Op<E>::apply(pre, cur, o);
  x[idx] = o;
}

template <class E, bool IDXREV, bool OPREV>
cudaError_t device_scan(E *x, int64_t M, E *scratch, cudaStream_t stream, int64_t *launches) {
  if (M <= 1) return cudaSuccess;
  const int64_t nb = (M + kScanBlock - 1) / kScanBlock;
  const size_t smem = 2 * kScanBlock * sizeof(E);
  cudaError_t e = cudaFuncSetAttribute(block_scan_kernel<E, IDXREV, OPREV>,
                                       cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  block_scan_kernel<E, IDXREV, OPREV><<<(unsigned)nb, kScanBlock, smem, stream>>>(
      x, M, nb > 1 ? scratch : nullptr);
  ++*launches;
  if (nb > 1) {
    e = device_scan<E, false, OPREV>(scratch, nb, scratch + nb, stream, launches);
    if (e != cudaSuccess) return e;
    add_prefix_kernel<E, IDXREV, OPREV><<<(unsigned)nb, kScanBlock, 0, stream>>>(x, M, scratch);
    ++*launches;
  }
  return cudaGetLastError();
}

// ---- apply, forward: sequential Kalman recursion from the scanned start state
template <int N>
__global__ void __launch_bounds__(128)
fwd_apply_kernel(const ScanModel<N> md, const double *__restrict__ y, int64_t T, int64_t M,
                 const FElem<N> *pre /* [M+1] inclusive scan with slot 0 = start */,
                 int keep_init, KfViews kf, int32_t *status) {
  const int64_t c = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (c >= M) return;
  const int64_t t0 = c * kSub, t1 = (t0 + kSub < T) ? t0 + kSub : T;
  double m[N], C[N * N], W[N * N];
  int st = 0;
#pragma unroll
  for (int k = 0; k < N; ++k) m[k] = pre[c].b[k];
#pragma unroll
  for (int k = 0; k < N * N; ++k) { C[k] = pre[c].C[k]; W[k] = md.W[k]; }
  auto store = [&](const View &v, int64_t row, const double *x, int K) {
    if (!v.ptr) return;
    double *p = v.ptr + row * v.sr;
    for (int k = 0; k < K; ++k) st_stream(p + k * v.sk, x[k]);
  };
  if (c == 0 && keep_init) {
    const double nanv = __longlong_as_double(0x7ff8000000000000LL);
    store(kf.m, 0, m, N); store(kf.C, 0, C, N * N);
    store(kf.a, 0, m, N); store(kf.R, 0, C, N * N);
    store(kf.f, 0, &nanv, 1); store(kf.Q, 0, &nanv, 1);
  }
  for (int64_t t = t0; t < t1; ++t) {
    double a[N], R[N * N], f, Q;
    advance<N, true>(md.G, W, 1.0, m, C, a, R);
    update<N>(md.F, md.V, ld_stream(y + t), a, R, f, Q, m, C, st);
    const int64_t row = t + keep_init;
    store(kf.a, row, a, N); store(kf.R, row, R, N * N);
    store(kf.f, row, &f, 1); store(kf.Q, row, &Q, 1);
    store(kf.m, row, m, N); store(kf.C, row, C, N * N);
  }
  if (status && st) atomicOr(status, st);
}

// ---- level 1, backward: per-thread aggregate of smoothing elements over rows
// [c*kSub, (c+1)*kSub) of the `nrows` rows that HAVE a successor.
template <int N>
__global__ void __launch_bounds__(128)
bwd_reduce_kernel(const ScanModel<N> md, View fm, View fC, int64_t nrows, int64_t M,
                  SElem<N> *agg /* [M + 1], slot M reserved for the terminal element */) {
  const int64_t c = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (c >= M) return;
  const int64_t r0 = c * kSub, r1 = (r0 + kSub < nrows) ? r0 + kSub : nrows;
  double W[N * N];
#pragma unroll
  for (int k = 0; k < N * N; ++k) W[k] = md.W[k];
  SElem<N> acc, e, tmp;
  for (int64_t r = r0; r < r1; ++r) {
    double m[N], C[N * N];
#pragma unroll
    for (int k = 0; k < N; ++k) m[k] = ld_stream(fm.ptr + r * fm.sr + k * fm.sk);
#pragma unroll
    for (int k = 0; k < N * N; ++k) C[k] = ld_stream(fC.ptr + r * fC.sr + k * fC.sk);
    s_element<N>(md.G, W, m, C, e);
    if (r == r0) acc = e;
    else { s_combine<N>(acc, e, tmp); acc = tmp; }
  }
  agg[c] = acc;
}

// ---- apply, backward: sequential (textbook) RTS recursion from the scanned successor state
template <int N>
__global__ void __launch_bounds__(128)
bwd_apply_kernel(const ScanModel<N> md, View fm, View fC, int64_t nrows, int64_t M,
                 const SElem<N> *suf /* [M+1] suffix-inclusive scan, slot M = terminal */,
                 View sv, View Sv, int32_t *status) {
  const int64_t c = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (c >= M) return;
  const int64_t r0 = c * kSub, r1 = (r0 + kSub < nrows) ? r0 + kSub : nrows;
  double W[N * N], s[N], S[N * N];
  int st = 0;
#pragma unroll
  for (int k = 0; k < N * N; ++k) { W[k] = md.W[k]; S[k] = suf[c + 1].L[k]; }
#pragma unroll
  for (int k = 0; k < N; ++k) s[k] = suf[c + 1].g[k];
  for (int64_t r = r1 - 1; r >= r0; --r) {
    double m[N], C[N * N], a1[N], R1[N * N];
#pragma unroll
    for (int k = 0; k < N; ++k) m[k] = ld_stream(fm.ptr + r * fm.sr + k * fm.sk);
#pragma unroll
    for (int k = 0; k < N * N; ++k) C[k] = ld_stream(fC.ptr + r * fC.sr + k * fC.sk);
    advance<N, true>(md.G, W, 1.0, m, C, a1, R1);
    rts_step<N>(md.G, m, C, a1, R1, /*textbook=*/true, s, S, st);
    if (sv.ptr)
      for (int k = 0; k < N; ++k) st_stream(sv.ptr + r * sv.sr + k * sv.sk, s[k]);
    if (Sv.ptr)
      for (int k = 0; k < N * N; ++k) st_stream(Sv.ptr + r * Sv.sr + k * Sv.sk, S[k]);
  }
  if (status && st) atomicOr(status, st);
}

template <int N>
__global__ void set_f_start(FElem<N> *slot, const double *mC, bool identity) {
  FElem<N> e;
  if (identity) f_identity<N>(e); else f_state<N>(e, mC, mC + N);
  *slot = e;
}
template <int N>
__global__ void set_s_terminal(SElem<N> *slot, const double *sS, bool identity) {
  SElem<N> e;
  if (identity) s_identity<N>(e); else s_state<N>(e, sS, sS + N);
  *slot = e;
}
template <int N>
__global__ void copy_last_row(View fm, View fC, int64_t row, View sv, View Sv, double *sS) {
  for (int k = 0; k < N; ++k) {
    const double v = fm.ptr[row * fm.sr + k * fm.sk];
    if (sv.ptr) sv.ptr[row * sv.sr + k * sv.sk] = v;
    sS[k] = v;
  }
  for (int k = 0; k < N * N; ++k) {
    const double v = fC.ptr[row * fC.sr + k * fC.sk];
    if (Sv.ptr) Sv.ptr[row * Sv.sr + k * Sv.sk] = v;
    sS[N + k] = v;
  }
}

template <int N>
ScanModel<N> make_model(const ScanArgs &a) {
  ScanModel<N> md;
  for (int k = 0; k < N * N; ++k) { md.G[k] = a.G[k]; md.W[k] = a.W[k]; }
  for (int k = 0; k < N; ++k) md.F[k] = a.F[k];
  md.V = a.V;
  return md;
}

#define CK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) return e_; } while (0)

template <int N>
cudaError_t scan_forward(const ScanArgs &a, cudaStream_t stream, int64_t *launches) {
  const ScanModel<N> md = make_model<N>(a);
  const int64_t T = a.T, M = (T + kSub - 1) / kSub;
  FElem<N> *X = reinterpret_cast<FElem<N> *>(a.workspace);       // [M + 1]
  FElem<N> *scratch = X + (M + 1);                                // block totals
  double *start_dev = reinterpret_cast<double *>(scratch + (M + 1) / kScanBlock * 2 + 8);
  const unsigned blocks = (unsigned)((M + 127) / 128);
  if (a.phase == kScanReduce) {
    set_f_start<N><<<1, 1, 0, stream>>>(X, nullptr, true);
    fwd_reduce_kernel<N><<<blocks, 128, 0, stream>>>(md, a.y, T, M, X);
    *launches += 2;
    CK(cudaGetLastError());
    CK((device_scan<FElem<N>, false, false>(X, M + 1, scratch, stream, launches)));
    CK(cudaMemcpyAsync(a.agg_out, X + M, sizeof(FElem<N>), cudaMemcpyDeviceToHost, stream));
    return cudaStreamSynchronize(stream);
  }
  // apply: prefix of (start (x) aggregates); a.start = host (m, C) of the state before t = 0
  CK(cudaMemcpyAsync(start_dev, a.start, sizeof(double) * (N + N * N), cudaMemcpyHostToDevice, stream));
  set_f_start<N><<<1, 1, 0, stream>>>(X, start_dev, false);
  fwd_reduce_kernel<N><<<blocks, 128, 0, stream>>>(md, a.y, T, M, X);
  *launches += 2;
  CK(cudaGetLastError());
  CK((device_scan<FElem<N>, false, false>(X, M + 1, scratch, stream, launches)));
  fwd_apply_kernel<N><<<blocks, 128, 0, stream>>>(md, a.y, T, M, X, a.keep_init, a.kf, a.status);
  ++*launches;
  return cudaGetLastError();
}

template <int N>
cudaError_t scan_backward(const ScanArgs &a, cudaStream_t stream, int64_t *launches) {
  const ScanModel<N> md = make_model<N>(a);
  const int64_t rows = a.T + a.keep_init;
  // rows with a successor inside this chunk: all but the last, plus the last when the chunk
  // is followed by another one (has_successor)
  const int64_t nrows = a.has_successor ? rows : rows - 1;
  const int64_t M = (nrows + kSub - 1) / kSub;
  SElem<N> *X = reinterpret_cast<SElem<N> *>(a.workspace);  // [M + 1]
  SElem<N> *scratch = X + (M + 1);
  double *term_dev = reinterpret_cast<double *>(scratch + (M + 1) / kScanBlock * 2 + 8);
  const unsigned blocks = (unsigned)((M + 127) / 128);
  const bool reduce = a.phase == kScanReduce;
  if (reduce) {
    set_s_terminal<N><<<1, 1, 0, stream>>>(X + M, nullptr, true);
  } else if (a.has_successor) {
    CK(cudaMemcpyAsync(term_dev, a.start, sizeof(double) * (N + N * N), cudaMemcpyHostToDevice, stream));
    set_s_terminal<N><<<1, 1, 0, stream>>>(X + M, term_dev, false);
  } else {  // last chunk: s_T = m_T, S_T = C_T (Smoothing.scala:59-61)
    copy_last_row<N><<<1, 1, 0, stream>>>(a.kf.m, a.kf.C, rows - 1, a.s, a.S, term_dev);
    set_s_terminal<N><<<1, 1, 0, stream>>>(X + M, term_dev, false);
    ++*launches;
  }
  ++*launches;
  if (M > 0) {
    bwd_reduce_kernel<N><<<blocks, 128, 0, stream>>>(md, a.kf.m, a.kf.C, nrows, M, X);
    ++*launches;
  }
  CK(cudaGetLastError());
  CK((device_scan<SElem<N>, true, true>(X, M + 1, scratch, stream, launches)));
  if (reduce) {
    CK(cudaMemcpyAsync(a.agg_out, X, sizeof(SElem<N>), cudaMemcpyDeviceToHost, stream));
    return cudaStreamSynchronize(stream);
  }
  if (M > 0) {
    bwd_apply_kernel<N><<<blocks, 128, 0, stream>>>(md, a.kf.m, a.kf.C, nrows, M, X, a.s, a.S, a.status);
    ++*launches;
  }
  return cudaGetLastError();
}

}  // namespace

size_t scan_workspace_bytes(int n, int64_t T) {
  const int64_t M = (T + 1 + kSub - 1) / kSub + 2;
  const size_t elem = sizeof(double) * (3 * n * n + 2 * n);
  return elem * (size_t)(M + 1 + (M + 1) / kScanBlock * 2 + 16) + 4096;
}

int scan_forward_elem_doubles(int n) { return 3 * n * n + 2 * n; }
int scan_backward_elem_doubles(int n) { return 2 * n * n + n; }

cudaError_t launch_scan(const ScanArgs &a, cudaStream_t stream, int64_t *launches) {
  switch (a.n) {
#define BDLM_SCAN_CASE(N_)                                                         \
  case N_:                                                                         \
    return a.backward ? scan_backward<N_>(a, stream, launches) : scan_forward<N_>(a, stream, launches);
    BDLM_SCAN_CASE(1)
    BDLM_SCAN_CASE(2)
    BDLM_SCAN_CASE(3)
    BDLM_SCAN_CASE(4)
#undef BDLM_SCAN_CASE
    default: return cudaErrorInvalidValue;
  }
}

// Host-side carry composition for multi-GPU runs: out = ei (x) ej (ei earlier in time).
void scan_combine_host(int n, bool backward, const double *ei, const double *ej, double *out) {
  switch (n) {
#define BDLM_COMB_CASE(N_)                                                                  \
  case N_:                                                                                  \
    if (backward) s_combine<N_>(*reinterpret_cast<const SElem<N_> *>(ei),                   \
                                *reinterpret_cast<const SElem<N_> *>(ej),                   \
                                *reinterpret_cast<SElem<N_> *>(out));                       \
    else f_combine<N_>(*reinterpret_cast<const FElem<N_> *>(ei),                            \
                       *reinterpret_cast<const FElem<N_> *>(ej),                            \
                       *reinterpret_cast<FElem<N_> *>(out));                                \
    break;
    BDLM_COMB_CASE(1)
    BDLM_COMB_CASE(2)
    BDLM_COMB_CASE(3)
    BDLM_COMB_CASE(4)
#undef BDLM_COMB_CASE
    default: break;
  }
}

}  // namespace bdlm
