// warp_linalg.cuh -- warp-cooperative dense linear algebra on shared-memory operands.
//
// One warp owns one series; its matrices (n <= 48, p <= 32) live in shared memory, column
// major, and the 32 lanes split the OUTPUT elements of every operation.  Each output
// element is produced by exactly one lane with the oracle's operation order (products
// summed in increasing inner index, no FMA), so results are bit-identical to
// oracle/bdlm_oracle.c regardless of how elements are distributed over lanes.
//
// Every routine ends with __syncwarp() so its results are visible to all lanes;
// outputs must not alias inputs unless stated.
#pragma once
#include "common.cuh"

namespace bdlm {

#define FULL 0xffffffffu

// Iterate idx = lane, lane+32, ... < rows*cols yielding (i, j) = (idx % rows, idx / rows)
// with one division at entry.
struct ElemIter {
  int i, j, di, dj, rows, left;
  __device__ __forceinline__ ElemIter(int lane, int rows_, int cols) : rows(rows_) {
    i = lane % rows_;
    j = lane / rows_;
    di = 32 % rows_;
    dj = 32 / rows_;
    left = rows_ * cols - lane;
  }
  __device__ __forceinline__ bool ok() const { return left > 0; }
  __device__ __forceinline__ void next() {
    i += di; j += dj;
    if (i >= rows) { i -= rows; ++j; }
    left -= 32;
  }
};

// out(ar x bc) = A(ar x ac) * B(ac x bc)
__device__ __forceinline__ void w_mm(int lane, int ar, int ac, int bc, const double *A,
                                     int lda, bool ta, const double *B, int ldb, bool tb,
                                     double *out, int ldo) {
  for (ElemIter it(lane, ar, bc); it.ok(); it.next()) {
    const int i = it.i, j = it.j;
    double acc = 0.0;
    for (int k = 0; k < ac; ++k) {
      const double a = ta ? A[k + i * lda] : A[i + k * lda];
      const double b = tb ? B[j + k * ldb] : B[k + j * ldb];
      const double prod = a * b;
      acc = (k == 0) ? prod : acc + prod;
    }
    out[i + j * ldo] = acc;
  }
  __syncwarp();
}

__device__ __forceinline__ void w_mv(int lane, int ar, int ac, const double *A, int lda,
                                     bool ta, const double *x, double *y) {
  w_mm(lane, ar, ac, 1, A, lda, ta, x, ac, false, y, ar);
}

__device__ __forceinline__ void w_copy(int lane, int cnt, const double *src, double *dst) {
  for (int k = lane; k < cnt; k += 32) dst[k] = src[k];
  __syncwarp();
}

// dst(n x n) = src^T
__device__ __forceinline__ void w_transpose(int lane, int n, const double *src, double *dst) {
  for (ElemIter it(lane, n, n); it.ok(); it.next()) dst[it.i + it.j * n] = src[it.j + it.i * n];
  __syncwarp();
}

// dgesv restatement (oracle lu_solve): A n x n (destroyed), Bm n x nrhs, in place.
__device__ __forceinline__ int w_lu_solve(int lane, int n, double *A, int nrhs, double *Bm) {
  int st = 0;
  for (int j = 0; j < n; ++j) {
    // idamax over rows j..n-1 of column j: first index of the maximum magnitude
    // (rows beyond 31 wrap onto the lanes again: n <= 64; the first index of the maximum wins)
    double v = -1.0;
    int idx = lane;
    for (int i = lane; i < n; i += 32)
      if (i >= j) {
        const double a = fabs(A[i + j * n]);
        if (a > v) { v = a; idx = i; }
      }
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) {
      const double ov = __shfl_xor_sync(FULL, v, off);
      const int oi = __shfl_xor_sync(FULL, idx, off);
      if (ov > v || (ov == v && oi < idx)) { v = ov; idx = oi; }
    }
    const int jp = idx;
    const double pv = A[jp + j * n];
    __syncwarp();
    if (pv != 0.0) {
      if (jp != j) {
        for (int c = lane; c < n; c += 32) {
          const double t = A[j + c * n]; A[j + c * n] = A[jp + c * n]; A[jp + c * n] = t;
        }
        for (int c = lane; c < nrhs; c += 32) {
          const double t = Bm[j + c * n]; Bm[j + c * n] = Bm[jp + c * n]; Bm[jp + c * n] = t;
        }
        __syncwarp();
      }
      const double r = 1.0 / A[j + j * n];
      for (int i = lane; i < n; i += 32)
        if (i > j) A[i + j * n] = A[i + j * n] * r;
      __syncwarp();
    } else {
      st = BDLM_ST_SINGULAR;
    }
    const int rem = n - j - 1;
    if (rem > 0) {
      for (ElemIter it(lane, rem, rem); it.ok(); it.next()) {
        const int i = j + 1 + it.i, c = j + 1 + it.j;
        A[i + c * n] = A[i + c * n] - A[i + j * n] * A[j + c * n];
      }
      __syncwarp();
    }
  }
  // unit-lower forward substitution, then non-unit upper back substitution
  for (int k = 0; k < n - 1; ++k) {
    const int rem = n - k - 1;
    for (ElemIter it(lane, rem, nrhs); it.ok(); it.next()) {
      const int i = k + 1 + it.i, c = it.j;
      Bm[i + c * n] = Bm[i + c * n] - Bm[k + c * n] * A[i + k * n];
    }
    __syncwarp();
  }
  for (int k = n - 1; k >= 0; --k) {
    const double dkk = A[k + k * n];
    for (int c = lane; c < nrhs; c += 32) Bm[k + c * n] = Bm[k + c * n] / dkk;
    __syncwarp();
    if (k > 0) {
      for (ElemIter it(lane, k, nrhs); it.ok(); it.next()) {
        const int i = it.i, c = it.j;
        Bm[i + c * n] = Bm[i + c * n] - Bm[k + c * n] * A[i + k * n];
      }
      __syncwarp();
    }
  }
  return st;
}

// Round-robin partner of index i in `round` (oracle rr_partners); -1 = idle.
__device__ __forceinline__ int rr_partner(int n, int round, int i) {
  const int m = (n + 1) & ~1, mm1 = m - 1;
  const int r = round % mm1;
  int q;
  if (i == mm1) {
    q = r;
  } else {
    int d = i - r;
    if (d < 0) d += mm1;
    if (d == 0) q = mm1;
    else { q = r - d; if (q < 0) q += mm1; }
  }
  return (q < n) ? q : -1;
}

// oracle sym_rot
__device__ __forceinline__ void sym_rot(double app, double aqq, double apq, double &c,
                                        double &s) {
  const double theta = (aqq - app) / (2.0 * apq);
  double t;
  if (fabs(theta) > 1e150) {
    t = 0.5 / theta;
  } else {
    const double r = sqrt(theta * theta + 1.0);
    t = 1.0 / (fabs(theta) + r);
    if (theta < 0.0) t = -t;
  }
  c = 1.0 / sqrt(t * t + 1.0);
  s = t * c;
}

// Stable rank of key[i] among key[0..n) (ascending, or descending when desc): the
// position element i takes after the oracle's stable insertion sort.
__device__ __forceinline__ int stable_rank(int n, const double *key, int i, bool desc) {
  const double ki = key[i];
  int rank = 0;
  for (int j = 0; j < n; ++j) {
    const double kj = key[j];
    const bool before = desc ? (kj > ki) : (kj < ki);
    rank += (before || (kj == ki && j < i)) ? 1 : 0;
  }
  return rank;
}

// oracle order_and_sign for column `src` of V (rows x ?): returns true when the column
// must be flipped (largest-magnitude component, first such, negative).
__device__ __forceinline__ bool col_flip(int rows, const double *col) {
  int im = 0;
  double best = fabs(col[0]);
  for (int i = 1; i < rows; ++i) {
    const double a = fabs(col[i]);
    if (a > best) { best = a; im = i; }
  }
  return col[im] < 0.0;
}

// Pair of round-robin slot k in `round` (oracle rr_partners enumeration): returns false
// when the slot touches a dummy index; p < q otherwise.
__device__ __forceinline__ bool rr_slot(int n, int round, int k, int &p, int &q) {
  const int m = (n + 1) & ~1, mm1 = m - 1;
  const int rr = round % mm1;
  int pa, pb;
  if (k == 0) { pa = mm1; pb = rr; }
  else {
    pa = rr + k; if (pa >= mm1) pa -= mm1;
    pb = rr - k; if (pb < 0) pb += mm1;
  }
  if (pa >= n || pb >= n) return false;
  p = pa < pb ? pa : pb;
  q = pa < pb ? pb : pa;
  return true;
}

// Per-warp scratch of the Jacobi routines, sized for n <= 48 (24 column pairs per round).
constexpr int kScrDoubles = 128;  // w_jacobi_svd: dots [72] | c [24] | s [24]; then norms [48] | flips [48]
constexpr int kIscrInts = 128;    // pair p [48] | pair q [48] | observed-component list [32] at kObsOff
constexpr int kObsOff = 96;

// One-sided (Hestenes) Jacobi iteration (oracle jacobi_onesided) on the columns of U (r x n, IN
// PLACE) with the rotations accumulated in V (n x n, set to I here).  n <= 48.
// scr: kScrDoubles doubles, iscr: the first kObsOff ints of the per-warp integer scratch.
__device__ __forceinline__ int w_jacobi_onesided(int lane, int r, int n, double *U, double *V,
                                                 double *scr, int *iscr) {
  for (ElemIter it(lane, n, n); it.ok(); it.next())
    V[it.i + it.j * n] = (it.i == it.j) ? 1.0 : 0.0;
  __syncwarp();
  int st = (n == 1) ? 0 : BDLM_ST_NOTCONVERGED;
  const int m = (n + 1) & ~1, half = m / 2;
  double *dots = scr, *pc = scr + 72, *ps = scr + 96;
  int *pp = iscr, *pq = iscr + 48;
  for (int sweep = 0; sweep < kJacobiMaxSweeps && n > 1; ++sweep) {
    bool rotated = false;
    for (int round = 0; round < m - 1; ++round) {
      // alpha = |u_p|^2, beta = |u_q|^2, gamma = u_p . u_q : one lane per (slot, which)
      for (int w = lane; w < 3 * half; w += 32) {
        const int k = w / 3, which = w - 3 * k;
        int p, q;
        double acc = 0.0;
        if (rr_slot(n, round, k, p, q)) {
          const double *x = U + ((which == 1) ? q : p) * r;
          const double *y = U + ((which == 0) ? p : q) * r;
          for (int i = 0; i < r; ++i) {
            const double prod = x[i] * y[i];
            acc = (i == 0) ? prod : acc + prod;
          }
        }
        dots[w] = acc;
      }
      __syncwarp();
      bool rot = false;
      if (lane < half) {  // half <= 24
        int p = -1, q = -1;
        double c = 1.0, s = 0.0;
        if (rr_slot(n, round, lane, p, q)) {
          const double alpha = dots[3 * lane], beta = dots[3 * lane + 1],
                       gamma = dots[3 * lane + 2];
          if (gamma * gamma > kJacobiThr2 * (alpha * beta)) {
            sym_rot(alpha, beta, gamma, c, s);
            rot = true;
          }
        }
        pp[lane] = rot ? p : -1;
        pq[lane] = q;
        pc[lane] = c;
        ps[lane] = s;
      }
      const bool any = __any_sync(FULL, rot);
      __syncwarp();
      if (!any) continue;
      rotated = true;
      const int per = r + n;
      for (ElemIter it(lane, per, half); it.ok(); it.next()) {
        const int k = it.j, i = it.i;
        const int p = pp[k];
        if (p < 0) continue;
        const int q = pq[k];
        const double c = pc[k], s = ps[k];
        double *xp, *xq;
        if (i < r) { xp = U + i + p * r; xq = U + i + q * r; }
        else { xp = V + (i - r) + p * n; xq = V + (i - r) + q * n; }
        const double up = *xp, uq = *xq;
        *xp = c * up - s * uq;
        *xq = s * up + c * uq;
      }
      __syncwarp();
    }
    if (!rotated) { st = 0; break; }
  }
  return st;
}

// Common finish of the SVD / eigSym restatements: keys scr[0..n) are ranked (descending for
// singular values, ascending for eigenvalues; stable), out_key[rank] = key, Vout = the columns of
// V in that order with the oracle's sign rule (largest-|component| positive).
__device__ __forceinline__ void w_order_and_sign(int lane, int n, const double *V, double *scr,
                                                 int *iscr, bool desc, double *out_key,
                                                 double *Vout) {
  for (int j = lane; j < n; j += 32) {
    const int rk = stable_rank(n, scr, j, desc);
    iscr[rk] = j;
    out_key[rk] = scr[j];
  }
  __syncwarp();
  for (int j = lane; j < n; j += 32) scr[48 + j] = col_flip(n, V + iscr[j] * n) ? -1.0 : 1.0;
  __syncwarp();
  for (ElemIter it(lane, n, n); it.ok(); it.next()) {
    const double v = V[it.i + iscr[it.j] * n];
    Vout[it.i + it.j * n] = (scr[48 + it.j] < 0.0) ? -v : v;
  }
  __syncwarp();
}

// svd restatement (oracle jacobi_svd): U (r x n) IN PLACE, destroyed; V: n x n work.  sv[n]
// descending, Vout (n x n) = right singular vectors (Breeze rightVectors.t).
__device__ __forceinline__ int w_jacobi_svd(int lane, int r, int n, double *U, double *V,
                                            double *scr, int *iscr, double *sv,
                                            double *Vout) {
  const int st = w_jacobi_onesided(lane, r, n, U, V, scr, iscr);
  for (int j = lane; j < n; j += 32) {
    const double *col = U + j * r;
    double acc = 0.0;
    for (int i = 0; i < r; ++i) {
      const double sq = col[i] * col[i];
      acc = (i == 0) ? sq : acc + sq;
    }
    scr[j] = sqrt(acc);
  }
  __syncwarp();
  w_order_and_sign(lane, n, V, scr, iscr, true, sv, Vout);
  return st;
}

// eigSym restatement (oracle jacobi_eigsym): one-sided Jacobi on U = A (lower triangle of Ain
// read), lam_j = sign(v_j . u_j) |u_j| ascending, Vout = sign-normalised eigenvectors.
// U, V: n*n work each (distinct from Ain and Vout).
__device__ __forceinline__ int w_jacobi_eigsym(int lane, int n, const double *Ain, double *U,
                                               double *V, double *scr, int *iscr, double *lam,
                                               double *Vout) {
  for (ElemIter it(lane, n, n); it.ok(); it.next()) {
    const int i = it.i, j = it.j;
    U[i + j * n] = (i >= j) ? Ain[i + j * n] : Ain[j + i * n];
  }
  __syncwarp();
  const int st = w_jacobi_onesided(lane, n, n, U, V, scr, iscr);
  for (int j = lane; j < n; j += 32) {
    const double *u = U + j * n, *v = V + j * n;
    double nn = 0.0, dot = 0.0;
    for (int i = 0; i < n; ++i) {
      const double sq = u[i] * u[i], vu = v[i] * u[i];
      nn = (i == 0) ? sq : nn + sq;
      dot = (i == 0) ? vu : dot + vu;
    }
    const double nrm = sqrt(nn);
    scr[j] = (dot < 0.0) ? -nrm : nrm;
  }
  __syncwarp();
  w_order_and_sign(lane, n, V, scr, iscr, false, lam, Vout);
  return st;
}

}  // namespace bdlm
