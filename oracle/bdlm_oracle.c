/*
 * bdlm_oracle.c -- CPU oracle for the Kalman hot path of jonnylaw/bayesian_dlms.
 *
 * TEST INFRASTRUCTURE ONLY.  This file is the checker the CUDA path is compared
 * against.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load it.  The product (libbdlm.so) never links,
 * calls or falls back to anything in oracle/.
 *
 * It is a plain-C, fp64, one-series-at-a-time restatement of the reference's
 * Scala algorithm, written from the maths in the cited lines (paths relative to
 * /root/reference/core/src/main/scala/dlm/model/).  Evaluation order follows
 * the Scala source literally: Breeze `*`, `+`, `-` are left-associative method
 * calls, products are summed in increasing inner index (reference dgemm/dgemv
 * order), the JVM never fuses multiply-add (build with -ffp-contract=off).
 *
 * Third-party arithmetic that is NOT under /root/reference (Breeze 0.13.2 ->
 * netlib-java 1.1.2 LAPACK, build.sbt:69-70) is restated from the published
 * algorithms:
 *   `\`      dgesv  = dgetf2 (partial pivoting, first-max pivot, reciprocal
 *                     scaling of the column) + dgetrs (dlaswp, unit-lower then
 *                     non-unit-upper dtrsm with divisions)            -> lu_solve()
 *   eigSym   dsyev  ('V','L'), eigenvalues ascending.  dsyev's QL iteration is
 *                     replaced by a round-robin cyclic Jacobi iteration whose
 *                     operation order is fixed and mirrored by the CUDA kernels;
 *                     eigenvector sign rule: largest-|component| positive.
 *                                                                      -> jacobi_eigsym()
 *   svd      dgesdd, singular values descending, rightVectors = V^T.  Only s and
 *                     V^T are ever read by the reference, so a one-sided
 *                     (Hestenes) Jacobi on the columns is used; same sign rule.
 *                                                                      -> jacobi_svd()
 *   cholesky dpotrf ('L')                                              -> chol_logdet()
 * Parity pins: golden CSVs (n = 1, bit exact), KalmanFilterTest / SmoothingTest
 * / SvdFilterTest known answers (1e-4 .. 1e-2) -- see tests/test_oracle_golden.py.
 * For n > 1 beyond 1e-4, every FFBS sample value and every SVD-sampler value the
 * reference's own tests pin nothing ("parity unpinned"); a second, independent
 * numpy/LAPACK restatement (oracle/lapack_flavour.py) bounds the difference.
 *
 * Layout of every array here: series-major, one row per time point, matrices
 * column-major inside a row (Breeze DenseMatrix.data order).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ORACLE_API __attribute__((visibility("default")))

#define ST_OK 0
#define ST_SINGULAR 1      /* dgesv info > 0  (Breeze MatrixSingularException) */
#define ST_NOTCONVERGED 2  /* Jacobi sweep cap (Breeze NotConvergedException)   */
#define ST_NOTPD 4         /* dpotrf info > 0 (Breeze NotSymmetric/NotPD)       */

#define JACOBI_MAX_SWEEPS 30
static const double JACOBI_THR2 = 1e-30; /* rotate iff a_pq^2 > (1e-15)^2 |a_pp a_qq| */

/* ---------------------------------------------------------------- helpers */

/* out(ar x bc) = A(ar x ac) * B(ac x bc); ta/tb: operand is stored transposed.
 * A element (i,k) is A[i + k*lda] (or A[k + i*lda] when ta). */
static void mm(int ar, int ac, int bc, const double *A, int lda, int ta,
               const double *B, int ldb, int tb, double *out, int ldo) {
  for (int j = 0; j < bc; ++j)
    for (int i = 0; i < ar; ++i) {
      double acc = 0.0;
      for (int k = 0; k < ac; ++k) {
        double a = ta ? A[k + i * lda] : A[i + k * lda];
        double b = tb ? B[j + k * ldb] : B[k + j * ldb];
        double prod = a * b;
        acc = (k == 0) ? prod : acc + prod;
      }
      out[i + j * ldo] = acc;
    }
}

/* y(ar) = A(ar x ac) * x(ac) */
static void mv(int ar, int ac, const double *A, int lda, int ta, const double *x,
               double *y) {
  mm(ar, ac, 1, A, lda, ta, x, ac, 0, y, ar);
}

/* Solve A X = B in place (B <- X), A is n x n (destroyed), B is n x nrhs.
 * dgesv semantics as listed in the header.  Returns ST_SINGULAR on zero pivot. */
static int lu_solve(int n, double *A, int nrhs, double *B) {
  int st = ST_OK;
  int piv[64];
  for (int j = 0; j < n; ++j) {
    int jp = j;
    double best = fabs(A[j + j * n]);
    for (int i = j + 1; i < n; ++i) {
      double v = fabs(A[i + j * n]);
      if (v > best) { best = v; jp = i; }
    }
    piv[j] = jp;
    if (A[jp + j * n] != 0.0) {
      if (jp != j)
        for (int c = 0; c < n; ++c) {
          double t = A[j + c * n]; A[j + c * n] = A[jp + c * n]; A[jp + c * n] = t;
        }
      double r = 1.0 / A[j + j * n];
      for (int i = j + 1; i < n; ++i) A[i + j * n] = A[i + j * n] * r;
    } else {
      st = ST_SINGULAR;
    }
    for (int c = j + 1; c < n; ++c)
      for (int i = j + 1; i < n; ++i)
        A[i + c * n] = A[i + c * n] - A[i + j * n] * A[j + c * n];
  }
  for (int j = 0; j < n; ++j)
    if (piv[j] != j)
      for (int c = 0; c < nrhs; ++c) {
        double t = B[j + c * n]; B[j + c * n] = B[piv[j] + c * n]; B[piv[j] + c * n] = t;
      }
  for (int c = 0; c < nrhs; ++c) {
    for (int k = 0; k < n; ++k)
      for (int i = k + 1; i < n; ++i)
        B[i + c * n] = B[i + c * n] - B[k + c * n] * A[i + k * n];
    for (int k = n - 1; k >= 0; --k) {
      B[k + c * n] = B[k + c * n] / A[k + k * n];
      for (int i = 0; i < k; ++i)
        B[i + c * n] = B[i + c * n] - B[k + c * n] * A[i + k * n];
    }
  }
  return st;
}

/* Round-robin ("circle") pairing: m = n rounded up to even, rounds 0..m-2, each
 * round m/2 disjoint pairs; index m-1 stays fixed.  Pairs touching a dummy
 * index (>= n) are dropped.  partner[i] = index paired with i in this round
 * or -1. */
static void rr_partners(int n, int round, int *partner) {
  int m = (n + 1) & ~1;
  for (int i = 0; i < n; ++i) partner[i] = -1;
  int a = m - 1, b = round % (m - 1);
  if (a < n && b < n) { partner[a] = b; partner[b] = a; }
  for (int k = 1; k < m / 2; ++k) {
    a = (round + k) % (m - 1);
    b = (round - k + (m - 1)) % (m - 1);
    if (a < n && b < n) { partner[a] = b; partner[b] = a; }
  }
}

/* Jacobi rotation (c, s) for the symmetric 2x2 [[app, apq],[apq, aqq]] such
 * that J^T A J is diagonal with J = [[c, s],[-s, c]]. */
static void sym_rot(double app, double aqq, double apq, double *c, double *s) {
  double theta = (aqq - app) / (2.0 * apq);
  double t;
  if (fabs(theta) > 1e150) {
    t = 0.5 / theta;
  } else {
    double r = sqrt(theta * theta + 1.0);
    t = 1.0 / (fabs(theta) + r);
    if (theta < 0.0) t = -t;
  }
  *c = 1.0 / sqrt(t * t + 1.0);
  *s = t * (*c);
}

/* sort helper: order[] = permutation such that key[order[k]] is ascending
 * (descending when desc), stable selection order. */
static void sort_perm(int n, const double *key, int desc, int *order) {
  for (int i = 0; i < n; ++i) order[i] = i;
  for (int i = 1; i < n; ++i) { /* insertion sort, stable */
    int oi = order[i];
    int j = i - 1;
    while (j >= 0 && (desc ? key[order[j]] < key[oi] : key[order[j]] > key[oi])) {
      order[j + 1] = order[j];
      --j;
    }
    order[j + 1] = oi;
  }
}

/* Apply column permutation + sign rule: out[:,k] = +-V[:,order[k]] with the
 * largest-magnitude component (first such) made positive. */
static void order_and_sign(int rows, int n, const double *V, const int *order,
                           double *out) {
  for (int k = 0; k < n; ++k) {
    const double *col = V + (size_t)order[k] * rows;
    int im = 0;
    double best = fabs(col[0]);
    for (int i = 1; i < rows; ++i)
      if (fabs(col[i]) > best) { best = fabs(col[i]); im = i; }
    int flip = col[im] < 0.0;
    for (int i = 0; i < rows; ++i) out[i + (size_t)k * rows] = flip ? -col[i] : col[i];
  }
}

/* One-sided (Hestenes) Jacobi on the columns of U (r x n, in place), accumulating the same
 * rotations in V (n x n, in place): round-robin pair order, rotate a pair iff
 * (u_p . u_q)^2 > 1e-30 |u_p|^2 |u_q|^2.  Pairs of a round are disjoint, so the order inside a
 * round does not matter (the GPU kernels rotate them concurrently).  Returns ST_OK once a whole
 * sweep saw no rotation. */
static int jacobi_onesided(int r, int n, double *U, double *V) {
  int partner[64];
  int st = ST_NOTCONVERGED;
  int m = (n + 1) & ~1;
  if (n == 1) st = ST_OK;
  for (int sweep = 0; sweep < JACOBI_MAX_SWEEPS && n > 1; ++sweep) {
    int rotated = 0;
    for (int round = 0; round < m - 1; ++round) {
      rr_partners(n, round, partner);
      for (int p = 0; p < n; ++p) {
        int q = partner[p];
        if (q < 0 || q < p) continue;
        double alpha = 0.0, beta = 0.0, gamma = 0.0;
        for (int i = 0; i < r; ++i) {
          double up = U[i + p * r], uq = U[i + q * r];
          double pp = up * up, qq = uq * uq, pq = up * uq;
          alpha = i == 0 ? pp : alpha + pp;
          beta = i == 0 ? qq : beta + qq;
          gamma = i == 0 ? pq : gamma + pq;
        }
        if (!(gamma * gamma > JACOBI_THR2 * (alpha * beta))) continue;
        rotated = 1;
        double c, s;
        sym_rot(alpha, beta, gamma, &c, &s);
        for (int i = 0; i < r; ++i) {
          double up = U[i + p * r], uq = U[i + q * r];
          U[i + p * r] = c * up - s * uq;
          U[i + q * r] = s * up + c * uq;
        }
        for (int i = 0; i < n; ++i) {
          double vp = V[i + p * n], vq = V[i + q * n];
          V[i + p * n] = c * vp - s * vq;
          V[i + q * n] = s * vp + c * vq;
        }
      }
    }
    if (!rotated) { st = ST_OK; break; }
  }
  return st;
}

/* eigSym restatement (MultivariateGaussianSvd.scala:13; Breeze -> LAPACK dsyev 'V','L', which is
 * not under /root/reference).  A: n x n symmetric, only the lower triangle is read.  lam
 * ascending, Vout columns = eigenvectors (sign rule: largest-|component| positive).
 *
 * Stand-in algorithm (round 2): ONE-SIDED Jacobi on the columns of U = A with V accumulated
 * from I.  At convergence U = A V has orthogonal columns, i.e. A v_j = lam_j v_j with
 * |lam_j| = |u_j| (the column norm, accurate to a few ulps RELATIVE to lam_j -- better than the
 * two-sided iteration it replaces, whose errors are relative to |A|) and sign(lam_j) =
 * sign(v_j . u_j).  Chosen because a round costs two length-n dot products and two column
 * rotations per pair with no dependence on the other pairs' rotation parameters: 3x fewer
 * instructions per round than the two-sided update J^T A J on the GPU.  Pinned against LAPACK
 * dsyev by tests/test_oracle_golden.py exactly as before; eigenvector signs (and bases inside
 * repeated eigenvalues) are implementation-defined in LAPACK itself. */
static int jacobi_eigsym(int n, const double *Ain, double *lam, double *Vout) {
  double U[64 * 64], V[64 * 64], d[64];
  int order[64];
  for (int j = 0; j < n; ++j)
    for (int i = 0; i < n; ++i) {
      U[i + j * n] = (i >= j) ? Ain[i + j * n] : Ain[j + i * n];
      V[i + j * n] = (i == j) ? 1.0 : 0.0;
    }
  int st = jacobi_onesided(n, n, U, V);
  for (int j = 0; j < n; ++j) {
    double nn = 0.0, dot = 0.0;
    for (int i = 0; i < n; ++i) {
      double u = U[i + j * n];
      double sq = u * u, vu = V[i + j * n] * u;
      nn = i == 0 ? sq : nn + sq;
      dot = i == 0 ? vu : dot + vu;
    }
    double nrm = sqrt(nn);
    d[j] = dot < 0.0 ? -nrm : nrm;
  }
  sort_perm(n, d, 0, order);
  for (int k = 0; k < n; ++k) lam[k] = d[order[k]];
  order_and_sign(n, n, V, order, Vout);
  return st;
}

/* svd restatement: M is r x n (r >= 1), returns singular values sv[n]
 * (descending) and Vout (n x n, columns = right singular vectors, i.e.
 * Breeze rightVectors.t).  One-sided Jacobi on the columns of a copy of M. */
static int jacobi_svd(int r, int n, const double *M, double *sv, double *Vout) {
  double U[128 * 64], V[64 * 64];
  int order[64];
  memcpy(U, M, sizeof(double) * r * n);
  for (int j = 0; j < n; ++j)
    for (int i = 0; i < n; ++i) V[i + j * n] = (i == j) ? 1.0 : 0.0;
  int st = jacobi_onesided(r, n, U, V);
  double nrm[64];
  for (int j = 0; j < n; ++j) {
    double acc = 0.0;
    for (int i = 0; i < r; ++i) {
      double sq = U[i + j * r] * U[i + j * r];
      acc = i == 0 ? sq : acc + sq;
    }
    nrm[j] = sqrt(acc);
  }
  sort_perm(n, nrm, 1, order);
  for (int k = 0; k < n; ++k) sv[k] = nrm[order[k]];
  order_and_sign(n, n, V, order, Vout);
  return st;
}

/* sum(log(diag(cholesky(S)))) for S n x n SPD (dpotrf 'L'). */
static int chol_logdet(int n, const double *S, double *sumlogdiag) {
  double L[64 * 64];
  int st = ST_OK;
  memcpy(L, S, sizeof(double) * n * n);
  double acc = 0.0;
  for (int j = 0; j < n; ++j) {
    double d = L[j + j * n];
    for (int k = 0; k < j; ++k) d = d - L[j + k * n] * L[j + k * n];
    if (!(d > 0.0)) st = ST_NOTPD;
    d = sqrt(d);
    L[j + j * n] = d;
    for (int i = j + 1; i < n; ++i) {
      double v = L[i + j * n];
      for (int k = 0; k < j; ++k) v = v - L[i + k * n] * L[j + k * n];
      L[i + j * n] = v / d;
    }
    acc = acc + log(d);
  }
  *sumlogdiag = acc;
  return st;
}

static const double LOG_2PI = 1.8378770664093453;

/* breeze MultivariateGaussian(mu, S).logPdf(x) = -(S\(x-mu)).(x-mu)/2
 *                                   - (n/2 log(2 pi) + sum log diag chol S) */
static int mvn_logpdf(int n, const double *x, const double *mu, const double *S,
                      double *out) {
  double c[64], slv[64], A[64 * 64];
  int st = ST_OK;
  for (int i = 0; i < n; ++i) { c[i] = x[i] - mu[i]; slv[i] = c[i]; }
  memcpy(A, S, sizeof(double) * n * n);
  st |= lu_solve(n, A, 1, slv);
  double dot = 0.0;
  for (int i = 0; i < n; ++i) dot = i == 0 ? slv[i] * c[i] : dot + slv[i] * c[i];
  double ld;
  st |= chol_logdet(n, S, &ld);
  *out = -dot / 2.0 - (n / 2.0 * LOG_2PI + ld);
  return st;
}

/* ------------------------------------------------------ Kalman filter step */

/* KalmanFilter.advState, KalmanFilter.scala:273-286 */
static void kf_advance(int n, const double *G, const double *W, double dt,
                       const double *m, const double *C, double *a, double *R,
                       double *tmp) {
  if (dt == 0.0) {
    memcpy(a, m, sizeof(double) * n);
    memcpy(R, C, sizeof(double) * n * n);
    return;
  }
  mv(n, n, G, n, 0, m, a);
  mm(n, n, n, G, n, 0, C, n, 0, tmp, n); /* G * C           */
  mm(n, n, n, tmp, n, 0, G, n, 1, R, n); /* (G * C) * G^T   */
  for (int k = 0; k < n * n; ++k) R[k] = R[k] + W[k] * dt;
}

/* KalmanFilter.oneStepPrediction :311-321 and updateState :64-94.
 * y[p] with NaN = None.  Writes f,Q (full p) and m,C. */
static int kf_update(int n, int p, const double *F, const double *V,
                     const double *a, const double *R, const double *y,
                     double *f, double *Q, double *m, double *C) {
  double t1[64 * 64], t2[64 * 64];
  int st = ST_OK;
  /* f = F^T a ; Q = (F^T R) F + V */
  mv(p, n, F, n, 1, a, f);
  mm(p, n, n, F, n, 1, R, n, 0, t1, p);
  mm(p, n, p, t1, p, 0, F, n, 0, Q, p);
  for (int k = 0; k < p * p; ++k) Q[k] = Q[k] + V[k];

  int obs[64], po = 0;
  for (int i = 0; i < p; ++i)
    if (!isnan(y[i])) obs[po++] = i;
  if (po == 0) { /* :74-75 */
    memcpy(m, a, sizeof(double) * n);
    memcpy(C, R, sizeof(double) * n * n);
    return st;
  }
  double Fm[64 * 64], Vm[64 * 64], Qm[64 * 64], fm[64], e[64], K[64 * 64],
      D[64 * 64], rhs[64 * 64], At[64 * 64];
  for (int k = 0; k < po; ++k) { /* missingF :202-208, missingV :213-218 */
    for (int i = 0; i < n; ++i) Fm[i + k * n] = F[i + obs[k] * n];
    for (int l = 0; l < po; ++l) Vm[l + k * po] = V[obs[l] + obs[k] * p];
  }
  /* oneStepMissing :44-53 */
  mv(po, n, Fm, n, 1, a, fm);
  mm(po, n, n, Fm, n, 1, R, n, 0, t1, po);
  mm(po, n, po, t1, po, 0, Fm, n, 0, Qm, po);
  for (int k = 0; k < po * po; ++k) Qm[k] = Qm[k] + Vm[k];
  for (int k = 0; k < po; ++k) e[k] = y[obs[k]] - fm[k];
  /* K = (Qm^T \ (Fm^T R^T))^T  :83 */
  mm(po, n, n, Fm, n, 1, R, n, 1, rhs, po);
  for (int j = 0; j < po; ++j)
    for (int i = 0; i < po; ++i) At[i + j * po] = Qm[j + i * po];
  st |= lu_solve(po, At, n, rhs);
  for (int j = 0; j < po; ++j)
    for (int i = 0; i < n; ++i) K[i + j * n] = rhs[j + i * po];
  /* m = a + K e  :84 */
  mv(n, po, K, n, 0, e, t1);
  for (int i = 0; i < n; ++i) m[i] = a[i] + t1[i];
  /* D = I - K Fm^T ; C = (D R) D^T + (K Vm) K^T  :89-90 */
  mm(n, po, n, K, n, 0, Fm, n, 1, D, n);
  for (int j = 0; j < n; ++j)
    for (int i = 0; i < n; ++i) D[i + j * n] = ((i == j) ? 1.0 : 0.0) - D[i + j * n];
  mm(n, n, n, D, n, 0, R, n, 0, t1, n);
  mm(n, n, n, t1, n, 0, D, n, 1, C, n);
  mm(n, po, po, K, n, 0, Vm, po, 0, t1, n);
  mm(n, po, n, t1, n, 0, K, n, 1, t2, n);
  for (int k = 0; k < n * n; ++k) C[k] = C[k] + t2[k];
  return st;
}

static double min_time(int T, const double *times) {
  double t0 = times[0];
  for (int i = 1; i < T; ++i) t0 = fmin(t0, times[i]);
  return t0;
}

/* KalmanFilter.filterDlm :291-294 (keep_init = 0, T rows) and
 * KalmanFilter(adv).filter, Filter.scala:41-45 (keep_init = 1, T+1 rows, row 0
 * = initialiseState :112-118 with f,Q = NaN standing for None).
 * F: n x p, or T of them when f_tv; G likewise (G[t] = g(dt_t) advances INTO
 * observation t).  times_out[rows] receives the state times. */
static int kf_filter_impl(int n, int p, int T, const double *F, int f_tv,
                          const double *G, int g_tv, const double *V, int v_tv,
                          const double *t_init,
                          const double *W, int w_tv, const double *m0,
                          const double *C0, const double *times,
                          const double *y, int keep_init, double *times_out,
                          double *m, double *C, double *a, double *R,
                          double *f, double *Q) {
  if (T <= 0) return -1; /* t0.get on empty data, KalmanFilter.scala:116-117 */
  int st = ST_OK, nn = n * n, pp = p * p;
  double mc[64], Cc[64 * 64], tmp[64 * 64];
  /* t_init: resume from a saved state (m0, C0) at that time -- what folding KalmanFilter.step
   * over later observations does (NoModel.scala:153-155), and Dlm.forecast's start
   * (Dlm.scala:322-338); NULL = initialiseState's min(time) - 1. */
  double tprev = t_init ? *t_init : min_time(T, times) - 1.0;
  memcpy(mc, m0, sizeof(double) * n);
  memcpy(Cc, C0, sizeof(double) * nn);
  int row = 0;
  if (keep_init) {
    times_out[0] = tprev;
    memcpy(m, m0, sizeof(double) * n); memcpy(a, m0, sizeof(double) * n);
    memcpy(C, C0, sizeof(double) * nn); memcpy(R, C0, sizeof(double) * nn);
    for (int k = 0; k < p; ++k) f[k] = NAN;
    for (int k = 0; k < pp; ++k) Q[k] = NAN;
    row = 1;
  }
  for (int t = 0; t < T; ++t, ++row) {
    const double *Ft = F + (f_tv ? (size_t)t * n * p : 0);
    const double *Gt = G + (g_tv ? (size_t)t * nn : 0);
    double dt = times[t] - tprev; /* KalmanFilter.step :99-107 */
    double *ar = a + (size_t)row * n, *Rr = R + (size_t)row * nn;
    double *mr = m + (size_t)row * n, *Cr = C + (size_t)row * nn;
    kf_advance(n, Gt, W + (w_tv ? (size_t)t * nn : 0), dt, mc, Cc, ar, Rr, tmp);
    st |= kf_update(n, p, Ft, V + (v_tv ? (size_t)t * pp : 0), ar, Rr, y + (size_t)t * p,
                    f + (size_t)row * p, Q + (size_t)row * pp, mr, Cr);
    memcpy(mc, mr, sizeof(double) * n);
    memcpy(Cc, Cr, sizeof(double) * nn);
    tprev = times[t];
    times_out[row] = tprev;
  }
  return st;
}

ORACLE_API int oracle_kf_filter(int n, int p, int T, const double *F, int f_tv,
                                const double *G, int g_tv, const double *V,
                                const double *W, const double *m0,
                                const double *C0, const double *times,
                                const double *y, int keep_init, double *times_out,
                                double *m, double *C, double *a, double *R,
                                double *f, double *Q) {
  return kf_filter_impl(n, p, T, F, f_tv, G, g_tv, V, 0, NULL, W, 0, m0, C0, times, y, keep_init,
                        times_out, m, C, a, R, f, Q);
}

/* Filter resumed from the state (m0, C0) at time t_init (see kf_filter_impl). */
ORACLE_API int oracle_kf_filter_from(int n, int p, int T, const double *F, int f_tv,
                                     const double *G, int g_tv, const double *V,
                                     const double *W, const double *m0,
                                     const double *C0, double t_init, const double *times,
                                     const double *y, int keep_init, double *times_out,
                                     double *m, double *C, double *a, double *R,
                                     double *f, double *Q) {
  return kf_filter_impl(n, p, T, F, f_tv, G, g_tv, V, 0, &t_init, W, 0, m0, C0, times, y,
                        keep_init, times_out, m, C, a, R, f, Q);
}

/* Time-varying observation variance V_t (next row f2): StudentTGibbs.filter
 * (StudentTGibbs.scala:100-119) runs KalmanFilter.step with params.copy(v = V_t) at step t.
 * V: T matrices of p x p. */
ORACLE_API int oracle_kf_filter_vt(int n, int p, int T, const double *F, int f_tv,
                                   const double *G, int g_tv, const double *V,
                                   const double *W, const double *m0,
                                   const double *C0, const double *times,
                                   const double *y, int keep_init, double *times_out,
                                   double *m, double *C, double *a, double *R,
                                   double *f, double *Q) {
  return kf_filter_impl(n, p, T, F, f_tv, G, g_tv, V, 1, NULL, W, 0, m0, C0, times, y, keep_init,
                        times_out, m, C, a, R, f, Q);
}

/* Time-varying V_t and / or W_t: DlmFsvSystem.ffbs (DlmFsvSystem.scala:137-167) runs
 * KalmanFilter.step with params.copy(w = W_t) at step t (W: T matrices of n x n). */
ORACLE_API int oracle_kf_filter_tv(int n, int p, int T, const double *F, int f_tv,
                                   const double *G, int g_tv, const double *V, int v_tv,
                                   const double *W, int w_tv, const double *m0,
                                   const double *C0, const double *times,
                                   const double *y, int keep_init, double *times_out,
                                   double *m, double *C, double *a, double *R,
                                   double *f, double *Q) {
  return kf_filter_impl(n, p, T, F, f_tv, G, g_tv, V, v_tv, NULL, W, w_tv, m0, C0, times, y,
                        keep_init, times_out, m, C, a, R, f, Q);
}

/* B = (R1^T \ (G C^T))^T  -- Smoothing.scala:41 and :85 */
static int smoothing_gain(int n, const double *G, const double *C,
                          const double *R1, double *B) {
  double rhs[64 * 64], At[64 * 64];
  mm(n, n, n, G, n, 0, C, n, 1, rhs, n);
  for (int j = 0; j < n; ++j)
    for (int i = 0; i < n; ++i) At[i + j * n] = R1[j + i * n];
  int st = lu_solve(n, At, n, rhs);
  for (int j = 0; j < n; ++j)
    for (int i = 0; i < n; ++i) B[i + j * n] = rhs[j + i * n];
  return st;
}

/* Smoothing.backwardsSmoother :57-64 with smoothStep :31-47.
 * rows = T + keep_init filtered states (as written by oracle_kf_filter).
 * textbook != 0 replaces the reference's  C - B (R1 - S1) B  (quirk Q1) by the
 * textbook  ... B^T. */
ORACLE_API int oracle_rts_smooth(int n, int T, int keep_init, const double *G,
                                 int g_tv, const double *m, const double *C,
                                 const double *a, const double *R, int textbook,
                                 double *s, double *S) {
  int rows = T + keep_init, nn = n * n, st = ST_OK;
  double B[64 * 64], d[64], D[64 * 64], t1[64 * 64], t2[64 * 64];
  memcpy(s + (size_t)(rows - 1) * n, m + (size_t)(rows - 1) * n, sizeof(double) * n);
  memcpy(S + (size_t)(rows - 1) * nn, C + (size_t)(rows - 1) * nn, sizeof(double) * nn);
  for (int r = rows - 2; r >= 0; --r) {
    int tobs = r + 1 - keep_init; /* observation index of row r+1 */
    const double *Gt = G + (g_tv ? (size_t)tobs * nn : 0);
    const double *a1 = a + (size_t)(r + 1) * n, *R1 = R + (size_t)(r + 1) * nn;
    const double *s1 = s + (size_t)(r + 1) * n, *S1 = S + (size_t)(r + 1) * nn;
    const double *mr = m + (size_t)r * n, *Cr = C + (size_t)r * nn;
    st |= smoothing_gain(n, Gt, Cr, R1, B);
    for (int i = 0; i < n; ++i) d[i] = s1[i] - a1[i];
    mv(n, n, B, n, 0, d, t1);
    for (int i = 0; i < n; ++i) s[(size_t)r * n + i] = mr[i] + t1[i];
    for (int k = 0; k < nn; ++k) D[k] = R1[k] - S1[k];
    mm(n, n, n, B, n, 0, D, n, 0, t1, n);
    mm(n, n, n, t1, n, 0, B, n, textbook ? 1 : 0, t2, n);
    for (int k = 0; k < nn; ++k) S[(size_t)r * nn + k] = Cr[k] - t2[k];
  }
  return st;
}

/* MultivariateGaussianSvd(mu, cov).draw :13-22 with injected normals z[n]. */
static int mvn_eig_draw(int n, const double *mu, const double *cov,
                        const double *z, double *out) {
  double lam[64], V[64 * 64], M[64 * 64], x[64];
  int st = jacobi_eigsym(n, cov, lam, V);
  for (int j = 0; j < n; ++j) {
    double sq = sqrt(lam[j]);
    for (int i = 0; i < n; ++i) M[i + j * n] = V[i + j * n] * sq;
  }
  mv(n, n, M, n, 0, z, x);
  for (int i = 0; i < n; ++i) out[i] = mu[i] + x[i];
  return st;
}

/* Smoothing.sample :114-122 with Smoothing.step :74-103 / initialise :105-109.
 * z[rows][n]: the N(0,1) values consumed for row r (the reference draws row
 * rows-1 first).  dts[r] for r < rows-1 is time[r+1]-time[r]. */
static int backward_sample_impl(int n, int T, int keep_init, const double *G,
                                int g_tv, const double *W0, int w_tv,
                                const double *times_rows, const double *m,
                                const double *C, const double *a,
                                const double *R, const double *z,
                                double *theta) {
  int rows = T + keep_init, nn = n * n, st = ST_OK;
  double B[64 * 64], d[64], h[64], D[64 * 64], t1[64 * 64], t2[64 * 64],
      H[64 * 64], Hs[64 * 64];
  st |= mvn_eig_draw(n, m + (size_t)(rows - 1) * n, C + (size_t)(rows - 1) * nn,
                     z + (size_t)(rows - 1) * n, theta + (size_t)(rows - 1) * n);
  for (int r = rows - 2; r >= 0; --r) {
    int tobs = r + 1 - keep_init;
    const double *Gt = G + (g_tv ? (size_t)tobs * nn : 0);
    /* W of the transition r -> r + 1 (DlmFsvSystem.scala:155-163 zips ps with filtered.init) */
    const double *W = W0 + (w_tv ? (size_t)tobs * nn : 0);
    double dt = times_rows[r + 1] - times_rows[r];
    const double *a1 = a + (size_t)(r + 1) * n, *R1 = R + (size_t)(r + 1) * nn;
    const double *th1 = theta + (size_t)(r + 1) * n;
    const double *mr = m + (size_t)r * n, *Cr = C + (size_t)r * nn;
    st |= smoothing_gain(n, Gt, Cr, R1, B);
    for (int i = 0; i < n; ++i) d[i] = th1[i] - a1[i];
    mv(n, n, B, n, 0, d, t1);
    for (int i = 0; i < n; ++i) h[i] = mr[i] + t1[i];
    /* diff = I - B G ; cov = (diff C) diff^T + ((B W) dt) B^T  :93-94 */
    mm(n, n, n, B, n, 0, Gt, n, 0, D, n);
    for (int j = 0; j < n; ++j)
      for (int i = 0; i < n; ++i) D[i + j * n] = ((i == j) ? 1.0 : 0.0) - D[i + j * n];
    mm(n, n, n, D, n, 0, Cr, n, 0, t1, n);
    mm(n, n, n, t1, n, 0, D, n, 1, H, n);
    mm(n, n, n, B, n, 0, W, n, 0, t1, n);
    for (int k = 0; k < nn; ++k) t1[k] = t1[k] * dt;
    mm(n, n, n, t1, n, 0, B, n, 1, t2, n);
    for (int k = 0; k < nn; ++k) H[k] = H[k] + t2[k];
    for (int j = 0; j < n; ++j) /* :95 */
      for (int i = 0; i < n; ++i) Hs[i + j * n] = (H[i + j * n] + H[j + i * n]) / 2.0;
    st |= mvn_eig_draw(n, h, Hs, z + (size_t)r * n, theta + (size_t)r * n);
  }
  return st;
}

ORACLE_API int oracle_backward_sample(int n, int T, int keep_init, const double *G,
                                      int g_tv, const double *W,
                                      const double *times_rows, const double *m,
                                      const double *C, const double *a,
                                      const double *R, const double *z,
                                      double *theta) {
  return backward_sample_impl(n, T, keep_init, G, g_tv, W, 0, times_rows, m, C, a, R, z, theta);
}

/* FFBS with time-varying V_t / W_t (StudentTGibbs.sampleState, DlmFsvSystem.ffbs). */
ORACLE_API int oracle_ffbs_tv(int n, int p, int T, const double *F, int f_tv,
                              const double *G, int g_tv, const double *V, int v_tv,
                              const double *W, int w_tv, const double *m0, const double *C0,
                              const double *times, const double *y, const double *z,
                              double *times_out, double *theta, double *m, double *C,
                              double *a, double *R) {
  int rows = T + 1;
  double *f = (double *)malloc(sizeof(double) * rows * p);
  double *Q = (double *)malloc(sizeof(double) * rows * p * p);
  int st = oracle_kf_filter_tv(n, p, T, F, f_tv, G, g_tv, V, v_tv, W, w_tv, m0, C0, times, y, 1,
                               times_out, m, C, a, R, f, Q);
  if (st >= 0)
    st |= backward_sample_impl(n, T, 1, G, g_tv, W, w_tv, times_out, m, C, a, R, z, theta);
  free(f); free(Q);
  return st;
}

/* Smoothing.ffbs :151-159 / ffbsDlm :173-180: filter keeping the initial state,
 * then backward sampling.  Outputs T+1 rows. */
ORACLE_API int oracle_ffbs(int n, int p, int T, const double *F, int f_tv,
                           const double *G, int g_tv, const double *V,
                           const double *W, const double *m0, const double *C0,
                           const double *times, const double *y, const double *z,
                           double *times_out, double *theta, double *m, double *C,
                           double *a, double *R) {
  int rows = T + 1;
  double *f = (double *)malloc(sizeof(double) * rows * p);
  double *Q = (double *)malloc(sizeof(double) * rows * p * p);
  int st = oracle_kf_filter(n, p, T, F, f_tv, G, g_tv, V, W, m0, C0, times, y, 1,
                            times_out, m, C, a, R, f, Q);
  if (st >= 0)
    st |= oracle_backward_sample(n, T, 1, G, g_tv, W, times_out, m, C, a, R, z, theta);
  free(f); free(Q);
  return st;
}

/* StudentTGibbs.sampleState (StudentTGibbs.scala:128-136): filter with V_t, then
 * Smoothing.sampleDlm. */
ORACLE_API int oracle_ffbs_vt(int n, int p, int T, const double *F, int f_tv,
                              const double *G, int g_tv, const double *V,
                              const double *W, const double *m0, const double *C0,
                              const double *times, const double *y, const double *z,
                              double *times_out, double *theta, double *m, double *C,
                              double *a, double *R) {
  int rows = T + 1;
  double *f = (double *)malloc(sizeof(double) * rows * p);
  double *Q = (double *)malloc(sizeof(double) * rows * p * p);
  int st = oracle_kf_filter_vt(n, p, T, F, f_tv, G, g_tv, V, W, m0, C0, times, y, 1,
                               times_out, m, C, a, R, f, Q);
  if (st >= 0)
    st |= oracle_backward_sample(n, T, 1, G, g_tv, W, times_out, m, C, a, R, z, theta);
  free(f); free(Q);
  return st;
}

/* KalmanFilter.likelihood :299-306 (quirk Q4: transition density of the
 * filtered means) and, in *innov, the innovations form built from
 * conditionalLikelihood :138-153 summed over t. */
ORACLE_API int oracle_loglik(int n, int p, int T, const double *F, int f_tv,
                             const double *G, int g_tv, const double *V,
                             const double *W, const double *m0, const double *C0,
                             const double *times, const double *y,
                             double *transition, double *innov) {
  if (T <= 0) return -1;
  int rows = T + 1, nn = n * n, pp = p * p, st;
  double *buf = (double *)malloc(sizeof(double) * rows * (2 * n + 2 * nn + p + pp + 1));
  double *tm = buf, *m = tm + rows, *C = m + rows * n, *a = C + rows * nn,
         *R = a + rows * n, *f = R + rows * nn, *Q = f + rows * p;
  st = oracle_kf_filter(n, p, T, F, f_tv, G, g_tv, V, W, m0, C0, times, y, 1, tm,
                        m, C, a, R, f, Q);
  double ll = 0.0, li = 0.0;
  for (int r = 1; r < rows; ++r) {
    const double *Gt = G + (g_tv ? (size_t)(r - 1) * nn : 0);
    double dt = tm[r] - tm[r - 1];
    double mu[64], S[64 * 64], v;
    mv(n, n, Gt, n, 0, m + (size_t)(r - 1) * n, mu);
    for (int k = 0; k < nn; ++k) S[k] = W[k] * dt;
    st |= mvn_logpdf(n, m + (size_t)r * n, mu, S, &v);
    ll = r == 1 ? v : ll + v;
    /* innovations */
    const double *yr = y + (size_t)(r - 1) * p;
    int obs[64], po = 0;
    for (int i = 0; i < p; ++i)
      if (!isnan(yr[i])) obs[po++] = i;
    if (po == 1) {
      double fo = f[(size_t)r * p + obs[0]];
      double sd = sqrt(Q[(size_t)r * pp + obs[0] + obs[0] * p]);
      double dd = (yr[obs[0]] - fo) / sd;
      li += -dd * dd / 2.0 - log(sqrt(2.0 * M_PI) * sd);
    } else if (po > 1) {
      double fo[64], yo[64], Qo[64 * 64];
      for (int k = 0; k < po; ++k) {
        fo[k] = f[(size_t)r * p + obs[k]];
        yo[k] = yr[obs[k]];
        for (int l = 0; l < po; ++l)
          Qo[l + k * po] = Q[(size_t)r * pp + obs[l] + obs[k] * p];
      }
      st |= mvn_logpdf(po, yo, fo, Qo, &v);
      li += v;
    }
  }
  *transition = ll;
  *innov = li;
  free(buf);
  return st;
}

/* ----------------------------------------------------------- SVD filter */

/* SvdFilter.sqrtInvSvd :210-216 (inv != 0) / sqrtSvd :224-227: diag(g(s)) V^T */
ORACLE_API int oracle_sqrt_svd(int n, const double *M, int inv, double *out) {
  double sv[64], V[64 * 64];
  int st = jacobi_svd(n, n, M, sv, V);
  for (int i = 0; i < n; ++i) {
    double d = inv ? 1.0 / sqrt(sv[i]) : sqrt(sv[i]);
    for (int j = 0; j < n; ++j) out[i + j * n] = d * V[j + i * n];
  }
  return st;
}

/* SvdFilter.advState :183-202.  Wadv is whatever the caller's closure holds
 * (raw W on the filterDlm/ffbsDlm/Gibbs paths -- quirk Q2). */
static int svd_advance(int n, const double *G, const double *Wadv, double dt,
                       const double *m, const double *dc, const double *uc,
                       double *a, double *dr, double *ur) {
  if (dt == 0.0) {
    memcpy(a, m, sizeof(double) * n);
    memcpy(dr, dc, sizeof(double) * n);
    memcpy(ur, uc, sizeof(double) * n * n);
    return ST_OK;
  }
  double X[64 * 64], t1[64 * 64], stack[128 * 64];
  mv(n, n, G, n, 0, m, a);
  for (int j = 0; j < n; ++j) /* diag(dc) * uc^T */
    for (int i = 0; i < n; ++i) X[i + j * n] = dc[i] * uc[j + i * n];
  mm(n, n, n, X, n, 0, G, n, 1, t1, n);
  double sq = sqrt(dt);
  for (int j = 0; j < n; ++j)
    for (int i = 0; i < n; ++i) {
      stack[i + j * 2 * n] = t1[i + j * n];
      stack[n + i + j * 2 * n] = Wadv[i + j * n] * sq;
    }
  return jacobi_svd(2 * n, n, stack, dr, ur);
}

/* SvdFilter.updateState :38-68; Vfac = V^{-1/2} from transformParams :232-236 */
static int svd_update(int n, int p, const double *F, const double *Vfac,
                      const double *a, const double *dr, const double *ur,
                      const double *y, double *m, double *dc, double *uc) {
  int obs[64], po = 0, st = ST_OK;
  for (int i = 0; i < p; ++i)
    if (!isnan(y[i])) obs[po++] = i;
  if (po == 0) {
    memcpy(m, a, sizeof(double) * n);
    memcpy(dc, dr, sizeof(double) * n);
    memcpy(uc, ur, sizeof(double) * n * n);
    return st;
  }
  double Fm[64 * 64], Vm[64 * 64], fm[64], e[64], t1[64 * 64], t2[64 * 64],
      stack[128 * 64], sv[64], Vr[64 * 64], X[64 * 64], XtX[64 * 64], fv[64 * 64],
      gain[64 * 64];
  for (int k = 0; k < po; ++k) {
    for (int i = 0; i < n; ++i) Fm[i + k * n] = F[i + obs[k] * n];
    for (int l = 0; l < po; ++l) Vm[l + k * po] = Vfac[obs[l] + obs[k] * p];
  }
  mv(po, n, Fm, n, 1, a, fm);
  int r = po + n;
  mm(po, po, n, Vm, po, 0, Fm, n, 1, t1, po); /* vm * fm^T        */
  mm(po, n, n, t1, po, 0, ur, n, 0, t2, po);  /* (vm fm^T) * ur   */
  for (int j = 0; j < n; ++j) {
    for (int i = 0; i < po; ++i) stack[i + j * r] = t2[i + j * po];
    for (int i = 0; i < n; ++i) stack[po + i + j * r] = (i == j) ? 1.0 / dr[i] : 0.0;
  }
  st |= jacobi_svd(r, n, stack, sv, Vr);
  mm(n, n, n, ur, n, 0, Vr, n, 0, uc, n); /* ur * rightVectors^T */
  for (int k = 0; k < po; ++k) e[k] = y[obs[k]] - fm[k];
  mm(n, po, po, Fm, n, 0, Vm, po, 1, t1, n);  /* fm * vm^T          */
  mm(n, po, po, t1, n, 0, Vm, po, 0, fv, n);  /* (fm vm^T) * vm     */
  for (int i = 0; i < n; ++i) dc[i] = 1.0 / sv[i];
  for (int j = 0; j < n; ++j)
    for (int i = 0; i < n; ++i) X[i + j * n] = dc[i] * uc[j + i * n];
  mm(n, n, n, X, n, 1, X, n, 0, XtX, n);
  mm(n, n, po, XtX, n, 0, fv, n, 0, gain, n);
  mv(n, po, gain, n, 0, e, t1);
  for (int i = 0; i < n; ++i) m[i] = a[i] + t1[i];
  return st;
}

/* SvdFilter.filter/filterTraverse :100-119 (transform = 1: V,W are the raw
 * parameters, V^{-1/2} is derived here) or filterDecomp :134-140 (transform =
 * 0: Vfac is passed in already transformed).  Wadv is the W factor used by the
 * advance closure.  keep_init as for the Kalman filter.  Outputs per row:
 * m[n], dc[n], uc[n*n], a[n], dr[n], ur[n*n], f[p].
 *
 * v_tv / w_tv (next row f2): V / Wadv hold T matrices, one per observation, as in
 * DlmFsv.ffbsSvd (DlmFsv.scala:208-229: ps = vs.map(vi => transformParams(p.copy(v = vi))),
 * the step for observation t runs with ps(t)) and DlmFsvSystem.ffbsSvd
 * (DlmFsvSystem.scala:177-207, the same with W_t).  With transform = 1 every V_t goes
 * through sqrtInvSvd at its own step. */
static int svd_filter_impl(int n, int p, int T, const double *F, int f_tv,
                           const double *G, int g_tv, const double *V, int v_tv,
                           int transform, const double *Wadv, int w_tv,
                           const double *m0, const double *C0,
                           const double *times, const double *y,
                           int keep_init, double *times_out, double *m,
                           double *dc, double *uc, double *a, double *dr,
                           double *ur, double *f) {
  if (T <= 0) return -1;
  int st = ST_OK, nn = n * n;
  double Vfac[64 * 64], mc[64], dcc[64], ucc[64 * 64], sv[64];
  if (!v_tv) {
    if (transform) st |= oracle_sqrt_svd(p, V, 1, Vfac);
    else memcpy(Vfac, V, sizeof(double) * p * p);
  }
  /* initialiseState :83-95 */
  st |= jacobi_svd(n, n, C0, sv, ucc);
  for (int i = 0; i < n; ++i) dcc[i] = sqrt(sv[i]);
  memcpy(mc, m0, sizeof(double) * n);
  double tmin = min_time(T, times);
  double tprev = tmin - 1;
  int row = 0;
  if (keep_init) {
    times_out[0] = tprev;
    memcpy(m, m0, sizeof(double) * n); memcpy(a, m0, sizeof(double) * n);
    memcpy(dc, dcc, sizeof(double) * n); memcpy(dr, dcc, sizeof(double) * n);
    memcpy(uc, ucc, sizeof(double) * nn); memcpy(ur, ucc, sizeof(double) * nn);
    mv(p, n, F, n, 1, m0, f); /* oneStepForecast(mod.f, m0, t0)  :92 */
    row = 1;
  }
  for (int t = 0; t < T; ++t, ++row) {
    const double *Ft = F + (f_tv ? (size_t)t * n * p : 0);
    const double *Gt = G + (g_tv ? (size_t)t * nn : 0);
    const double *Wt = Wadv + (w_tv ? (size_t)t * nn : 0);
    if (v_tv) {
      const double *Vt = V + (size_t)t * p * p;
      if (transform) st |= oracle_sqrt_svd(p, Vt, 1, Vfac);
      else memcpy(Vfac, Vt, sizeof(double) * p * p);
    }
    double dt = times[t] - tprev;
    double *ar = a + (size_t)row * n, *drr = dr + (size_t)row * n,
           *urr = ur + (size_t)row * nn;
    st |= svd_advance(n, Gt, Wt, dt, mc, dcc, ucc, ar, drr, urr);
    mv(p, n, Ft, n, 1, ar, f + (size_t)row * p); /* :74 */
    st |= svd_update(n, p, Ft, Vfac, ar, drr, urr, y + (size_t)t * p,
                     m + (size_t)row * n, dc + (size_t)row * n, uc + (size_t)row * nn);
    memcpy(mc, m + (size_t)row * n, sizeof(double) * n);
    memcpy(dcc, dc + (size_t)row * n, sizeof(double) * n);
    memcpy(ucc, uc + (size_t)row * nn, sizeof(double) * nn);
    tprev = times[t];
    times_out[row] = tprev;
  }
  return st;
}

ORACLE_API int oracle_svd_filter(int n, int p, int T, const double *F, int f_tv,
                                 const double *G, int g_tv, const double *V,
                                 int transform, const double *Wadv,
                                 const double *m0, const double *C0,
                                 const double *times, const double *y,
                                 int keep_init, double *times_out, double *m,
                                 double *dc, double *uc, double *a, double *dr,
                                 double *ur, double *f) {
  return svd_filter_impl(n, p, T, F, f_tv, G, g_tv, V, 0, transform, Wadv, 0, m0, C0, times,
                         y, keep_init, times_out, m, dc, uc, a, dr, ur, f);
}

/* SvdSampler.sample :54-60 with step :15-36, initialise :38-45, rnorm :94-102.
 * sqrtW is the matrix handed to SvdSampler.step (W^{1/2} on the ffbs path). */
static int svd_backward_sample_impl(int n, int T, int keep_init,
                                    const double *G, int g_tv,
                                    const double *sqrtW_all, int w_tv, const double *m,
                                    const double *dc, const double *uc,
                                    const double *a, const double *z,
                                    double *theta) {
  int rows = T + keep_init, nn = n * n, st = ST_OK;
  double M[64 * 64], x[64], t1[64 * 64], t2[64 * 64], stack[128 * 64], sv[64],
      Vr[64 * 64], uh[64 * 64], dh[64], gW[64 * 64], du[64 * 64], dd[64 * 64],
      big[64 * 64], d[64];
  {
    int r = rows - 1; /* theta_T = m + (uc diag(dc)) z */
    const double *ucr = uc + (size_t)r * nn, *dcr = dc + (size_t)r * n;
    for (int j = 0; j < n; ++j)
      for (int i = 0; i < n; ++i) M[i + j * n] = ucr[i + j * n] * dcr[j];
    mv(n, n, M, n, 0, z + (size_t)r * n, x);
    for (int i = 0; i < n; ++i) theta[(size_t)r * n + i] = m[(size_t)r * n + i] + x[i];
  }
  for (int r = rows - 2; r >= 0; --r) {
    int tobs = r + 1 - keep_init;
    const double *Gt = G + (g_tv ? (size_t)tobs * nn : 0);
    /* w_tv: (ps, filtered.init).zipped.scanRight -- the state of row r is stepped with the
     * parameters of observation r, i.e. of the transition r -> r + 1 (DlmFsvSystem.scala:193-201) */
    const double *sqrtW = sqrtW_all + (w_tv ? (size_t)tobs * nn : 0);
    const double *ucr = uc + (size_t)r * nn, *dcr = dc + (size_t)r * n;
    mm(n, n, n, sqrtW, n, 0, Gt, n, 0, t1, n); /* sqrtW * G        */
    mm(n, n, n, t1, n, 0, ucr, n, 0, t2, n);   /* (sqrtW G) * uc   */
    for (int j = 0; j < n; ++j)
      for (int i = 0; i < n; ++i) {
        stack[i + j * 2 * n] = t2[i + j * n];
        stack[n + i + j * 2 * n] = (i == j) ? 1.0 / dcr[i] : 0.0;
      }
    st |= jacobi_svd(2 * n, n, stack, sv, Vr);
    mm(n, n, n, ucr, n, 0, Vr, n, 0, uh, n);
    for (int i = 0; i < n; ++i) dh[i] = 1.0 / sv[i];
    mm(n, n, n, Gt, n, 1, sqrtW, n, 1, t1, n); /* G^T * sqrtW^T         */
    mm(n, n, n, t1, n, 0, sqrtW, n, 0, gW, n); /* (G^T sqrtW^T) * sqrtW */
    for (int j = 0; j < n; ++j)
      for (int i = 0; i < n; ++i) du[i + j * n] = dh[i] * uh[j + i * n];
    mm(n, n, n, du, n, 1, du, n, 0, dd, n);    /* du^T * du             */
    mm(n, n, n, dd, n, 0, gW, n, 0, big, n);   /* (du^T du) * gWinv     */
    for (int i = 0; i < n; ++i)
      d[i] = theta[(size_t)(r + 1) * n + i] - a[(size_t)(r + 1) * n + i];
    mv(n, n, big, n, 0, d, x);
    double h[64];
    for (int i = 0; i < n; ++i) h[i] = m[(size_t)r * n + i] + x[i];
    for (int j = 0; j < n; ++j)
      for (int i = 0; i < n; ++i) M[i + j * n] = uh[i + j * n] * dh[j];
    mv(n, n, M, n, 0, z + (size_t)r * n, x);
    for (int i = 0; i < n; ++i) theta[(size_t)r * n + i] = h[i] + x[i];
  }
  return st;
}

ORACLE_API int oracle_svd_backward_sample(int n, int T, int keep_init,
                                          const double *G, int g_tv,
                                          const double *sqrtW, const double *m,
                                          const double *dc, const double *uc,
                                          const double *a, const double *z,
                                          double *theta) {
  return svd_backward_sample_impl(n, T, keep_init, G, g_tv, sqrtW, 0, m, dc, uc, a, z, theta);
}

/* SvdSampler.ffbs :66-74 as called by ffbsDlm :79-82 / Gibbs.stepSvd
 * (Gibbs.scala:187-190): params transformed, advance closure holds RAW W (Q2)
 * unless consistent != 0, sampler gets W^{1/2}. */
ORACLE_API int oracle_svd_ffbs(int n, int p, int T, const double *F, int f_tv,
                               const double *G, int g_tv, const double *V,
                               const double *W, const double *m0, const double *C0,
                               const double *times, const double *y,
                               const double *z, int consistent, double *times_out,
                               double *theta, double *m, double *dc, double *uc,
                               double *a, double *dr, double *ur) {
  int rows = T + 1, st = ST_OK;
  double Vfac[64 * 64], sqrtW[64 * 64];
  st |= oracle_sqrt_svd(p, V, 1, Vfac);
  st |= oracle_sqrt_svd(n, W, 0, sqrtW);
  double *f = (double *)malloc(sizeof(double) * rows * p);
  int s2 = oracle_svd_filter(n, p, T, F, f_tv, G, g_tv, Vfac, 0,
                             consistent ? sqrtW : W, m0, C0, times, y, 1, times_out,
                             m, dc, uc, a, dr, ur, f);
  free(f);
  if (s2 < 0) return s2;
  st |= s2;
  st |= oracle_svd_backward_sample(n, T, 1, G, g_tv, sqrtW, m, dc, uc, a, z, theta);
  return st;
}

/* Next row f2 on the SVD path.  V: [T][p*p] when v_tv, W: [T][n*n] when w_tv (raw matrices).
 *  - DlmFsv.ffbsSvd (DlmFsv.scala:208-229): v_tv = 1, w_tv = 0, consistent = 1 -- every step runs
 *    with transformParams(p.copy(v = V_t)), the advance closure is built from the TRANSFORMED
 *    parameters (so it stacks sqrtSvd(W), not the raw W of quirk Q2), the sampler gets ps.head.w.
 *  - DlmFsvSystem.ffbsSvd (DlmFsvSystem.scala:177-207): w_tv = 1, consistent = 1; the backward
 *    step of row r uses sqrtSvd(W) of observation r.
 * consistent = 0 keeps the raw-W advance of filterDlm / ffbsDlm for callers that mix the two. */
ORACLE_API int oracle_svd_filter_tv(int n, int p, int T, const double *F, int f_tv,
                                    const double *G, int g_tv, const double *V, int v_tv,
                                    const double *W, int w_tv, int consistent,
                                    const double *m0, const double *C0,
                                    const double *times, const double *y, int keep_init,
                                    double *times_out, double *m, double *dc, double *uc,
                                    double *a, double *dr, double *ur, double *f) {
  int nn = n * n, st = ST_OK, nW = w_tv ? T : 1;
  double *Wadv = (double *)malloc(sizeof(double) * nn * nW);
  for (int t = 0; t < nW; ++t) {
    if (consistent) st |= oracle_sqrt_svd(n, W + (size_t)t * nn, 0, Wadv + (size_t)t * nn);
    else memcpy(Wadv + (size_t)t * nn, W + (size_t)t * nn, sizeof(double) * nn);
  }
  int s2 = svd_filter_impl(n, p, T, F, f_tv, G, g_tv, V, v_tv, 1, Wadv, w_tv, m0, C0, times, y,
                           keep_init, times_out, m, dc, uc, a, dr, ur, f);
  free(Wadv);
  return s2 < 0 ? s2 : (st | s2);
}

ORACLE_API int oracle_svd_ffbs_tv(int n, int p, int T, const double *F, int f_tv,
                                  const double *G, int g_tv, const double *V, int v_tv,
                                  const double *W, int w_tv, int consistent,
                                  const double *m0, const double *C0, const double *times,
                                  const double *y, const double *z, double *times_out,
                                  double *theta, double *m, double *dc, double *uc,
                                  double *a, double *dr, double *ur) {
  int rows = T + 1, nn = n * n, st = ST_OK, nW = w_tv ? T : 1;
  double *sqrtW = (double *)malloc(sizeof(double) * nn * nW);
  for (int t = 0; t < nW; ++t) st |= oracle_sqrt_svd(n, W + (size_t)t * nn, 0, sqrtW + (size_t)t * nn);
  double *f = (double *)malloc(sizeof(double) * rows * p);
  int s2 = svd_filter_impl(n, p, T, F, f_tv, G, g_tv, V, v_tv, 1, consistent ? sqrtW : W, w_tv,
                           m0, C0, times, y, 1, times_out, m, dc, uc, a, dr, ur, f);
  free(f);
  if (s2 < 0) { free(sqrtW); return s2; }
  st |= s2;
  st |= svd_backward_sample_impl(n, T, 1, G, g_tv, sqrtW, w_tv, m, dc, uc, a, z, theta);
  free(sqrtW);
  return st;
}

/* --------------------------------------------------- Gibbs sufficient stats */

/* GibbsSampling.sampleObservationMatrix (Gibbs.scala:29-43), sampleSystemMatrix
 * (:63-73) and GibbsWishart.sampleSystemMatrix (GibbsWishart.scala:22-29).
 * theta[T+1][n] (row 0 = theta_0), times_rows[T+1].
 * ssy[p], ny[p] (counts), ssw[n], scatter[n*n]. */
ORACLE_API void oracle_gibbs_stats(int n, int p, int T, const double *F, int f_tv,
                                   const double *G, int g_tv,
                                   const double *times_rows, const double *y,
                                   const double *theta, double *ssy, double *ny,
                                   double *ssw, double *scatter) {
  int nn = n * n;
  double ft[64], gx[64], diff[64];
  for (int i = 0; i < p; ++i) { ssy[i] = 0.0; ny[i] = 0.0; }
  for (int t = 0; t < T; ++t) {
    const double *Ft = F + (f_tv ? (size_t)t * n * p : 0);
    const double *th = theta + (size_t)(t + 1) * n;
    mv(p, n, Ft, n, 1, th, ft);
    for (int i = 0; i < p; ++i) {
      double yi = y[(size_t)t * p + i];
      double res = 0.0;
      if (!isnan(yi)) { res = (yi - ft[i]) * (yi - ft[i]); ny[i] += 1.0; }
      ssy[i] = t == 0 ? res : ssy[i] + res;
    }
  }
  for (int t = 0; t < T; ++t) {
    const double *Gt = G + (g_tv ? (size_t)t * nn : 0);
    double dt = times_rows[t + 1] - times_rows[t];
    mv(n, n, Gt, n, 0, theta + (size_t)t * n, gx);
    for (int i = 0; i < n; ++i) diff[i] = theta[(size_t)(t + 1) * n + i] - gx[i];
    for (int i = 0; i < n; ++i) {
      double v = (diff[i] * diff[i]) / dt;
      ssw[i] = t == 0 ? v : ssw[i] + v;
    }
    for (int j = 0; j < n; ++j)
      for (int i = 0; i < n; ++i) {
        double v = (diff[i] * diff[j]) / dt;
        scatter[i + j * n] = t == 0 ? v : scatter[i + j * n] + v;
      }
  }
}

/* ------------------------------------------- conjugate parameter draws (f1)
 * GibbsSampling.sampleObservationMatrix (Gibbs.scala:41-49) and sampleSystemMatrix
 * (:72-77): posterior InverseGamma(shape, rate) per diagonal element and
 * InverseGamma.draw (InverseGamma.scala:14) = 1.0 / Gamma(shape, 1.0 / rate).draw.
 * Breeze's Gamma(shape, scale).draw is scale * g with g a standard Gamma(shape, 1)
 * variate (Marsaglia-Tsang: g = d * v); g is INJECTED here, so
 *     draw = 1.0 / ((1.0 / rate) * g).
 * count[k] = number of observed y_i (sampleObservationMatrix) or NULL with
 * count_all = theta.size - 1 = T (sampleSystemMatrix). */
ORACLE_API void oracle_gibbs_invgamma(int k, double prior_shape, double prior_scale,
                                      const double *count, double count_all,
                                      const double *ss, const double *g,
                                      double *shape, double *rate, double *draw) {
  for (int i = 0; i < k; ++i) {
    double c = count ? count[i] : count_all;
    shape[i] = prior_shape + c * 0.5;
    rate[i] = prior_scale + ss[i] * 0.5;
    if (draw) draw[i] = 1.0 / ((1.0 / rate[i]) * g[i]);
  }
}

/* dpotrf 'L' with the strict upper triangle zeroed (Breeze cholesky). */
static int chol_lower(int n, const double *S, double *L) {
  int st = ST_OK;
  memcpy(L, S, sizeof(double) * n * n);
  for (int j = 0; j < n; ++j) {
    double d = L[j + j * n];
    for (int k = 0; k < j; ++k) d = d - L[j + k * n] * L[j + k * n];
    if (!(d > 0.0)) st = ST_NOTPD;
    d = sqrt(d);
    L[j + j * n] = d;
    for (int i = j + 1; i < n; ++i) {
      double v = L[i + j * n];
      for (int k = 0; k < j; ++k) v = v - L[i + k * n] * L[j + k * n];
      L[i + j * n] = v / d;
    }
    for (int i = 0; i < j; ++i) L[i + j * n] = 0.0;
  }
  return st;
}

/* Breeze inv(M) (dgetrf + dgetri) restated as dgesv against the identity. */
static int inv_lu(int n, const double *M, double *out) {
  double A[64 * 64];
  memcpy(A, M, sizeof(double) * n * n);
  for (int k = 0; k < n * n; ++k) out[k] = (k % (n + 1) == 0) ? 1.0 : 0.0;
  return lu_solve(n, A, n, out);
}

/* GibbsWishart.sampleSystemMatrix (GibbsWishart.scala:16-35) followed by
 * InverseWishart.draw (InverseWishart.scala:17-25): dof = nu + T,
 * scale = psi + scatter; l = cholesky(inv(scale)); a = Bartlett factor of
 * Wishart(dof, scale) (Wishart.scala:34-43: lower triangular, a_ii = sqrt(chi2(dof - i)),
 * a_ij = N(0,1) for i > j) -- INJECTED as A; draw = inv(l)^T inv(a)^T inv(a) inv(l). */
ORACLE_API int oracle_inverse_wishart(int n, const double *psi, const double *scatter,
                                      const double *A, double *scale_out, double *W) {
  double sc[64 * 64], isc[64 * 64], l[64 * 64], il[64 * 64], ia[64 * 64],
      t1[64 * 64], t2[64 * 64];
  int st = ST_OK;
  for (int k = 0; k < n * n; ++k) sc[k] = psi[k] + (scatter ? scatter[k] : 0.0);
  if (scale_out) memcpy(scale_out, sc, sizeof(double) * n * n);
  st |= inv_lu(n, sc, isc);
  st |= chol_lower(n, isc, l);
  st |= inv_lu(n, l, il);
  st |= inv_lu(n, A, ia);
  mm(n, n, n, il, n, 1, ia, n, 1, t1, n);   /* invl.t * inva.t */
  mm(n, n, n, t1, n, 0, ia, n, 0, t2, n);   /* ... * inva      */
  mm(n, n, n, t2, n, 0, il, n, 0, W, n);    /* ... * invl      */
  return st;
}

/* ---------------------------------- scalar AR(1) / OU filters and samplers (f3)
 * FilterAr.stepUni / filterUnivariate (FilterAr.scala:15-47), FilterOu.stepUni /
 * filterUnivariate (FilterOu.scala:7-45), backStepUni / univariateSample
 * (FilterAr.scala:56-75, FilterOu.scala:47-71).  SvParameters(phi, mu, sigmaEta).
 * ou = 0: AR(1) on a unit grid; ou = 1: Ornstein-Uhlenbeck on times[].
 * v[T]: per-step observation variances.  Outputs have T + 1 rows (row 0 = prior).
 * Math.pow(x, 2) is evaluated as x * x (HotSpot's pow intrinsic special-cases y == 2). */
ORACLE_API void oracle_ar_filter(int ou, int T, double phi, double mu, double sigma,
                                 const double *times, const double *v, const double *y,
                                 double *tm, double *m, double *C, double *a, double *R) {
  double m0 = mu;
  double c0 = ou ? sigma * sigma / phi * phi : sigma * sigma / (1 - phi * phi);
  tm[0] = ou ? times[0] : times[0] - 1.0;
  m[0] = m0; C[0] = c0; a[0] = m0; R[0] = c0;
  for (int t = 0; t < T; ++t) {
    double at, rt;
    if (ou) {
      double dt = times[t] - tm[t];
      double variance = ((sigma * sigma) * (1 - exp(-2 * phi * dt))) / (2 * phi);
      at = mu + exp(-phi * dt) * (m[t] - mu);
      rt = exp(-2 * phi * dt) * C[t] + variance;
    } else {
      at = mu + phi * (m[t] - mu);
      rt = phi * phi * C[t] + sigma * sigma;
    }
    tm[t + 1] = times[t];
    a[t + 1] = at; R[t + 1] = rt;
    if (isnan(y[t])) {
      m[t + 1] = at; C[t + 1] = rt;
    } else {
      double kt = rt / (rt + v[t]);
      double et = y[t] - at;
      m[t + 1] = at + kt * et;
      C[t + 1] = kt * v[t];
    }
  }
}

/* z[T + 1]: injected N(0,1); Gaussian(mean, sd).draw = mean + sd * z (Breeze). */
ORACLE_API void oracle_ar_backward_sample(int ou, int T, double phi, const double *tm,
                                          const double *m, const double *C,
                                          const double *a, const double *R,
                                          const double *z, double *theta) {
  theta[T] = m[T] + sqrt(C[T]) * z[T];
  for (int t = T - 1; t >= 0; --t) {
    double ph = ou ? exp(-phi * (tm[t + 1] - tm[t])) : phi;
    double mean = m[t] + (C[t] * ph / R[t + 1]) * (theta[t + 1] - a[t + 1]);
    double cov = C[t] - ((C[t] * C[t]) * (ph * ph)) / R[t + 1];
    theta[t] = mean + sqrt(cov) * z[t];
  }
}

/* ------------------------------------------------- conjugate filter (f4)
 * ConjugateFilter.step / updateStats / initialiseState (ConjugateFilter.scala:23-94)
 * for p = 1 (the reference's own use: FirstOrderDlm.scala:144-172): unknown observation
 * variance with an InverseGamma(shape, scale) prior updated alongside the Kalman step.
 * Note the reference's m = mt + k e (it adds the gain term to the PREVIOUS mean, :83).
 * Outputs T + 1 rows: m[n], C[n*n], shape, scale of the variance posterior. */
ORACLE_API int oracle_conjugate_filter(int n, int T, const double *F, int f_tv,
                                       const double *G, int g_tv, const double *W,
                                       const double *m0, const double *C0,
                                       double prior_shape, double prior_scale,
                                       const double *times, const double *y,
                                       double *m, double *C, double *shape, double *scale) {
  int nn = n * n, st = ST_OK;
  double a[64], R[64 * 64], fr[64], rhs[64], K[64], D[64 * 64], t1[64 * 64],
      C1[64 * 64], kv[64], C2[64 * 64];
  double tprev = min_time(T, times) - 1.0;
  memcpy(m, m0, sizeof(double) * n);
  memcpy(C, C0, sizeof(double) * nn);
  shape[0] = prior_shape; scale[0] = prior_scale;
  for (int t = 0; t < T; ++t) {
    const double *Ft = F + (f_tv ? (size_t)t * n : 0);
    const double *Gt = G + (g_tv ? (size_t)t * nn : 0);
    const double *mp = m + (size_t)t * n, *Cp = C + (size_t)t * nn;
    double *mn = m + (size_t)(t + 1) * n, *Cn = C + (size_t)(t + 1) * nn;
    double dt = times[t] - tprev;
    kf_advance(n, Gt, W, dt, mp, Cp, a, R, t1);
    double v = scale[t] / (shape[t] - 1);          /* meanVariance: InverseGamma.mean */
    double ft, qt;
    mv(1, n, Ft, n, 1, a, &ft);
    mm(1, n, n, Ft, n, 1, R, n, 0, fr, 1);
    mm(1, n, 1, fr, 1, 0, Ft, n, 0, &qt, 1);
    qt = qt + v;
    double e = y[t] - ft;
    mm(1, n, n, Ft, n, 1, R, n, 1, rhs, 1);          /* f.t * rt.t */
    if (qt == 0.0) st |= ST_SINGULAR;
    for (int i = 0; i < n; ++i) K[i] = rhs[i] / qt;  /* (qt.t \ ...).t */
    /* updateStats: shape + 1, scale + qt.t \ (v * (e * e.t)) */
    shape[t + 1] = shape[t] + 1;
    scale[t + 1] = scale[t] + (v * (e * e)) / qt;
    for (int j = 0; j < n; ++j)
      for (int i = 0; i < n; ++i)
        D[i + j * n] = (i == j ? 1.0 : 0.0) - K[i] * Ft[j];
    mm(n, n, n, D, n, 0, R, n, 0, t1, n);
    mm(n, n, n, t1, n, 0, D, n, 1, C1, n);
    for (int i = 0; i < n; ++i) kv[i] = K[i] * v;
    for (int j = 0; j < n; ++j)
      for (int i = 0; i < n; ++i) C2[i + j * n] = kv[i] * K[j];
    for (int k = 0; k < nn; ++k) Cn[k] = C1[k] + C2[k];
    for (int i = 0; i < n; ++i) mn[i] = mp[i] + K[i] * e;   /* sic: mt, not at (:83) */
    tprev = times[t];
  }
  return st;
}

/* -------------------------------------------------- exposed small helpers */
ORACLE_API int oracle_eigsym(int n, const double *A, double *lam, double *V) {
  return jacobi_eigsym(n, A, lam, V);
}
ORACLE_API int oracle_svd(int r, int n, const double *M, double *sv, double *V) {
  return jacobi_svd(r, n, M, sv, V);
}
ORACLE_API int oracle_solve(int n, const double *A, int nrhs, const double *B,
                            double *X) {
  double Ac[64 * 64];
  memcpy(Ac, A, sizeof(double) * n * n);
  memcpy(X, B, sizeof(double) * n * nrhs);
  return lu_solve(n, Ac, nrhs, X);
}

/* ------------------------------------------------------------ batch drivers
 * The reference's only concurrency is "N independent copies on N threads"
 * (Streaming.scala:162-173); the batch drivers below do exactly that with
 * OpenMP and exist for the cpu_baseline / --impl reference timing legs.
 * Per-series parameters (stride 0 = shared).  Arrays are [B][rows][k]. */
ORACLE_API int oracle_batch_filter_smooth(
    int64_t B, int nthreads, int n, int p, int T, const double *F, const double *G,
    const double *V, int64_t v_stride, const double *W, int64_t w_stride,
    const double *m0, int64_t m0_stride, const double *C0, int64_t c0_stride,
    const double *times, const double *y, int keep_init, double *m, double *C,
    double *a, double *R, double *f, double *Q, double *s, double *S) {
  int rows = T + keep_init, nn = n * n, pp = p * p;
  int bad = 0;
#pragma omp parallel for num_threads(nthreads) schedule(static) reduction(+ : bad)
  for (int64_t b = 0; b < B; ++b) {
    double *tm = (double *)malloc(sizeof(double) * rows);
    size_t o = (size_t)b * rows;
    int st = oracle_kf_filter(n, p, T, F, 0, G, 0, V + b * v_stride, W + b * w_stride,
                              m0 + b * m0_stride, C0 + b * c0_stride, times,
                              y + (size_t)b * T * p, keep_init, tm, m + o * n,
                              C + o * nn, a + o * n, R + o * nn, f + o * p, Q + o * pp);
    st |= oracle_rts_smooth(n, T, keep_init, G, 0, m + o * n, C + o * nn, a + o * n,
                            R + o * nn, 0, s + o * n, S + o * nn);
    free(tm);
    bad += st != 0;
  }
  return bad;
}

ORACLE_API int oracle_batch_ffbs(int64_t B, int nthreads, int n, int p, int T,
                                 const double *F, const double *G, const double *V,
                                 int64_t v_stride, const double *W, int64_t w_stride,
                                 const double *m0, int64_t m0_stride,
                                 const double *C0, int64_t c0_stride,
                                 const double *times, const double *y,
                                 const double *z, int svd, double *theta) {
  int rows = T + 1, nn = n * n;
  int bad = 0;
#pragma omp parallel for num_threads(nthreads) schedule(static) reduction(+ : bad)
  for (int64_t b = 0; b < B; ++b) {
    double *buf = (double *)malloc(sizeof(double) * rows * (1 + 4 * n + 2 * nn));
    double *tm = buf, *m = tm + rows, *a = m + rows * n, *d1 = a + rows * n,
           *d2 = d1 + rows * n, *M1 = d2 + rows * n, *M2 = M1 + rows * nn;
    int st;
    if (svd)
      st = oracle_svd_ffbs(n, p, T, F, 0, G, 0, V + b * v_stride, W + b * w_stride,
                           m0 + b * m0_stride, C0 + b * c0_stride, times,
                           y + (size_t)b * T * p, z + (size_t)b * rows * n, 0, tm,
                           theta + (size_t)b * rows * n, m, d1, M1, a, d2, M2);
    else
      st = oracle_ffbs(n, p, T, F, 0, G, 0, V + b * v_stride, W + b * w_stride,
                       m0 + b * m0_stride, C0 + b * c0_stride, times,
                       y + (size_t)b * T * p, z + (size_t)b * rows * n, tm,
                       theta + (size_t)b * rows * n, m, M1, a, M2);
    free(buf);
    bad += st != 0;
  }
  return bad;
}
