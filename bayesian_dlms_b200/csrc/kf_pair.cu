// kf_pair.cu -- fused Kalman filter + RTS smoother for n = 4, p = 1 with TWO LANES PER SERIES.
//
// The thread-per-series kernel (kf_small.cu) holds m, C, a, R, W, s, S and the temporaries of a
// 4 x 4 step in 255 registers: 8 warps per SM, and measured at 54 % of HBM with the FP64 pipe
// and the issue slots both under half busy -- a latency-bound kernel.  Here every series is
// walked by a PAIR of adjacent lanes; lane h owns columns 2h, 2h+1 of each matrix (pair_steps.cuh),
// which roughly halves the registers and the dependent instruction chain per lane and doubles
// the warps that can be resident.  The arithmetic per output element is unchanged, so the kernel
// is bit-identical to kf_small.cu and to the oracle.
//
// Memory side, same design as kf_small.cu: device-native time-major SoA [rows][k][B]; the 16
// pairs of a warp read / write 16 consecutive series = one full 128-byte line per field element
// and half-warp; y prefetched one 4-step chunk ahead; the backward pass's (m_t, C_t) rows staged
// through a cp.async ring in shared memory (each lane fetches its own two columns of C and two
// elements of m; the partner's part is read from the partner's slot after a pair barrier).
//
// Two transports for the gathers between the lanes of a pair (template parameter):
//   kShfl  __shfl_xor_sync + selects: no shared memory, no barrier
//   kSmem  both lanes store their half into a per-pair shared-memory buffer, one pair barrier,
//          both read the whole: a third of the instructions; buffers rotate by gather site so
//          that no buffer is rewritten before the barrier of the following gather
//
// Selection: BDLM_KF_PAIR = 0 (off: kf_small.cu) | 1 (kShfl) | 2 (kSmem) at 3 resident blocks per SM,
// 3 | 4 the same at 4 blocks (128 registers); bdlm_debug_set_pair_mode overrides it at run time.
#include <atomic>
#include <cmath>
#include <cstdlib>
#include <limits>
#include <thread>

#include "common.cuh"
#include "launch.h"
#include "pair_steps.cuh"

namespace bdlm {

namespace {

using namespace pairk;

enum { kShfl = 1, kSmem = 2 };

// ---------------------------------------------------------------- pair contexts (device)
struct DevPairShfl {
  int h;
  unsigned mask;  // the two lanes of the pair: every exchange involves these two lanes only, so
                  // pairs of one warp may take different paths (missing y, dt == 0)
  template <int SITE>
  __device__ __forceinline__ void gather_mat(const double (&loc)[NL], double (&full)[N * N]) const {
#pragma unroll
    for (int k = 0; k < NL; ++k) {
      const double o = __shfl_xor_sync(mask, loc[k], 1);
      full[k] = h ? o : loc[k];
      full[NL + k] = h ? loc[k] : o;
    }
  }
  template <int SITE>
  __device__ __forceinline__ void gather_vec(double x0, double x1, double (&full)[N]) const {
    const double o0 = __shfl_xor_sync(mask, x0, 1), o1 = __shfl_xor_sync(mask, x1, 1);
    full[0] = h ? o0 : x0; full[1] = h ? o1 : x1;
    full[2] = h ? x0 : o0; full[3] = h ? x1 : o1;
  }
  __device__ __forceinline__ int or_int(int v) const { return v | __shfl_xor_sync(mask, v, 1); }
  __device__ __forceinline__ void barrier() const { __syncwarp(mask); }
};

// Per pair: three matrix buffers of 18 doubles (144-byte pitch: the 4 pairs of a quarter-warp hit
// distinct banks with 16-byte accesses) and two vector buffers of 4 doubles.
constexpr int kMatPitch = 18, kPairDoubles = 3 * kMatPitch + 2 * 4;  // 62 doubles = 496 bytes
template <int SITE> struct SiteBuf;
// consecutive gathers on every path use different buffers:
//   forward  AdvT1(M0) Fr(V0) R(M1) [K redundant] UpdT1(M2) | AdvT1(M0) ...;  missing y: AdvT1(M0) Fr(V0) | AdvT1(M0)
//   backward AdvT1(M0) R1(M1) T(V1) X(M2) RtsT1(M1) | AdvT1(M0) ...
template <> struct SiteBuf<kSiteAdvT1> { static constexpr int off = 0; };
template <> struct SiteBuf<kSiteR> { static constexpr int off = kMatPitch; };
template <> struct SiteBuf<kSiteUpdT1> { static constexpr int off = 2 * kMatPitch; };
template <> struct SiteBuf<kSiteR1> { static constexpr int off = kMatPitch; };
template <> struct SiteBuf<kSiteX> { static constexpr int off = 2 * kMatPitch; };
template <> struct SiteBuf<kSiteRtsT1> { static constexpr int off = kMatPitch; };
template <> struct SiteBuf<kSiteFr> { static constexpr int off = 3 * kMatPitch; };
template <> struct SiteBuf<kSiteT> { static constexpr int off = 3 * kMatPitch + 4; };

struct DevPairSmem {
  int h;
  unsigned mask;
  double *buf;  // this pair's kPairDoubles doubles, 16-byte aligned
  template <int SITE>
  __device__ __forceinline__ void gather_mat(const double (&loc)[NL], double (&full)[N * N]) const {
    double2 *w = reinterpret_cast<double2 *>(buf + SiteBuf<SITE>::off + NL * h);
#pragma unroll
    for (int k = 0; k < NL / 2; ++k) w[k] = make_double2(loc[2 * k], loc[2 * k + 1]);
    __syncwarp(mask);
    const double2 *r = reinterpret_cast<const double2 *>(buf + SiteBuf<SITE>::off);
#pragma unroll
    for (int k = 0; k < N * N / 2; ++k) {
      const double2 v = r[k];
      full[2 * k] = v.x; full[2 * k + 1] = v.y;
    }
  }
  template <int SITE>
  __device__ __forceinline__ void gather_vec(double x0, double x1, double (&full)[N]) const {
    *reinterpret_cast<double2 *>(buf + SiteBuf<SITE>::off + 2 * h) = make_double2(x0, x1);
    __syncwarp(mask);
    const double2 *r = reinterpret_cast<const double2 *>(buf + SiteBuf<SITE>::off);
    const double2 u = r[0], v = r[1];
    full[0] = u.x; full[1] = u.y; full[2] = v.x; full[3] = v.y;
    // two steps with dt == 0 and a missing observation gather through this buffer back to back
    // (no other exchange in between): the trailing barrier keeps the second write behind this read
    __syncwarp(mask);
  }
  __device__ __forceinline__ int or_int(int v) const { return v | __shfl_xor_sync(mask, v, 1); }
  __device__ __forceinline__ void barrier() const { __syncwarp(mask); }
};

template <int PXK> struct PxOf;
template <> struct PxOf<kShfl> { using type = DevPairShfl; };
template <> struct PxOf<kSmem> { using type = DevPairSmem; };

// ---------------------------------------------------------------- loads / stores
// own columns of a matrix field: elements 8h .. 8h + 7
__device__ __forceinline__ void store_cols(const View &v, int64_t b, int64_t row, int h,
                                           const double (&x)[NL]) {
  if (v.ptr == nullptr) return;
  double *p = v.ptr + b * v.sb + row * v.sr + (int64_t)(NL * h) * v.sk;
#pragma unroll
  for (int k = 0; k < NL; ++k) st_stream(p + k * v.sk, x[k]);
}
// a vector field is held by both lanes: lane h stores elements 2h, 2h + 1
__device__ __forceinline__ void store_vec_half(const View &v, int64_t b, int64_t row, int h,
                                               const double (&x)[N]) {
  if (v.ptr == nullptr) return;
  double *p = v.ptr + b * v.sb + row * v.sr + (int64_t)(2 * h) * v.sk;
  st_stream(p, h ? x[2] : x[0]);
  st_stream(p + v.sk, h ? x[3] : x[1]);
}
__device__ __forceinline__ void load_cols(const View &v, int64_t b, int64_t row, int h, double (&x)[NL]) {
  const double *p = v.ptr + b * v.sb + row * v.sr + (int64_t)(NL * h) * v.sk;
#pragma unroll
  for (int k = 0; k < NL; ++k) x[k] = ld_stream(p + k * v.sk);
}
__device__ __forceinline__ void load_pcols(const PView &v, int64_t b, int h, double (&x)[NL]) {
  const double *p = v.ptr + b * v.sb + (int64_t)(NL * h) * v.sk;
#pragma unroll
  for (int k = 0; k < NL; ++k) x[k] = p[k * v.sk];
}

struct PairModel {
  double G[N * N];
  double F[N];
};

template <bool REG>
__device__ __forceinline__ void load_model(const Batch &bt, const PairModel &mdl, int64_t b, int t,
                                           double (&G)[N * N], double (&F)[N]) {
  if (!REG && bt.g_tv) {
    const double *g = bt.G + b * bt.G_sb + (int64_t)t * bt.G_sr;
#pragma unroll
    for (int k = 0; k < N * N; ++k) G[k] = __ldg(g + k * bt.G_sk);
  } else {
#pragma unroll
    for (int k = 0; k < N * N; ++k) G[k] = mdl.G[k];
  }
  if (!REG && bt.f_tv) {
    const double *f = bt.F + b * bt.F_sb + (int64_t)t * bt.F_sr;
#pragma unroll
    for (int k = 0; k < N; ++k) F[k] = __ldg(f + k * bt.F_sk);
  } else {
#pragma unroll
    for (int k = 0; k < N; ++k) F[k] = mdl.F[k];
  }
}

constexpr int kChunk = 4;   // y prefetch distance (steps)
constexpr int kRing = 4;    // backward-pass spill rows in flight per lane (power of two)
constexpr int kRow = 2 + NL;  // doubles per lane and spill row: two elements of m, two columns of C
constexpr int kThreads = 128, kSeriesPerBlock = kThreads / 2;

__device__ __forceinline__ void cp_async8(double *smem_dst, const double *gmem_src) {
  const unsigned sa = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(sa), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int K>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(K) : "memory"); }

constexpr int kDoFilter = 1, kDoSmooth = 2;

// OCC = resident blocks asked of ptxas: 3 -> 168 registers (12 warps per SM, a few dozen bytes of
// spill), 4 -> 128 registers (16 warps per SM, ~0.5 KB of spill per lane).
template <bool REG, int MODE, int PXK, int OCC>
__global__ void __launch_bounds__(kThreads, OCC)
kf_pair_kernel(const Batch bt, const PairModel mdl, const KfViews kf, const View sv, const View Sv) {
  extern __shared__ __align__(16) double dsm[];
  const int h = threadIdx.x & 1;
  const int lane = threadIdx.x & 31;
  const int64_t b = blockIdx.x * (int64_t)kSeriesPerBlock + (threadIdx.x >> 1);
  if (b >= bt.B) return;  // both lanes of a pair leave together; nothing below spans pairs
  typename PxOf<PXK>::type px;
  px.h = h;
  px.mask = 3u << (lane & ~1);
  double *ring = dsm;
  if constexpr (PXK == kSmem) {
    px.buf = dsm + (size_t)(threadIdx.x >> 1) * kPairDoubles;
    ring = dsm + ((size_t)kSeriesPerBlock * kPairDoubles + 1) / 2 * 2;
  }
  const int T = bt.T, ki = bt.keep_init, rows = T + ki;
  int st = 0;

  double Wl[NL], m[N], Cl[NL];
  double V = bt.V.ptr[b * bt.V.sb];
  load_pcols(bt.W, b, h, Wl);
  // REG: the model is fixed; rows 2h, 2h+1 of G are kept for the G C G^T products
  double Grc[NL];
  if (REG) own_rows(px, mdl.G, Grc);

  if (MODE & kDoFilter) {
#pragma unroll
    for (int i = 0; i < N; ++i) m[i] = bt.m0.ptr[b * bt.m0.sb + i * bt.m0.sk];
    load_pcols(bt.C0, b, h, Cl);
    if (ki) {  // initialiseState (KalmanFilter.scala:112-118): f, Q = None -> NaN
      const double nanv = __longlong_as_double(0x7ff8000000000000LL);
      store_vec_half(kf.m, b, 0, h, m);
      store_cols(kf.C, b, 0, h, Cl);
      store_vec_half(kf.a, b, 0, h, m);
      store_cols(kf.R, b, 0, h, Cl);
      if (h == 0) { if (kf.f.ptr) st_stream(kf.f.ptr + b * kf.f.sb, nanv); }
      else { if (kf.Q.ptr) st_stream(kf.Q.ptr + b * kf.Q.sb, nanv); }
    }
    const double *yp = bt.y.ptr + b * bt.y.sb;
    double ycur[kChunk], ynxt[kChunk];
#pragma unroll
    for (int u = 0; u < kChunk; ++u) ycur[u] = (u < T) ? ld_stream(yp + u * bt.y.sr) : 0.0;
    for (int t0 = 0; t0 < T; t0 += kChunk) {
#pragma unroll
      for (int u = 0; u < kChunk; ++u) {
        const int t = t0 + kChunk + u;
        ynxt[u] = (t < T) ? ld_stream(yp + (int64_t)t * bt.y.sr) : 0.0;
      }
#pragma unroll
      for (int u = 0; u < kChunk; ++u) {
        const int t = t0 + u;
        if (t < T) {
          double a[N], Rl[NL], f, Q;
          const double dt = REG ? 1.0 : dt_at(bt, b, t);
          double G[N * N], F[N], Gr[NL];
          load_model<REG>(bt, mdl, b, t, G, F);
          if (REG) {
#pragma unroll
            for (int k = 0; k < NL; ++k) Gr[k] = Grc[k];
          } else {
            own_rows(px, G, Gr);
          }
          pairk::advance<REG>(px, G, Gr, Wl, dt, m, Cl, a, Rl);
          pairk::update(px, F, V, ycur[u], a, Rl, f, Q, m, Cl, st);
          const int64_t row = t + ki;
          store_vec_half(kf.a, b, row, h, a);
          store_cols(kf.R, b, row, h, Rl);
          if (h == 0) { if (kf.f.ptr) st_stream(kf.f.ptr + b * kf.f.sb + row * kf.f.sr, f); }
          else { if (kf.Q.ptr) st_stream(kf.Q.ptr + b * kf.Q.sb + row * kf.Q.sr, Q); }
          store_vec_half(kf.m, b, row, h, m);
          store_cols(kf.C, b, row, h, Cl);
        }
      }
#pragma unroll
      for (int u = 0; u < kChunk; ++u) ycur[u] = ynxt[u];
    }
  } else {
#pragma unroll
    for (int i = 0; i < N; ++i) m[i] = ld_stream(kf.m.ptr + b * kf.m.sb + (int64_t)(rows - 1) * kf.m.sr + i * kf.m.sk);
    load_cols(kf.C, b, rows - 1, h, Cl);
  }

  if (MODE & kDoSmooth) {
    // backwardsSmoother (Smoothing.scala:57-64): s_T = m_T, S_T = C_T
    double s[N], Sl[NL];
#pragma unroll
    for (int i = 0; i < N; ++i) s[i] = m[i];
#pragma unroll
    for (int k = 0; k < NL; ++k) Sl[k] = Cl[k];
    store_vec_half(sv, b, rows - 1, h, s);
    store_cols(Sv, b, rows - 1, h, Sl);
    const bool textbook = (bt.compat & BDLM_TEXTBOOK_SMOOTHER) != 0;
    // ring[(row & 3)][k][thread]: k = 0, 1: m[2h], m[2h + 1]; k = 2 .. 9: own columns of C
    const int t0i = threadIdx.x & ~1, t1i = t0i | 1;
    auto slot = [&](int r, int k, int thr) { return ring + ((size_t)(r & (kRing - 1)) * kRow + k) * kThreads + thr; };
    auto issue = [&](int r) {
      if (r >= 0) {
        const double *pm = kf.m.ptr + b * kf.m.sb + (int64_t)r * kf.m.sr + (int64_t)(2 * h) * kf.m.sk;
        const double *pc = kf.C.ptr + b * kf.C.sb + (int64_t)r * kf.C.sr + (int64_t)(NL * h) * kf.C.sk;
        cp_async8(slot(r, 0, threadIdx.x), pm);
        cp_async8(slot(r, 1, threadIdx.x), pm + kf.m.sk);
#pragma unroll
        for (int k = 0; k < NL; ++k) cp_async8(slot(r, 2 + k, threadIdx.x), pc + k * kf.C.sk);
      }
      cp_async_commit();
    };
#pragma unroll
    for (int d = 0; d < kRing - 1; ++d) issue(rows - 2 - d);
    for (int r = rows - 2; r >= 0; --r) {
      double a1[N], R1l[NL], Cr[NL];
      // the slot refilled here held row r + 1: the partner finished reading it before it joined
      // the last exchange of the previous iteration; the barrier makes that explicit
      px.barrier();
      issue(r - (kRing - 1));
      cp_async_wait<kRing - 1>();  // this lane's group of row r has landed ...
      px.barrier();                // ... and so has the partner's
      m[0] = *slot(r, 0, t0i); m[1] = *slot(r, 1, t0i);
      m[2] = *slot(r, 0, t1i); m[3] = *slot(r, 1, t1i);
#pragma unroll
      for (int k = 0; k < NL; ++k) Cl[k] = *slot(r, 2 + k, threadIdx.x);
      // rows 2h, 2h+1 of C: element (2h + jj, k) sits in the slot of lane k / 2 at column k % 2
#pragma unroll
      for (int k = 0; k < N; ++k)
#pragma unroll
        for (int jj = 0; jj < 2; ++jj)
          Cr[jj + 2 * k] = *slot(r, 2 + 2 * h + jj + 4 * (k & 1), (k >> 1) ? t1i : t0i);
      const int tobs = r + 1 - ki;  // observation index of row r + 1
      const double dt = REG ? 1.0 : dt_at(bt, b, tobs);
      double G[N * N], F[N], Gr[NL];
      load_model<REG>(bt, mdl, b, tobs, G, F);
      if (REG) {
#pragma unroll
        for (int k = 0; k < NL; ++k) Gr[k] = Grc[k];
      } else {
        own_rows(px, G, Gr);
      }
      pairk::advance<REG>(px, G, Gr, Wl, dt, m, Cl, a1, R1l);  // bit-identical to the forward a, R
      pairk::rts_step(px, G, m, Cl, Cr, a1, R1l, textbook, s, Sl, st);
      store_vec_half(sv, b, r, h, s);
      store_cols(Sv, b, r, h, Sl);
    }
#pragma unroll
    for (int i = 0; i < N; ++i) m[i] = s[i];
#pragma unroll
    for (int k = 0; k < NL; ++k) Cl[k] = Sl[k];
  }

  if (bt.status) {
    bool finite = true;
#pragma unroll
    for (int i = 0; i < N; ++i) finite = finite && isfinite(m[i]);
#pragma unroll
    for (int k = 0; k < NL; ++k) finite = finite && isfinite(Cl[k]);
    const int bad = px.or_int(finite ? 0 : 1);
    if (bad) st |= BDLM_ST_NONFINITE;
    if (h == 0) bt.status[b] = st;
  }
}

template <int PXK>
constexpr size_t pair_smem_bytes(bool smooth) {
  size_t d = 0;
  if (PXK == kSmem) d += ((size_t)kSeriesPerBlock * kPairDoubles + 1) / 2 * 2;
  if (smooth) d += (size_t)kRing * kRow * kThreads;
  return d * sizeof(double);
}

template <bool REG, int MODE, int PXK, int OCC>
cudaError_t launch_t(const Batch &bt, const PairModel &mdl, const KfViews &kf, const View &sv,
                     const View &Sv, cudaStream_t stream, int *wave_series) {
  const size_t smem = pair_smem_bytes<PXK>((MODE & kDoSmooth) != 0);
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(kf_pair_kernel<REG, MODE, PXK, OCC>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
  }
  if (wave_series) {  // occupancy query only
    int blocks = 0, dev = 0, sms = 0;
    cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(
        &blocks, kf_pair_kernel<REG, MODE, PXK, OCC>, kThreads, smem);
    if (e != cudaSuccess) return e;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    *wave_series = blocks * sms * kSeriesPerBlock;
    return cudaSuccess;
  }
  const int64_t blocks = (bt.B + kSeriesPerBlock - 1) / kSeriesPerBlock;
  if (blocks <= 0) return cudaSuccess;
  kf_pair_kernel<REG, MODE, PXK, OCC><<<(unsigned)blocks, kThreads, smem, stream>>>(bt, mdl, kf, sv, Sv);
  return cudaGetLastError();
}

template <int PXK, int OCC>
cudaError_t launch_px(const Batch &bt, const double *hG, const double *hF, const KfViews &kf,
                      const View &sv, const View &Sv, int mode, cudaStream_t stream, int *wave_series) {
  PairModel mdl;
  for (int k = 0; k < N * N; ++k) mdl.G[k] = hG ? hG[k] : 0.0;
  for (int k = 0; k < N; ++k) mdl.F[k] = hF ? hF[k] : 0.0;
  const bool reg = bt.dt == nullptr && !bt.g_tv && !bt.f_tv;
#define BDLM_GO(REG_, MODE_) return launch_t<REG_, MODE_, PXK, OCC>(bt, mdl, kf, sv, Sv, stream, wave_series)
  if (mode == kDoFilter) { if (reg) BDLM_GO(true, kDoFilter); else BDLM_GO(false, kDoFilter); }
  if (mode == (kDoFilter | kDoSmooth)) {
    if (reg) BDLM_GO(true, kDoFilter | kDoSmooth); else BDLM_GO(false, kDoFilter | kDoSmooth);
  }
  if (reg) BDLM_GO(true, kDoSmooth); else BDLM_GO(false, kDoSmooth);
#undef BDLM_GO
}

// ---------------------------------------------------------------- host build of the arithmetic
// Two host threads stand in for the two lanes; an exchange is a store into a shared block and a
// two-thread barrier.  Used only by bdlm_debug_pair_filter_smooth_host (CPU pin of the split
// operation order against the oracle).
struct HostShared {
  double mat[N * N];
  double vec[N];
  int flag[2];
  std::atomic<int> arrived{0};
  std::atomic<int> generation{0};
  void barrier() {
    const int gen = generation.load(std::memory_order_acquire);
    if (arrived.fetch_add(1, std::memory_order_acq_rel) == 1) {
      arrived.store(0, std::memory_order_relaxed);
      generation.store(gen + 1, std::memory_order_release);
    } else {
      while (generation.load(std::memory_order_acquire) == gen) std::this_thread::yield();
    }
  }
};

struct HostPair {
  int h;
  HostShared *sh;
  template <int SITE>
  void gather_mat(const double (&loc)[NL], double (&full)[N * N]) const {
    for (int k = 0; k < NL; ++k) sh->mat[NL * h + k] = loc[k];
    sh->barrier();
    for (int k = 0; k < N * N; ++k) full[k] = sh->mat[k];
    sh->barrier();
  }
  template <int SITE>
  void gather_vec(double x0, double x1, double (&full)[N]) const {
    sh->vec[2 * h] = x0; sh->vec[2 * h + 1] = x1;
    sh->barrier();
    for (int k = 0; k < N; ++k) full[k] = sh->vec[k];
    sh->barrier();
  }
  int or_int(int v) const {
    sh->flag[h] = v;
    sh->barrier();
    const int r = sh->flag[0] | sh->flag[1];
    sh->barrier();
    return r;
  }
};

struct HostCall {
  const double *G, *F, *W, *m0, *C0, *dt, *y;
  double V;
  int T, textbook;
  double *m, *C, *a, *R, *f, *Q, *s, *S;
  int st[2];
};

void host_lane(HostCall *c, HostShared *sh, int h) {
  HostPair px{h, sh};
  const int T = c->T, rows = T + 1;
  const bool reg = c->dt == nullptr;
  int st = 0;
  double Wl[NL], m[N], Cl[NL], Gr[NL];
  for (int k = 0; k < NL; ++k) { Wl[k] = c->W[NL * h + k]; Cl[k] = c->C0[NL * h + k]; }
  for (int i = 0; i < N; ++i) m[i] = c->m0[i];
  own_rows(px, c->G, Gr);
  auto put_cols = [&](double *dst, int row, const double (&x)[NL]) {
    for (int k = 0; k < NL; ++k) dst[(size_t)row * N * N + NL * h + k] = x[k];
  };
  auto put_half = [&](double *dst, int row, const double (&x)[N]) {
    dst[(size_t)row * N + 2 * h] = x[2 * h]; dst[(size_t)row * N + 2 * h + 1] = x[2 * h + 1];
  };
  const double nanv = std::numeric_limits<double>::quiet_NaN();
  put_half(c->m, 0, m); put_cols(c->C, 0, Cl); put_half(c->a, 0, m); put_cols(c->R, 0, Cl);
  if (h == 0) c->f[0] = nanv; else c->Q[0] = nanv;
  for (int t = 0; t < T; ++t) {
    double a[N], Rl[NL], f, Q;
    if (reg) pairk::advance<true>(px, c->G, Gr, Wl, 1.0, m, Cl, a, Rl);
    else pairk::advance<false>(px, c->G, Gr, Wl, c->dt[t], m, Cl, a, Rl);
    pairk::update(px, c->F, c->V, c->y[t], a, Rl, f, Q, m, Cl, st);
    const int row = t + 1;
    put_half(c->a, row, a); put_cols(c->R, row, Rl);
    if (h == 0) c->f[row] = f; else c->Q[row] = Q;
    put_half(c->m, row, m); put_cols(c->C, row, Cl);
  }
  double s[N], Sl[NL];
  for (int i = 0; i < N; ++i) s[i] = m[i];
  for (int k = 0; k < NL; ++k) Sl[k] = Cl[k];
  put_half(c->s, rows - 1, s); put_cols(c->S, rows - 1, Sl);
  sh->barrier();  // the partner's halves of every stored row are in place
  for (int r = rows - 2; r >= 0; --r) {
    double a1[N], R1l[NL], Cr[NL];
    const double *Cf = c->C + (size_t)r * N * N;
    for (int i = 0; i < N; ++i) m[i] = c->m[(size_t)r * N + i];
    for (int k = 0; k < NL; ++k) Cl[k] = Cf[NL * h + k];
    own_rows(px, Cf, Cr);
    if (reg) pairk::advance<true>(px, c->G, Gr, Wl, 1.0, m, Cl, a1, R1l);
    else pairk::advance<false>(px, c->G, Gr, Wl, c->dt[r], m, Cl, a1, R1l);
    pairk::rts_step(px, c->G, m, Cl, Cr, a1, R1l, c->textbook != 0, s, Sl, st);
    put_half(c->s, r, s); put_cols(c->S, r, Sl);
  }
  bool finite = true;
  for (int i = 0; i < N; ++i) finite = finite && std::isfinite(s[i]);
  for (int k = 0; k < NL; ++k) finite = finite && std::isfinite(Sl[k]);
  if (px.or_int(finite ? 0 : 1)) st |= BDLM_ST_NONFINITE;
  c->st[h] = st;
}

// 0 = off; 1 | 2 = shuffle | shared-memory exchange at 3 resident blocks; 3 | 4 = the same at 4
std::atomic<int> g_pair_mode{-1};

int pair_mode_env() {
  const char *e = std::getenv("BDLM_KF_PAIR");
  const int v = e ? std::atoi(e) : 0;
  return (v >= 1 && v <= 4) ? v : 0;
}

}  // namespace

// n = 4, p = 1 filter / smoother / fused calls served by the pair kernel instead of kf_small.cu
// (smoother-only calls that reload a, R stay on kf_small.cu).  0 = not in use.
int pair_kernel_mode() {
  int m = g_pair_mode.load(std::memory_order_relaxed);
  if (m < 0) { m = pair_mode_env(); g_pair_mode.store(m, std::memory_order_relaxed); }
  return m;
}
void set_pair_kernel_mode(int mode) { g_pair_mode.store((mode >= 0 && mode <= 4) ? mode : 0); }

cudaError_t launch_kf_pair(int pxk, const Batch &bt, const double *hG, const double *hF,
                           const KfViews &kf, const View &sv, const View &Sv, bool do_filter,
                           bool do_smooth, cudaStream_t stream, int *wave_series) {
  const int mode = (do_filter ? kDoFilter : 0) | (do_smooth ? kDoSmooth : 0);
  if (bt.n != N || bt.p != 1 || mode == 0) return cudaErrorInvalidValue;
  switch (pxk) {
    case 1: return launch_px<kShfl, 3>(bt, hG, hF, kf, sv, Sv, mode, stream, wave_series);
    case 2: return launch_px<kSmem, 3>(bt, hG, hF, kf, sv, Sv, mode, stream, wave_series);
    case 3: return launch_px<kShfl, 4>(bt, hG, hF, kf, sv, Sv, mode, stream, wave_series);
    case 4: return launch_px<kSmem, 4>(bt, hG, hF, kf, sv, Sv, mode, stream, wave_series);
    default: return cudaErrorInvalidValue;
  }
}

int pair_filter_smooth_host(const double *G, const double *F, double V, const double *W,
                            const double *m0, const double *C0, const double *dt, const double *y,
                            int T, int textbook, double *m, double *C, double *a, double *R,
                            double *f, double *Q, double *s, double *S) {
  HostCall c{G, F, W, m0, C0, dt, y, V, T, textbook, m, C, a, R, f, Q, s, S, {0, 0}};
  HostShared sh;
  std::thread partner(host_lane, &c, &sh, 1);
  host_lane(&c, &sh, 0);
  partner.join();
  return c.st[0] | c.st[1];
}

}  // namespace bdlm
