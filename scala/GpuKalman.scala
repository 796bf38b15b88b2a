// GpuKalman.scala -- reference-side facade over libbdlm.so (include/bdlm.h).
//
// SOURCE ONLY: this build image has no JVM / sbt, so this file is neither compiled nor tested
// here (see INTEGRATION.md).  It lives in the reference's package so that switching a call
// site is a one-word change (KalmanFilter.filterDlm -> GpuKalman.filterDlm).  The marshalling
// below is the same as bayesian_dlms_b200/reference_api.py, which IS tested against the oracle.
package com.github.jonnylaw.dlm

import breeze.linalg.{DenseMatrix, DenseVector}
import breeze.stats.distributions.{Rand, RandBasis}
import java.lang.foreign._
import java.lang.foreign.ValueLayout._

object GpuKalman {
  import BdlmNative._ // downcall handles + struct layout, INTEGRATION.md section 3

  private lazy val ctx: MemorySegment = {
    val out = Arena.global().allocate(ADDRESS)
    val rc = create.invoke(0, out).asInstanceOf[Int]
    if (rc != 0) throw new IllegalStateException(s"bdlm_create failed ($rc): no CUDA device, no CPU fallback")
    out.get(ADDRESS, 0)
  }

  // ---- flattening ------------------------------------------------------------------------
  private def cm(m: DenseMatrix[Double]): Array[Double] = m.toDenseMatrix.copy.data // column-major

  private case class Flat(n: Int, p: Int, times: Array[Double], y: Array[Double],
                          f: Array[Double], fTv: Boolean, g: Array[Double], gTv: Boolean)

  private def flatten(mod: Dlm, ys: Vector[Data]): Flat = {
    if (ys.isEmpty) throw new NoSuchElementException("None.get") // KalmanFilter.scala:116-117
    val times = ys.map(_.time).toArray
    val t0 = times.min - 1.0
    val dts = times.zip(t0 +: times.init).map { case (t, prev) => t - prev }
    val fs = times.map(t => cm(mod.f(t)))
    val gs = dts.map(dt => cm(mod.g(dt)))
    val fTv = fs.exists(!_.sameElements(fs.head))
    val gTv = gs.exists(!_.sameElements(gs.head))
    val f0 = mod.f(times.head)
    val y = ys.flatMap(_.observation.data.map(_.getOrElse(Double.NaN))).toArray
    Flat(f0.rows, f0.cols, times, y, if (fTv) fs.flatten else fs.head, fTv,
         if (gTv) gs.flatten else gs.head, gTv)
  }

  private def seg(a: Arena, xs: Array[Double]): MemorySegment = {
    val s = a.allocate(JAVA_DOUBLE, xs.length.toLong.max(1L))
    MemorySegment.copy(xs, 0, s, JAVA_DOUBLE, 0, xs.length)
    s
  }

  private def problem(a: Arena, fl: Flat, p: DlmParameters, keepInit: Boolean): MemorySegment = {
    val pr = a.allocate(BdlmNative.problem)
    def setL(name: String, v: Long) = pr.set(JAVA_LONG, BdlmNative.problem.byteOffset(MemoryLayout.PathElement.groupElement(name)), v)
    def setI(name: String, v: Int) = pr.set(JAVA_INT, BdlmNative.problem.byteOffset(MemoryLayout.PathElement.groupElement(name)), v)
    def setP(name: String, s: MemorySegment) = pr.set(ADDRESS, BdlmNative.problem.byteOffset(MemoryLayout.PathElement.groupElement(name)), s)
    val regular = fl.times.zipWithIndex.forall { case (t, i) => t == (i + 1).toDouble }
    setL("B", 1L); setI("T", fl.times.length); setI("n", fl.n); setI("p", fl.p)
    setI("layout", 1 /* SERIES_MAJOR */); setI("mem", 1 /* HOST */)
    setI("keep_init", if (keepInit) 1 else 0)
    setI("f_tv", if (fl.fTv) 1 else 0); setI("g_tv", if (fl.gTv) 1 else 0)
    setI("per_series", 0); setI("compat", 0)
    setP("F", seg(a, fl.f)); setP("G", seg(a, fl.g))
    setP("times", if (regular) MemorySegment.NULL else seg(a, fl.times))
    setP("V", seg(a, cm(p.v))); setP("W", seg(a, cm(p.w)))
    setP("m0", seg(a, p.m0.toArray)); setP("C0", seg(a, cm(p.c0)))
    setP("y", seg(a, fl.y))
    pr
  }

  private def check(rc: Int, status: Int): Unit = {
    if (rc == -2) throw new NoSuchElementException("None.get")
    if (rc < 0) throw new IllegalArgumentException(lastErr.invoke(ctx).asInstanceOf[MemorySegment].reinterpret(1024).getString(0))
    if ((status & 1) != 0) throw new breeze.linalg.MatrixSingularException("")
    if ((status & 2) != 0) throw new breeze.linalg.NotConvergedException(breeze.linalg.NotConvergedException.Iterations)
  }

  private def rowsOf(s: MemorySegment, rows: Int, k: Int): Array[Array[Double]] =
    Array.tabulate(rows)(r => Array.tabulate(k)(j => s.getAtIndex(JAVA_DOUBLE, r.toLong * k + j)))

  // ---- KalmanFilter ------------------------------------------------------------------------

  /** KalmanFilter(advanceState(p, mod.g)).filter (Filter.scala:41-45): T+1 states. */
  def filter(mod: Dlm, ys: Vector[Data], p: DlmParameters): Vector[KfState] = run(mod, ys, p, keepInit = true)

  /** KalmanFilter.filterDlm (KalmanFilter.scala:291-294): T states. */
  def filterDlm(mod: Dlm, ys: Vector[Data], p: DlmParameters): Vector[KfState] = run(mod, ys, p, keepInit = false)

  private def run(mod: Dlm, ys: Vector[Data], p: DlmParameters, keepInit: Boolean): Vector[KfState] = {
    val fl = flatten(mod, ys)
    val a = Arena.ofConfined()
    try {
      val rows = fl.times.length + (if (keepInit) 1 else 0)
      val (n, q) = (fl.n, fl.p)
      val out = a.allocate(ADDRESS, 6)
      val sizes = Array(n, n * n, n, n * n, q, q * q)
      val bufs = sizes.map(k => a.allocate(JAVA_DOUBLE, rows.toLong * k))
      bufs.zipWithIndex.foreach { case (b, i) => out.setAtIndex(ADDRESS, i, b) }
      val status = a.allocate(JAVA_INT)
      check(kfFilter.invoke(ctx, problem(a, fl, p, keepInit), out, status).asInstanceOf[Int], status.get(JAVA_INT, 0))
      val Array(m, c, at, rt, f, qq) = bufs.zip(sizes).map { case (b, k) => rowsOf(b, rows, k) }
      val tm = if (keepInit) (fl.times.min - 1.0) +: fl.times else fl.times
      Vector.tabulate(rows) { r =>
        val init = keepInit && r == 0
        KfState(tm(r), DenseVector(m(r)), new DenseMatrix(n, n, c(r)), DenseVector(at(r)),
                new DenseMatrix(n, n, rt(r)),
                if (init) None else Some(DenseVector(f(r))),
                if (init) None else Some(new DenseMatrix(q, q, qq(r))))
      }
    } finally a.close()
  }

  /** KalmanFilter.likelihood(mod, ys)(p) (KalmanFilter.scala:299-306). */
  def likelihood(mod: Dlm, ys: Vector[Data])(p: DlmParameters): Double = {
    val fl = flatten(mod, ys)
    val a = Arena.ofConfined()
    try {
      val tr = a.allocate(JAVA_DOUBLE); val st = a.allocate(JAVA_INT)
      check(loglik.invoke(ctx, problem(a, fl, p, true), tr, MemorySegment.NULL, st).asInstanceOf[Int], 0)
      tr.get(JAVA_DOUBLE, 0)
    } finally a.close()
  }

  // ---- Smoothing ----------------------------------------------------------------------------

  /** Smoothing.ffbsDlm (Smoothing.scala:173-180); consumes the same N(0,1) stream as the reference:
    * (T+1)*n draws, last row first. */
  def ffbsDlm(mod: Dlm, ys: Vector[Data], p: DlmParameters)(implicit rand: RandBasis = Rand): Rand[Vector[SamplingState]] = {
    val fl = flatten(mod, ys)
    val rows = fl.times.length + 1
    val n = fl.n
    val draws = Array.fill(rows)(Array.fill(n)(rand.gaussian(0, 1).draw)) // draw order: last row first
    val z = draws.reverse.flatten
    val a = Arena.ofConfined()
    try {
      val theta = a.allocate(JAVA_DOUBLE, rows.toLong * n)
      val kf = a.allocate(ADDRESS, 6)
      val sizes = Array(n, n * n, n, n * n)
      val bufs = sizes.map(k => a.allocate(JAVA_DOUBLE, rows.toLong * k))
      bufs.zipWithIndex.foreach { case (b, i) => kf.setAtIndex(ADDRESS, i, b) }
      kf.setAtIndex(ADDRESS, 4, MemorySegment.NULL); kf.setAtIndex(ADDRESS, 5, MemorySegment.NULL)
      val st = a.allocate(JAVA_INT)
      check(ffbs.invoke(ctx, problem(a, fl, p, true), seg(a, z), theta, kf, MemorySegment.NULL, st).asInstanceOf[Int], st.get(JAVA_INT, 0))
      val th = rowsOf(theta, rows, n)
      val Array(m, c, at, rt) = bufs.zip(sizes).map { case (b, k) => rowsOf(b, rows, k) }
      val tm = (fl.times.min - 1.0) +: fl.times
      Rand.always(Vector.tabulate(rows)(r =>
        SamplingState(tm(r), DenseVector(th(r)), DenseVector(m(r)), new DenseMatrix(n, n, c(r)),
                      DenseVector(at(r)), new DenseMatrix(n, n, rt(r)))))
    } finally a.close()
  }

  // backwardsSmoother, svdFilterDlm, svdFfbsDlm and the batched overloads follow the same
  // pattern over bdlm_rts_smooth, bdlm_svd_filter and bdlm_svd_ffbs (B > 1, per_series mask).
}

/** FilterAr.filterUnivariate (FilterAr.scala:37-47) over bdlm_ar_filter.
  * struct bdlm_ar_problem { int64 B; int32 T, layout, mem, process, per_series, v_mode;
  *                          const double *phi, *mu, *sigma_eta, *times, *v, *y; }  (80 bytes) */
object GpuFilterAr {
  import BdlmNative._
  import ValueLayout._
  val arProblem: StructLayout = MemoryLayout.structLayout(
    JAVA_LONG.withName("B"), JAVA_INT.withName("T"), JAVA_INT.withName("layout"),
    JAVA_INT.withName("mem"), JAVA_INT.withName("process"), JAVA_INT.withName("per_series"),
    JAVA_INT.withName("v_mode"),
    ADDRESS.withName("phi"), ADDRESS.withName("mu"), ADDRESS.withName("sigma_eta"),
    ADDRESS.withName("times"), ADDRESS.withName("v"), ADDRESS.withName("y"))

  def filterUnivariate(ys: Vector[(Double, Option[Double])], vs: Vector[Double],
                       p: SvParameters, ou: Boolean = false): Vector[FilterAr.FilterState] = {
    val a = Arena.ofConfined()
    try {
      val T = ys.size
      val rows = T + 1
      val pr = a.allocate(arProblem)
      pr.set(JAVA_LONG, 0, 1L); pr.set(JAVA_INT, 8, T); pr.set(JAVA_INT, 12, 1 /* SERIES_MAJOR */)
      pr.set(JAVA_INT, 16, 1 /* HOST */); pr.set(JAVA_INT, 20, if (ou) 1 else 0)
      pr.set(JAVA_INT, 24, 0); pr.set(JAVA_INT, 28, 1 /* BDLM_V_PER_STEP */)
      pr.set(ADDRESS, 32, GpuKalman.seg(a, Array(p.phi))); pr.set(ADDRESS, 40, GpuKalman.seg(a, Array(p.mu)))
      pr.set(ADDRESS, 48, GpuKalman.seg(a, Array(p.sigmaEta)))
      pr.set(ADDRESS, 56, GpuKalman.seg(a, ys.map(_._1).toArray))
      pr.set(ADDRESS, 64, GpuKalman.seg(a, vs.toArray))
      pr.set(ADDRESS, 72, GpuKalman.seg(a, ys.map(_._2.getOrElse(Double.NaN)).toArray))
      val out = a.allocate(ADDRESS, 4)
      val bufs = Array.fill(4)(a.allocate(JAVA_DOUBLE, rows.toLong))
      bufs.zipWithIndex.foreach { case (b, i) => out.setAtIndex(ADDRESS, i, b) }
      GpuKalman.check(arFilter.invoke(GpuKalman.ctx, pr, out).asInstanceOf[Int], 0)
      val t0 = if (ou) ys.head._1 else ys.head._1 - 1.0
      val tm = t0 +: ys.map(_._1)
      Vector.tabulate(rows)(r => FilterAr.FilterState(tm(r), bufs(0).getAtIndex(JAVA_DOUBLE, r),
        bufs(1).getAtIndex(JAVA_DOUBLE, r), bufs(2).getAtIndex(JAVA_DOUBLE, r), bufs(3).getAtIndex(JAVA_DOUBLE, r)))
    } finally a.close()
  }
  // ffbs(p, ys, vs): same problem struct + z (T + 1 normals drawn from the caller's RandBasis,
  // last row first) over bdlm_ar_ffbs; GpuFilterOu = the same with ou = true.
}

/** The conjugate half of a Gibbs sweep on the device (Gibbs.scala:41-49,72-77;
  * GibbsWishart.scala:16-35): GpuGibbs.sample keeps the per-chain V, W arrays in device memory and
  * alternates bdlm_ffbs(stats) / bdlm_gibbs_draw, copying back only the recorded draws. */
object GpuGibbs {
  // struct bdlm_gibbs_prior { double v_shape, v_scale, w_shape, w_scale, w_nu; const double *w_psi; }
  // struct bdlm_gibbs_rng   { uint64 seed, sweep; const double *gamma_v, *gamma_w, *bartlett; }
}

/** All GPUs of a node behind one object: `bdlm_comm_*` (include/bdlm.h).  The reference's only
  * parallel driver maps chains over futures on CPU cores (Streaming.scala:162-173); here one JVM
  * hands the whole batch to the library, which cuts it into contiguous blocks of series, one per
  * GPU, and runs them on its own host threads.  NCCL lives inside libbdlm.so (dlopen): the only
  * collectives are the sums of log-likelihoods / pooled Gibbs statistics and the chunk aggregates
  * of the time-sharded scan.  NOT compiled in this repository (no JVM in the image); the same calls
  * are exercised by bayesian_dlms_b200/comm.py in tests/test_gpu_comm.py. */
final class GpuComm(devices: Seq[Int]) extends AutoCloseable {
  import java.lang.foreign._
  import ValueLayout._
  private val linker = Linker.nativeLinker()
  private val lib    = SymbolLookup.libraryLookup("libbdlm.so", Arena.global())
  private def h(name: String, fd: FunctionDescriptor) = linker.downcallHandle(lib.find(name).get, fd)
  private val commCreate  = h("bdlm_comm_create", FunctionDescriptor.of(JAVA_INT, ADDRESS, JAVA_INT, JAVA_INT, JAVA_INT, ADDRESS, ADDRESS))
  private val commDestroy = h("bdlm_comm_destroy", FunctionDescriptor.ofVoid(ADDRESS))
  private val commErr     = h("bdlm_comm_last_error", FunctionDescriptor.of(ADDRESS, ADDRESS))
  private val commFs      = h("bdlm_comm_kf_filter_smooth", FunctionDescriptor.of(JAVA_INT, ADDRESS, ADDRESS, ADDRESS, ADDRESS, ADDRESS))
  private val commLoglik  = h("bdlm_comm_loglik", FunctionDescriptor.of(JAVA_INT, ADDRESS, ADDRESS, ADDRESS, ADDRESS, ADDRESS, ADDRESS))
  private val commFfbs    = h("bdlm_comm_ffbs", FunctionDescriptor.of(JAVA_INT, ADDRESS, ADDRESS, ADDRESS, ADDRESS, ADDRESS, ADDRESS, ADDRESS, ADDRESS))
  private val commScan    = h("bdlm_comm_scan_filter_smooth", FunctionDescriptor.of(JAVA_INT, ADDRESS, ADDRESS, ADDRESS, ADDRESS, ADDRESS))
  private val commSync    = h("bdlm_comm_sync", FunctionDescriptor.of(JAVA_INT, ADDRESS))

  private val handle: MemorySegment = {
    val a = Arena.ofConfined()
    try {
      val devs = a.allocateFrom(JAVA_INT, devices.toArray: _*)
      val out  = a.allocate(ADDRESS)
      val rc = commCreate.invoke(devs, devices.size, 0, devices.size, MemorySegment.NULL, out).asInstanceOf[Int]
      if (rc != 0) throw new IllegalStateException(
        commErr.invoke(MemorySegment.NULL).asInstanceOf[MemorySegment].reinterpret(4096).getString(0))
      out.get(ADDRESS, 0)
    } finally a.close()
  }

  /** Σ_b KalmanFilter.likelihood over EVERY series of the batch (all GPUs), e.g. the pooled
    * log-likelihood of a Metropolis step over parameters shared by the series
    * (MetropolisHastings.scala:126-137).  `problem` is a host-memory bdlm_problem for the whole
    * batch built exactly like GpuKalman.problem (B = number of series). */
  def loglikSums(problem: MemorySegment, transition: MemorySegment, innovations: MemorySegment,
                 status: MemorySegment): (Double, Double) = {
    val a = Arena.ofConfined()
    try {
      val sums = a.allocate(JAVA_DOUBLE, 2)
      val rc = commLoglik.invoke(handle, problem, transition, innovations, status, sums).asInstanceOf[Int]
      if (rc < 0) throw new IllegalArgumentException(
        commErr.invoke(handle).asInstanceOf[MemorySegment].reinterpret(4096).getString(0))
      (sums.getAtIndex(JAVA_DOUBLE, 0), sums.getAtIndex(JAVA_DOUBLE, 1))
    } finally a.close()
  }

  /** bdlm_comm_kf_filter_smooth: host arrays for the whole batch, cut over the GPUs. */
  def filterSmooth(problem: MemorySegment, kf: MemorySegment, sm: MemorySegment, status: MemorySegment): Int =
    commFs.invoke(handle, problem, kf, sm, status).asInstanceOf[Int]

  /** bdlm_comm_ffbs with pooled sufficient statistics (ss_y, n_y, ss_w, scatter summed over all chains). */
  def ffbs(problem: MemorySegment, z: MemorySegment, theta: MemorySegment, kf: MemorySegment,
           stats: MemorySegment, status: MemorySegment, pooled: MemorySegment): Int =
    commFfbs.invoke(handle, problem, z, theta, kf, stats, status, pooled).asInstanceOf[Int]

  /** ONE long series cut along time, one chunk per GPU (BASELINE config 5): enqueue-only. */
  def scanFilterSmooth(chunkProblems: MemorySegment, kfOuts: MemorySegment, smOuts: MemorySegment,
                       status: MemorySegment): Int = {
    val rc = commScan.invoke(handle, chunkProblems, kfOuts, smOuts, status).asInstanceOf[Int]
    if (rc == 0) commSync.invoke(handle).asInstanceOf[Int] else rc
  }

  override def close(): Unit = commDestroy.invoke(handle)
}
