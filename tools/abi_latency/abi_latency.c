/* Per-call latency of the C ABI itself (what a JVM host pays through Panama / JNI, without any
 * Python in the way): the reference's JMH shape -- ONE series, polynomial(1), V = 3, W = 1,
 * m0 = 0, C0 = 1, T = 10 (benchmark/src/main/scala/bench/KalmanFilter.scala:10-33) -- and T = 1000,
 * host buffers, bdlm_kf_filter / bdlm_kf_filter_smooth / bdlm_ffbs.
 *
 *   gcc -O2 -Iinclude tools/abi_latency/abi_latency.c -o /tmp/abi_latency -ldl && /tmp/abi_latency bayesian_dlms_b200/libbdlm.so
 */
#include <dlfcn.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include "bdlm.h"

static double now_us(void) {
  struct timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return ts.tv_sec * 1e6 + ts.tv_nsec * 1e-3;
}

int main(int argc, char **argv) {
  void *h = dlopen(argc > 1 ? argv[1] : "bayesian_dlms_b200/libbdlm.so", RTLD_NOW);
  if (!h) { fprintf(stderr, "%s\n", dlerror()); return 1; }
  int (*create)(int, bdlm_ctx **) = dlsym(h, "bdlm_create");
  void (*destroy)(bdlm_ctx *) = dlsym(h, "bdlm_destroy");
  int (*filter)(bdlm_ctx *, const bdlm_problem *, const bdlm_kf_out *, int32_t *) = dlsym(h, "bdlm_kf_filter");
  int (*fsm)(bdlm_ctx *, const bdlm_problem *, const bdlm_kf_out *, const bdlm_smooth_out *, int32_t *) =
      dlsym(h, "bdlm_kf_filter_smooth");
  int (*ffbs)(bdlm_ctx *, const bdlm_problem *, const double *, double *, const bdlm_kf_out *,
              const bdlm_gibbs_stats *, int32_t *) = dlsym(h, "bdlm_ffbs");
  const char *(*last_error)(bdlm_ctx *) = dlsym(h, "bdlm_last_error");
  bdlm_ctx *ctx = NULL;
  if (create(0, &ctx)) { fprintf(stderr, "bdlm_create: %s\n", last_error(NULL)); return 2; }
  const int Ts[2] = {10, 1000};
  for (int c = 0; c < 2; ++c) {
    const int T = Ts[c], rows = T + 1;
    double one = 1.0, V = 3.0, W = 1.0, m0 = 0.0, C0 = 1.0;
    double *y = malloc(sizeof(double) * T), *z = malloc(sizeof(double) * rows);
    double *buf = malloc(sizeof(double) * rows * 9);
    for (int t = 0; t < T; ++t) y[t] = 0.1 * t + (t % 3);
    for (int t = 0; t < rows; ++t) z[t] = ((t * 7919) % 13 - 6) / 4.0;
    bdlm_problem pr;
    memset(&pr, 0, sizeof pr);
    pr.B = 1; pr.T = T; pr.n = 1; pr.p = 1; pr.layout = BDLM_SERIES_MAJOR; pr.mem = BDLM_HOST; pr.keep_init = 1;
    pr.F = &one; pr.G = &one; pr.V = &V; pr.W = &W; pr.m0 = &m0; pr.C0 = &C0; pr.y = y;
    bdlm_kf_out ko = {buf, buf + rows, buf + 2 * rows, buf + 3 * rows, buf + 4 * rows, buf + 5 * rows};
    bdlm_smooth_out so = {buf + 6 * rows, buf + 7 * rows};
    int32_t st = 0;
    const char *names[3] = {"bdlm_kf_filter", "bdlm_kf_filter_smooth", "bdlm_ffbs"};
    for (int op = 0; op < 3; ++op) {
      const int reps = T == 10 ? 2000 : 300;
      int rc = 0;
      for (int w = 0; w < 20 + reps && !rc; ++w) {
        static double t0;
        if (w == 20) t0 = now_us();
        if (op == 0) rc = filter(ctx, &pr, &ko, &st);
        else if (op == 1) rc = fsm(ctx, &pr, &ko, &so, &st);
        else rc = ffbs(ctx, &pr, z, buf + 8 * rows, NULL, NULL, &st);
        if (w == 20 + reps - 1) printf("T=%-5d %-22s %8.1f us/call (host buffers, %d calls, status %d)\n", T,
                                       names[op], (now_us() - t0) / reps, reps, st);
      }
      if (rc) { fprintf(stderr, "%s: %s\n", names[op], last_error(ctx)); return 3; }
    }
    free(y); free(z); free(buf);
  }
  destroy(ctx);
  return 0;
}
