/* ORACLE (test infrastructure, not product code) -- multi-threaded CPU sparse x dense product used to time the
 * reference's CPU path "with all the host threads it can use" (bench.py cpu_baseline / --impl reference).
 * It computes exactly what the reference's `Y*C` / `U*C` does (src/primal/ManiSDP_onlyunitdiag.m:118,128) in the
 * oracle's row layout: out(i,:) = sum_e val[e] * U(col[e],:) over the CSR row i of the symmetric C.
 * Same FP64 arithmetic as SciPy's csr_matvecs (one accumulation per entry, in entry order); only the rows are spread
 * over POSIX threads (the image has no libgomp).   gcc -O3 -shared -fPIC -pthread -o liboracle_spmm.so spmm_omp.c */
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>

static int g_threads = 0;

int oracle_spmm_threads(void) {
  if (g_threads <= 0) {
    const char* e = getenv("ORACLE_THREADS");
    long t = e ? atol(e) : sysconf(_SC_NPROCESSORS_ONLN);
    if (t < 1) t = 1;
    if (t > 256) t = 256;
    g_threads = (int)t;
  }
  return g_threads;
}

typedef struct {
  int64_t r0, r1, p;
  const int32_t *rowptr, *col;
  const double *val, *U;
  double* out;
} job_t;

static void* worker(void* arg) {
  const job_t* j = (const job_t*)arg;
  const int64_t p = j->p;
  for (int64_t i = j->r0; i < j->r1; ++i) {
    double* o = j->out + i * p;
    memset(o, 0, (size_t)p * sizeof(double));
    for (int32_t e = j->rowptr[i]; e < j->rowptr[i + 1]; ++e) {
      const double w = j->val[e];
      const double* u = j->U + (int64_t)j->col[e] * p;
      for (int64_t c = 0; c < p; ++c) o[c] += w * u[c];
    }
  }
  return 0;
}

void oracle_csr_spmm(int64_t n, const int32_t* rowptr, const int32_t* col, const double* val, const double* U,
                     int64_t p, double* out) {
  const int T = oracle_spmm_threads();
  pthread_t th[256];
  job_t jobs[256];
  const int64_t chunk = (n + T - 1) / T;
  int started = 0;
  for (int t = 0; t < T; ++t) {
    job_t* j = &jobs[t];
    j->r0 = t * chunk < n ? t * chunk : n;
    j->r1 = (t + 1) * chunk < n ? (t + 1) * chunk : n;
    j->p = p; j->rowptr = rowptr; j->col = col; j->val = val; j->U = U; j->out = out;
    if (t == T - 1 || pthread_create(&th[t], 0, worker, j) != 0) {
      worker(j); /* last slice (or a failed create) runs on the calling thread */
      if (t != T - 1) th[t] = 0;
    } else {
      started |= 0; /* joined below */
    }
  }
  for (int t = 0; t < T - 1; ++t)
    if (th[t]) pthread_join(th[t], 0);
  (void)started;
}
