/*
 * oracle/smx_driver.c — bulk-call helpers and the pthread CPU benchmark harness.
 *
 * TEST / BASELINE INFRASTRUCTURE ONLY.  Drives ANY library exporting the 8-function smatrix C
 * API (src/smatrix.h:87-94) through function pointers handed in by the caller (ctypes), so the
 * same loops run over oracle/_ref/libsmatrix_ref.so (the unmodified reference), over
 * oracle/liboracle.so (our restatement) and over the CUDA library's single-op API.
 *
 * The benchmark harness follows the shape of the reference's own driver
 * (src/smatrix_benchmark.c:98-132): T pthreads, the op stream split statically by index,
 * wall clock around create -> join.  Streams come from the counter-based generator
 * r = splitmix64(seed + i) that SURVEY.md 8(d) defines, so host and device see identical ops.
 */
#define _GNU_SOURCE
#include <pthread.h>
#include <stddef.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

typedef uint32_t (*fn_xyv)(void*, uint32_t, uint32_t, uint32_t);
typedef uint32_t (*fn_xy)(void*, uint32_t, uint32_t);
typedef uint32_t (*fn_x)(void*, uint32_t);
typedef uint32_t (*fn_row)(void*, uint32_t, uint32_t*, size_t);

/* ---------------------------------------------------------------- bulk helpers */

void drv_apply(fn_xyv op, void* h, const uint32_t* xs, const uint32_t* ys, const uint32_t* vs,
               size_t n, uint32_t* out) {
  for (size_t i = 0; i < n; i++) {
    uint32_t r = op(h, xs[i], ys[i], vs[i]);
    if (out) out[i] = r;
  }
}

void drv_get_many(fn_xy get, void* h, const uint32_t* xs, const uint32_t* ys, size_t n,
                  uint32_t* out) {
  for (size_t i = 0; i < n; i++) out[i] = get(h, xs[i], ys[i]);
}

void drv_rowlen_many(fn_x rowlen, void* h, const uint32_t* xs, size_t n, uint32_t* out) {
  for (size_t i = 0; i < n; i++) out[i] = rowlen(h, xs[i]);
}

/* Full rows as CSR.  The buffer handed to getrow is (rowlen + 2) pairs so that a column-0 entry
 * the running counter does not include is never truncated away.
 * Returns the number of pairs written, or (uint64_t)-1 if `cap_pairs` was too small. */
uint64_t drv_getrow_many(fn_x rowlen, fn_row getrow, void* h, const uint32_t* xs, size_t n,
                         uint64_t* offsets, uint32_t* pairs, uint64_t cap_pairs) {
  uint64_t at = 0;
  for (size_t i = 0; i < n; i++) {
    offsets[i] = at;
    uint64_t want = (uint64_t)rowlen(h, xs[i]) + 2;
    if (at + want > cap_pairs) return (uint64_t)-1;
    at += getrow(h, xs[i], pairs + 2 * at, (size_t)want * 8);
  }
  offsets[n] = at;
  return at;
}

/* ---------------------------------------------------------------- streams */

static inline uint64_t splitmix64(uint64_t z) {
  z += 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}

/* C2 build stream: op i of the uniform-random incr workload. */
static inline void c2_op(uint64_t seed, uint64_t i, uint32_t rows, uint32_t ycols, uint32_t* x,
                         uint32_t* y) {
  uint64_t r = splitmix64(seed + i);
  *x = (uint32_t)((r >> 32) % rows) * 2654435761u;
  *y = 1u + (uint32_t)(r & 0xFFFFFFFFull) % ycols;
}

/* C2 get stream: query j re-generates build op k = r % n_build; odd j shifts y out of range
 * (a guaranteed miss inside an existing row), so exactly half the queries hit. */
static inline void c2_query(uint64_t seed_get, uint64_t seed_build, uint64_t j, uint64_t n_build,
                            uint32_t rows, uint32_t ycols, uint32_t* x, uint32_t* y) {
  uint64_t k = splitmix64(seed_get + j) % n_build;
  c2_op(seed_build, k, rows, ycols, x, y);
  if (j & 1) *y += ycols;
}

void drv_gen_c2_ops(uint64_t seed, uint64_t first, size_t count, uint32_t rows, uint32_t ycols,
                    uint32_t* xs, uint32_t* ys) {
  for (size_t i = 0; i < count; i++) c2_op(seed, first + i, rows, ycols, &xs[i], &ys[i]);
}

void drv_gen_c2_queries(uint64_t seed_get, uint64_t seed_build, uint64_t first, size_t count,
                        uint64_t n_build, uint32_t rows, uint32_t ycols, uint32_t* xs,
                        uint32_t* ys) {
  for (size_t i = 0; i < count; i++)
    c2_query(seed_get, seed_build, first + i, n_build, rows, ycols, &xs[i], &ys[i]);
}

/* ---------------------------------------------------------------- pthread harness */

typedef struct {
  int kind; /* 0 = incr over the build stream, 1 = get over the query stream */
  void* fn;
  void* h;
  uint64_t seed, seed_build, first, count, n_build;
  uint32_t rows, ycols;
  uint64_t sink;
} job_t;

static void* worker(void* arg) {
  job_t* j = (job_t*)arg;
  uint32_t x, y;
  uint64_t acc = 0;
  if (j->kind == 0) {
    fn_xyv incr = (fn_xyv)j->fn;
    for (uint64_t i = 0; i < j->count; i++) {
      c2_op(j->seed, j->first + i, j->rows, j->ycols, &x, &y);
      acc += incr(j->h, x, y, 1);
    }
  } else {
    fn_xy get = (fn_xy)j->fn;
    for (uint64_t i = 0; i < j->count; i++) {
      c2_query(j->seed, j->seed_build, j->first + i, j->n_build, j->rows, j->ycols, &x, &y);
      acc += get(j->h, x, y);
    }
  }
  j->sink = acc;
  return NULL;
}

static double run_threads(job_t proto, int threads) {
  pthread_t* tid = malloc(sizeof(pthread_t) * (size_t)threads);
  job_t* jobs = malloc(sizeof(job_t) * (size_t)threads);
  struct timespec t0, t1;
  uint64_t per = proto.count / (uint64_t)threads;
  clock_gettime(CLOCK_MONOTONIC, &t0);
  for (int t = 0; t < threads; t++) {
    jobs[t] = proto;
    jobs[t].first = proto.first + per * (uint64_t)t;
    jobs[t].count = (t == threads - 1) ? proto.count - per * (uint64_t)t : per;
    pthread_create(&tid[t], NULL, worker, &jobs[t]);
  }
  for (int t = 0; t < threads; t++) pthread_join(tid[t], NULL);
  clock_gettime(CLOCK_MONOTONIC, &t1);
  free(tid);
  free(jobs);
  return (double)(t1.tv_sec - t0.tv_sec) + 1e-9 * (double)(t1.tv_nsec - t0.tv_nsec);
}

/* Apply build ops [first, first+count) with `threads` pthreads; returns wall seconds. */
double drv_bench_c2_incr(void* incr, void* h, int threads, uint64_t seed, uint64_t first,
                         uint64_t count, uint32_t rows, uint32_t ycols) {
  job_t p;
  memset(&p, 0, sizeof p);
  p.kind = 0;
  p.fn = incr;
  p.h = h;
  p.seed = seed;
  p.first = first;
  p.count = count;
  p.rows = rows;
  p.ycols = ycols;
  return run_threads(p, threads);
}

/* Run queries [first, first+count) against a table built from n_build ops; returns seconds. */
double drv_bench_c2_get(void* get, void* h, int threads, uint64_t seed_get, uint64_t seed_build,
                        uint64_t first, uint64_t count, uint64_t n_build, uint32_t rows,
                        uint32_t ycols) {
  job_t p;
  memset(&p, 0, sizeof p);
  p.kind = 1;
  p.fn = get;
  p.h = h;
  p.seed = seed_get;
  p.seed_build = seed_build;
  p.first = first;
  p.count = count;
  p.n_build = n_build;
  p.rows = rows;
  p.ycols = ycols;
  return run_threads(p, threads);
}

/* ---------------------------------------------------------------- full-scale parity digests */
/* Per row: {rowlen, number of pairs, sum of columns, sum of values, sum of column*value} (uint64,
 * wrapping), computed through the library's own rowlen / getrow calls.  Order independent, so it
 * can be compared with digests computed from another implementation's getrow output. */
void drv_row_digests(fn_x rowlen, fn_row getrow, void* h, const uint32_t* xs, size_t n, uint64_t* out5) {
  size_t cap = 1024;
  uint32_t* buf = malloc(cap * 8);
  for (size_t i = 0; i < n; i++) {
    uint64_t len = rowlen(h, xs[i]);
    if (len + 2 > cap) {
      cap = (len + 2) * 2;
      buf = realloc(buf, cap * 8);
    }
    uint32_t got = getrow(h, xs[i], buf, (size_t)(len + 2) * 8);
    uint64_t sc = 0, sv = 0, sp = 0;
    for (uint32_t k = 0; k < got; k++) {
      sc += buf[2 * k];
      sv += buf[2 * k + 1];
      sp += (uint64_t)buf[2 * k] * (uint64_t)buf[2 * k + 1];
    }
    out5[5 * i + 0] = len;
    out5[5 * i + 1] = got;
    out5[5 * i + 2] = sc;
    out5[5 * i + 3] = sv;
    out5[5 * i + 4] = sp;
  }
  free(buf);
}
