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

/* ---- C3 / C4 streams: the same definitions as the device generators (csrc/smx_kernels.cu) ---- */
static inline uint32_t draw(const uint64_t* thr, uint32_t m, uint64_t r) {
  uint32_t lo = 0, hi = m - 1u;
  while (lo < hi) {
    uint32_t mid = lo + ((hi - lo) >> 1);
    if (thr[mid] < r) lo = mid + 1u; else hi = mid;
  }
  return lo + 1u;
}
/* examples/cf_recommender.c:35-47: per basket of 8 ids, for n: incr(ids[n],0,1); for i != n:
 * incr(ids[n], ids[i], 1) — op k = i % 64 of basket i / 64 */
static inline void c3_op(uint64_t seed, uint64_t i, const uint64_t* thr, uint32_t items, uint32_t* x,
                         uint32_t* y) {
  uint64_t b = i >> 6;
  uint32_t k = (uint32_t)(i & 63u), n = k >> 3, j = k & 7u;
  *x = draw(thr, items, splitmix64(seed + b * 8u + n));
  if (j == 0) { *y = 0; return; }
  uint32_t other = (j - 1u < n) ? j - 1u : j;
  *y = draw(thr, items, splitmix64(seed + b * 8u + other));
}
void drv_gen_c3_ops(uint64_t seed, uint64_t first, size_t count, const uint64_t* thr, uint32_t items,
                    uint32_t* xs, uint32_t* ys) {
  for (size_t i = 0; i < count; i++) c3_op(seed, first + i, thr, items, &xs[i], &ys[i]);
}
void drv_gen_c3_queries(uint64_t seed_get, uint64_t seed_build, uint64_t first, size_t count,
                        uint64_t n_build, const uint64_t* thr, uint32_t items, uint32_t* xs, uint32_t* ys) {
  for (size_t i = 0; i < count; i++) {
    uint64_t j = first + i;
    c3_op(seed_build, splitmix64(seed_get + j) % n_build, thr, items, &xs[i], &ys[i]);
    if (j & 1) ys[i] += items + 1u;
  }
}
void drv_gen_c4_lens(uint64_t seed, uint64_t first, size_t count, const uint64_t* thr, uint32_t kmax,
                     uint32_t* lens) {
  for (size_t i = 0; i < count; i++) lens[i] = draw(thr, kmax, splitmix64(seed + first + i));
}
static inline uint32_t fmix32(uint32_t h) {
  h ^= h >> 16; h *= 0x85ebca6bu; h ^= h >> 13; h *= 0xc2b2ae35u; h ^= h >> 16;
  return h;
}
void drv_gen_c4_ops(uint64_t seed, uint64_t first, size_t count, const uint64_t* offs, uint32_t rows,
                    uint32_t* xs, uint32_t* ys, uint32_t* vs) {
  for (size_t q = 0; q < count; q++) {
    uint64_t i = first + q;
    uint32_t lo = 0, hi = rows - 1u;
    while (lo < hi) {
      uint32_t mid = lo + ((hi - lo + 1u) >> 1);
      if (offs[mid] <= i) lo = mid; else hi = mid - 1u;
    }
    uint32_t j = (uint32_t)(i - offs[lo]);
    uint32_t salt = 1u + (uint32_t)(splitmix64((seed ^ 0xC4C4C4C4ull) + lo) % 0xFFE00000ull);
    xs[q] = lo * 2654435761u;
    ys[q] = fmix32(j + salt);
    vs[q] = (uint32_t)(splitmix64(seed + 0x5EED0000ull + i) >> 32) | 1u;
  }
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

/* ---- the same harness over PRE-GENERATED arrays (any workload): thread t applies the ops
 * [t*n/T, (t+1)*n/T) in order; vs == NULL means every value is 1.  Returns wall seconds. */
typedef struct {
  int kind; /* 0 = write op (incr/set/decr), 1 = get, 2 = rowlen + getrow of whole rows */
  void *fn, *fn2, *h;
  const uint32_t *xs, *ys, *vs;
  size_t first, count;
  uint64_t sink;
} ajob_t;
static void* aworker(void* arg) {
  ajob_t* j = (ajob_t*)arg;
  uint64_t acc = 0;
  if (j->kind == 0) {
    fn_xyv op = (fn_xyv)j->fn;
    for (size_t i = j->first; i < j->first + j->count; i++) acc += op(j->h, j->xs[i], j->ys[i], j->vs ? j->vs[i] : 1u);
  } else if (j->kind == 1) {
    fn_xy get = (fn_xy)j->fn;
    for (size_t i = j->first; i < j->first + j->count; i++) acc += get(j->h, j->xs[i], j->ys[i]);
  } else {
    fn_x rowlen = (fn_x)j->fn;
    fn_row getrow = (fn_row)j->fn2;
    size_t cap = 1024;
    uint32_t* buf = malloc(cap * 8);
    for (size_t i = j->first; i < j->first + j->count; i++) {
      uint64_t len = rowlen(j->h, j->xs[i]);
      if (len + 2 > cap) { cap = (len + 2) * 2; buf = realloc(buf, cap * 8); }
      acc += getrow(j->h, j->xs[i], buf, (size_t)(len + 2) * 8); /* pairs returned */
    }
    free(buf);
  }
  j->sink = acc;
  return NULL;
}
static double run_athreads(ajob_t proto, size_t n, int threads, uint64_t* sink) {
  pthread_t* tid = malloc(sizeof(pthread_t) * (size_t)threads);
  ajob_t* jobs = malloc(sizeof(ajob_t) * (size_t)threads);
  struct timespec t0, t1;
  size_t per = n / (size_t)threads;
  clock_gettime(CLOCK_MONOTONIC, &t0);
  for (int t = 0; t < threads; t++) {
    jobs[t] = proto;
    jobs[t].first = per * (size_t)t;
    jobs[t].count = (t == threads - 1) ? n - per * (size_t)t : per;
    pthread_create(&tid[t], NULL, aworker, &jobs[t]);
  }
  uint64_t acc = 0;
  for (int t = 0; t < threads; t++) { pthread_join(tid[t], NULL); acc += jobs[t].sink; }
  clock_gettime(CLOCK_MONOTONIC, &t1);
  if (sink) *sink = acc;
  free(tid);
  free(jobs);
  return (double)(t1.tv_sec - t0.tv_sec) + 1e-9 * (double)(t1.tv_nsec - t0.tv_nsec);
}
double drv_bench_apply(void* op, void* h, int threads, const uint32_t* xs, const uint32_t* ys,
                       const uint32_t* vs, size_t n) {
  ajob_t p;
  memset(&p, 0, sizeof p);
  p.kind = 0; p.fn = op; p.h = h; p.xs = xs; p.ys = ys; p.vs = vs;
  return run_athreads(p, n, threads, NULL);
}
double drv_bench_get(void* get, void* h, int threads, const uint32_t* xs, const uint32_t* ys, size_t n) {
  ajob_t p;
  memset(&p, 0, sizeof p);
  p.kind = 1; p.fn = get; p.h = h; p.xs = xs; p.ys = ys;
  return run_athreads(p, n, threads, NULL);
}
/* rowlen + getrow (full-size buffer) of n rows; *pairs_out = pairs returned in total */
double drv_bench_getrow(void* rowlen, void* getrow, void* h, int threads, const uint32_t* xs, size_t n,
                        uint64_t* pairs_out) {
  ajob_t p;
  memset(&p, 0, sizeof p);
  p.kind = 2; p.fn = rowlen; p.fn2 = getrow; p.h = h; p.xs = xs;
  return run_athreads(p, n, threads, pairs_out);
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
