/*
 * examples/c_multi_gpu_example.c — a plain C host driving SEVERAL GPUs through include/smatrix_shard.h:
 * one thread per GPU (rank), no MPI / NCCL / Python — the ranks meet in a POSIX shared-memory segment,
 * batches are routed to the owners of their rows by the partition kernel's own stores over NVLink.
 * The same calls work with one PROCESS per GPU (give every process the same name and its own rank).
 *
 * Every rank pushes its slice of a counter-based random incr stream, then asks for a slice of ALL
 * keys (point gets), all row lengths and whole rows, and checks them against counts the host keeps in
 * a dense array.  Exit code 0 and "c_multi_gpu_example: OK" on success.
 *
 *   gcc -O2 -pthread -I include examples/c_multi_gpu_example.c -L libsmatrix_b200/lib -lsmatrix_b200 \
 *       -Wl,-rpath,$PWD/libsmatrix_b200/lib -o c_multi_gpu_example
 *   ./c_multi_gpu_example [ranks = 2] [n_ops = 4000000] [gpus = ranks]
 */
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>

#include "smatrix.h"
#include "smatrix_shard.h"

#define ROWS 3000u
#define COLS 96u /* columns 1 .. COLS; column 0 carries a per-row total like examples/cf_recommender.c */

static uint64_t mix64(uint64_t z) { /* splitmix64, the stream generator of SURVEY.md 8(d) */
  z += 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}
static void op_of(uint64_t i, uint32_t* x, uint32_t* y) {
  const uint64_t r = mix64(77 + i);
  *x = (uint32_t)((r >> 32) % ROWS) * 2654435761u;
  *y = (uint32_t)(r % (COLS + 1)); /* 0 .. COLS */
}

typedef struct {
  int rank, world, device;
  size_t n;
  const char* name;
  const uint32_t* expect; /* [ROWS][COLS + 1] */
  int failed;
} job_t;

#define CHECK(cond, ...)                                                     \
  do {                                                                       \
    if (!(cond)) {                                                           \
      printf("c_multi_gpu_example: rank %d FAILED %s:%d: ", j->rank, __FILE__, __LINE__); \
      printf(__VA_ARGS__);                                                   \
      printf("\n");                                                          \
      j->failed = 1;                                                         \
      return NULL;                                                           \
    }                                                                        \
  } while (0)

static void* rank_main(void* arg) {
  job_t* j = (job_t*)arg;
  smatrix_shard_t* sh = smatrix_b200_shard_open(j->name, j->rank, j->world, j->device);
  CHECK(sh != NULL, "smatrix_b200_shard_open returned NULL");
  /* my slice of the stream: ops [lo, hi) */
  const size_t lo = j->n * (size_t)j->rank / (size_t)j->world, hi = j->n * (size_t)(j->rank + 1) / (size_t)j->world;
  uint32_t* xs = malloc((hi - lo + 1) * 4);
  uint32_t* ys = malloc((hi - lo + 1) * 4);
  for (size_t i = lo; i < hi; i++) op_of(i, &xs[i - lo], &ys[i - lo]);
  /* the stream writes column 0 too, so the batch is ordered: rowlen is then bit-exact with a
   * sequential application of the whole collective batch */
  smatrix_b200_shard_incr_batch(sh, xs, ys, NULL, hi - lo, 1);

  /* every rank asks for a slice of ALL keys */
  const size_t keys = (size_t)ROWS * (COLS + 1);
  const size_t klo = keys * (size_t)j->rank / (size_t)j->world, khi = keys * (size_t)(j->rank + 1) / (size_t)j->world;
  uint32_t* qx = malloc((khi - klo) * 4);
  uint32_t* qy = malloc((khi - klo) * 4);
  uint32_t* out = malloc((khi - klo) * 4);
  for (size_t k = klo; k < khi; k++) {
    qx[k - klo] = (uint32_t)(k / (COLS + 1)) * 2654435761u;
    qy[k - klo] = (uint32_t)(k % (COLS + 1));
  }
  smatrix_b200_shard_get_batch(sh, qx, qy, khi - klo, out);
  for (size_t k = klo; k < khi; k++)
    CHECK(out[k - klo] == j->expect[k], "get(%u, %u) = %u, expected %u", qx[k - klo], qy[k - klo], out[k - klo], j->expect[k]);

  /* whole rows: rank r takes rows r, r + world, ...  (pairs compared as sums: order is the table's) */
  size_t nrows = 0;
  uint32_t* rows = malloc(ROWS * 4);
  for (uint32_t r = (uint32_t)j->rank; r < ROWS; r += (uint32_t)j->world) rows[nrows++] = r * 2654435761u;
  uint64_t* offs = malloc((nrows + 1) * 8);
  const uint64_t total = smatrix_b200_shard_getrow_batch(sh, rows, nrows, offs, NULL, 0); /* size query (collective) */
  uint32_t* pairs = malloc((total + 1) * 8);
  const uint64_t got = smatrix_b200_shard_getrow_batch(sh, rows, nrows, offs, pairs, total + 1);
  CHECK(got == total && offs[nrows] == total, "getrow totals differ: %llu / %llu", (unsigned long long)got, (unsigned long long)total);
  uint32_t* lens = malloc(nrows * 4);
  smatrix_b200_shard_rowlen_batch(sh, rows, nrows, lens);
  for (size_t i = 0; i < nrows; i++) {
    const uint32_t r = (uint32_t)j->rank + (uint32_t)i * (uint32_t)j->world;
    uint64_t want_n = 0, want_sum = 0, have_sum = 0;
    for (uint32_t c = 0; c <= COLS; c++)
      if (j->expect[(size_t)r * (COLS + 1) + c]) { want_n++; want_sum += (uint64_t)c * 1000003u + j->expect[(size_t)r * (COLS + 1) + c]; }
    for (uint64_t p = offs[i]; p < offs[i + 1]; p++) have_sum += (uint64_t)pairs[2 * p] * 1000003u + pairs[2 * p + 1];
    CHECK(offs[i + 1] - offs[i] == want_n && have_sum == want_sum, "row %u: %llu pairs (expected %llu) or wrong contents", r,
          (unsigned long long)(offs[i + 1] - offs[i]), (unsigned long long)want_n);
    CHECK(lens[i] == want_n || lens[i] + 1 == want_n, "rowlen(%u) = %u with %llu pairs", r, lens[i], (unsigned long long)want_n);
  }
  /* the shards together hold every op exactly once */
  const uint64_t applied = smatrix_b200_shard_sum(sh, hi - lo);
  CHECK(applied == j->n, "ops applied %llu != %zu", (unsigned long long)applied, j->n);
  free(xs); free(ys); free(qx); free(qy); free(out); free(rows); free(offs); free(pairs); free(lens);
  smatrix_b200_shard_close(sh);
  return NULL;
}

int main(int argc, char** argv) {
  const int world = argc > 1 ? atoi(argv[1]) : 2;
  const size_t n = argc > 2 ? (size_t)strtoull(argv[2], NULL, 10) : 4000000;
  const int gpus = argc > 3 ? atoi(argv[3]) : world;
  if (world < 1 || world > 64 || gpus < 1) { printf("usage: c_multi_gpu_example [ranks] [n_ops] [gpus]\n"); return 2; }
  uint32_t* expect = calloc((size_t)ROWS * (COLS + 1), 4);
  for (size_t i = 0; i < n; i++) {
    uint32_t x, y;
    op_of(i, &x, &y);
    expect[(size_t)((x * 244002641u) % ROWS) * (COLS + 1) + y]++; /* 244002641 = 2654435761^-1 mod 2^32: x -> row number */
  }
  char name[64];
  snprintf(name, sizeof name, "smx_c_example_%d", (int)getpid());
  pthread_t tid[64];
  job_t jobs[64];
  for (int r = 0; r < world; r++) {
    jobs[r] = (job_t){r, world, r % gpus, n, name, expect, 0};
    pthread_create(&tid[r], NULL, rank_main, &jobs[r]);
  }
  int failed = 0;
  for (int r = 0; r < world; r++) { pthread_join(tid[r], NULL); failed |= jobs[r].failed; }
  free(expect);
  if (failed) return 1;
  printf("c_multi_gpu_example: OK (%d ranks on %d GPU(s), %zu ops)\n", world, gpus, n);
  return 0;
}
