/*
 * examples/c_host_example.c — a plain C host program on the drop-in boundary: nothing but
 * include/smatrix.h (the reference's eight functions, src/smatrix.h:87-94) and
 * include/smatrix_batch.h.  It checks known answers through the single-op API, then pushes a
 * counter-based random stream through the batched calls and verifies size-independent
 * properties (checksum of checksums, rowlen == row size, last writer wins, running return
 * values).  Exit code 0 and "c_host_example: OK" on success.
 *
 *   gcc -O2 -I include examples/c_host_example.c -L libsmatrix_b200/lib -lsmatrix_b200 \
 *       -Wl,-rpath,$PWD/libsmatrix_b200/lib -o c_host_example && ./c_host_example [n_ops]
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "smatrix.h"
#include "smatrix_batch.h"

#define CHECK(cond, ...)                                  \
  do {                                                    \
    if (!(cond)) {                                        \
      printf("c_host_example: FAILED %s:%d: ", __FILE__, __LINE__); \
      printf(__VA_ARGS__);                                \
      printf("\n");                                       \
      return 1;                                           \
    }                                                     \
  } while (0)

static uint64_t mix64(uint64_t z) { /* splitmix64, the stream generator of SURVEY.md 8(d) */
  z += 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}

int main(int argc, char** argv) {
  const size_t n = argc > 1 ? (size_t)strtoull(argv[1], NULL, 10) : 2000000;
  const uint32_t rows = 5000, cols = 64;

  smatrix_t* m = smatrix_open(NULL);
  CHECK(m != NULL, "smatrix_open(NULL) returned NULL");

  /* ---- single-op API: return values and wrap-around as the reference defines them */
  CHECK(smatrix_set(m, 42, 23, 17) == 17 && smatrix_get(m, 42, 23) == 17, "set/get");
  CHECK(smatrix_incr(m, 42, 23, 5) == 22 && smatrix_decr(m, 42, 23, 2) == 20, "incr/decr return the new value");
  CHECK(smatrix_decr(m, 7, 9, 3) == 0xFFFFFFFDu, "decr of a missing cell wraps to 2^32 - v");
  CHECK(smatrix_get(m, 1000000, 1) == 0 && smatrix_rowlen(m, 1000000) == 0, "missing row reads as 0");
  CHECK(smatrix_rowlen(m, 42) == 1 && smatrix_rowlen(m, 7) == 1, "rowlen");
  {
    uint32_t pair[2] = {0, 0};
    CHECK(smatrix_getrow(m, 42, pair, sizeof pair) == 1 && pair[0] == 23 && pair[1] == 20, "getrow");
  }
  smatrix_close(m);

  /* ---- batched calls on a fresh matrix */
  m = smatrix_open(NULL);
  CHECK(m != NULL, "second smatrix_open");
  uint32_t* xs = malloc(n * 4);
  uint32_t* ys = malloc(n * 4);
  uint32_t* out = malloc(n * 4);
  uint32_t* ids = malloc(rows * 4);
  uint32_t* lens = malloc(rows * 4);
  uint64_t* offs = malloc(((size_t)rows + 1) * 8);
  CHECK(xs && ys && out && ids && lens && offs, "out of host memory");
  for (size_t i = 0; i < n; i++) {
    const uint64_t r = mix64(2 + i);
    xs[i] = (uint32_t)((r >> 32) % rows) * 2654435761u; /* sparse 32-bit row ids */
    ys[i] = 1u + (uint32_t)(r & 0xFFFFFFFFu) % cols;    /* never column 0 */
  }
  smatrix_incr_batch(m, xs, ys, NULL, n); /* NULL: every value is 1 */

  smatrix_get_batch(m, xs, ys, n, out);
  for (size_t i = 0; i < n; i++) CHECK(out[i] >= 1, "op %zu: its cell reads %u", i, out[i]);

  for (uint32_t k = 0; k < rows; k++) ids[k] = k * 2654435761u;
  smatrix_rowlen_batch(m, ids, rows, lens);
  const uint64_t total = smatrix_getrow_batch(m, ids, rows, offs, NULL, 0); /* size query */
  uint32_t* pairs = malloc((size_t)total * 8 + 8);
  CHECK(pairs != NULL, "out of host memory");
  CHECK(smatrix_getrow_batch(m, ids, rows, offs, pairs, total) == total, "getrow_batch fill");
  uint64_t sum = 0;
  for (uint32_t k = 0; k < rows; k++) {
    CHECK(offs[k + 1] - offs[k] == lens[k], "row %u: %llu pairs but rowlen %u", k,
          (unsigned long long)(offs[k + 1] - offs[k]), lens[k]);
    for (uint64_t j = offs[k]; j < offs[k + 1]; j++) {
      CHECK(pairs[2 * j] >= 1 && pairs[2 * j] <= cols, "row %u: column %u out of range", k, pairs[2 * j]);
      sum += pairs[2 * j + 1];
    }
  }
  CHECK(sum == n, "sum of all values %llu != number of increments %zu", (unsigned long long)sum, n);
  CHECK(total <= (uint64_t)rows * cols, "more pairs than cells");

  /* ---- duplicate keys: set resolves to the last writer, *_batch_out returns the running values */
  {
    uint32_t dx[6] = {9, 9, 9, 9, 9, 9}, dy[6] = {4, 4, 5, 4, 5, 4}, dv[6] = {10, 20, 30, 40, 50, 60}, r[6];
    smatrix_set_batch(m, dx, dy, dv, 6);
    CHECK(smatrix_get(m, 9, 4) == 60 && smatrix_get(m, 9, 5) == 50, "set_batch: last writer in input order wins");
    smatrix_incr_batch_out(m, dx, dy, dv, 6, r);
    const uint32_t want[6] = {70, 90, 80, 130, 130, 190};
    for (int i = 0; i < 6; i++) CHECK(r[i] == want[i], "incr_batch_out[%d] = %u, want %u", i, r[i], want[i]);
  }
  smatrix_close(m);
  free(xs); free(ys); free(out); free(ids); free(lens); free(offs); free(pairs);
  printf("c_host_example: OK (%zu increments, %llu cells)\n", n, (unsigned long long)total);
  return 0;
}
