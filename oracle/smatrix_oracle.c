/*
 * oracle/smatrix_oracle.c — single-threaded CPU restatement of the reference's nested
 * open-addressing hash maps (row directory "cmap" -> per-row column map "rmap").
 *
 * TEST INFRASTRUCTURE ONLY (see smatrix_oracle.h).  Written from the reference's behaviour,
 * not from its text: no locks, no file mode, index-based instead of pointer-based.  Every
 * function names the reference lines whose behaviour it restates (paths relative to the
 * reference checkout).
 *
 * What has to be reproduced for bit-exact results:
 *   - column map: identity hash `y % size`, linear probing, EMPTY <=> key==0 && value==0
 *     (src/smatrix.c:363-380); grow x2 when used > size/2 BEFORE placing (:343-360); growth
 *     re-places every non-empty cell in table order, recounting `used` (:383-416)
 *   - a write to column 0 never goes through the placing path, because the probe for key 0
 *     stops on any key==0 cell and the caller treats `cell.key == y` as a hit (:297-300)
 *   - rowlen returns the running `used` counter (:212-223), getrow walks the table in order
 *     and stops once `++num * 8 >= ret_len` (:189-210)
 *   - directory: `x % size`, linear probing over {used flag, key}, grow x2 when
 *     used*4 >= size*3 (:673-741), initial 65536 entries (src/smatrix.h:24), rows start with
 *     16 cells (src/smatrix.h:21)
 */
#include "smatrix_oracle.h"

#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define ROW_CELLS0 16u    /* src/smatrix.h:21 SMATRIX_RMAP_INITIAL_SIZE */
#define DIR_ENTRIES0 65536u /* src/smatrix.h:24 SMATRIX_CMAP_INITIAL_SIZE */

typedef struct {
  uint32_t key;
  uint32_t val;
} cell_t;

typedef struct {
  uint32_t size; /* capacity in cells, power of two */
  uint32_t used; /* the reference's running counter, NOT a live count */
  cell_t* cell;
} row_t;

typedef struct {
  uint32_t taken;
  uint32_t x;
  row_t* row;
} dirent_t;

struct smx_oracle_s {
  uint64_t dsize;
  uint64_t dused;
  dirent_t* dir;
};

static void* zalloc(size_t bytes) {
  void* p = calloc(1, bytes);
  if (!p) {
    fprintf(stderr, "smx_oracle: out of memory\n");
    abort();
  }
  return p;
}

static inline int cell_is_empty(const cell_t* c) { return c->key == 0 && c->val == 0; }

/* src/smatrix.c:363-380 — index of the cell where the probe for `y` stops. */
static uint32_t row_probe(const row_t* r, uint32_t y) {
  uint64_t at = y % r->size;
  for (uint64_t step = 0; step < r->size; step++) {
    const cell_t* c = &r->cell[at];
    if (c->key == y || cell_is_empty(c)) break;
    at = (at + 1) % r->size;
  }
  return (uint32_t)at;
}

static uint32_t row_place(row_t* r, uint32_t y);

/* src/smatrix.c:383-416 — double the row and re-place every non-empty cell in table order. */
static void row_double(row_t* r) {
  row_t bigger;
  bigger.size = r->size * 2;
  bigger.used = 0;
  bigger.cell = zalloc(sizeof(cell_t) * bigger.size);
  for (uint32_t i = 0; i < r->size; i++) {
    if (cell_is_empty(&r->cell[i])) continue;
    uint32_t at = row_place(&bigger, r->cell[i].key);
    bigger.cell[at].val = r->cell[i].val;
  }
  free(r->cell);
  *r = bigger;
}

/* src/smatrix.c:343-360 — make sure `y` owns a cell; counts it in `used` when newly claimed. */
static uint32_t row_place(row_t* r, uint32_t y) {
  if (r->used > r->size / 2) row_double(r);
  uint32_t at = row_probe(r, y);
  cell_t* c = &r->cell[at];
  if (c->key == 0 || c->key != y) {
    r->used++;
    c->key = y;
    c->val = 0;
  }
  return at;
}

/* src/smatrix.c:673-693 — directory probe: stops on a free entry or on the key. */
static uint64_t dir_probe(const smx_oracle_t* m, uint32_t x) {
  uint32_t walk = x; /* the reference walks an `unsigned` that wraps at 2^32 */
  for (;;) {
    uint64_t at = walk % m->dsize;
    if (!m->dir[at].taken || m->dir[at].x == x) return at;
    walk++;
  }
}

static uint64_t dir_place(smx_oracle_t* m, uint32_t x);

/* src/smatrix.c:715-741 */
static void dir_double(smx_oracle_t* m) {
  smx_oracle_t bigger;
  bigger.dsize = m->dsize * 2;
  bigger.dused = 0;
  bigger.dir = zalloc(sizeof(dirent_t) * bigger.dsize);
  for (uint64_t i = 0; i < m->dsize; i++) {
    if (!m->dir[i].taken) continue;
    uint64_t at = dir_place(&bigger, m->dir[i].x);
    bigger.dir[at].row = m->dir[i].row;
  }
  free(m->dir);
  *m = bigger;
}

/* src/smatrix.c:695-713 */
static uint64_t dir_place(smx_oracle_t* m, uint32_t x) {
  if (m->dused * 4 >= m->dsize * 3) dir_double(m);
  uint64_t at = dir_probe(m, x);
  dirent_t* e = &m->dir[at];
  if (!e->taken || e->x != x) {
    m->dused++;
    e->taken = 1;
    e->x = x;
    e->row = NULL;
  }
  return at;
}

/* src/smatrix.c:621-670 without the locking: find the row, optionally creating it. */
static row_t* find_row(smx_oracle_t* m, uint32_t x, int create) {
  uint64_t at = dir_probe(m, x);
  if (m->dir[at].taken && m->dir[at].x == x) return m->dir[at].row;
  if (!create) return NULL;
  row_t* r = zalloc(sizeof(row_t));
  r->size = ROW_CELLS0;
  r->used = 0;
  r->cell = zalloc(sizeof(cell_t) * ROW_CELLS0);
  at = dir_place(m, x);
  m->dir[at].row = r;
  return r;
}

/* src/smatrix.c:258-304 — resolve (x,y) to a cell; writers create row and cell. */
static cell_t* resolve(smx_oracle_t* m, uint32_t x, uint32_t y, int write, row_t** row_out) {
  row_t* r = find_row(m, x, write);
  if (row_out) *row_out = r;
  if (!r) return NULL;
  uint32_t at = row_probe(r, y);
  if (r->cell[at].key == y) return &r->cell[at];
  if (!write) return NULL;
  at = row_place(r, y);
  return &r->cell[at];
}

smx_oracle_t* smx_oracle_open(const char* fname) {
  if (fname) return NULL; /* memory mode only */
  smx_oracle_t* m = zalloc(sizeof(*m));
  m->dsize = DIR_ENTRIES0; /* src/smatrix.c:598-612 */
  m->dused = 0;
  m->dir = zalloc(sizeof(dirent_t) * m->dsize);
  return m;
}

void smx_oracle_close(smx_oracle_t* m) { /* src/smatrix.c:113-133 */
  if (!m) return;
  for (uint64_t i = 0; i < m->dsize; i++) {
    if (m->dir[i].taken && m->dir[i].row) {
      free(m->dir[i].row->cell);
      free(m->dir[i].row);
    }
  }
  free(m->dir);
  free(m);
}

uint32_t smx_oracle_get(smx_oracle_t* m, uint32_t x, uint32_t y) { /* src/smatrix.c:174-185 */
  cell_t* c = resolve(m, x, y, 0, NULL);
  return c ? c->val : 0;
}

uint32_t smx_oracle_set(smx_oracle_t* m, uint32_t x, uint32_t y, uint32_t v) { /* :225-234 */
  cell_t* c = resolve(m, x, y, 1, NULL);
  c->val = v;
  return c->val;
}

uint32_t smx_oracle_incr(smx_oracle_t* m, uint32_t x, uint32_t y, uint32_t v) { /* :236-245 */
  cell_t* c = resolve(m, x, y, 1, NULL);
  c->val += v;
  return c->val;
}

uint32_t smx_oracle_decr(smx_oracle_t* m, uint32_t x, uint32_t y, uint32_t v) { /* :247-256 */
  cell_t* c = resolve(m, x, y, 1, NULL);
  c->val -= v;
  return c->val;
}

uint32_t smx_oracle_rowlen(smx_oracle_t* m, uint32_t x) { /* src/smatrix.c:212-223 */
  row_t* r = find_row(m, x, 0);
  return r ? r->used : 0;
}

uint32_t smx_oracle_getrow(smx_oracle_t* m, uint32_t x, uint32_t* ret, size_t ret_len) { /* :189-210 */
  row_t* r = find_row(m, x, 0);
  uint32_t n = 0;
  if (!r) return 0;
  for (uint32_t i = 0; i < r->size; i++) {
    if (cell_is_empty(&r->cell[i])) continue;
    ret[2 * n] = r->cell[i].key;
    ret[2 * n + 1] = r->cell[i].val;
    n++;
    if ((size_t)n * 8 >= ret_len) break;
  }
  return n;
}

uint64_t smx_oracle_nrows(smx_oracle_t* m) { return m->dused; }
uint64_t smx_oracle_dirsize(smx_oracle_t* m) { return m->dsize; }
uint32_t smx_oracle_rowsize(smx_oracle_t* m, uint32_t x) {
  row_t* r = find_row(m, x, 0);
  return r ? r->size : 0;
}
