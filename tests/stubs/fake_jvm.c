/* tests/stubs/fake_jvm.c — TEST SCAFFOLDING: a toy JNIEnv (tests/stubs/jni.h) that is just enough to
 * EXECUTE the reference's unchanged src/smatrix_jni.c natives and examples/jni/smatrix_jni_batch.c
 * against our library without a JVM: one "SparseMatrix object" = a struct with the `ptr` long field,
 * arrays = {length, data}, the row map of getRowNative = an array that putIntTuple appends to.
 * main() plays the Java side: init -> incrBatch / setBatch -> get / getBatch / getRowNative /
 * getRowsNative must agree with each other and with a host-side tally.  Prints "fake_jvm: OK". */
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <jni.h>

typedef struct { jlong ptr; int thrown; } obj_t;          /* a SparseMatrix instance */
typedef struct { jsize len; int elem; void* data; } arr_t; /* int[] / long[] */
typedef struct { jint n, cap; jint* kv; } map_t;           /* the SparseMatrix row map (putIntTuple) */
static int g_thrown;

static jclass f_FindClass(JNIEnv* e, const char* n) { (void)e; return (jclass)n; }
static jint f_ThrowNew(JNIEnv* e, jclass c, const char* m) { (void)e; (void)c; fprintf(stderr, "fake_jvm: exception: %s\n", m); g_thrown++; return 0; }
static jfieldID f_GetFieldID(JNIEnv* e, jclass c, const char* n, const char* s) { (void)e; (void)c; (void)s; return (jfieldID)n; }
static void f_SetLongField(JNIEnv* e, jobject o, jfieldID f, jlong v) { (void)e; (void)f; ((obj_t*)o)->ptr = v; }
static jlong f_GetLongField(JNIEnv* e, jobject o, jfieldID f) { (void)e; (void)f; return ((obj_t*)o)->ptr; }
static const char* f_GetStringUTFChars(JNIEnv* e, jstring s, jboolean* c) { (void)e; (void)c; return (const char*)s; }
static void f_ReleaseStringUTFChars(JNIEnv* e, jstring s, const char* c) { (void)e; (void)s; (void)c; }
static jclass f_GetObjectClass(JNIEnv* e, jobject o) { (void)e; return o; }
static jmethodID f_GetMethodID(JNIEnv* e, jclass c, const char* n, const char* s) { (void)e; (void)c; (void)s; return (jmethodID)n; }
static void f_CallVoidMethod(JNIEnv* e, jobject o, jmethodID m, ...) { /* map.putIntTuple(int, int) */
  (void)e; (void)m;
  map_t* map = (map_t*)o;
  va_list ap;
  va_start(ap, m);
  jint k = va_arg(ap, jint), v = va_arg(ap, jint);
  va_end(ap);
  if (map->n == map->cap) { map->cap = map->cap ? 2 * map->cap : 64; map->kv = realloc(map->kv, sizeof(jint) * 2 * (size_t)map->cap); }
  map->kv[2 * map->n] = k; map->kv[2 * map->n + 1] = v; map->n++;
}
static jsize f_GetArrayLength(JNIEnv* e, jarray a) { (void)e; return ((arr_t*)a)->len; }
static void* f_GetCritical(JNIEnv* e, jarray a, jboolean* c) { (void)e; (void)c; return ((arr_t*)a)->data; }
static void f_ReleaseCritical(JNIEnv* e, jarray a, void* p, jint m) { (void)e; (void)a; (void)p; (void)m; }
static arr_t* new_arr(jsize n, int elem) { arr_t* a = malloc(sizeof *a); a->len = n; a->elem = elem; a->data = calloc((size_t)n + 1, (size_t)elem); return a; }
static jintArray f_NewIntArray(JNIEnv* e, jsize n) { (void)e; return new_arr(n, 4); }
static jint* f_GetInts(JNIEnv* e, jintArray a, jboolean* c) { (void)e; (void)c; return ((arr_t*)a)->data; }
static void f_ReleaseInts(JNIEnv* e, jintArray a, jint* p, jint m) { (void)e; (void)a; (void)p; (void)m; }
static jlong* f_GetLongs(JNIEnv* e, jlongArray a, jboolean* c) { (void)e; (void)c; return ((arr_t*)a)->data; }
static void f_ReleaseLongs(JNIEnv* e, jlongArray a, jlong* p, jint m) { (void)e; (void)a; (void)p; (void)m; }

static const struct JNINativeInterface_ g_table = {
    f_FindClass, f_ThrowNew, f_GetFieldID, f_SetLongField, f_GetLongField, f_GetStringUTFChars,
    f_ReleaseStringUTFChars, f_GetObjectClass, f_GetMethodID, f_CallVoidMethod, f_GetArrayLength,
    f_GetCritical, f_ReleaseCritical, f_NewIntArray, f_GetInts, f_ReleaseInts, f_GetLongs, f_ReleaseLongs};

/* the natives under test */
#define _JM(X) Java_com_paulasmuth_libsmatrix_SparseMatrix_##X
void _JM(init)(JNIEnv*, jobject, jstring);
void _JM(close)(JNIEnv*, jobject);
jint _JM(get)(JNIEnv*, jobject, jint, jint);
void _JM(incr)(JNIEnv*, jobject, jint, jint, jint);
jint _JM(getRowLength)(JNIEnv*, jobject, jint);
void _JM(getRowNative)(JNIEnv*, jobject, jint, jobject, jint);
void _JM(incrBatch)(JNIEnv*, jobject, jintArray, jintArray, jintArray);
void _JM(setBatch)(JNIEnv*, jobject, jintArray, jintArray, jintArray);
jintArray _JM(getBatch)(JNIEnv*, jobject, jintArray, jintArray);
jintArray _JM(getRowsNative)(JNIEnv*, jobject, jintArray, jlongArray);

#define CHECK(c, ...) do { if (!(c)) { printf("fake_jvm: FAILED %s:%d: ", __FILE__, __LINE__); printf(__VA_ARGS__); printf("\n"); return 1; } } while (0)
enum { ROWS = 300, COLS = 40, N = 60000 };

int main(void) {
  JNIEnv env_ = &g_table;
  JNIEnv* env = &env_;
  obj_t m = {0, 0};
  _JM(init)(env, &m, NULL);                                  /* reference glue, unchanged: smatrix_open(NULL) */
  CHECK(m.ptr != 0 && !g_thrown, "init");
  static jint tally[ROWS][COLS + 1];
  arr_t *xs = new_arr(N, 4), *ys = new_arr(N, 4), *vs = new_arr(N, 4);
  unsigned long long z = 12345;
  for (int i = 0; i < N; i++) {
    z = z * 6364136223846793005ull + 1442695040888963407ull;
    jint x = (jint)((z >> 33) % ROWS), y = (jint)((z >> 13) % COLS) + 1, v = (jint)((z >> 50) % 7);
    ((jint*)xs->data)[i] = x * 7 + 1; ((jint*)ys->data)[i] = y; ((jint*)vs->data)[i] = v;
    tally[x][y] += v;
  }
  _JM(incrBatch)(env, &m, xs, ys, vs);                       /* new native: one call for the whole batch */
  _JM(incrBatch)(env, &m, xs, ys, NULL);                     /* vals == null: every value is 1 */
  for (int i = 0; i < N; i++) tally[(((jint*)xs->data)[i] - 1) / 7][((jint*)ys->data)[i]] += 1;
  _JM(incr)(env, &m, 1, 5, 1000);                            /* the reference's single-op native still works */
  tally[0][5] += 1000;
  /* getBatch == get == tally */
  arr_t* got = _JM(getBatch)(env, &m, xs, ys);
  CHECK(got && got->len == N, "getBatch length");
  for (int i = 0; i < N; i += 97) {
    jint x = ((jint*)xs->data)[i], y = ((jint*)ys->data)[i];
    CHECK(((jint*)got->data)[i] == tally[(x - 1) / 7][y], "getBatch[%d]", i);
    CHECK(_JM(get)(env, &m, x, y) == tally[(x - 1) / 7][y], "get(%d, %d)", x, y);
  }
  /* getRowsNative (one call, CSR) == getRowNative (reference glue, one up-call per pair) == tally */
  arr_t *rows = new_arr(ROWS, 4), *offs = new_arr(ROWS + 1, 8);
  for (int r = 0; r < ROWS; r++) ((jint*)rows->data)[r] = r * 7 + 1;
  arr_t* pairs = _JM(getRowsNative)(env, &m, rows, offs);
  CHECK(pairs && !g_thrown, "getRowsNative");
  const jlong* o = offs->data;
  for (int r = 0; r < ROWS; r++) {
    jint want = 0;
    for (int c = 1; c <= COLS; c++) want += tally[r][c] != 0 || 0;
    map_t map = {0, 0, NULL};
    _JM(getRowNative)(env, &m, r * 7 + 1, &map, 0);
    jint live = 0;                                           /* columns ever written stay live, also with value 0 (Q2) */
    for (jlong p = o[r]; p < o[r + 1]; p++) {
      jint c = ((jint*)pairs->data)[2 * p], v = ((jint*)pairs->data)[2 * p + 1];
      CHECK(c >= 1 && c <= COLS && v == tally[r][c], "row %d col %d: %d != %d", r, c, v, tally[r][c]);
      live++;
    }
    CHECK(live >= want && live == map.n && live == _JM(getRowLength)(env, &m, r * 7 + 1), "row %d: %d pairs, map %d", r, live, map.n);
    long long s1 = 0, s2 = 0;
    for (jint k = 0; k < map.n; k++) s1 += (long long)map.kv[2 * k] * 1000003 + map.kv[2 * k + 1];
    for (jlong p = o[r]; p < o[r + 1]; p++) s2 += (long long)((jint*)pairs->data)[2 * p] * 1000003 + ((jint*)pairs->data)[2 * p + 1];
    CHECK(s1 == s2, "row %d: getRowNative and getRowsNative disagree", r);
    free(map.kv);
  }
  /* setBatch: last writer wins */
  arr_t *sx = new_arr(3, 4), *sy = new_arr(3, 4), *sv = new_arr(3, 4);
  for (int i = 0; i < 3; i++) { ((jint*)sx->data)[i] = 8; ((jint*)sy->data)[i] = 2; ((jint*)sv->data)[i] = 100 + i; }
  _JM(setBatch)(env, &m, sx, sy, sv);
  CHECK(_JM(get)(env, &m, 8, 2) == 102, "setBatch: last writer wins");
  _JM(close)(env, &m);
  CHECK(m.ptr == 0, "close clears ptr");
  _JM(close)(env, &m);                                       /* closed object: the glue throws, like the reference */
  CHECK(g_thrown == 1, "closed object must throw");
  printf("fake_jvm: OK\n");
  return 0;
}
