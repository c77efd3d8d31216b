/*
 * examples/jni/smatrix_jni_batch.c — the batched natives a maintainer adds NEXT TO the reference's
 * src/smatrix_jni.c (which keeps working unchanged) so that the Java binding reaches the bulk C-ABI
 * of include/smatrix_batch.h:
 *
 *   public native void  incrBatch(int[] xs, int[] ys, int[] vals);      // vals == null: all ones
 *   public native void  setBatch(int[] xs, int[] ys, int[] vals);
 *   public native int[] getBatch(int[] xs, int[] ys);
 *   public native int[] getRowsNative(int[] xs, long[] offsets);        // -> [col, val, col, val, ...]
 *
 * getRowsNative replaces the reference's per-row getRowNative for bulk reads: that one does rowlen +
 * malloc + getrow and ONE JNI up-call (putIntTuple) per pair (src/smatrix_jni.c:114-149); this one
 * returns all rows of a batch as one CSR (offsets[n + 1] filled in place, pairs as the result).
 *
 * Compiled and EXECUTED by tests/test_bindings_link.py against tests/stubs/jni.h + a toy JNIEnv
 * (no JDK in the image); with a real JDK: gcc -shared -fPIC -I include -I $JAVA_HOME/include ...
 */
#include <stdint.h>
#include <stdlib.h>

#include <jni.h>

#include "smatrix.h"
#include "smatrix_batch.h"

#define _JM(X) Java_com_paulasmuth_libsmatrix_SparseMatrix_##X

/* the handle lives in the Java object's `ptr` field, exactly like src/smatrix_jni.c:21-49 */
static int batch_get_ptr(JNIEnv* env, jobject self, void** ptr) {
  jclass cls = (*env)->FindClass(env, "com/paulasmuth/libsmatrix/SparseMatrix");
  jfieldID fid = (*env)->GetFieldID(env, cls, "ptr", "J");
  jlong p = (*env)->GetLongField(env, self, fid);
  if (p > 0) {
    *ptr = (void*)p;
    return 0;
  }
  (*env)->ThrowNew(env, (*env)->FindClass(env, "java/lang/IllegalArgumentException"),
                   "can't find native object. maybe close() was already called");
  return 1;
}

static void write_batch(JNIEnv* env, jobject self, jintArray xs_, jintArray ys_, jintArray vs_, int is_set) {
  void* ptr = NULL;
  if (batch_get_ptr(env, self, &ptr)) return;
  const jsize n = (*env)->GetArrayLength(env, xs_);
  /* no copy: the library streams straight from the (pinned-by-the-JVM) host arrays */
  jint* xs = (*env)->GetPrimitiveArrayCritical(env, xs_, 0);
  jint* ys = (*env)->GetPrimitiveArrayCritical(env, ys_, 0);
  jint* vs = vs_ ? (*env)->GetPrimitiveArrayCritical(env, vs_, 0) : NULL; /* NULL = every value is 1 */
  if (is_set) smatrix_set_batch(ptr, (uint32_t*)xs, (uint32_t*)ys, (uint32_t*)vs, (size_t)n);
  else smatrix_incr_batch(ptr, (uint32_t*)xs, (uint32_t*)ys, (uint32_t*)vs, (size_t)n);
  if (vs) (*env)->ReleasePrimitiveArrayCritical(env, vs_, vs, JNI_ABORT);
  (*env)->ReleasePrimitiveArrayCritical(env, ys_, ys, JNI_ABORT);
  (*env)->ReleasePrimitiveArrayCritical(env, xs_, xs, JNI_ABORT);
}

JNIEXPORT void JNICALL _JM(incrBatch)(JNIEnv* env, jobject self, jintArray xs, jintArray ys, jintArray vals) {
  write_batch(env, self, xs, ys, vals, 0);
}
JNIEXPORT void JNICALL _JM(setBatch)(JNIEnv* env, jobject self, jintArray xs, jintArray ys, jintArray vals) {
  write_batch(env, self, xs, ys, vals, 1);
}

JNIEXPORT jintArray JNICALL _JM(getBatch)(JNIEnv* env, jobject self, jintArray xs_, jintArray ys_) {
  void* ptr = NULL;
  if (batch_get_ptr(env, self, &ptr)) return NULL;
  const jsize n = (*env)->GetArrayLength(env, xs_);
  jintArray out_ = (*env)->NewIntArray(env, n);
  jint* xs = (*env)->GetIntArrayElements(env, xs_, 0);
  jint* ys = (*env)->GetIntArrayElements(env, ys_, 0);
  jint* out = (*env)->GetIntArrayElements(env, out_, 0);
  smatrix_get_batch(ptr, (uint32_t*)xs, (uint32_t*)ys, (size_t)n, (uint32_t*)out);
  (*env)->ReleaseIntArrayElements(env, out_, out, 0);
  (*env)->ReleaseIntArrayElements(env, ys_, ys, JNI_ABORT);
  (*env)->ReleaseIntArrayElements(env, xs_, xs, JNI_ABORT);
  return out_;
}

JNIEXPORT jintArray JNICALL _JM(getRowsNative)(JNIEnv* env, jobject self, jintArray xs_, jlongArray offs_) {
  void* ptr = NULL;
  if (batch_get_ptr(env, self, &ptr)) return NULL;
  const jsize n = (*env)->GetArrayLength(env, xs_);
  if ((*env)->GetArrayLength(env, offs_) < n + 1) {
    (*env)->ThrowNew(env, (*env)->FindClass(env, "java/lang/IllegalArgumentException"), "offsets must hold n + 1 longs");
    return NULL;
  }
  jint* xs = (*env)->GetIntArrayElements(env, xs_, 0);
  jlong* offs = (*env)->GetLongArrayElements(env, offs_, 0);
  const uint64_t total = smatrix_getrow_batch(ptr, (uint32_t*)xs, (size_t)n, (uint64_t*)offs, NULL, 0); /* size query */
  jintArray out_ = (*env)->NewIntArray(env, (jsize)(2 * total));
  if (total) {
    jint* pairs = (*env)->GetIntArrayElements(env, out_, 0);
    smatrix_getrow_batch(ptr, (uint32_t*)xs, (size_t)n, (uint64_t*)offs, (uint32_t*)pairs, total);
    (*env)->ReleaseIntArrayElements(env, out_, pairs, 0);
  }
  (*env)->ReleaseLongArrayElements(env, offs_, offs, 0);
  (*env)->ReleaseIntArrayElements(env, xs_, xs, JNI_ABORT);
  return out_;
}
