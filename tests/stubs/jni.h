/* tests/stubs/jni.h — TEST SCAFFOLDING: the slice of the JNI C interface that the reference's
 * src/smatrix_jni.c and examples/jni/smatrix_jni_batch.c use, enough to COMPILE AND LINK them against
 * include/smatrix.h and our library (no JDK in this image), and — with
 * tests/stubs/fake_jvm.c, a toy implementation of this table — to EXECUTE them.  Not a JVM. */
#ifndef SMX_STUB_JNI_H
#define SMX_STUB_JNI_H
#include <stdint.h>
#define JNIEXPORT __attribute__((visibility("default")))
#define JNICALL
typedef int32_t jint;
typedef int64_t jlong;
typedef void* jobject;
typedef jobject jclass;
typedef jobject jstring;
typedef void* jfieldID;
typedef void* jmethodID;
typedef unsigned char jboolean;
typedef jint jsize;
typedef jobject jarray;
typedef jarray jintArray;
typedef jarray jlongArray;
#define JNI_ABORT 2
struct JNINativeInterface_;
typedef const struct JNINativeInterface_* JNIEnv;
struct JNINativeInterface_ {
  jclass (*FindClass)(JNIEnv*, const char*);
  jint (*ThrowNew)(JNIEnv*, jclass, const char*);
  jfieldID (*GetFieldID)(JNIEnv*, jclass, const char*, const char*);
  void (*SetLongField)(JNIEnv*, jobject, jfieldID, jlong);
  jlong (*GetLongField)(JNIEnv*, jobject, jfieldID);
  const char* (*GetStringUTFChars)(JNIEnv*, jstring, jboolean*);
  void (*ReleaseStringUTFChars)(JNIEnv*, jstring, const char*);
  jclass (*GetObjectClass)(JNIEnv*, jobject);
  jmethodID (*GetMethodID)(JNIEnv*, jclass, const char*, const char*);
  void (*CallVoidMethod)(JNIEnv*, jobject, jmethodID, ...);
  /* primitive arrays (examples/jni/smatrix_jni_batch.c) */
  jsize (*GetArrayLength)(JNIEnv*, jarray);
  void* (*GetPrimitiveArrayCritical)(JNIEnv*, jarray, jboolean*);
  void (*ReleasePrimitiveArrayCritical)(JNIEnv*, jarray, void*, jint);
  jintArray (*NewIntArray)(JNIEnv*, jsize);
  jint* (*GetIntArrayElements)(JNIEnv*, jintArray, jboolean*);
  void (*ReleaseIntArrayElements)(JNIEnv*, jintArray, jint*, jint);
  jlong* (*GetLongArrayElements)(JNIEnv*, jlongArray, jboolean*);
  void (*ReleaseLongArrayElements)(JNIEnv*, jlongArray, jlong*, jint);
};
#endif
