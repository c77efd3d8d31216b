/* tests/stubs/jni.h — TEST SCAFFOLDING: the slice of the JNI C interface that the reference's
 * src/smatrix_jni.c uses, enough to COMPILE AND LINK that file unchanged against include/smatrix.h
 * and our library (no JDK in this image).  Not a working JVM interface. */
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
};
#endif
