/*
 * tests/jni_stub/jni.h — COMPILE-CHECK STUB, not the JDK header and NOT ABI-compatible with a JVM.
 * The image has no JDK, so falcon-genome_b200/jni/fcs_pairhmm_jni.c cannot be built for real here; this
 * stub declares just the JNI names the shim uses so that tests/test_jni_shim.py can at least type-check
 * it (gcc -fsyntax-only).  A real build uses $JAVA_HOME/include/jni.h.
 */
#ifndef FCS_TEST_JNI_STUB_H
#define FCS_TEST_JNI_STUB_H
#include <stdint.h>
typedef int32_t jint;
typedef int32_t jsize;
typedef int8_t jbyte;
typedef uint8_t jboolean;
typedef double jdouble;
typedef struct _jobject* jobject;
typedef jobject jclass;
typedef jobject jarray;
typedef jarray jobjectArray;
typedef jarray jbyteArray;
typedef jarray jdoubleArray;
typedef struct _jfieldID* jfieldID;
#define JNIEXPORT __attribute__((visibility("default")))
#define JNICALL
#define JNI_ABORT 2
struct JNINativeInterface_;
typedef const struct JNINativeInterface_* JNIEnv;
struct JNINativeInterface_ {
  jfieldID (*GetFieldID)(JNIEnv*, jclass, const char*, const char*);
  jobject (*GetObjectField)(JNIEnv*, jobject, jfieldID);
  jsize (*GetArrayLength)(JNIEnv*, jarray);
  jobject (*GetObjectArrayElement)(JNIEnv*, jobjectArray, jsize);
  jbyte* (*GetByteArrayElements)(JNIEnv*, jbyteArray, jboolean*);
  void (*ReleaseByteArrayElements)(JNIEnv*, jbyteArray, jbyte*, jint);
  jdouble* (*GetDoubleArrayElements)(JNIEnv*, jdoubleArray, jboolean*);
  void (*ReleaseDoubleArrayElements)(JNIEnv*, jdoubleArray, jdouble*, jint);
  jclass (*FindClass)(JNIEnv*, const char*);
  jint (*ThrowNew)(JNIEnv*, jclass, const char*);
  jboolean (*ExceptionCheck)(JNIEnv*);
  void (*DeleteLocalRef)(JNIEnv*, jobject);
  jint (*EnsureLocalCapacity)(JNIEnv*, jint);
  void (*GetByteArrayRegion)(JNIEnv*, jbyteArray, jsize, jsize, jbyte*);
  void (*SetDoubleArrayRegion)(JNIEnv*, jdoubleArray, jsize, jsize, const jdouble*);
};
#endif
