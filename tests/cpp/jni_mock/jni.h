/*
 * Minimal stand-in for <jni.h>: just the types and the JNIEnv function-table entries jni/subrecon_jni.c uses, with the
 * signatures of the JNI specification. This image has no JDK; the mock lets the test-suite at least COMPILE the shim
 * against include/subrecon_b200.h (argument counts and types of every sr_* call) — it proves nothing about a real JVM.
 */
#ifndef SUBRECON_JNI_MOCK_H
#define SUBRECON_JNI_MOCK_H
#include <stdint.h>

typedef int32_t jint;
typedef int64_t jlong;
typedef double jdouble;
typedef uint8_t jboolean;
typedef jint jsize;
typedef void *jobject;
typedef jobject jclass;
typedef jobject jstring;
typedef jobject jarray;
typedef jarray jdoubleArray;
typedef jarray jintArray;

#define JNIEXPORT __attribute__((visibility("default")))
#define JNICALL
#define JNI_ABORT 2

struct JNINativeInterface_;
typedef const struct JNINativeInterface_ *JNIEnv;
struct JNINativeInterface_ {
    jsize (*GetArrayLength)(JNIEnv *, jarray);
    void *(*GetDirectBufferAddress)(JNIEnv *, jobject);
    jlong (*GetDirectBufferCapacity)(JNIEnv *, jobject);
    jdouble *(*GetDoubleArrayElements)(JNIEnv *, jdoubleArray, jboolean *);
    jint *(*GetIntArrayElements)(JNIEnv *, jintArray, jboolean *);
    const char *(*GetStringUTFChars)(JNIEnv *, jstring, jboolean *);
    jobject (*NewDirectByteBuffer)(JNIEnv *, void *, jlong);
    jstring (*NewStringUTF)(JNIEnv *, const char *);
    void (*ReleaseDoubleArrayElements)(JNIEnv *, jdoubleArray, jdouble *, jint);
    void (*ReleaseIntArrayElements)(JNIEnv *, jintArray, jint *, jint);
    void (*ReleaseStringUTFChars)(JNIEnv *, jstring, const char *);
};
#endif
