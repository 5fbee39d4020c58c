/*
 * jni/subrecon_jni.c — JNI shim between SubRecon's Java host (java/subrecon/recon/NativeRecon.java) and
 * libsubrecon_b200.so. Every function forwards to one entry of include/subrecon_b200.h after checking what a C callee
 * cannot trust from a JVM caller: NULL references, array lengths, direct-buffer capacities. A failed check returns
 * SR_ERR_ARG (NativeRecon.ck turns any non-zero status into one RuntimeException); nothing here can crash the JVM on a
 * short array.
 *
 * NOT BUILT AGAINST A REAL JDK IN THIS REPOSITORY'S ENVIRONMENT (the image has none). It is compiled with -Wall -Wextra
 * -Werror against tests/cpp/jni_mock/jni.h (tests/test_abi.py), which checks every sr_* call against the header, and
 * tests/cpp/host_mirror.cpp makes the same sr_* calls in the same order against the real library on a GPU.
 *   gcc -shared -fPIC -I$JAVA_HOME/include -I$JAVA_HOME/include/linux -I../include -o libsubrecon_jni.so subrecon_jni.c \
 *       -L../subrecon_b200 -lsubrecon_b200
 */
#include <jni.h>
#include <stdint.h>

#include "subrecon_b200.h"

#define MULTI(h) ((sr_multi *)(intptr_t)(h))

JNIEXPORT jlong JNICALL Java_subrecon_recon_NativeRecon_multiCreate(JNIEnv *env, jclass cls, jint nDevices) {
    sr_multi *m = NULL;
    return sr_multi_create(&m, nDevices, NULL) == SR_OK ? (jlong)(intptr_t)m : 0;
}

JNIEXPORT void JNICALL Java_subrecon_recon_NativeRecon_multiDestroy(JNIEnv *env, jclass cls, jlong h) { sr_multi_destroy(MULTI(h)); }

JNIEXPORT jstring JNICALL Java_subrecon_recon_NativeRecon_multiLastError(JNIEnv *env, jclass cls, jlong h) {
    return (*env)->NewStringUTF(env, sr_multi_last_error(MULTI(h)));
}

JNIEXPORT jint JNICALL Java_subrecon_recon_NativeRecon_multiCount(JNIEnv *env, jclass cls, jlong h) { return sr_multi_count(MULTI(h)); }

/* a double[] of exactly `want` elements (want < 0: any length), or NULL */
static jdouble *get_doubles(JNIEnv *env, jdoubleArray a, jsize want) {
    if (!a || (want >= 0 && (*env)->GetArrayLength(env, a) != want)) return NULL;
    return (*env)->GetDoubleArrayElements(env, a, NULL);
}
static jint *get_ints(JNIEnv *env, jintArray a, jsize want) {
    if (!a || (want >= 0 && (*env)->GetArrayLength(env, a) != want)) return NULL;
    return (*env)->GetIntArrayElements(env, a, NULL);
}
static void put_doubles(JNIEnv *env, jdoubleArray a, jdouble *p) { if (p) (*env)->ReleaseDoubleArrayElements(env, a, p, JNI_ABORT); }
static void put_ints(JNIEnv *env, jintArray a, jint *p) { if (p) (*env)->ReleaseIntArrayElements(env, a, p, JNI_ABORT); }

JNIEXPORT jint JNICALL Java_subrecon_recon_NativeRecon_multiSetModel(JNIEnv *env, jclass cls, jlong h, jdoubleArray eval,
                                                                     jdoubleArray evec, jdoubleArray ievc, jdoubleArray pi) {
    jdouble *a = get_doubles(env, eval, 20), *b = get_doubles(env, evec, 400), *c = get_doubles(env, ievc, 400), *d = get_doubles(env, pi, 20);
    int rc = (a && b && c && d) ? sr_multi_set_model(MULTI(h), a, b, c, d) : SR_ERR_ARG;   /* sr_set_* copy their arguments */
    put_doubles(env, eval, a); put_doubles(env, evec, b); put_doubles(env, ievc, c); put_doubles(env, pi, d);
    return rc;
}

JNIEXPORT jint JNICALL Java_subrecon_recon_NativeRecon_multiSetRates(JNIEnv *env, jclass cls, jlong h, jdoubleArray rates) {
    jdouble *r = get_doubles(env, rates, -1);
    int rc = r ? sr_multi_set_rates(MULTI(h), (int)(*env)->GetArrayLength(env, rates), r) : SR_ERR_ARG;
    put_doubles(env, rates, r);
    return rc;
}

JNIEXPORT jint JNICALL Java_subrecon_recon_NativeRecon_multiSetTree(JNIEnv *env, jclass cls, jlong h, jintArray childOff,
                                                                    jintArray childIdx, jdoubleArray blen, jintArray leafRow, jint root) {
    if (!blen) return SR_ERR_ARG;
    const jsize n = (*env)->GetArrayLength(env, blen);                       /* nodes */
    if (n < 3) return SR_ERR_ARG;
    jint *off = get_ints(env, childOff, n + 1), *idx = get_ints(env, childIdx, n - 1), *lr = get_ints(env, leafRow, n);
    jdouble *bl = get_doubles(env, blen, n);
    int rc = SR_ERR_ARG;
    if (off && idx && lr && bl && off[n] == n - 1)                           /* the library validates the rest (sr_set_tree) */
        rc = sr_multi_set_tree(MULTI(h), (int32_t)n, (const int32_t *)off, (const int32_t *)idx, bl, (const int32_t *)lr, root);
    put_ints(env, childOff, off); put_ints(env, childIdx, idx); put_ints(env, leafRow, lr); put_doubles(env, blen, bl);
    return rc;
}

/* sr_set_option on every context: "input_chars", "compact_cap", "strict_root", "compress_patterns", ... */
JNIEXPORT jint JNICALL Java_subrecon_recon_NativeRecon_multiSetOption(JNIEnv *env, jclass cls, jlong h, jstring name, jlong value) {
    if (!name) return SR_ERR_ARG;
    const char *n = (*env)->GetStringUTFChars(env, name, NULL);
    if (!n) return SR_ERR_OOM;
    const int rc = sr_multi_set_option(MULTI(h), n, (int64_t)value);
    (*env)->ReleaseStringUTFChars(env, name, n);
    return rc;
}

/* address of a direct buffer that holds at least `need` bytes, or NULL */
static void *direct(JNIEnv *env, jobject buf, int64_t need) {
    if (!buf || need < 0) return NULL;
    void *p = (*env)->GetDirectBufferAddress(env, buf);
    if (!p || (*env)->GetDirectBufferCapacity(env, buf) < (jlong)need) return NULL;
    return p;
}

/* states / lnl / joint / compact are direct ByteBuffers over sr_host_alloc memory (page-locked: DMA without a bounce copy).
 * joint and compact may each be null; compactCap is the capacity the caller set with "compact_cap" (it sizes the buffer check). */
JNIEXPORT jint JNICALL Java_subrecon_recon_NativeRecon_multiRunStreamed(JNIEnv *env, jclass cls, jlong h, jint nTaxa, jlong nSites,
                                                                        jobject states, jlong pitch, jlong site0, jlong n,
                                                                        jobject lnl, jobject joint, jobject compact,
                                                                        jint compactCap, jdouble threshold) {
    if (nTaxa < 1 || nSites < 0 || pitch < nSites || site0 < 0 || n < 0 || site0 + n > nSites) return SR_ERR_ARG;
    const uint8_t *st = (const uint8_t *)direct(env, states, (int64_t)pitch * (nTaxa - 1) + nSites);
    double *l = (double *)direct(env, lnl, 8 * (int64_t)n);
    double *j = NULL;
    sr_compact *c = NULL;
    if (!st || !l) return SR_ERR_ARG;
    if (joint && !(j = (double *)direct(env, joint, 3200 * (int64_t)n))) return SR_ERR_ARG;
    if (compact) {
        const size_t stride = sr_compact_stride(compactCap);
        if (!stride || !(c = (sr_compact *)direct(env, compact, (int64_t)stride * (int64_t)n))) return SR_ERR_ARG;
    }
    return sr_multi_run_streamed(MULTI(h), nTaxa, nSites, st, pitch, site0, n, l, j, c, threshold);
}

JNIEXPORT jint JNICALL Java_subrecon_recon_NativeRecon_compactCapFor(JNIEnv *env, jclass cls, jdouble threshold) {
    return sr_compact_cap_for(threshold);
}

JNIEXPORT jlong JNICALL Java_subrecon_recon_NativeRecon_compactStride(JNIEnv *env, jclass cls, jint cap) {
    return (jlong)sr_compact_stride(cap);
}

JNIEXPORT jobject JNICALL Java_subrecon_recon_NativeRecon_hostAlloc(JNIEnv *env, jclass cls, jlong bytes) {
    if (bytes < 0) return NULL;
    void *p = sr_host_alloc((size_t)bytes);
    if (!p) return NULL;
    jobject buf = (*env)->NewDirectByteBuffer(env, p, bytes);
    if (!buf) sr_host_free(p);
    return buf;
}

JNIEXPORT void JNICALL Java_subrecon_recon_NativeRecon_hostFree(JNIEnv *env, jclass cls, jobject buf) {
    if (buf) sr_host_free((*env)->GetDirectBufferAddress(env, buf));
}
