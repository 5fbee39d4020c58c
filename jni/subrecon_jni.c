/*
 * jni/subrecon_jni.c — JNI shim between SubRecon's Java host and libsubrecon_b200.so.
 *
 * NOT COMPILED IN THIS REPOSITORY'S ENVIRONMENT: the build image has no JDK (no jni.h). It is the ~100 lines a
 * SubRecon maintainer adds next to src/subrecon/recon/NativeRecon.java; every function forwards
 * to one entry of include/subrecon_b200.h, which is the boundary that IS exercised here (ctypes + C++ tests).
 *   gcc -shared -fPIC -I$JAVA_HOME/include -I$JAVA_HOME/include/linux -I../include -o libsubrecon_jni.so subrecon_jni.c \
 *       -L../subrecon_b200 -lsubrecon_b200
 */
#include <jni.h>
#include <stdint.h>

#include "subrecon_b200.h"

#define CTX(h) ((sr_ctx *)(intptr_t)(h))

JNIEXPORT jlong JNICALL Java_subrecon_recon_NativeRecon_create(JNIEnv *env, jclass cls, jint device) {
    sr_ctx *ctx = NULL;
    return sr_create(&ctx, device) == SR_OK ? (jlong)(intptr_t)ctx : 0;
}

JNIEXPORT void JNICALL Java_subrecon_recon_NativeRecon_destroy(JNIEnv *env, jclass cls, jlong ctx) { sr_destroy(CTX(ctx)); }

JNIEXPORT jstring JNICALL Java_subrecon_recon_NativeRecon_lastError(JNIEnv *env, jclass cls, jlong ctx) {
    return (*env)->NewStringUTF(env, sr_last_error(CTX(ctx)));
}

JNIEXPORT jint JNICALL Java_subrecon_recon_NativeRecon_setModel(JNIEnv *env, jclass cls, jlong ctx, jdoubleArray eval,
                                                                jdoubleArray evec, jdoubleArray ievc, jdoubleArray pi) {
    jdouble *a = (*env)->GetDoubleArrayElements(env, eval, NULL), *b = (*env)->GetDoubleArrayElements(env, evec, NULL);
    jdouble *c = (*env)->GetDoubleArrayElements(env, ievc, NULL), *d = (*env)->GetDoubleArrayElements(env, pi, NULL);
    int rc = sr_set_model(CTX(ctx), a, b, c, d);          /* sr_set_* copy their arguments */
    (*env)->ReleaseDoubleArrayElements(env, eval, a, JNI_ABORT); (*env)->ReleaseDoubleArrayElements(env, evec, b, JNI_ABORT);
    (*env)->ReleaseDoubleArrayElements(env, ievc, c, JNI_ABORT); (*env)->ReleaseDoubleArrayElements(env, pi, d, JNI_ABORT);
    return rc;
}

JNIEXPORT jint JNICALL Java_subrecon_recon_NativeRecon_setRates(JNIEnv *env, jclass cls, jlong ctx, jdoubleArray rates) {
    jsize k = (*env)->GetArrayLength(env, rates);
    jdouble *r = (*env)->GetDoubleArrayElements(env, rates, NULL);
    int rc = sr_set_rates(CTX(ctx), (int)k, r);
    (*env)->ReleaseDoubleArrayElements(env, rates, r, JNI_ABORT);
    return rc;
}

JNIEXPORT jint JNICALL Java_subrecon_recon_NativeRecon_setTree(JNIEnv *env, jclass cls, jlong ctx, jintArray childOff,
                                                               jintArray childIdx, jdoubleArray blen, jintArray leafRow, jint root) {
    jsize n = (*env)->GetArrayLength(env, blen);
    jint *off = (*env)->GetIntArrayElements(env, childOff, NULL), *idx = (*env)->GetIntArrayElements(env, childIdx, NULL);
    jint *lr = (*env)->GetIntArrayElements(env, leafRow, NULL);
    jdouble *bl = (*env)->GetDoubleArrayElements(env, blen, NULL);
    int rc = sr_set_tree(CTX(ctx), (int32_t)n, (const int32_t *)off, (const int32_t *)idx, bl, (const int32_t *)lr, root);
    (*env)->ReleaseIntArrayElements(env, childOff, off, JNI_ABORT); (*env)->ReleaseIntArrayElements(env, childIdx, idx, JNI_ABORT);
    (*env)->ReleaseIntArrayElements(env, leafRow, lr, JNI_ABORT); (*env)->ReleaseDoubleArrayElements(env, blen, bl, JNI_ABORT);
    return rc;
}

/* states / lnl / joint are direct ByteBuffers over sr_host_alloc memory (page-locked: DMA without a bounce copy) */
JNIEXPORT jint JNICALL Java_subrecon_recon_NativeRecon_runStreamed(JNIEnv *env, jclass cls, jlong ctx, jint nTaxa, jlong nSites,
                                                                   jobject states, jlong pitch, jlong site0, jlong n,
                                                                   jobject lnl, jobject joint) {
    const uint8_t *st = (const uint8_t *)(*env)->GetDirectBufferAddress(env, states);
    double *l = (double *)(*env)->GetDirectBufferAddress(env, lnl);
    double *j = joint ? (double *)(*env)->GetDirectBufferAddress(env, joint) : NULL;
    if (!st || !l) return SR_ERR_ARG;
    return sr_run_streamed(CTX(ctx), nTaxa, nSites, st, pitch, site0, n, l, j, NULL, 0.5);
}

/* sr_set_option: "input_chars", "cherry_tables", "compress_patterns", ... (include/subrecon_b200.h) */
JNIEXPORT jint JNICALL Java_subrecon_recon_NativeRecon_setOption(JNIEnv *env, jclass cls, jlong ctx, jstring name, jlong value) {
    const char *n = (*env)->GetStringUTFChars(env, name, NULL);
    if (!n) return SR_ERR_OOM;
    const int rc = sr_set_option(CTX(ctx), n, (int64_t)value);
    (*env)->ReleaseStringUTFChars(env, name, n);
    return rc;
}

JNIEXPORT jobject JNICALL Java_subrecon_recon_NativeRecon_hostAlloc(JNIEnv *env, jclass cls, jlong bytes) {
    void *p = sr_host_alloc((size_t)bytes);
    return p ? (*env)->NewDirectByteBuffer(env, p, bytes) : NULL;
}

JNIEXPORT void JNICALL Java_subrecon_recon_NativeRecon_hostFree(JNIEnv *env, jclass cls, jobject buf) {
    sr_host_free((*env)->GetDirectBufferAddress(env, buf));
}
