/*
 * filmyou_rm2_jni.c -- the JNI stub between filmyou-core's Java classes and libfilmyou_rm2.so.
 *
 * Every native method unwraps direct NIO buffers and forwards to ONE C-ABI call of include/filmyou_rm2.h /
 * include/filmyou_nmf.h; no logic lives here.  Java side: integration/java/es/udc/fi/dc/irlab/rm/RM2Native.java and
 * integration/java/es/udc/fi/dc/irlab/nmf/ppc/NmfNative.java.  Build (on a machine with a JDK):
 *     gcc -shared -fPIC -I$JAVA_HOME/include -I$JAVA_HOME/include/linux -Iinclude \
 *         integration/jni/filmyou_rm2_jni.c -Lfilmyou_core_b200 -lfilmyou_rm2 -o libfilmyou_rm2_jni.so
 * No JDK exists in the build image: tests/test_jni_stub.py compiles this file against tests/mock_jni/jni.h (the few
 * JNI declarations used here) and drives it through a fake JNIEnv on the GPU box.
 */
#include <jni.h>
#include <stddef.h>
#include <stdint.h>

#include "filmyou_nmf.h"
#include "filmyou_rm2.h"

#include <stdio.h>

#define BUF(b) ((b) ? (*env)->GetDirectBufferAddress(env, (b)) : NULL)
/* A sizing mistake on the Java side must not become an out-of-bounds access inside the JVM: every buffer's capacity
 * (in elements of its own type, as GetDirectBufferCapacity reports it for a view buffer) is checked against the count
 * the call will touch; a non-direct buffer (address NULL, capacity -1) fails the same way. */
#define NEED(b, n) do { if ((b) && ((*env)->GetDirectBufferAddress(env, (b)) == NULL || (*env)->GetDirectBufferCapacity(env, (b)) < (jlong)(n))) return FY_E_ARG; } while (0)
#define NEED_NN(b, n) do { if (!(b)) return FY_E_ARG; NEED(b, n); } while (0)

static void throw_runtime(JNIEnv* env, const char* what, int code) {
    char msg[160];
    snprintf(msg, sizeof(msg), "%s failed: fy_status %d (no usable B200, or bad parameters)", what, code);
    jclass ex = (*env)->FindClass(env, "java/lang/RuntimeException");       /* RM2Job.java:265-268: a failed job throws */
    if (ex) (*env)->ThrowNew(env, ex, msg);
}
#define RM2(h) ((fy_rm2_ctx*)(intptr_t)(h))
#define NMF(h) ((fy_nmf_ctx*)(intptr_t)(h))

/* ---- es.udc.fi.dc.irlab.rm.RM2Native ---- */
JNIEXPORT jlong JNICALL Java_es_udc_fi_dc_irlab_rm_RM2Native_create(JNIEnv* env, jclass cls, jdouble lambda, jint numberOfItems,
                                                                    jint numberOfRecommendations, jint filterUsers, jint device,
                                                                    jint shardRank, jint shardCount, jint nGpus) {
    (void)cls;
    fy_rm2_params p;
    fy_rm2_default_params(&p);
    p.lambda = lambda; p.number_of_items = numberOfItems; p.top_n = numberOfRecommendations; p.filter_users = filterUsers;
    p.device = device; p.shard_rank = shardRank; p.shard_count = shardCount; p.n_gpus = nGpus;
    fy_rm2_ctx* ctx = NULL;
    const int rc = fy_rm2_create(&ctx, &p);
    if (rc != FY_OK) { throw_runtime(env, "fy_rm2_create", rc); return 0; }
    return (jlong)(intptr_t)ctx;
}

JNIEXPORT jint JNICALL Java_es_udc_fi_dc_irlab_rm_RM2Native_setRatings(JNIEnv* env, jclass cls, jlong h, jobject user, jobject item,
                                                                       jobject score, jlong nnz) {
    (void)cls;
    if (nnz < 0) return FY_E_ARG;
    NEED_NN(user, nnz); NEED_NN(item, nnz); NEED_NN(score, nnz);
    return fy_rm2_set_ratings(RM2(h), (const int32_t*)BUF(user), (const int32_t*)BUF(item), (const float*)BUF(score), nnz);
}

JNIEXPORT jint JNICALL Java_es_udc_fi_dc_irlab_rm_RM2Native_setClustering(JNIEnv* env, jclass cls, jlong h, jobject user,
                                                                          jobject cluster, jlong nUsers, jobject clusterSize,
                                                                          jint nClusters) {
    (void)cls;
    if (nUsers < 0 || nClusters < 0) return FY_E_ARG;
    NEED_NN(user, nUsers); NEED_NN(cluster, nUsers); NEED_NN(clusterSize, nClusters);
    return fy_rm2_set_clustering(RM2(h), (const int32_t*)BUF(user), (const int32_t*)BUF(cluster), nUsers,
                                 (const int32_t*)BUF(clusterSize), nClusters);
}

JNIEXPORT jint JNICALL Java_es_udc_fi_dc_irlab_rm_RM2Native_run(JNIEnv* env, jclass cls, jlong h) {
    (void)env; (void)cls;
    return fy_rm2_run(RM2(h));
}

JNIEXPORT jint JNICALL Java_es_udc_fi_dc_irlab_rm_RM2Native_scoreGroup(JNIEnv* env, jclass cls, jlong h, jint cluster, jint split,
                                                                       jint nSplits, jobject groupUser, jobject groupUserSum,
                                                                       jint nGroupUsers, jobject rUser, jobject rItem, jobject rScore,
                                                                       jlong nnz, jobject itemProb, jint maxItem) {
    (void)cls;
    if (nGroupUsers < 0 || nnz < 0 || maxItem < 0) return FY_E_ARG;
    NEED_NN(groupUser, nGroupUsers); NEED_NN(groupUserSum, nGroupUsers);
    NEED_NN(rUser, nnz); NEED_NN(rItem, nnz); NEED_NN(rScore, nnz); NEED_NN(itemProb, (jlong)maxItem + 1);
    return fy_rm2_score_group(RM2(h), cluster, split, nSplits, (const int32_t*)BUF(groupUser), (const double*)BUF(groupUserSum),
                              nGroupUsers, (const int32_t*)BUF(rUser), (const int32_t*)BUF(rItem), (const float*)BUF(rScore), nnz,
                              (const double*)BUF(itemProb), maxItem);
}

JNIEXPORT jlong JNICALL Java_es_udc_fi_dc_irlab_rm_RM2Native_resultCount(JNIEnv* env, jclass cls, jlong h) {
    (void)env; (void)cls;
    return fy_rm2_result_count(RM2(h));
}

JNIEXPORT jint JNICALL Java_es_udc_fi_dc_irlab_rm_RM2Native_results(JNIEnv* env, jclass cls, jlong h, jobject user, jobject item,
                                                                    jobject score64, jobject score32, jobject cluster) {
    (void)cls;
    const jlong n = fy_rm2_result_count(RM2(h));
    if (n < 0) return FY_E_STATE;
    NEED(user, n); NEED(item, n); NEED(score64, n); NEED(score32, n); NEED(cluster, n);
    return fy_rm2_results(RM2(h), (int32_t*)BUF(user), (int32_t*)BUF(item), (double*)BUF(score64), (float*)BUF(score32),
                          (int32_t*)BUF(cluster));
}

JNIEXPORT jint JNICALL Java_es_udc_fi_dc_irlab_rm_RM2Native_stats(JNIEnv* env, jclass cls, jlong h, jobject userSum, jobject itemProb,
                                                                  jobject total) {
    (void)cls;
    const jlong nu = fy_rm2_user_count(RM2(h));
    if (nu < 0) return FY_E_STATE;
    NEED(userSum, nu); NEED(itemProb, (jlong)fy_rm2_max_item(RM2(h)) + 1); NEED(total, 1);
    return fy_rm2_stats(RM2(h), (double*)BUF(userSum), (double*)BUF(itemProb), (double*)BUF(total));
}

JNIEXPORT jlong JNICALL Java_es_udc_fi_dc_irlab_rm_RM2Native_resultRowCount(JNIEnv* env, jclass cls, jlong h) {
    (void)env; (void)cls;
    return fy_rm2_result_row_count(RM2(h));
}

JNIEXPORT jint JNICALL Java_es_udc_fi_dc_irlab_rm_RM2Native_resultRows(JNIEnv* env, jclass cls, jlong h, jobject user, jobject cluster,
                                                                       jobject count) {
    (void)cls;
    const jlong n = fy_rm2_result_row_count(RM2(h));
    if (n < 0) return FY_E_STATE;
    NEED(user, n); NEED(cluster, n); NEED(count, n);
    return fy_rm2_result_rows(RM2(h), (int32_t*)BUF(user), (int32_t*)BUF(cluster), (int32_t*)BUF(count));
}

JNIEXPORT jint JNICALL Java_es_udc_fi_dc_irlab_rm_RM2Native_maxItem(JNIEnv* env, jclass cls, jlong h) {
    (void)env; (void)cls;
    return fy_rm2_max_item(RM2(h));
}

JNIEXPORT jstring JNICALL Java_es_udc_fi_dc_irlab_rm_RM2Native_lastError(JNIEnv* env, jclass cls, jlong h) {
    (void)cls;
    return (*env)->NewStringUTF(env, fy_rm2_last_error(RM2(h)));
}

JNIEXPORT void JNICALL Java_es_udc_fi_dc_irlab_rm_RM2Native_destroy(JNIEnv* env, jclass cls, jlong h) {
    (void)env; (void)cls;
    fy_rm2_destroy(RM2(h));
}

/* ---- es.udc.fi.dc.irlab.nmf.ppc.NmfNative ---- */
JNIEXPORT jlong JNICALL Java_es_udc_fi_dc_irlab_nmf_ppc_NmfNative_create(JNIEnv* env, jclass cls, jint mode, jint numberOfUsers,
                                                                         jint numberOfItems, jint numberOfClusters,
                                                                         jint numberOfIterations, jint normalizationFrequency,
                                                                         jint idBase, jint device) {
    (void)cls;
    fy_nmf_params p;
    fy_nmf_default_params(&p);
    p.mode = mode; p.number_of_users = numberOfUsers; p.number_of_items = numberOfItems; p.number_of_clusters = numberOfClusters;
    p.number_of_iterations = numberOfIterations; p.normalization_frequency = normalizationFrequency; p.id_base = idBase;
    p.device = device;
    fy_nmf_ctx* ctx = NULL;
    const int rc = fy_nmf_create(&ctx, &p);
    if (rc != FY_OK) { throw_runtime(env, "fy_nmf_create", rc); return 0; }
    return (jlong)(intptr_t)ctx;
}

JNIEXPORT jint JNICALL Java_es_udc_fi_dc_irlab_nmf_ppc_NmfNative_setRatings(JNIEnv* env, jclass cls, jlong h, jobject user, jobject item,
                                                                            jobject score, jlong nnz) {
    (void)cls;
    if (nnz < 0) return FY_E_ARG;
    NEED_NN(user, nnz); NEED_NN(item, nnz); NEED_NN(score, nnz);
    return fy_nmf_set_ratings(NMF(h), (const int32_t*)BUF(user), (const int32_t*)BUF(item), (const float*)BUF(score), nnz);
}

JNIEXPORT jint JNICALL Java_es_udc_fi_dc_irlab_nmf_ppc_NmfNative_setFactors(JNIEnv* env, jclass cls, jlong h, jobject H, jobject W) {
    (void)cls;
    return fy_nmf_set_factors(NMF(h), (const double*)BUF(H), (const double*)BUF(W));
}

JNIEXPORT jint JNICALL Java_es_udc_fi_dc_irlab_nmf_ppc_NmfNative_initRandom(JNIEnv* env, jclass cls, jlong h, jlong seed) {
    (void)env; (void)cls;
    return fy_nmf_init_random(NMF(h), (uint64_t)seed);
}

JNIEXPORT jint JNICALL Java_es_udc_fi_dc_irlab_nmf_ppc_NmfNative_run(JNIEnv* env, jclass cls, jlong h) {
    (void)env; (void)cls;
    return fy_nmf_run(NMF(h));
}

JNIEXPORT jint JNICALL Java_es_udc_fi_dc_irlab_nmf_ppc_NmfNative_getFactors(JNIEnv* env, jclass cls, jlong h, jobject H, jobject W) {
    (void)cls;
    return fy_nmf_get_factors(NMF(h), (double*)BUF(H), (double*)BUF(W));
}

JNIEXPORT jint JNICALL Java_es_udc_fi_dc_irlab_nmf_ppc_NmfNative_clusterAssignment(JNIEnv* env, jclass cls, jlong h, jobject cluster,
                                                                                   jobject clusterSize) {
    (void)cls;
    return fy_nmf_cluster_assignment(NMF(h), (int32_t*)BUF(cluster), (int32_t*)BUF(clusterSize));
}

JNIEXPORT jstring JNICALL Java_es_udc_fi_dc_irlab_nmf_ppc_NmfNative_lastError(JNIEnv* env, jclass cls, jlong h) {
    (void)cls;
    return (*env)->NewStringUTF(env, fy_nmf_last_error(NMF(h)));
}

JNIEXPORT void JNICALL Java_es_udc_fi_dc_irlab_nmf_ppc_NmfNative_destroy(JNIEnv* env, jclass cls, jlong h) {
    (void)env; (void)cls;
    fy_nmf_destroy(NMF(h));
}
