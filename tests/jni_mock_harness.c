/*
 * Drives integration/jni/filmyou_rm2_jni.c through a fake JNIEnv (tests/mock_jni/jni.h): a "direct buffer" is a
 * struct holding an address and a capacity, exactly what GetDirectBufferAddress exposes of a java.nio direct buffer.
 * Input on stdin (text): the RM2 golden job and a PPC job; output: what a Java caller would read back.
 * Compiled and run by tests/test_jni_stub.py.
 */
#include <jni.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

typedef struct { void* addr; jlong cap; } DirectBuffer;
static void* get_addr(JNIEnv* env, jobject b) { (void)env; return ((DirectBuffer*)b)->addr; }
static jlong get_cap(JNIEnv* env, jobject b) { (void)env; return ((DirectBuffer*)b)->cap; }
static jstring new_utf(JNIEnv* env, const char* s) { (void)env; return (jstring)strdup(s); }
static char g_thrown[256];                      /* the pending Java exception, as the JVM would hold it */
static jclass find_class(JNIEnv* env, const char* name) { (void)env; return (jclass)strdup(name); }
static jint throw_new(JNIEnv* env, jclass cls, const char* msg) { (void)env; snprintf(g_thrown, sizeof(g_thrown), "%s: %s", (char*)cls, msg); return 0; }
static const struct JNINativeInterface_ TABLE = {get_addr, get_cap, new_utf, find_class, throw_new};

#define DECL(ret, name, ...) ret name(JNIEnv*, jclass, __VA_ARGS__)
DECL(jlong, Java_es_udc_fi_dc_irlab_rm_RM2Native_create, jdouble, jint, jint, jint, jint, jint, jint, jint);
DECL(jint, Java_es_udc_fi_dc_irlab_rm_RM2Native_scoreGroup, jlong, jint, jint, jint, jobject, jobject, jint, jobject, jobject, jobject, jlong, jobject, jint);
DECL(jint, Java_es_udc_fi_dc_irlab_rm_RM2Native_stats, jlong, jobject, jobject, jobject);
DECL(jint, Java_es_udc_fi_dc_irlab_rm_RM2Native_maxItem, jlong);
DECL(jint, Java_es_udc_fi_dc_irlab_rm_RM2Native_setRatings, jlong, jobject, jobject, jobject, jlong);
DECL(jint, Java_es_udc_fi_dc_irlab_rm_RM2Native_setClustering, jlong, jobject, jobject, jlong, jobject, jint);
DECL(jint, Java_es_udc_fi_dc_irlab_rm_RM2Native_run, jlong);
DECL(jlong, Java_es_udc_fi_dc_irlab_rm_RM2Native_resultCount, jlong);
DECL(jint, Java_es_udc_fi_dc_irlab_rm_RM2Native_results, jlong, jobject, jobject, jobject, jobject, jobject);
DECL(jstring, Java_es_udc_fi_dc_irlab_rm_RM2Native_lastError, jlong);
DECL(void, Java_es_udc_fi_dc_irlab_rm_RM2Native_destroy, jlong);
DECL(jlong, Java_es_udc_fi_dc_irlab_nmf_ppc_NmfNative_create, jint, jint, jint, jint, jint, jint, jint, jint);
DECL(jint, Java_es_udc_fi_dc_irlab_nmf_ppc_NmfNative_setRatings, jlong, jobject, jobject, jobject, jlong);
DECL(jint, Java_es_udc_fi_dc_irlab_nmf_ppc_NmfNative_setFactors, jlong, jobject, jobject);
DECL(jint, Java_es_udc_fi_dc_irlab_nmf_ppc_NmfNative_run, jlong);
DECL(jint, Java_es_udc_fi_dc_irlab_nmf_ppc_NmfNative_getFactors, jlong, jobject, jobject);
DECL(jint, Java_es_udc_fi_dc_irlab_nmf_ppc_NmfNative_clusterAssignment, jlong, jobject, jobject);
DECL(jstring, Java_es_udc_fi_dc_irlab_nmf_ppc_NmfNative_lastError, jlong);
DECL(void, Java_es_udc_fi_dc_irlab_nmf_ppc_NmfNative_destroy, jlong);

static DirectBuffer* direct_n(size_t n, size_t elem) {  /* ByteBuffer.allocateDirect(n * elem).asXxxBuffer(): capacity in ELEMENTS */
    DirectBuffer* b = (DirectBuffer*)malloc(sizeof(DirectBuffer));
    b->addr = calloc((n * elem) > 0 ? n * elem : 1, 1); b->cap = (jlong)n;
    return b;
}
#define direct(bytes) direct_n((size_t)(bytes) / 4, 4)      /* every 4-byte-element buffer below */
#define direct8(n) direct_n((size_t)(n), 8)
#define I32(b) ((int32_t*)(b)->addr)
#define F32(b) ((float*)(b)->addr)
#define F64(b) ((double*)(b)->addr)

int main(void) {
    const struct JNINativeInterface_* table = &TABLE;
    JNIEnv* env = &table;
    long nnz, n_users, n_clusters, n_items, top_n;
    double lambda;
    /* ---- RM2 job ---- */
    if (scanf("%ld %ld %ld %ld %ld %lf", &nnz, &n_users, &n_clusters, &n_items, &top_n, &lambda) != 6) return 2;
    DirectBuffer *u = direct(nnz * 4), *i = direct(nnz * 4), *s = direct(nnz * 4);
    for (long k = 0; k < nnz; k++) if (scanf("%d %d %f", &I32(u)[k], &I32(i)[k], &F32(s)[k]) != 3) return 2;
    DirectBuffer *cu = direct(n_users * 4), *cc = direct(n_users * 4), *cs = direct(n_clusters * 4);
    for (long k = 0; k < n_users; k++) if (scanf("%d %d", &I32(cu)[k], &I32(cc)[k]) != 2) return 2;
    for (long k = 0; k < n_clusters; k++) if (scanf("%d", &I32(cs)[k]) != 1) return 2;
    jlong h = Java_es_udc_fi_dc_irlab_rm_RM2Native_create(env, NULL, lambda, (jint)n_items, (jint)top_n, 0, 0, 0, 1, 0);
    if (!h) { printf("create failed\n"); return 3; }
    int rc = Java_es_udc_fi_dc_irlab_rm_RM2Native_setRatings(env, NULL, h, u, i, s, nnz);
    if (!rc) rc = Java_es_udc_fi_dc_irlab_rm_RM2Native_setClustering(env, NULL, h, cu, cc, n_users, cs, (jint)n_clusters);
    if (!rc) rc = Java_es_udc_fi_dc_irlab_rm_RM2Native_run(env, NULL, h);
    if (rc) { printf("rm2 error %d %s\n", rc, (char*)Java_es_udc_fi_dc_irlab_rm_RM2Native_lastError(env, NULL, h)); return 4; }
    const jlong n = Java_es_udc_fi_dc_irlab_rm_RM2Native_resultCount(env, NULL, h);
    DirectBuffer *ou = direct(n * 4), *oi = direct(n * 4), *o64 = direct8(n), *o32 = direct(n * 4), *oc = direct(n * 4);
    rc = Java_es_udc_fi_dc_irlab_rm_RM2Native_results(env, NULL, h, ou, oi, o64, o32, oc);
    printf("rm2 %ld\n", (long)n);
    for (jlong k = 0; k < n; k++) printf("%d %d %.17g %d\n", I32(ou)[k], I32(oi)[k], F64(o64)[k], I32(oc)[k]);
    /* an error path: run before set_* on a fresh context must give FY_E_STATE and a message */
    jlong h2 = Java_es_udc_fi_dc_irlab_rm_RM2Native_create(env, NULL, lambda, (jint)n_items, (jint)top_n, 0, 0, 0, 1, 0);
    const int rc2 = Java_es_udc_fi_dc_irlab_rm_RM2Native_run(env, NULL, h2);
    printf("state %d %s\n", rc2, (char*)Java_es_udc_fi_dc_irlab_rm_RM2Native_lastError(env, NULL, h2));
    /* a buffer one element too small must be refused by the stub (FY_E_ARG), not read out of bounds */
    DirectBuffer* small = direct((nnz - 1) * 4);
    printf("capacity %d\n", Java_es_udc_fi_dc_irlab_rm_RM2Native_setRatings(env, NULL, h2, u, small, s, nnz));
    /* create() on a device that does not exist: 0 AND a pending RuntimeException */
    const jlong h3 = Java_es_udc_fi_dc_irlab_rm_RM2Native_create(env, NULL, lambda, (jint)n_items, (jint)top_n, 0, 4096, 0, 1, 0);
    printf("create %ld %s\n", (long)h3, g_thrown);
    /* ---- fine seam: the reduce() groups of RM2GpuHDFSReducer, one scoreGroup call per "c-split-nSplits" key.  The
     * group's records in arrival order (AbstractRM2Reducer.java:149-174): K (user, userSum) records, then the ratings
     * of the cluster; userSum / itemColl come from the statistics jobs (here: the coarse run's stats) ---- */
    DirectBuffer *usum = direct8(n_users), *iprob = direct8(n_items + 2), *tot = direct8(1);
    const jint max_item = Java_es_udc_fi_dc_irlab_rm_RM2Native_maxItem(env, NULL, h);
    rc = Java_es_udc_fi_dc_irlab_rm_RM2Native_stats(env, NULL, h, usum, iprob, tot);
    if (rc) { printf("stats error %d\n", rc); return 4; }
    long cluster_split, split_size, n_fine = 0;
    if (scanf("%ld %ld", &cluster_split, &split_size) != 2) return 2;
    DirectBuffer *gu = direct(n_users * 4), *gs = direct8(n_users), *ru = direct(nnz * 4), *ri = direct(nnz * 4), *rs = direct(nnz * 4);
    printf("fine\n");
    for (long c = 0; c < n_clusters; c++) {
        long K = 0, m = 0;
        for (long k = 0; k < n_users; k++) if (I32(cc)[k] == c) { I32(gu)[K] = I32(cu)[k]; F64(gs)[K] = F64(usum)[k]; K++; }
        if (K == 0) continue;
        for (long e = 0; e < nnz; e++) {
            int in = 0;
            for (long k = 0; k < K; k++) if (I32(gu)[k] == I32(u)[e]) { in = 1; break; }
            if (in) { I32(ru)[m] = I32(u)[e]; I32(ri)[m] = I32(i)[e]; F32(rs)[m] = F32(s)[e]; m++; }
        }
        const long n_splits = K >= cluster_split ? (K + split_size - 1) / split_size : 1;   /* AbstractByClusterAndCountMapper.java:86-102 */
        for (long sp = 0; sp < n_splits; sp++) {
            rc = Java_es_udc_fi_dc_irlab_rm_RM2Native_scoreGroup(env, NULL, h2, (jint)c, (jint)sp, (jint)n_splits, gu, gs, (jint)K, ru, ri, rs,
                                                                 m, iprob, max_item);
            if (rc) { printf("scoreGroup error %d %s\n", rc, (char*)Java_es_udc_fi_dc_irlab_rm_RM2Native_lastError(env, NULL, h2)); return 4; }
            const jlong nn = Java_es_udc_fi_dc_irlab_rm_RM2Native_resultCount(env, NULL, h2);
            DirectBuffer *fu = direct(nn * 4), *fi = direct(nn * 4), *f64 = direct8(nn), *fc = direct(nn * 4);
            rc = Java_es_udc_fi_dc_irlab_rm_RM2Native_results(env, NULL, h2, fu, fi, f64, NULL, fc);
            for (jlong k = 0; k < nn; k++) printf("%d %d %.17g %d %ld\n", I32(fu)[k], I32(fi)[k], F64(f64)[k], I32(fc)[k], sp);
            n_fine += nn;
        }
    }
    printf("fine-end %ld\n", n_fine);
    Java_es_udc_fi_dc_irlab_rm_RM2Native_destroy(env, NULL, h2);
    Java_es_udc_fi_dc_irlab_rm_RM2Native_destroy(env, NULL, h);
    /* ---- PPC job ---- */
    long pu, pi, pk, pit, pnnz;
    if (scanf("%ld %ld %ld %ld %ld", &pu, &pi, &pk, &pit, &pnnz) != 5) return 2;
    DirectBuffer *xu = direct(pnnz * 4), *xi = direct(pnnz * 4), *xs = direct(pnnz * 4);
    for (long k = 0; k < pnnz; k++) if (scanf("%d %d %f", &I32(xu)[k], &I32(xi)[k], &F32(xs)[k]) != 3) return 2;
    DirectBuffer *H = direct8(pu * pk), *W = direct8(pi * pk);
    for (long k = 0; k < pu * pk; k++) if (scanf("%lf", &F64(H)[k]) != 1) return 2;
    for (long k = 0; k < pi * pk; k++) if (scanf("%lf", &F64(W)[k]) != 1) return 2;
    jlong g = Java_es_udc_fi_dc_irlab_nmf_ppc_NmfNative_create(env, NULL, 1, (jint)pu, (jint)pi, (jint)pk, (jint)pit, 12, 1, 0);
    if (!g) { printf("nmf create failed\n"); return 3; }
    rc = Java_es_udc_fi_dc_irlab_nmf_ppc_NmfNative_setRatings(env, NULL, g, xu, xi, xs, pnnz);
    if (!rc) rc = Java_es_udc_fi_dc_irlab_nmf_ppc_NmfNative_setFactors(env, NULL, g, H, W);
    if (!rc) rc = Java_es_udc_fi_dc_irlab_nmf_ppc_NmfNative_run(env, NULL, g);
    if (!rc) rc = Java_es_udc_fi_dc_irlab_nmf_ppc_NmfNative_getFactors(env, NULL, g, H, W);
    DirectBuffer *cl = direct(pu * 4), *cnt = direct(pk * 4);
    if (!rc) rc = Java_es_udc_fi_dc_irlab_nmf_ppc_NmfNative_clusterAssignment(env, NULL, g, cl, cnt);
    if (rc) { printf("nmf error %d %s\n", rc, (char*)Java_es_udc_fi_dc_irlab_nmf_ppc_NmfNative_lastError(env, NULL, g)); return 4; }
    printf("ppc %ld %ld\n", pu, pk);
    for (long k = 0; k < pu * pk; k++) printf("%.17g\n", F64(H)[k]);
    for (long k = 0; k < pu; k++) printf("%d\n", I32(cl)[k]);
    Java_es_udc_fi_dc_irlab_nmf_ppc_NmfNative_destroy(env, NULL, g);
    return 0;
}
