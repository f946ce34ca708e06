package es.udc.fi.dc.irlab.rm;

import java.nio.DoubleBuffer;
import java.nio.FloatBuffer;
import java.nio.IntBuffer;

/**
 * Native side of {@link RM2GpuJob}: one static method per entry point of include/filmyou_rm2.h (B200 engine).
 * Every buffer must be DIRECT (ByteBuffer.allocateDirect(..).order(ByteOrder.nativeOrder())): the stub
 * (integration/jni/filmyou_rm2_jni.c) passes GetDirectBufferAddress straight to the C ABI.
 * A context handle is single-caller; 0 from create means "no usable B200" (there is no CPU fallback).
 */
public final class RM2Native {

    static {
        System.loadLibrary("filmyou_rm2_jni");
    }

    private RM2Native() {
    }

    /**
     * @param nGpus 0/1: this context drives <code>device</code>; n &gt; 1: ONE context drives devices device..device+n-1
     *              (users sharded over them, results concatenated) -- shardRank/shardCount must then be 0/1.
     * @throws RuntimeException when no usable B200 exists or a parameter is invalid (also returns 0)
     */
    public static native long create(double lambda, int numberOfItems, int numberOfRecommendations,
            int filterUsers, int device, int shardRank, int shardCount, int nGpus);

    public static native int setRatings(long ctx, IntBuffer user, IntBuffer item, FloatBuffer score, long nnz);

    public static native int setClustering(long ctx, IntBuffer user, IntBuffer cluster, long nUsers,
            IntBuffer clusterSize, int nClusters);

    public static native int run(long ctx);

    /** Fine seam: one AbstractRM2Reducer.reduce() group. */
    public static native int scoreGroup(long ctx, int cluster, int split, int nSplits, IntBuffer groupUser,
            DoubleBuffer groupUserSum, int nGroupUsers, IntBuffer rUser, IntBuffer rItem, FloatBuffer rScore,
            long nnz, DoubleBuffer itemProb, int maxItem);

    public static native long resultCount(long ctx);

    public static native int results(long ctx, IntBuffer user, IntBuffer item, DoubleBuffer score64,
            FloatBuffer score32, IntBuffer cluster);

    /** One (user, cluster, count) record per scored user, in the order of the triples of {@link #results}. */
    public static native long resultRowCount(long ctx);

    public static native int resultRows(long ctx, IntBuffer user, IntBuffer cluster, IntBuffer count);

    public static native int stats(long ctx, DoubleBuffer userSum, DoubleBuffer itemProb, DoubleBuffer total);

    public static native int maxItem(long ctx);

    public static native String lastError(long ctx);

    public static native void destroy(long ctx);
}
