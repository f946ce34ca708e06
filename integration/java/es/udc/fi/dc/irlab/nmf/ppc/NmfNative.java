package es.udc.fi.dc.irlab.nmf.ppc;

import java.nio.DoubleBuffer;
import java.nio.FloatBuffer;
import java.nio.IntBuffer;

/**
 * Native side of the PPC / NMF clustering step (include/filmyou_nmf.h): replaces the 9 MapReduce jobs per iteration
 * of PPCDriver / NMFDriver plus ClusterAssignmentJob(false) and CountClustersJob.  Direct buffers only.
 * mode: 0 = NMFDriver, 1 = PPCDriver.  H is [numberOfUsers x numberOfClusters], W [numberOfItems x numberOfClusters],
 * row-major, row r = id idBase + r (1 for the files DataInitialization.createMatrix writes).
 */
public final class NmfNative {

    static {
        System.loadLibrary("filmyou_rm2_jni");
    }

    private NmfNative() {
    }

    public static native long create(int mode, int numberOfUsers, int numberOfItems, int numberOfClusters,
            int numberOfIterations, int normalizationFrequency, int idBase, int device);

    public static native int setRatings(long ctx, IntBuffer user, IntBuffer item, FloatBuffer score, long nnz);

    public static native int setFactors(long ctx, DoubleBuffer H, DoubleBuffer W);

    public static native int initRandom(long ctx, long seed);

    public static native int run(long ctx);

    public static native int getFactors(long ctx, DoubleBuffer H, DoubleBuffer W);

    /** cluster[r] = arg max of row r of H; clusterSize[c] = users assigned to c. */
    public static native int clusterAssignment(long ctx, IntBuffer cluster, IntBuffer clusterSize);

    public static native String lastError(long ctx);

    public static native void destroy(long ctx);
}
