package es.udc.fi.dc.irlab.rm;

import java.io.File;
import java.nio.ByteBuffer;
import java.nio.ByteOrder;
import java.nio.DoubleBuffer;
import java.nio.FloatBuffer;
import java.nio.IntBuffer;

import org.apache.hadoop.conf.Configuration;
import org.apache.hadoop.fs.FileSystem;
import org.apache.hadoop.fs.Path;
import org.apache.hadoop.io.FloatWritable;
import org.apache.hadoop.io.IntWritable;
import org.apache.hadoop.io.SequenceFile;
import org.apache.mahout.common.IntPairWritable;

import es.udc.fi.dc.irlab.rmrecommender.RMRecommenderDriver;
import es.udc.fi.dc.irlab.util.HadoopUtils;

/**
 * Drop-in for {@link RM2Job}: same Configuration keys, same input and output paths, one native call instead of the
 * three MapReduce jobs RM2-1..3.  Selected by RMRecommenderDriver.run when -Drm2.gpu=true:
 *
 * <pre>
 * final Tool rm2 = conf.getBoolean("rm2.gpu", false) ? new RM2GpuJob() : new RM2Job();
 * if (conf.getInt(numberOfRecommendations, -1) &gt; 0 &amp;&amp; ToolRunner.run(conf, rm2, args) &lt; 0) { ... }
 * </pre>
 *
 * Not compiled in the build image (no JDK); the native half is exercised by tests/test_jni_stub.py.
 */
public class RM2GpuJob extends RM2Job {

    private static IntBuffer ints(final int n) {
        return ByteBuffer.allocateDirect(4 * Math.max(n, 1)).order(ByteOrder.nativeOrder()).asIntBuffer();
    }

    private static FloatBuffer floats(final int n) {
        return ByteBuffer.allocateDirect(4 * Math.max(n, 1)).order(ByteOrder.nativeOrder()).asFloatBuffer();
    }

    private static DoubleBuffer doubles(final int n) {
        return ByteBuffer.allocateDirect(8 * Math.max(n, 1)).order(ByteOrder.nativeOrder()).asDoubleBuffer();
    }

    @Override
    public int run(final String[] args) throws Exception {
        final Configuration conf = getConf();
        final String directory = conf.get(RMRecommenderDriver.directory);
        final int numberOfClusters = conf.getInt(RMRecommenderDriver.numberOfClusters, -1);

        /* 1. ratings: the records the three mappers of RM2Job read */
        int nnz = 0;
        for (final SequenceFile.Reader reader : HadoopUtils.getSequenceReaders(HadoopUtils.getInputPath(conf), conf)) {
            final IntPairWritable key = new IntPairWritable();
            final FloatWritable val = new FloatWritable();
            while (reader.next(key, val)) {
                nnz++;
            }
        }
        final IntBuffer user = ints(nnz), item = ints(nnz);
        final FloatBuffer score = floats(nnz);
        for (final SequenceFile.Reader reader : HadoopUtils.getSequenceReaders(HadoopUtils.getInputPath(conf), conf)) {
            final IntPairWritable key = new IntPairWritable();
            final FloatWritable val = new FloatWritable();
            while (reader.next(key, val)) {
                user.put(key.getFirst());
                item.put(key.getSecond());
                score.put(val.get());
            }
        }

        /* 2. clustering / clusteringCount: the two DistributedCache files of RM2-3 */
        final Path clustering = new Path(directory + File.separator + conf.get(RMRecommenderDriver.clustering));
        final Path clusteringCount = new Path(directory + File.separator + conf.get(RMRecommenderDriver.clusteringCount));
        int nUsers = 0;
        for (final SequenceFile.Reader reader : HadoopUtils.getSequenceReaders(clustering, conf)) {
            final IntWritable k = new IntWritable(), v = new IntWritable();
            while (reader.next(k, v)) {
                nUsers++;
            }
        }
        final IntBuffer clUser = ints(nUsers), clCluster = ints(nUsers), clusterSize = ints(numberOfClusters);
        for (final SequenceFile.Reader reader : HadoopUtils.getSequenceReaders(clustering, conf)) {
            final IntWritable k = new IntWritable(), v = new IntWritable();
            while (reader.next(k, v)) {
                clUser.put(k.get());
                clCluster.put(v.get());
            }
        }
        for (final SequenceFile.Reader reader : HadoopUtils.getSequenceReaders(clusteringCount, conf)) {
            final IntWritable k = new IntWritable(), v = new IntWritable();
            while (reader.next(k, v)) {
                clusterSize.put(k.get(), v.get());
            }
        }

        /* 3. the whole of RM2-1..3 */
        final long ctx = RM2Native.create(Double.valueOf(conf.get(RMRecommenderDriver.lambda)),
                conf.getInt(RMRecommenderDriver.numberOfItems, -1),
                conf.getInt(RMRecommenderDriver.numberOfRecommendations, -1),
                conf.getInt(RMRecommenderDriver.filterUsers, 0), conf.getInt("rm2.gpu.device", 0), 0, 1);
        if (ctx == 0) {
            throw new RuntimeException("RM2-GPU failed! no usable B200 / libfilmyou_rm2.so");
        }
        try {
            check(ctx, RM2Native.setRatings(ctx, user, item, score, nnz), "RM2-1");
            check(ctx, RM2Native.setClustering(ctx, clUser, clCluster, nUsers, clusterSize, numberOfClusters), "RM2-3");
            check(ctx, RM2Native.run(ctx), "RM2-3");
            final int n = (int) RM2Native.resultCount(ctx);
            final IntBuffer outUser = ints(n), outItem = ints(n);
            final FloatBuffer outScore = floats(n);
            check(ctx, RM2Native.results(ctx, outUser, outItem, null, outScore, null), "RM2-3");

            /* 4. the sink of RM2HDFSReducer (RM2HDFSReducer.java:44-50) */
            final Path out = new Path(HadoopUtils.getOutputPath(conf), "part-r-00000");
            final FileSystem fs = out.getFileSystem(conf);
            final SequenceFile.Writer writer = SequenceFile.createWriter(fs, conf, out, IntPairWritable.class,
                    FloatWritable.class);
            try {
                for (int k = 0; k < n; k++) {
                    writer.append(new IntPairWritable(outUser.get(k), outItem.get(k)),
                            new FloatWritable(outScore.get(k)));
                }
            } finally {
                writer.close();
            }
        } finally {
            RM2Native.destroy(ctx);
        }
        return 0;
    }

    private static void check(final long ctx, final int rc, final String job) {
        if (rc != 0) {
            throw new RuntimeException(job + " failed! " + RM2Native.lastError(ctx)); // RM2Job.java:265-268
        }
    }
}
