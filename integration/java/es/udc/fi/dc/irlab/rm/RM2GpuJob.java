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
import org.apache.hadoop.io.DoubleWritable;
import org.apache.hadoop.io.FloatWritable;
import org.apache.hadoop.io.IntWritable;
import org.apache.hadoop.io.MapFile;
import org.apache.hadoop.io.SequenceFile;
import org.apache.mahout.common.IntPairWritable;

import es.udc.fi.dc.irlab.rmrecommender.RMRecommenderDriver;
import es.udc.fi.dc.irlab.util.HadoopUtils;

/**
 * Coarse seam of the B200 engine, a drop-in for {@link RM2Job}: same Configuration keys, same input and output
 * paths, the same three outputs (rm2/userSum, rm2/itemColl, the recommendations), one native call instead of the
 * three MapReduce jobs RM2-1..3.  Selected by RMRecommenderDriver.run when -Drm2.gpu=true:
 *
 * <pre>
 * final Tool rm2 = conf.getBoolean("rm2.gpu", false) ? new RM2GpuJob() : new RM2Job();
 * if (conf.getInt(numberOfRecommendations, -1) &gt; 0 &amp;&amp; ToolRunner.run(conf, rm2, args) &lt; 0) { ... }
 * </pre>
 *
 * Differences from RM2Job, all deliberate:
 * <ul>
 * <li>Cassandra on either side (useCassandraInput / useCassandraOutput, both <code>true</code> by default inside
 * RM2Job) is not read or written natively: those runs are delegated to <code>super.run</code>, the stock job; the GPU
 * option there is the fine seam, {@link RM2GpuCassandraReducer} selected by a one-line change in
 * RM2Job.runItemRecommendation (INTEGRATION.md, section 3).</li>
 * <li><code>rm2.gpu.count</code> (default 1) devices starting at <code>rm2.gpu.device</code> (default 0) score the users
 * in one native call (fy_rm2_params.n_gpus); the reduce-task fan-out of RM2-3 (numReduceTasks = numberOfClusters)
 * has no other counterpart here.</li>
 * <li>One output part file per job instead of one per reduce task.</li>
 * </ul>
 *
 * Not compiled in the build image (no JDK); the native half is exercised by tests/test_jni_stub.py.
 */
public class RM2GpuJob extends RM2Job {

    private static ByteBuffer direct(final long bytes) {
        if (bytes > Integer.MAX_VALUE) { // one direct ByteBuffer holds at most 2^31-1 bytes
            throw new IllegalArgumentException(bytes + " bytes do not fit one direct buffer; shard the job (rm2.gpu.count) "
                    + "or lower numberOfRecommendations");
        }
        return ByteBuffer.allocateDirect((int) Math.max(bytes, 8L)).order(ByteOrder.nativeOrder());
    }

    private static IntBuffer ints(final long n) {
        return direct(4L * n).asIntBuffer();
    }

    private static FloatBuffer floats(final long n) {
        return direct(4L * n).asFloatBuffer();
    }

    private static DoubleBuffer doubles(final long n) {
        return direct(8L * n).asDoubleBuffer();
    }

    @Override
    public int run(final String[] args) throws Exception {
        final Configuration conf = getConf();
        /* RM2Job reads both flags with default TRUE (RM2Job.java:120,228,237); HadoopUtils.getInputPath / getOutputPath
         * both test useCassandraInput (HadoopUtils.java:143-162) */
        if (conf.getBoolean(RMRecommenderDriver.useCassandraInput, true)
                || conf.getBoolean(RMRecommenderDriver.useCassandraOutput, true)) {
            return super.run(args); // the Cassandra formats stay the reference's own code path
        }
        final String baseDirectory = conf.get(RMRecommenderDriver.directory);
        final String directory = baseDirectory + "/rm2";
        HadoopUtils.removeData(conf, directory); // RM2Job.java:82-84
        final int numberOfClusters = conf.getInt(RMRecommenderDriver.numberOfClusters, -1);

        /* 1. ratings: the records the three mappers of RM2Job read */
        long nnz = 0;
        for (final SequenceFile.Reader reader : HadoopUtils.getSequenceReaders(HadoopUtils.getInputPath(conf), conf)) {
            final IntPairWritable key = new IntPairWritable();
            final FloatWritable val = new FloatWritable();
            while (reader.next(key, val)) {
                nnz++;
            }
            reader.close();
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
            reader.close();
        }

        /* 2. clustering / clusteringCount: the two DistributedCache files of RM2-3 */
        final Path clustering = new Path(baseDirectory + File.separator + conf.get(RMRecommenderDriver.clustering));
        final Path clusteringCount = new Path(
                baseDirectory + File.separator + conf.get(RMRecommenderDriver.clusteringCount));
        int nUsers = 0;
        for (final SequenceFile.Reader reader : HadoopUtils.getSequenceReaders(clustering, conf)) {
            final IntWritable k = new IntWritable(), v = new IntWritable();
            while (reader.next(k, v)) {
                nUsers++;
            }
            reader.close();
        }
        final IntBuffer clUser = ints(nUsers), clCluster = ints(nUsers), clusterSize = ints(numberOfClusters);
        for (final SequenceFile.Reader reader : HadoopUtils.getSequenceReaders(clustering, conf)) {
            final IntWritable k = new IntWritable(), v = new IntWritable();
            while (reader.next(k, v)) {
                clUser.put(k.get());
                clCluster.put(v.get());
            }
            reader.close();
        }
        for (final SequenceFile.Reader reader : HadoopUtils.getSequenceReaders(clusteringCount, conf)) {
            final IntWritable k = new IntWritable(), v = new IntWritable();
            while (reader.next(k, v)) {
                clusterSize.put(k.get(), v.get());
            }
            reader.close();
        }

        /* 3. the whole of RM2-1..3 on rm2.gpu.count devices */
        final long ctx = RM2Native.create(Double.valueOf(conf.get(RMRecommenderDriver.lambda)),
                conf.getInt(RMRecommenderDriver.numberOfItems, -1),
                conf.getInt(RMRecommenderDriver.numberOfRecommendations, -1),
                conf.getInt(RMRecommenderDriver.filterUsers, 0), conf.getInt("rm2.gpu.device", 0), 0, 1,
                conf.getInt("rm2.gpu.count", 1));
        if (ctx == 0) {
            throw new RuntimeException("RM2-GPU failed! no usable B200 / libfilmyou_rm2.so");
        }
        try {
            check(ctx, RM2Native.setRatings(ctx, user, item, score, nnz), "RM2-1");
            check(ctx, RM2Native.setClustering(ctx, clUser, clCluster, nUsers, clusterSize, numberOfClusters), "RM2-3");
            check(ctx, RM2Native.run(ctx), "RM2-3");

            /* 4. rm2/userSum (SequenceFile<IntWritable, DoubleWritable>, RM2Job.java:138-142) and rm2/itemColl
             * (MapFile<IntWritable, DoubleWritable>, RM2Job.java:190-196): what TestHDFSRM2.java:70-71 reads back */
            final int maxItem = RM2Native.maxItem(ctx);
            final DoubleBuffer userSum = doubles(nUsers), itemProb = doubles(maxItem + 1L), total = doubles(1);
            check(ctx, RM2Native.stats(ctx, userSum, itemProb, total), "RM2-2");
            final FileSystem fs = FileSystem.get(conf);
            final SequenceFile.Writer sums = SequenceFile.createWriter(fs, conf,
                    new Path(directory + File.separator + RM2Job.USER_SUM, "part-r-00000"), IntWritable.class,
                    DoubleWritable.class);
            try {
                /* ascending user id, as one reducer of RM2-1 emits them; users arrive in `clustering` order */
                final Integer[] order = new Integer[nUsers];
                for (int k = 0; k < nUsers; k++) {
                    order[k] = k;
                }
                java.util.Arrays.sort(order, new java.util.Comparator<Integer>() { // the pom compiles with -source 1.7: no lambdas
                    @Override
                    public int compare(final Integer a, final Integer b) {
                        return Integer.compare(clUser.get(a), clUser.get(b));
                    }
                });
                for (final int k : order) {
                    sums.append(new IntWritable(clUser.get(k)), new DoubleWritable(userSum.get(k)));
                }
            } finally {
                sums.close();
            }
            final MapFile.Writer coll = new MapFile.Writer(conf, fs,
                    new Path(directory + File.separator + RM2Job.ITEMM_COLL, "part-r-00000").toString(),
                    IntWritable.class, DoubleWritable.class);
            try {
                for (int i = 0; i <= maxItem; i++) { // ascending keys, as MapFile requires; unrated ids have no entry
                    if (itemProb.get(i) > 0.0) {
                        coll.append(new IntWritable(i), new DoubleWritable(itemProb.get(i)));
                    }
                }
            } finally {
                coll.close();
            }

            /* 5. the sink of RM2HDFSReducer (RM2HDFSReducer.java:44-50): the 12-byte (item, score) stream plus one
             * (user, cluster, count) record per user, expanded here */
            final long n = RM2Native.resultCount(ctx);
            final long rows = RM2Native.resultRowCount(ctx);
            final IntBuffer outItem = ints(n), rowUser = ints(rows), rowCount = ints(rows);
            final DoubleBuffer outScore = doubles(n);
            check(ctx, RM2Native.results(ctx, null, outItem, outScore, null, null), "RM2-3");
            check(ctx, RM2Native.resultRows(ctx, rowUser, null, rowCount), "RM2-3");
            final Path out = new Path(HadoopUtils.getOutputPath(conf), "part-r-00000");
            final SequenceFile.Writer writer = SequenceFile.createWriter(out.getFileSystem(conf), conf, out,
                    IntPairWritable.class, FloatWritable.class);
            try {
                int t = 0;
                for (int r = 0; r < rows; r++) {
                    for (int k = 0; k < rowCount.get(r); k++, t++) {
                        writer.append(new IntPairWritable(rowUser.get(r), outItem.get(t)),
                                new FloatWritable((float) outScore.get(t)));
                    }
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
