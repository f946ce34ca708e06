package es.udc.fi.dc.irlab.rm;

import java.io.FileNotFoundException;
import java.io.IOException;
import java.nio.ByteBuffer;
import java.nio.ByteOrder;
import java.nio.DoubleBuffer;
import java.nio.FloatBuffer;
import java.nio.IntBuffer;
import java.util.Iterator;
import java.util.regex.Matcher;
import java.util.regex.Pattern;

import org.apache.hadoop.conf.Configuration;
import org.apache.hadoop.filecache.DistributedCache;
import org.apache.hadoop.fs.Path;
import org.apache.hadoop.io.DoubleWritable;
import org.apache.hadoop.io.IntWritable;
import org.apache.hadoop.io.MapFile.Reader;
import org.apache.hadoop.io.SequenceFile;
import org.apache.hadoop.mapreduce.Reducer;

import es.udc.fi.dc.irlab.rmrecommender.RMRecommenderDriver;
import es.udc.fi.dc.irlab.util.HadoopUtils;
import es.udc.fi.dc.irlab.util.IntDoubleOrPrefWritable;
import es.udc.fi.dc.irlab.util.MapFileOutputFormat;
import es.udc.fi.dc.irlab.util.StringIntPairWritable;

/**
 * Fine seam of the B200 engine: the same generic signature as {@link AbstractRM2Reducer} (key = cluster or
 * "cluster-split-nSplits", values = the K user-sum records followed by the cluster's rating records), so that
 * {@link RM2Job#runItemRecommendation} can select {@link RM2GpuHDFSReducer} / {@link RM2GpuCassandraReducer} instead of
 * RM2HDFSReducer / RM2CassandraReducer and nothing else of job RM2-3 changes (mappers, partitioner, comparators,
 * DistributedCache files, output formats).
 *
 * One reduce() call = one native call (fy_rm2_score_group): the records are copied into direct buffers in arrival
 * order, the GPU scores the users whose id falls into this split and the sink method is called for every returned
 * triple, in descending score order per user.  The coarse seam ({@link RM2GpuJob}) is the fast path; this class
 * exists for deployments that must keep the three MapReduce jobs.
 *
 * Not compiled in the build image (no JDK); the native half of reduce() is exercised group by group through the JNI
 * stub in tests/test_jni_stub.py.
 *
 * @param <A> output key
 * @param <B> output value
 */
public abstract class AbstractRM2GpuReducer<A, B> extends Reducer<StringIntPairWritable, IntDoubleOrPrefWritable, A, B> {

    private static final Pattern SPLIT_KEY = Pattern.compile("([0-9]+)-([0-9]+)-([0-9]+)");

    private int[] clusterSizes;
    private double[] itemColl; // p(i|C) by item id, the content of the rm2/itemColl MapFile
    private int maxItem;
    private long ctx;

    private static IntBuffer ints(final long n) {
        return direct(4L * Math.max(n, 1)).asIntBuffer();
    }

    private static FloatBuffer floats(final long n) {
        return direct(4L * Math.max(n, 1)).asFloatBuffer();
    }

    private static DoubleBuffer doubles(final long n) {
        return direct(8L * Math.max(n, 1)).asDoubleBuffer();
    }

    private static ByteBuffer direct(final long bytes) {
        if (bytes > Integer.MAX_VALUE) {
            throw new IllegalArgumentException("reduce group of " + bytes + " bytes exceeds one direct buffer");
        }
        return ByteBuffer.allocateDirect((int) bytes).order(ByteOrder.nativeOrder());
    }

    @Override
    public void setup(final Context context) throws IOException, InterruptedException {
        final Configuration conf = context.getConfiguration();
        final Path[] paths = DistributedCache.getLocalCacheFiles(conf);
        if (paths == null || paths.length != 3) {
            throw new FileNotFoundException(); // same contract as AbstractRM2Reducer.setup
        }
        clusterSizes = new int[conf.getInt(RMRecommenderDriver.numberOfClusters, -1)];
        final IntWritable key = new IntWritable();
        final IntWritable val = new IntWritable();
        for (final SequenceFile.Reader reader : HadoopUtils.getLocalSequenceReaders(paths[1], conf)) {
            while (reader.next(key, val)) {
                clusterSizes[key.get()] = val.get();
            }
        }
        /* rm2/itemColl once per task (the reference looks every item up per reduce call): a dense table by item id */
        final int numberOfItems = conf.getInt(RMRecommenderDriver.numberOfItems, -1);
        double[] table = new double[Math.max(numberOfItems, 0) + 2];
        maxItem = -1;
        final IntWritable item = new IntWritable();
        final DoubleWritable prob = new DoubleWritable();
        for (final Reader reader : MapFileOutputFormat.getLocalReaders(paths[2], conf)) {
            while (reader.next(item, prob)) {
                if (item.get() >= table.length) {
                    table = java.util.Arrays.copyOf(table, Math.max(item.get() + 1, 2 * table.length));
                }
                table[item.get()] = prob.get();
                maxItem = Math.max(maxItem, item.get());
            }
        }
        itemColl = table;
        ctx = RM2Native.create(Double.valueOf(conf.get(RM2Job.LAMBDA_NAME)), numberOfItems,
                conf.getInt(RMRecommenderDriver.numberOfRecommendations, -1),
                conf.getInt(RMRecommenderDriver.filterUsers, 0), conf.getInt("rm2.gpu.device", 0), 0, 1, 0);
    }

    @Override
    protected void reduce(final StringIntPairWritable key, final Iterable<IntDoubleOrPrefWritable> values,
            final Context context) throws IOException, InterruptedException {
        /* "c" or "c-split-nSplits" (AbstractByClusterAndCountMapper.getSplits) */
        final String name = key.getKey();
        int cluster, split = 0, numberOfSplits = 1;
        if (name.contains("-")) {
            final Matcher m = SPLIT_KEY.matcher(name);
            m.find();
            cluster = Integer.valueOf(m.group(1));
            split = Integer.valueOf(m.group(2));
            numberOfSplits = Integer.valueOf(m.group(3));
        } else {
            cluster = Integer.valueOf(name);
        }
        final int k = clusterSizes[cluster];
        final Iterator<IntDoubleOrPrefWritable> it = values.iterator();

        /* tag-0 records first, exactly clusterSizes[c] of them (the sort comparator guarantees it) */
        final IntBuffer groupUser = ints(k);
        final DoubleBuffer groupUserSum = doubles(k);
        for (int j = 0; j < k; j++) {
            final IntDoubleOrPrefWritable entry = it.next();
            groupUser.put(j, entry.getKey());
            groupUserSum.put(j, entry.getValue());
        }
        /* then the rating records; their number is not known in advance */
        int cap = 1 << 16, nnz = 0;
        IntBuffer rUser = ints(cap), rItem = ints(cap);
        FloatBuffer rScore = floats(cap);
        while (it.hasNext()) {
            final IntDoubleOrPrefWritable entry = it.next();
            if (nnz == cap) {
                final long grown = 2L * cap;
                final IntBuffer u2 = ints(grown), i2 = ints(grown);
                final FloatBuffer s2 = floats(grown);
                rUser.rewind(); rItem.rewind(); rScore.rewind();
                u2.put(rUser); i2.put(rItem); s2.put(rScore);
                rUser = u2; rItem = i2; rScore = s2;
                cap = (int) grown;
            }
            rUser.put(nnz, entry.getUserId());
            rItem.put(nnz, entry.getItemId());
            rScore.put(nnz, entry.getScore());
            nnz++;
        }
        final DoubleBuffer prob = doubles(maxItem + 1);
        prob.put(itemColl, 0, maxItem + 1);

        final int rc = RM2Native.scoreGroup(ctx, cluster, split, numberOfSplits, groupUser, groupUserSum, k, rUser, rItem,
                rScore, nnz, prob, maxItem);
        if (rc != 0) {
            throw new RuntimeException("RM2-3 (GPU) failed for group " + name + ": " + RM2Native.lastError(ctx));
        }
        final long n = RM2Native.resultCount(ctx);
        final IntBuffer outUser = ints(n), outItem = ints(n);
        final DoubleBuffer outScore = doubles(n);
        if (RM2Native.results(ctx, outUser, outItem, outScore, null, null) != 0) {
            throw new RuntimeException("RM2-3 (GPU) failed for group " + name + ": " + RM2Native.lastError(ctx));
        }
        for (int t = 0; t < n; t++) {
            writePreference(context, outUser.get(t), outItem.get(t), outScore.get(t), cluster);
            if ((t & 0xffff) == 0) {
                context.progress(); // AbstractRM2Reducer.java:218
            }
        }
    }

    @Override
    protected void cleanup(final Context context) throws IOException, InterruptedException {
        RM2Native.destroy(ctx);
        ctx = 0;
    }

    /** Same sink contract as AbstractRM2Reducer.writePreference. */
    protected abstract void writePreference(final Context context, final int userId, final int itemId,
            final double score, final int cluster) throws IOException, InterruptedException;
}
