package es.udc.fi.dc.irlab.rm;

import java.io.IOException;

import org.apache.hadoop.io.FloatWritable;
import org.apache.mahout.common.IntPairWritable;

/**
 * GPU twin of {@link RM2HDFSReducer}: same output records, &lt;(user, item), (float) score&gt; into the
 * SequenceFile&lt;IntPairWritable, FloatWritable&gt; of RM2-3.  Selected in RM2Job.runItemRecommendation by
 * <code>job.setReducerClass(conf.getBoolean("rm2.gpu", false) ? RM2GpuHDFSReducer.class : RM2HDFSReducer.class)</code>.
 */
public class RM2GpuHDFSReducer extends AbstractRM2GpuReducer<IntPairWritable, FloatWritable> {

    @Override
    protected void writePreference(final Context context, final int userId, final int itemId, final double score,
            final int cluster) throws IOException, InterruptedException {
        context.write(new IntPairWritable(userId, itemId), new FloatWritable((float) score));
    }

}
