package es.udc.fi.dc.irlab.rm;

import java.io.IOException;
import java.nio.ByteBuffer;
import java.util.LinkedHashMap;
import java.util.LinkedList;
import java.util.List;
import java.util.Map;

import org.apache.cassandra.utils.ByteBufferUtil;

/**
 * GPU twin of {@link RM2CassandraReducer}: the CQL row of the recommendations table, keys (user, item, relevance) and
 * the bound value (cluster) of <code>UPDATE ... SET cluster = ?</code> (CassandraSetup.updateConfForOutput).
 */
public class RM2GpuCassandraReducer extends AbstractRM2GpuReducer<Map<String, ByteBuffer>, List<ByteBuffer>> {

    @Override
    protected void writePreference(final Context context, final int userId, final int itemId, final double score,
            final int cluster) throws IOException, InterruptedException {
        final Map<String, ByteBuffer> keys = new LinkedHashMap<String, ByteBuffer>();
        keys.put("user", ByteBufferUtil.bytes(userId));
        keys.put("item", ByteBufferUtil.bytes(itemId));
        keys.put("relevance", ByteBufferUtil.bytes((float) score));
        final List<ByteBuffer> value = new LinkedList<ByteBuffer>();
        value.add(ByteBufferUtil.bytes(cluster));
        context.write(keys, value);
    }

}
