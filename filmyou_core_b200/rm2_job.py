"""Host-side mirror of the reference's job interface for the RM2 path, over the C ABI.

Names, configuration keys and error behaviour follow the reference so that tests read like its own:

  RM2Job.run()                         M/rm/RM2Job.java:76-100   (jobs RM2-1, RM2-2, RM2-3)
  conf keys                            M/rmrecommender/RMRecommenderDriver.java:89-120
  PreferenceSink.writePreference(...)  M/rm/AbstractRM2Reducer.java:405-407
      HDFSSink       -> (IntPairWritable(user,item), FloatWritable((float) score))   M/rm/RM2HDFSReducer.java:44-50
      CassandraSink  -> (user, relevance float, item, cluster)                       M/rm/RM2CassandraReducer.java:49-63
  a failed job raises (the reference throws RuntimeException(job + " failed!"),      M/rm/RM2Job.java:265-268)

The GPU does jobs RM2-1..3 in one fy_rm2_run call; the Hadoop drivers, record formats and sinks stay
on the host (off the timed path), exactly as BASELINE.json's north_star prescribes.
"""
import numpy as np

from .engine import Rm2Engine, Rm2Error

# defaults of RMRecommenderDriver (:95, :114, :116-119)
DEFAULTS = {"numberOfRecommendations": 1000, "lambda": "0.1", "clusterSplit": 400, "splitSize": 100, "filterUsers": 0}


class PreferenceSink:
    """writePreference(userId, itemId, score, cluster) -- the reducer's abstract sink."""

    def writePreference(self, userId, itemId, score, cluster):      # noqa: N802,N803 (reference names)
        raise NotImplementedError


class HDFSSink(PreferenceSink):
    """Collects what RM2HDFSReducer would append to the output SequenceFile."""

    def __init__(self):
        self.records = []           # ((user, item), float32 score)

    def writePreference(self, userId, itemId, score, cluster):      # noqa: N802,N803
        self.records.append(((int(userId), int(itemId)), np.float32(score)))


class CassandraSink(PreferenceSink):
    """Collects the bound variables of RM2CassandraReducer's UPDATE ... SET cluster = ?."""

    def __init__(self):
        self.rows = []              # (user, relevance float32, item, cluster)

    def writePreference(self, userId, itemId, score, cluster):      # noqa: N802,N803
        self.rows.append((int(userId), np.float32(score), int(itemId), int(cluster)))


class RM2Job:
    """conf: dict with the reference's keys (numberOfItems and numberOfClusters are required,
    RMRecommenderDriver.java:91-92)."""

    def __init__(self, conf, device=0, shard_rank=0, shard_count=1, n_gpus=None):
        # n_gpus (or conf["rm2.gpu.count"], the key integration/java/.../RM2GpuJob.java reads): one context over that many
        # devices from `device` on -- the reduce-task fan-out of RM2Job.java:251 inside one native call
        for k in ("numberOfItems", "numberOfClusters"):
            if k not in conf:
                raise KeyError("missing required option --%s" % k)
        self.conf = dict(DEFAULTS)
        self.conf.update(conf)
        self.userSum = None          # rm2/userSum
        self.itemColl = None         # rm2/itemColl
        self.totalSum = None
        self._eng = Rm2Engine(lam=float(self.conf["lambda"]),            # Double.valueOf(conf.get("lambda"))
                              number_of_items=int(self.conf["numberOfItems"]),
                              top_n=int(self.conf["numberOfRecommendations"]),
                              filter_users=int(self.conf["filterUsers"]),
                              device=device, shard_rank=shard_rank, shard_count=shard_count,
                              n_gpus=int(n_gpus if n_gpus is not None else self.conf.get("rm2.gpu.count", 0)))

    def close(self):
        self._eng.close()

    def run(self, ratings_user, ratings_item, ratings_score, clustering_user, clustering_cluster, clusteringCount,
            sink=None):
        """Inputs are the contents of the ratings SequenceFile and of the `clustering` /
        `clusteringCount` files.  Returns the packed triples; with a sink, calls
        sink.writePreference for every triple in emission order."""
        k = int(self.conf["numberOfClusters"])
        csize = np.zeros(k, np.int32)
        cc = np.asarray(clusteringCount, np.int32)
        if len(cc) > k:
            raise Rm2Error(-1, "clusteringCount has more entries than numberOfClusters")
        csize[:len(cc)] = cc
        self._eng.set_ratings(ratings_user, ratings_item, ratings_score)
        self._eng.set_clustering(clustering_user, clustering_cluster, csize)
        try:
            self._eng.run()
        except Rm2Error as e:
            raise RuntimeError("RM2-3 failed! " + str(e)) from e
        self.userSum, self.itemColl, self.totalSum = self._eng.stats()
        out = self._eng.results()
        if sink is not None:
            for u, i, s, c in zip(out["user"], out["item"], out["score64"], out["cluster"]):
                sink.writePreference(u, i, s, c)
        return out
