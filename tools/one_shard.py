#!/usr/bin/env python3
"""One rank's share of the ML-20M job (shard_rank / shard_count given), single process: the command
ncu profiles to see a multi-GPU rank's kernels (ncu must not wrap a multi-rank launch)."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import filmyou_core_b200 as fy
from filmyou_core_b200 import datagen

rank, world, runs = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]) if len(sys.argv) > 3 else 2
r = datagen.generate("ml-20m")
with fy.Rm2Engine(lam=0.1, number_of_items=r.n_items, top_n=100, shard_rank=rank, shard_count=world) as eng:
    eng.set_ratings(r.user, r.item, r.score)
    eng.set_clustering(r.cl_user, r.cl_cluster, r.cluster_size)
    for _ in range(runs):
        eng.run()
    print(json.dumps({"results": eng.result_count(), **eng.profile()}))
