#!/usr/bin/env python3
"""Run one named synthetic shape (datagen.SHAPES) through the C ABI on one GPU and print the profile."""
import sys, os, json, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import filmyou_core_b200 as fy
from filmyou_core_b200 import datagen

shape, runs = sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 3
t = time.time(); r = datagen.generate(shape); t_gen = time.time() - t
with fy.Rm2Engine(lam=0.1, number_of_items=r.n_items, top_n=100) as eng:
    t = time.time()
    eng.set_ratings(r.user, r.item, r.score)
    eng.set_clustering(r.cl_user, r.cl_cluster, r.cluster_size)
    t_set = time.time() - t
    for _ in range(runs):
        eng.run()
    p = eng.profile()
    res = eng.results()
scores = res["score64"].reshape(-1, 100) if len(res["user"]) == r.n_users * 100 else None
print(json.dumps({"shape": shape, "nnz": r.nnz, "datagen_s": t_gen, "set_s": t_set, "results": len(res["user"]),
                  "users_per_s": p["users_scored"] / (p["ms_total"] * 1e-3),
                  "descending": bool(scores is not None and np.all(np.diff(scores, axis=1) <= 0)), **p}))
