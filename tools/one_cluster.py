#!/usr/bin/env python3
"""One ML-20M-sized cluster (2770 users x 26744 items, 400k ratings, k=1) through the C ABI:
the short command ncu profiles (k_build_H, k_score, k_topn once each per run)."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import filmyou_core_b200 as fy
from filmyou_core_b200 import datagen

runs = int(sys.argv[1]) if len(sys.argv) > 1 else 3
r = datagen.generate("one-cluster", n_users=2770, n_items=26744, nnz=400_000, n_clusters=1, seed=77)
with fy.Rm2Engine(lam=0.1, number_of_items=r.n_items, top_n=100) as eng:
    eng.set_ratings(r.user, r.item, r.score)
    eng.set_clustering(r.cl_user, r.cl_cluster, r.cluster_size)
    hist = []
    for _ in range(runs):
        eng.run()
        hist.append(eng.profile())
    p = hist[-1]
    n = eng.result_count()
med = lambda k: float(np.median([h[k] for h in hist[1:] or hist]))
print(json.dumps({"results": n, **p, "median_ms_score": med("ms_score"), "median_ms_gram": med("ms_gram"), "median_ms_total": med("ms_total")}))
