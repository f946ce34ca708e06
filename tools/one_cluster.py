#!/usr/bin/env python3
"""One cluster through the C ABI: the short command ncu profiles (the H build, score, top-N once each per run).
  one_cluster.py [runs] [ml20m|netflix]     ML-20M-sized: 2770 users x 26744 items, 400k ratings;
                                            Netflix-sized: 9604 users x 17770 items, 2.01M ratings
Prints the profile, the median stage times and a digest of the packed results (equal across kernel variants)."""
import sys, os, json, hashlib
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import filmyou_core_b200 as fy
from filmyou_core_b200 import datagen

runs = int(sys.argv[1]) if len(sys.argv) > 1 else 3
shape = sys.argv[2] if len(sys.argv) > 2 else "ml20m"
if shape == "netflix":
    r = datagen.generate("one-cluster-nf", n_users=9604, n_items=17770, nnz=2_009_610, n_clusters=1, seed=78)
else:
    r = datagen.generate("one-cluster", n_users=2770, n_items=26744, nnz=400_000, n_clusters=1, seed=77)
with fy.Rm2Engine(lam=0.1, number_of_items=r.n_items, top_n=100) as eng:
    eng.set_ratings(r.user, r.item, r.score)
    eng.set_clustering(r.cl_user, r.cl_cluster, r.cluster_size)
    hist = []
    for _ in range(runs):
        eng.run()
        hist.append(eng.profile())
    p = hist[-1]
    res = eng.results()
h = hashlib.sha256()
for k in ("user", "item", "score64"):
    h.update(np.ascontiguousarray(res[k]).tobytes())
med = lambda k: float(np.median([x[k] for x in hist[1:] or hist]))
print(json.dumps({"shape": shape, "variant": os.environ.get("FY_BUILD_H", "2"), "cfg": os.environ.get("FY_H2_CFG", "0"),
                  "results": len(res["user"]), "sha": h.hexdigest()[:16], "median_ms_gram": med("ms_gram"),
                  "median_ms_score": med("ms_score"), "median_ms_refine": med("ms_refine"), "median_ms_topn": med("ms_topn"),
                  "median_ms_total": med("ms_total"), "gram_bytes": p["gram_bytes"]}))
