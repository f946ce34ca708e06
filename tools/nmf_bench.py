"""PPC clustering timing at a BASELINE.json shape: python tools/nmf_bench.py [shape] [k] [iterations]
Prints one JSON line (CUDA-event times from fy_nmf_get_profile)."""
import json
import sys
import time

import numpy as np

sys.path.insert(0, __file__.rsplit("/", 2)[0])
from filmyou_core_b200 import datagen                      # noqa: E402
from filmyou_core_b200.nmf import PPC, NmfEngine           # noqa: E402


def main():
    shape = sys.argv[1] if len(sys.argv) > 1 else "ml-1m"
    k = int(sys.argv[2]) if len(sys.argv) > 2 else 10
    iters = int(sys.argv[3]) if len(sys.argv) > 3 else 10
    r = datagen.generate(shape)
    ids, inv = np.unique(r.item, return_inverse=True)
    item = (inv + 1).astype(np.int32)
    out = {"shape": shape, "users": r.n_users, "items": int(len(ids)), "ratings": r.nnz, "k": k, "iterations": iters}
    with NmfEngine(PPC, r.n_users, len(ids), k, iters) as eng:
        eng.set_ratings(r.user, item, r.score)
        runs = []
        for rep in range(3):
            eng.init_random(rep)
            t0 = time.perf_counter()
            eng.run()
            wall = time.perf_counter() - t0
            p = eng.profile()
            p["wall_s"] = wall
            runs.append(p)
        cl, cnt = eng.cluster_assignment()
    best = min(runs, key=lambda p: p["ms_per_iteration"])
    out.update(best)
    out["ms_index_first_run"] = runs[0]["ms_index"]
    out["join_gbs"] = best["join_bytes"] / (best["ms_per_iteration"] * 1e-3) / 1e9
    out["cluster_sizes_min_max"] = [int(cnt.min()), int(cnt.max())]
    print(json.dumps(out))


if __name__ == "__main__":
    main()
