#!/usr/bin/env python3
"""Config-3 measurement: the int8 tcgen05 co-occurrence GEMM at ML-1M and ML-20M shape, next to an
int8 yardstick (torch._int_mm = cuBLASLt, library) on the same GPU."""
import sys, os, json, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import filmyou_core_b200 as fy
from filmyou_core_b200 import datagen

out = {}
a = torch.randint(-3, 3, (8192, 8192), dtype=torch.int8, device="cuda"); b = torch.randint(-3, 3, (8192, 8192), dtype=torch.int8, device="cuda")
for _ in range(3): torch._int_mm(a, b)
torch.cuda.synchronize(); e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
best = 1e9
for _ in range(10):
    e0.record(); torch._int_mm(a, b); e1.record(); torch.cuda.synchronize(); best = min(best, e0.elapsed_time(e1))
out["int8_yardstick_tops_cublaslt_8192"] = 2 * 8192**3 / (best * 1e-3) / 1e12
for shape in [a for a in sys.argv[1:] if not a.startswith("--")] or ["ml-1m"]:
    r = datagen.generate(shape)
    n_u, n_i = r.n_users + 1, r.n_items + 1
    with fy.Rm2Engine(number_of_items=r.n_items) as eng:
        eng.set_ratings(r.user, r.item, r.score)
        ms = []
        for _ in range(4):
            _, m = eng.cooc_counts(n_u, n_i, want_counts=False); ms.append(m)
    ops = 2.0 * n_i * n_i * n_u
    mt, nt = -(-n_i // 128), -(-n_i // 256)
    tiles = sum(1 for m in range(mt) for n in range(nt) if (n + 1) * 256 > m * 128)     # upper-triangle tiles only
    pad_ops = 2.0 * tiles * 128 * 256 * (-(-n_u // 128) * 128)
    out[shape] = {"ms": ms, "algorithmic_int8_tops": ops / (min(ms) * 1e-3) / 1e12, "executed_int8_tops": pad_ops / (min(ms) * 1e-3) / 1e12,
                  "tiles_computed": tiles, "tiles_total": mt * nt}
if "--knn" in sys.argv:
    r = datagen.generate("ml-20m")
    n_u, n_i = r.n_users + 1, r.n_items + 1
    with fy.Rm2Engine(number_of_items=r.n_items) as eng:
        eng.set_ratings(r.user, r.item, r.score)
        t = time.time(); nb, cnt, n, ms = eng.knn_neighbours(n_u, n_i, 100); wall = time.time() - t
        t = time.time(); nb, cnt, n, ms = eng.knn_neighbours(n_u, n_i, 100); wall = time.time() - t
    out["knn_ml-20m"] = {"users": n_u, "k": 100, "ms_gemm": ms, "wall_s": wall, "executed_int8_tops": 2.0 * n_u * n_u * n_i / (ms * 1e-3) / 1e12,
                         "users_with_100_neighbours": int((n == 100).sum())}
print(json.dumps(out))
