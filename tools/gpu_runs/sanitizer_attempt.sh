set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_neighbours.py tests/test_jni_stub.py tests/test_gpu_cooc.py tests/test_gpu_nmf.py tests/test_seqfile.py -x -q -m gpu > gpurun_out/v3_pytest.log 2>&1
tail -15 gpurun_out/v3_pytest.log
# memcheck on the small parity cases (one tool per call)
cat > /tmp/san_small.py <<'PY'
import sys, numpy as np
sys.path.insert(0, ".")
import filmyou_core_b200 as fy
from filmyou_core_b200 import datagen
from oracle import rm2_oracle as orc
for shape, lam, n, mode in (("tiny", 0.1, 10, 0), ("small", 0.1, 100, 0), ("small", 0.9, 7, 1), ("tiny", 0.0, 10, 0)):
    r = datagen.generate(shape)
    with fy.Rm2Engine(lam=lam, number_of_items=r.n_items, top_n=n, score_mode=mode) as eng:
        eng.set_ratings(r.user, r.item, r.score); eng.set_clustering(r.cl_user, r.cl_cluster, r.cluster_size); eng.run()
        got = eng.results()
    want = orc.run(r.user, r.item, r.score, r.cl_user, r.cl_cluster, r.cluster_size, lam, r.n_items, n)
    assert np.array_equal(got["item"], want["item"]), shape
    for rank in range(3):
        with fy.Rm2Engine(lam=lam, number_of_items=r.n_items, top_n=n, shard_rank=rank, shard_count=3) as eng:
            eng.set_ratings(r.user, r.item, r.score); eng.set_clustering(r.cl_user, r.cl_cluster, r.cluster_size); eng.run()
r = datagen.generate("small")
with fy.Rm2Engine(number_of_items=r.n_items, top_n=20) as eng:
    eng.set_ratings(r.user, r.item, r.score)
    eng.cooc_counts(r.n_users + 1, r.n_items + 1); eng.cooc_topk(r.n_items + 1, 20)
    nb, cnt, n, _ = eng.knn_neighbours(r.n_users + 1, r.n_items + 1, 10)
    eng.run_neighbours(r.cl_user[:40], nb[r.cl_user[:40]])
from filmyou_core_b200.nmf import PPC, NmfEngine
ids, inv = np.unique(r.item, return_inverse=True)
with NmfEngine(PPC, r.n_users, len(ids), 6, 5) as e:
    e.set_ratings(r.user, (inv + 1).astype(np.int32), r.score); e.init_random(1); e.run()
print("sanitizer workload ok")
PY
cp /tmp/san_small.py gpurun_out/san_small.py
python gpurun_out/san_small.py > gpurun_out/san_plain.log 2>&1 && \
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 7 python gpurun_out/san_small.py > gpurun_out/san_memcheck.log 2>&1
echo "memcheck rc=$?"
tail -5 gpurun_out/san_memcheck.log
