"""GPU suite: the CUDA path, called through the C ABI, against the oracle and the golden vectors.

Bar (BASELINE.json north_star): top-N item ids bit-exact and in the same order under the canonical
tie-break, scores within 1e-6 relative in the log domain; statistics bit-exact."""
import numpy as np
import pytest

import filmyou_core_b200 as fy
from filmyou_core_b200 import datagen
from oracle import rm2_oracle as orc

from conftest import assert_parity, assert_parity_near_ties, by_user

pytestmark = pytest.mark.gpu
REL = 1e-6          # north_star: "scores within 1e-6 relative in the log domain against Java doubles"


def gpu_run(r, lam, n_items, top_n, filter_users=0, **kw):
    with fy.Rm2Engine(lam=lam, number_of_items=n_items, top_n=top_n, filter_users=filter_users, **kw) as eng:
        eng.set_ratings(r.user, r.item, r.score)
        eng.set_clustering(r.cl_user, r.cl_cluster, r.cluster_size)
        eng.run()
        out = eng.results()
        out["stats"] = eng.stats()
        out["profile"] = eng.profile()
        out["users_scored"] = eng.users_scored()
    return out


def cpu_run(r, lam, n_items, top_n, **kw):
    return orc.run(r.user, r.item, r.score, r.cl_user, r.cl_cluster, r.cluster_size, lam, n_items, top_n, **kw)


def test_golden_507_triples(golden, golden_ratings):
    # the reference's own assertion (T/util/HadoopIntegrationTest.java:407-438) on the GPU output
    out = gpu_run(golden_ratings, golden["lambda"], golden["numberOfItems"], golden["numberOfRecommendations"])
    gold = {(int(u), int(i)): s for u, i, s in golden["recommendations"]}
    assert len(out["user"]) == 507
    for u, i, s in zip(out["user"], out["item"], out["score32"]):
        assert abs(gold[(int(u), int(i))] - float(s)) <= golden["accuracy"]
    us, ip, tot = out["stats"]
    assert tot == golden["totalSum"]
    assert np.array_equal(us, np.array(golden["userSum"]))
    assert np.max(np.abs(ip[1:] - np.array(golden["itemColl"]))) <= 1e-18


@pytest.mark.parametrize("top_n", [1, 5, 10, 1000])
def test_golden_vs_oracle_ids_and_scores(golden_ratings, top_n):
    got = gpu_run(golden_ratings, 0.5, 100, top_n)
    want = cpu_run(golden_ratings, 0.5, 100, top_n)
    assert_parity(got, want, REL, "golden N=%d" % top_n)
    assert np.array_equal(got["cluster"], want["cluster"])
    assert np.array_equal(got["score32"], got["score64"].astype(np.float32))


def test_golden_exact_tie(golden_ratings):
    items, scores = by_user(gpu_run(golden_ratings, 0.5, 100, 1000))[24]
    k38, k43 = int(np.flatnonzero(items == 38)[0]), int(np.flatnonzero(items == 43)[0])
    assert scores[k38] == scores[k43] and k43 == k38 + 1


@pytest.mark.parametrize("score_mode", [0, 1])
@pytest.mark.parametrize("shape,lam,top_n", [("tiny", 0.1, 10), ("small", 0.1, 100), ("small", 0.9, 7),
                                              ("ml-100k", 0.1, 100)])
def test_synthetic_vs_oracle(shape, lam, top_n, score_mode):
    # score_mode 0 = hi-word stream + exact fp64 re-score of the candidate set; 1 = fp64 stream
    r = datagen.generate(shape)
    got = gpu_run(r, lam, r.n_items, top_n, score_mode=score_mode)
    assert got["profile"]["bytes_per_term"] == (4.0 if score_mode == 0 else 8.0)
    want = cpu_run(r, lam, r.n_items, top_n)
    worst = assert_parity(got, want, REL, shape)
    assert worst < 1e-9          # the engine is fp64 end to end; 1e-6 is the contract, this is the margin
    us, ip, tot = got["stats"]
    ous, _, oip, otot = orc.stats(r.user, r.item, r.score, r.cl_user)
    assert tot == otot and np.array_equal(us, ous) and np.array_equal(ip, oip[:len(ip)])
    assert got["users_scored"] == want["users_scored"]


def test_ml1m_sampled_users_vs_oracle():
    # config[1] shape; the oracle scores a seeded sample of users (the literal loop is ~3e12 flops)
    r = datagen.generate("ml-1m")
    got = gpu_run(r, 0.1, r.n_items, 100)
    sample = np.random.default_rng(5).choice(r.cl_user, size=48, replace=False)
    want = cpu_run(r, 0.1, r.n_items, 100, only_users=sample)
    g = by_user(got)
    sub = {k: np.concatenate([np.full(len(g[int(u)][0]), int(u)) if k == "user" else
                              (g[int(u)][0] if k == "item" else g[int(u)][1])
                              for u in by_user(want)]) for k in ("user", "item", "score64")}
    assert_parity(sub, want, REL, "ml-1m sample")
    assert got["users_scored"] == r.n_users


def test_candidate_overflow_falls_back_to_the_exact_stream():
    # 600 items that nobody else rated identically: every user sees hundreds of candidates with
    # (near-)identical scores, far more than the candidate capacity for N=5 -> the run must notice
    # and redo itself in exact mode, with the same answer as score_mode=1
    rng = np.random.default_rng(3)
    users = np.arange(1, 41)
    user, item, score = [], [], []
    for u in users:                               # everybody rates items 1..5, nobody rates 6..605 twice
        for i in range(1, 6):
            user.append(u); item.append(i); score.append(float(rng.integers(1, 6)))
    for i in range(6, 606):                       # each cold item rated once, all by user 1, same score
        user.append(1); item.append(i); score.append(3.0)
    r = datagen.Ratings("ovf", 40, 605, np.array(user, np.int32), np.array(item, np.int32), np.array(score, np.float32),
                        users.astype(np.int32), np.zeros(40, np.int32), np.array([40], np.int32), 0)
    a = gpu_run(r, 0.1, 605, 5, score_mode=0)
    b = gpu_run(r, 0.1, 605, 5, score_mode=1)
    assert a["profile"]["exact_rerun"] == 1 and b["profile"]["exact_rerun"] == 0
    assert all(np.array_equal(a[k], b[k]) for k in ("user", "item", "score64"))
    assert_parity(a, cpu_run(r, 0.1, 605, 5), REL, "overflow")


def _subset(got, users):
    g = by_user(got)
    return {k: np.concatenate([np.full(len(g[int(u)][0]), int(u)) if k == "user" else
                               (g[int(u)][0] if k == "item" else g[int(u)][1]) for u in users])
            for k in ("user", "item", "score64")}


def _full_cluster_and_heaviest(r, got, what, n_heavy):
    """Parity where the numbers are quoted: EVERY user of the cluster that holds the most active user of the workload
    against the oracle's GRAM mode (G - self algebra in fp64, ~1e12 multiply-adds: the literal loop would be ~1e14), and
    the n_heavy most active users of the whole workload against MODE_LITERAL_FAST (the reducer's own loop nest and order).
    These are the users where the candidate margin (eps_u = n_u * 2.5e-7), the candidate cap, the fp32 exponent-peel range
    and the most-active-first processing order matter."""
    n_u = np.bincount(r.user, minlength=r.n_users + 1)
    order = r.cl_user[np.argsort(-n_u[r.cl_user], kind="stable")]
    heavy = order[:n_heavy].astype(np.int32)
    cl_of = np.zeros(r.n_users + 1, np.int64); cl_of[r.cl_user] = r.cl_cluster
    c = int(cl_of[heavy[0]])
    members = r.cl_user[r.cl_cluster == c].astype(np.int32)
    want = cpu_run(r, 0.1, r.n_items, 100, only_users=members, mode=orc.MODE_GRAM)
    assert want["users_scored"] == len(members)
    worst, swapped = assert_parity_near_ties(_subset(got, list(by_user(want))), want, REL, 1e-12, what + " cluster %d (GRAM)" % c)
    assert worst < 1e-9 and swapped <= max(2, len(members) // 200), (worst, swapped)
    want = cpu_run(r, 0.1, r.n_items, 100, only_users=heavy, mode=orc.MODE_LITERAL_FAST)
    worst_h, swapped_h = assert_parity_near_ties(_subset(got, list(by_user(want))), want, REL, 1e-12, what + " heaviest users (literal)")
    assert worst_h < 1e-9
    return {"cluster": c, "cluster_users": int(len(members)), "max_n_u": int(n_u[heavy[0]]), "worst_rel": max(worst, worst_h),
            "near_tie_swaps": swapped + swapped_h}


def test_ml20m_full_cluster_and_heaviest_users_vs_oracle():
    # BASELINE.json headline shape (138 493 x 26 744, 20 M half-star ratings, 50 clusters), the whole job on the GPU
    r = datagen.generate("ml-20m")
    got = gpu_run(r, 0.1, r.n_items, 100)
    assert got["users_scored"] == r.n_users and len(got["user"]) == r.n_users * 100
    assert got["profile"]["exact_rerun"] == 0
    info = _full_cluster_and_heaviest(r, got, "ml-20m", 4)
    assert info["cluster_users"] > 2000 and info["max_n_u"] > 5000
    print("ml-20m parity:", info)
    # light and medium users spread over other clusters, the reducer's literal loop
    n_u = np.bincount(r.user, minlength=r.n_users + 1)
    rng = np.random.default_rng(20)
    light = rng.choice(np.flatnonzero((n_u >= 20) & (n_u <= 30)), size=6, replace=False)
    medium = rng.choice(np.flatnonzero((n_u >= 100) & (n_u <= 140)), size=2, replace=False)
    want = cpu_run(r, 0.1, r.n_items, 100, only_users=np.concatenate([light, medium]).astype(np.int32))
    assert assert_parity(_subset(got, list(by_user(want))), want, REL, "ml-20m sample") < 1e-9
    # size-independent properties over ALL users
    items = got["item"].reshape(r.n_users, 100); scores = got["score64"].reshape(r.n_users, 100)
    assert np.all(np.diff(scores, axis=1) <= 0)
    assert np.all(np.sort(items, axis=1)[:, 1:] != np.sort(items, axis=1)[:, :-1])      # distinct per user
    us, ip, tot = got["stats"]
    ous, _, oip, otot = orc.stats(r.user, r.item, r.score, r.cl_user)
    assert tot == otot and np.array_equal(us, ous) and np.array_equal(ip, oip[:len(ip)])


def test_netflix_shape_full_cluster_and_properties():
    # BASELINE.json configs[4] shape (480 189 x 17 770, 100 M ratings, 50 clusters) on ONE GPU: the whole job; every user of
    # one full cluster (~9 600 users) and the heaviest users against the oracle; size-independent properties over all users
    r = datagen.generate("netflix")
    got = gpu_run(r, 0.1, r.n_items, 100)
    assert got["users_scored"] == r.n_users and len(got["user"]) == r.n_users * 100
    info = _full_cluster_and_heaviest(r, got, "netflix", 1)    # the heaviest user twice: GRAM with its cluster, then literal
    print("netflix parity:", info)
    items = got["item"].reshape(r.n_users, 100); scores = got["score64"].reshape(r.n_users, 100)
    assert np.all(np.diff(scores, axis=1) <= 0)
    srt = np.sort(items, axis=1)
    assert np.all(srt[:, 1:] != srt[:, :-1])                           # distinct per user
    rated = set(zip(r.user[:200000].tolist(), r.item[:200000].tolist()))
    users = got["user"].reshape(r.n_users, 100)[:, 0]
    pos = {int(u): k for k, u in enumerate(users)}
    assert not any(i in items[pos[u]] for u, i in list(rated)[:20000])   # never recommends a rated item
    us, ip, tot = got["stats"]
    ous, _, oip, otot = orc.stats(r.user, r.item, r.score, r.cl_user)
    assert tot == otot and np.array_equal(us, ous) and np.array_equal(ip, oip[:len(ip)])


def test_compact_result_stream_rebuilds_the_packed_triples():
    # the 12-byte read-back (item, score64) + one (user, cluster, count) record per row = the five packed arrays
    r = datagen.generate("small")
    with fy.Rm2Engine(lam=0.1, number_of_items=r.n_items, top_n=600) as eng:       # N > items: ragged row counts
        eng.set_ratings(r.user, r.item, r.score)
        eng.set_clustering(r.cl_user, r.cl_cluster, r.cluster_size)
        eng.run()
        full = eng.results()
        comp = eng.results_compact()
        rebuilt = fy.Rm2Engine.expand_compact(comp)
    assert len(set(comp["row_count"].tolist())) > 1
    for k in ("user", "item", "score64", "score32", "cluster"):
        assert np.array_equal(full[k], rebuilt[k]), k


def test_n_gpus_context_equals_one_gpu():
    # fy_rm2_params.n_gpus: ONE context, one host process, several devices -- the result a JVM host gets from one native call
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    r = datagen.generate("ml-100k")
    ref = gpu_run(r, 0.1, r.n_items, 100)
    n = min(4, torch.cuda.device_count())
    got = gpu_run(r, 0.1, r.n_items, 100, n_gpus=n)
    for k in ("user", "item", "score64", "score32", "cluster"):
        assert np.array_equal(ref[k], got[k]), k
    assert got["users_scored"] == ref["users_scored"]
    assert got["stats"][2] == ref["stats"][2] and np.array_equal(got["stats"][0], ref["stats"][0])
    with fy.Rm2Engine(lam=0.1, number_of_items=r.n_items, top_n=100, n_gpus=n) as eng:
        eng.set_ratings(r.user, r.item, r.score)
        eng.set_clustering(r.cl_user, r.cl_cluster, r.cluster_size)
        eng.run()
        rebuilt = fy.Rm2Engine.expand_compact(eng.results_compact())
    assert all(np.array_equal(ref[k], rebuilt[k]) for k in ("user", "item", "score64", "cluster"))


def test_multi_process_exchange_digest_equals_one_gpu(tmp_path):
    # one process per GPU, the NCCL exchange inside the library (fy_rm2_comm_init): rank 0's digest of the whole job's
    # (user, item, score64 bits) must equal the 1-GPU digest -- the check SCALE relies on
    import json, os, subprocess, sys, torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    def bench(n):
        cmd = [sys.executable, os.path.join(root, "bench.py")] if n == 1 else \
              [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(n), "--master-addr", "127.0.0.1",
               "--master-port", "29653", os.path.join(root, "bench.py")]
        out = subprocess.check_output(cmd + ["--gpus", str(n), "--steps", "1", "--warmup", "3", "--workload", "ml-1m",
                                             "--no-cpu-baseline", "--no-secondary"], text=True, cwd=root)
        return json.loads(out.strip().splitlines()[-1])
    one, two = bench(1), bench(2)
    assert one["result_digest"]["triples"] == 6040 * 100
    assert two["result_digest"] == one["result_digest"]
    assert two["n_gpus"] == 2 and two["users_scored"] == one["users_scored"] == 6040
    assert two["e2e"]["d2h_bytes_per_step"] == one["e2e"]["d2h_bytes_per_step"]


def test_more_than_4096_recommendations_per_user():
    # min(numberOfRecommendations, items of the cluster) beyond the shared-memory select bound: the engine
    # switches to a whole-row stable segmented sort; everything a user has not rated is emitted, in order
    r = datagen.generate("big-n", n_users=100, n_items=5200, nnz=26000, n_clusters=1, seed=5)
    got = gpu_run(r, 0.1, r.n_items, 6000)
    want = cpu_run(r, 0.1, r.n_items, 6000)
    assert max(len(v[0]) for v in by_user(want).values()) > 4096
    assert_parity(got, want, REL, "N > 4096")


def test_lambda_edge_cases():
    r = datagen.generate("tiny")
    for lam in (0.0, 1.0):
        got = gpu_run(r, lam, r.n_items, 10)
        want = cpu_run(r, lam, r.n_items, 10)
        assert_parity(got, want, REL, "lambda=%g" % lam)
    # lambda = 0 produces log(0) = -inf scores exactly like Math.log does
    assert np.isneginf(gpu_run(r, 0.0, r.n_items, 1000)["score64"]).any()


def test_singleton_cluster_and_user_with_everything_rated():
    # cluster 0 = one user (no candidates -> skipped, AbstractRM2Reducer.java:210-213);
    # cluster 1: user 2 rated every item of the cluster (skipped), users 3,4 scored
    user = np.array([1, 1, 2, 2, 2, 3, 4, 4], np.int32)
    item = np.array([1, 2, 1, 2, 3, 1, 2, 3], np.int32)
    score = np.array([5, 3, 4, 2, 1, 3, 5, 4], np.float32)
    r = datagen.Ratings("edge", 4, 3, user, item, score, np.array([1, 2, 3, 4], np.int32),
                        np.array([0, 1, 1, 1], np.int32), np.array([1, 3], np.int32), 0)
    got = gpu_run(r, 0.3, 3, 10)
    want = cpu_run(r, 0.3, 3, 10)
    assert set(by_user(want)) == {3, 4}
    assert_parity(got, want, REL, "edge")


def test_unordered_ids_sparse_ids_and_nonpositive_scores():
    r = datagen.generate("tiny")
    # remap ids to sparse, non-monotone values; add ignored (<= 0) ratings
    umap = np.random.default_rng(1).permutation(5000)[:r.n_users + 1] + 10
    imap = np.random.default_rng(2).permutation(3000)[:r.n_items + 1] + 1
    r2 = datagen.Ratings("remap", r.n_users, r.n_items, umap[r.user].astype(np.int32), imap[r.item].astype(np.int32),
                         r.score, umap[r.cl_user].astype(np.int32), r.cl_cluster, r.cluster_size, 0)
    u = np.append(r2.user, r2.user[:3]); i = np.append(r2.item, [2999, 2998, 2997]); s = np.append(r2.score, [0, -1, 0]).astype(np.float32)
    r3 = datagen.Ratings("remap", r.n_users, r.n_items, u, i, s, r2.cl_user, r2.cl_cluster, r2.cluster_size, 0)
    got = gpu_run(r3, 0.2, 3000, 12)
    want = cpu_run(r2, 0.2, 3000, 12)
    assert_parity(got, want, REL, "remap")


def test_non_dyadic_scores_take_the_sequential_statistics_path():
    # scores like 3.7f are not multiples of 2^-16: sums depend on the order, so the engine must add
    # them in the oracle's (= canonical) order; statistics stay bit-exact
    r = datagen.generate("small")
    sc = (r.score * np.float32(0.74) + np.float32(0.013)).astype(np.float32)
    r2 = datagen.Ratings("nd", r.n_users, r.n_items, r.user, r.item, sc, r.cl_user, r.cl_cluster, r.cluster_size, 0)
    got = gpu_run(r2, 0.1, r.n_items, 15)
    want = cpu_run(r2, 0.1, r.n_items, 15)
    assert_parity(got, want, REL, "non-dyadic")
    us, ip, tot = got["stats"]
    ous, _, oip, otot = orc.stats(r2.user, r2.item, r2.score, r2.cl_user)
    assert tot == otot and np.array_equal(us, ous)


def test_filter_users(golden_ratings):
    got = gpu_run(golden_ratings, 0.5, 100, 5, filter_users=20)
    want = cpu_run(golden_ratings, 0.5, 100, 5, filter_users=20)
    assert_parity(got, want, REL, "filterUsers")


def test_error_codes(golden_ratings):
    r = golden_ratings

    def expect(code, user, item, score, cl_user=r.cl_user, cl_cluster=r.cl_cluster, csize=r.cluster_size):
        with fy.Rm2Engine(lam=0.5, number_of_items=100, top_n=10) as eng:
            with pytest.raises(fy.Rm2Error) as e:
                eng.set_ratings(user, item, score)
                eng.set_clustering(cl_user, cl_cluster, csize)
                eng.run()
            assert e.value.code == code, str(e.value)

    bad = r.cluster_size.copy(); bad[0] += 1
    expect(-4, r.user, r.item, r.score, csize=bad)
    expect(-3, np.append(r.user, r.user[0]), np.append(r.item, r.item[0]), np.append(r.score, 1.0))
    expect(-5, np.append(r.user, 99), np.append(r.item, 1), np.append(r.score, 1.0))
    keep = r.user != 7
    expect(-2, r.user[keep], r.item[keep], r.score[keep])
    with fy.Rm2Engine(lam=0.5, number_of_items=100, top_n=10) as eng:
        with pytest.raises(fy.Rm2Error) as e:
            eng.run()
        assert e.value.code == -8


def test_fine_seam_score_group_matches_coarse(golden, golden_ratings):
    # one reduce() group per cluster split, as TestHDFSRM2 exercises (clusterSplit=5, splitSize=3)
    r = golden_ratings
    coarse = by_user(gpu_run(r, 0.5, 100, 1000))
    us, ip, tot = gpu_run(r, 0.5, 100, 1)["stats"]
    seen = {}
    with fy.Rm2Engine(lam=0.5, number_of_items=100, top_n=1000) as eng:
        for c, size in enumerate(golden["clusteringCount"]):
            members = r.cl_user[r.cl_cluster == c]
            n_splits = int(np.ceil(size / golden["splitSize"])) if size >= golden["clusterSplit"] else 1
            sel = np.isin(r.user, members)
            for split in range(n_splits):
                eng.score_group(c, split, n_splits, members, us[members - 1], r.user[sel], r.item[sel], r.score[sel], ip)
                out = eng.results()
                assert (out["cluster"] == c).all()
                for u, (it, sc) in by_user(out).items():
                    assert u % n_splits == split and u not in seen
                    seen[u] = (it, sc)
    assert set(seen) == set(coarse)
    for u in coarse:
        assert np.array_equal(seen[u][0], coarse[u][0]) and np.array_equal(seen[u][1], coarse[u][1])


@pytest.mark.parametrize("dyadic", [True, False])
def test_sharded_contexts_cover_all_users_once(dyadic):
    # dyadic scores -> sharded index (global statistics by exact atomics, local sort only);
    # non-dyadic -> every rank keeps the replicated, order-preserving index
    r = datagen.generate("small")
    if not dyadic:
        sc = (r.score * np.float32(0.74) + np.float32(0.013)).astype(np.float32)
        r = datagen.Ratings("nd", r.n_users, r.n_items, r.user, r.item, sc, r.cl_user, r.cl_cluster, r.cluster_size, 0)
    ref = gpu_run(r, 0.1, r.n_items, 20)
    full = by_user(ref)
    merged = {}
    for rank in range(3):
        out = gpu_run(r, 0.1, r.n_items, 20, shard_rank=rank, shard_count=3)
        assert out["stats"][2] == ref["stats"][2] and np.array_equal(out["stats"][0], ref["stats"][0]) \
            and np.array_equal(out["stats"][1], ref["stats"][1])            # statistics are global on every rank
        part = by_user(out)
        assert not (set(part) & set(merged))
        merged.update(part)
    assert set(merged) == set(full)
    for u in full:
        assert np.array_equal(merged[u][0], full[u][0]) and np.array_equal(merged[u][1], full[u][1])
    # more shards than clusters, including empty shards
    merged = {}
    for rank in range(16):
        merged.update(by_user(gpu_run(r, 0.1, r.n_items, 20, shard_rank=rank, shard_count=16)))
    assert set(merged) == set(full) and all(np.array_equal(merged[u][0], full[u][0]) for u in full)


@pytest.mark.parametrize("world", [3, 8])
def test_shard_bounds_follow_the_documented_rule(world):
    # the partition computed on the device (k_shard_bounds) = sharding.plan_shards, the Python statement of the rule:
    # equal cost where a rank pays its users' n_u * I_c plus 0.55 x the work of every cluster it touches
    from filmyou_core_b200 import sharding
    r = datagen.generate("ml-100k")
    order = np.lexsort((r.cl_user, r.cl_cluster))
    n_u = np.bincount(r.user, minlength=r.n_users + 1)
    cl_of = np.zeros(r.n_users + 1, np.int64); cl_of[r.cl_user] = r.cl_cluster
    i_c = np.array([len(np.unique(r.item[cl_of[r.user] == c])) for c in range(r.n_clusters)])
    work = n_u[r.cl_user[order]] * i_c[r.cl_cluster[order]]
    cs = np.concatenate([[0], np.cumsum(r.cluster_size)])
    want = sharding.plan_shards(work, world, cluster_start=cs)
    for rank in (0, world - 1):
        with fy.Rm2Engine(lam=0.1, number_of_items=r.n_items, top_n=10, shard_rank=rank, shard_count=world) as eng:
            eng.set_ratings(r.user, r.item, r.score)
            eng.set_clustering(r.cl_user, r.cl_cluster, r.cluster_size)
            eng.run()
            assert eng.shard_bounds().tolist() == want
            assert eng.users_scored() == want[rank + 1] - want[rank]


def test_size_independent_properties_ml100k():
    r = datagen.generate("ml-100k")
    out = gpu_run(r, 0.1, r.n_items, 100)
    g = by_user(out)
    assert len(g) == r.n_users
    rated = {}
    for u, i in zip(r.user, r.item):
        rated.setdefault(int(u), set()).add(int(i))
    for u, (items, scores) in g.items():
        assert len(items) == 100 and len(set(items.tolist())) == 100            # distinct
        assert not (set(items.tolist()) & rated[u])                             # never a rated item
        assert np.all(np.diff(scores) <= 0)                                     # descending
        tie = np.flatnonzero(np.diff(scores) == 0)
        assert np.all(items[tie] < items[tie + 1])                              # ties by ascending item id
    # idempotence: a second run on the same context gives the same bits
    with fy.Rm2Engine(lam=0.1, number_of_items=r.n_items, top_n=100) as eng:
        eng.set_ratings(r.user, r.item, r.score)
        eng.set_clustering(r.cl_user, r.cl_cluster, r.cluster_size)
        eng.run(); a = eng.results()
        eng.run(); b = eng.results()
    assert all(np.array_equal(a[k], b[k]) for k in a)
    assert np.array_equal(a["score64"], out["score64"])


def test_reference_style_job_and_sinks(golden, golden_ratings):
    # reads like T/rm/TestHDFSRM2.java:39-75: build conf, run RM2Job, compare userSum / itemColl / output
    from filmyou_core_b200.rm2_job import RM2Job, HDFSSink, CassandraSink
    r = golden_ratings
    conf = {"numberOfItems": golden["numberOfItems"], "numberOfClusters": golden["numberOfClusters"],
            "numberOfRecommendations": 1000, "lambda": "0.5", "clusterSplit": 5, "splitSize": 3}
    job = RM2Job(conf)
    sink = HDFSSink()
    job.run(r.user, r.item, r.score, r.cl_user, r.cl_cluster, golden["clusteringCount"], sink=sink)
    assert np.array_equal(job.userSum, np.array(golden["userSum"]))
    assert np.max(np.abs(job.itemColl[1:] - np.array(golden["itemColl"]))) <= 1e-18
    gold = {(int(u), int(i)): s for u, i, s in golden["recommendations"]}
    assert len(sink.records) == len(gold)
    for key, val in sink.records:
        assert abs(gold[key] - float(val)) <= golden["accuracy"]
    cs = CassandraSink()
    job.run(r.user, r.item, r.score, r.cl_user, r.cl_cluster, golden["clusteringCount"], sink=cs)
    # CLUSTERING ORDER BY (relevance DESC, item ASC) inside a user  (T/util/CassandraUtils.java:144-147)
    for a, b in zip(cs.rows, cs.rows[1:]):
        if a[0] == b[0]:
            assert (a[1] > b[1]) or (a[1] == b[1] and a[2] < b[2]) or (np.float32(a[1]) == np.float32(b[1]))
    job.close()
    with pytest.raises(RuntimeError):
        bad = RM2Job(conf)
        bad.run(r.user, r.item, r.score, r.cl_user, r.cl_cluster, [8, 7, 7, 7, 1])


def test_cpp_host_mirror(tmp_path):
    import os, subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = str(tmp_path / "cpp_host_smoke")
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-I", os.path.join(root, "include"),
                           os.path.join(root, "tests", "cpp_host_smoke.cpp"), "-o", exe,
                           "-L", os.path.join(root, "filmyou_core_b200"), "-lfilmyou_rm2",
                           "-Wl,-rpath," + os.path.join(root, "filmyou_core_b200")])
    out = subprocess.check_output([exe], text=True).splitlines()
    assert out[0] == "totalSum 27.0"                                  # RMTestData2.java:45
    assert [l for l in out if l.startswith("userSum")] == ["userSum %.1f" % v for v in (9, 3, 3, 5, 7)]
    r = datagen.from_dense([[5, 0, 0, 1, 2], [4, 3, 1, 0, 0], [0, 0, 2, 4, 5]], [0, 0, 1, 1, 1], [2, 3])
    want = cpu_run(r, 0.5, 3, 10)
    recs = [l.split() for l in out if l.startswith("rec")]
    assert [(int(a[1]), int(a[2])) for a in recs] == list(zip(want["user"].tolist(), want["item"].tolist()))
    assert np.allclose([float(a[3]) for a in recs], want["score32"], rtol=0, atol=1e-5)
    # include/filmyou_nmf_job.hpp: one PPC iteration + arg-max on the same toy, against the NMF oracle
    from oracle import nmf_oracle as norc
    H0 = np.array([0.2, 0.8, 0.6, 0.4, 0.5, 0.5, 0.9, 0.1, 0.3, 0.7]).reshape(5, 2)
    W0 = np.array([0.7, 0.3, 0.4, 0.6, 0.1, 0.9]).reshape(3, 2)
    Ho, _ = norc.run(norc.PPC, r.user, r.item, r.score, H0, W0, 1, combine_len=1024, split_rows=256)
    H = np.array([float(l.split()[1]) for l in out if l.startswith("H ")]).reshape(5, 2)
    assert np.array_equal(H, Ho)
    cl, cnt = norc.cluster_assign(Ho)
    assert [int(l.split()[1]) for l in out if l.startswith("cluster")] == cl.tolist()
    assert [l for l in out if l.startswith("count")] == ["count %d %d" % tuple(cnt)]


def test_file_level_job_reads_and_writes_sequence_files(golden, golden_ratings, tmp_path):
    # T/rm/TestHDFSRM2.java:39-75 at the file level: fixtures written the way DataInitialization writes them
    # (A/data, clustering/data with keys from 1, clusteringCount/data with keys from 0), job run through
    # fy_rm2_run_files, outputs read back from rm2/output, rm2/userSum and the MapFile rm2/itemColl
    from filmyou_core_b200 import seqfile
    r = golden_ratings
    base = tmp_path / "integrationTest"
    for d in ("A", "clustering", "clusteringCount"):
        (base / d).mkdir(parents=True)
    seqfile.write_intpair_float(str(base / "A" / "data"), r.user, r.item, r.score)
    seqfile.write_int_int(str(base / "clustering" / "data"), np.arange(1, 31), golden["clustering"])
    seqfile.write_int_int(str(base / "clusteringCount" / "data"), np.arange(5), golden["clusteringCount"])
    with fy.Rm2Engine(lam=0.5, number_of_items=100, top_n=1000) as eng:
        eng.run_files(str(base / "A"), str(base / "clustering"), str(base / "clusteringCount"), golden["numberOfClusters"],
                      str(base / "rm2" / "output"), str(base / "rm2"))
    u, i, s = seqfile.read_intpair_float(str(base / "rm2" / "output"))
    gold = {(int(a), int(b)): c for a, b, c in golden["recommendations"]}
    assert len(u) == 507
    for a, b, c in zip(u, i, s):
        assert abs(gold[(int(a), int(b))] - float(c)) <= golden["accuracy"]          # compareIntPairFloatData
    ku, vu = seqfile.read_int_double(str(base / "rm2" / "userSum"))
    assert ku.tolist() == list(range(1, 31)) and np.array_equal(vu, np.array(golden["userSum"]))   # compareIntDoubleData
    ki, vi = seqfile.read_int_double(str(base / "rm2" / "itemColl"))
    assert ki.tolist() == list(range(1, 101)) and np.max(np.abs(vi - np.array(golden["itemColl"]))) <= 1e-18
