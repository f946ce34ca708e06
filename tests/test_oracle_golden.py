"""CPU suite: the oracle against every golden vector the reference's tests hold for the RM2 path
(SURVEY.md 8c).  This is what pins the oracle; the GPU parity tests then compare against it."""
import numpy as np
import pytest

from oracle import rm2_oracle as orc
from filmyou_core_b200 import datagen

from conftest import by_user


def _run(r, lam, n_items, top_n, **kw):
    return orc.run(r.user, r.item, r.score, r.cl_user, r.cl_cluster, r.cluster_size, lam, n_items, top_n, **kw)


@pytest.mark.parametrize("mode", [orc.MODE_LITERAL, orc.MODE_LITERAL_FAST, orc.MODE_GRAM])
def test_recommendations_match_507_golden_triples(golden, golden_ratings, mode):
    # T/rm/TestHDFSRM2.java:39-75 + T/util/HadoopIntegrationTest.java:407-438: every emitted
    # (user,item) is in the golden map within `accuracy`, and the count matches.
    out = _run(golden_ratings, golden["lambda"], golden["numberOfItems"], golden["numberOfRecommendations"],
               mode=mode, threads=2)
    gold = {(int(u), int(i)): s for u, i, s in golden["recommendations"]}
    assert len(gold) == 507
    assert len(out["user"]) == 507
    for u, i, s in zip(out["user"], out["item"], out["score32"]):
        assert abs(gold[(int(u), int(i))] - float(s)) <= golden["accuracy"]
    worst = max(abs(gold[(int(u), int(i))] - float(s)) for u, i, s in zip(out["user"], out["item"], out["score64"]))
    assert worst < 3e-5      # goldens are float32 printed to 6 decimals


def test_literal_fast_is_bit_identical_and_gram_agrees(golden, golden_ratings):
    a = _run(golden_ratings, 0.5, 100, 1000, mode=orc.MODE_LITERAL, threads=1)
    b = _run(golden_ratings, 0.5, 100, 1000, mode=orc.MODE_LITERAL_FAST, threads=3)
    c = _run(golden_ratings, 0.5, 100, 1000, mode=orc.MODE_GRAM, threads=3)
    assert np.array_equal(a["item"], b["item"]) and np.array_equal(a["score64"], b["score64"])
    assert np.array_equal(a["item"], c["item"])
    assert np.max(np.abs(a["score64"] - c["score64"]) / np.abs(a["score64"])) < 1e-13


def test_blocked_literal_fast_equals_the_scalar_loop(monkeypatch):
    # LITERAL_FAST scores 4 G candidates side by side (AVX2 lanes, one sequential sum per lane); ORC_LITERAL_SCALAR=1
    # forces the one-pair-at-a-time loop.  Same bits, for every lambda (lambda = 0 reaches log 0 = -inf), ragged blocks
    # (candidate counts that are not multiples of 4) and the strided timing mode; and both equal the row-major literal loop.
    r = datagen.generate("small")
    for lam, top_n, kw in ((0.1, 100, {}), (0.0, 7, {}), (1.0, 100, {}), (0.5, 100, {"cand_stride": 7})):
        monkeypatch.delenv("ORC_LITERAL_SCALAR", raising=False)
        a = _run(r, lam, r.n_items, top_n, mode=orc.MODE_LITERAL_FAST, threads=3, **kw)
        monkeypatch.setenv("ORC_LITERAL_SCALAR", "1")
        b = _run(r, lam, r.n_items, top_n, mode=orc.MODE_LITERAL_FAST, threads=3, **kw)
        monkeypatch.delenv("ORC_LITERAL_SCALAR", raising=False)
        assert np.array_equal(a["item"], b["item"]) and a["score64"].tobytes() == b["score64"].tobytes(), (lam, kw)
        if not kw:
            c = _run(r, lam, r.n_items, top_n, mode=orc.MODE_LITERAL, threads=1)
            assert np.array_equal(a["item"], c["item"]) and a["score64"].tobytes() == c["score64"].tobytes(), lam


def test_statistics_match_golden(golden, golden_ratings):
    # T/rm/TestHDFSRM2.java:70-71: userSum and itemColl; RMTestData.java:426 totalSum
    r = golden_ratings
    us, isum, ip, tot = orc.stats(r.user, r.item, r.score, r.cl_user)
    assert tot == golden["totalSum"] == 7577.0
    assert np.array_equal(us, np.array(golden["userSum"]))
    assert np.array_equal(isum[1:], np.array(golden["itemSum"]))
    assert np.max(np.abs(ip[1:] - np.array(golden["itemColl"]))) <= 1e-18


def test_statistics_match_golden2(golden2):
    r = datagen.from_dense(golden2["A_item_by_user"], [c - 1 for c in golden2["clustering"]], golden2["clusteringCount"])
    us, isum, ip, tot = orc.stats(r.user, r.item, r.score, r.cl_user)
    assert tot == golden2["totalSum"] == 27.0
    assert np.array_equal(us, np.array(golden2["userSum"]))
    assert np.array_equal(isum[1:], np.array(golden2["itemSum"]))
    assert np.max(np.abs(ip[1:] - np.array(golden2["itemColl"]))) < 1e-9   # goldens carry 9 digits


def test_item_prob_divide_kat():
    # T/rm/TestItemProbInCollectionMapper.java:39-54: 10.0 / 2.0 = 5.0
    us, isum, ip, tot = orc.stats([1, 1], [3, 4], [2.0, 8.0], [1])
    assert tot == 10.0 and ip[3] == 0.2 and isum[4] / 2.0 == 4.0


def test_truncated_total_quirk():
    # (long) sum * OFFSET truncates each user's sum before scaling (DoubleSumAndCountReducer.java:41)
    us, isum, ip, tot = orc.stats([1, 1, 2], [1, 2, 1], [0.5, 1.0, 2.5], [1, 2])
    assert list(us) == [1.5, 2.5]
    assert tot == 3.0                       # 1 + 2, not 4.0
    assert ip[1] == 3.0 / 3.0 and ip[2] == 1.0 / 3.0


def test_golden_tie_is_broken_by_item_id(golden, golden_ratings):
    # RMTestData.java:357: user 24, items 38 and 43 carry the same double score
    out = _run(golden_ratings, 0.5, 100, 1000)
    items, scores = by_user(out)[24]
    k38, k43 = int(np.flatnonzero(items == 38)[0]), int(np.flatnonzero(items == 43)[0])
    assert scores[k38] == scores[k43]
    assert k43 == k38 + 1


def test_top_n_truncation_and_filter_users(golden_ratings):
    full = by_user(_run(golden_ratings, 0.5, 100, 1000))
    top5 = by_user(_run(golden_ratings, 0.5, 100, 5))
    for u in full:
        assert np.array_equal(top5[u][0], full[u][0][:5])
    flt = by_user(_run(golden_ratings, 0.5, 100, 5, filter_users=20))
    assert set(flt) == {u for u in full if u >= 20}


def test_only_users_subset(golden_ratings):
    full = by_user(_run(golden_ratings, 0.5, 100, 10))
    sub = by_user(_run(golden_ratings, 0.5, 100, 10, only_users=[3, 17, 24]))
    assert set(sub) == {3, 17, 24}
    for u in sub:
        assert np.array_equal(sub[u][0], full[u][0]) and np.array_equal(sub[u][1], full[u][1])


def test_error_codes(golden_ratings):
    r = golden_ratings
    with pytest.raises(orc.OracleError) as e:      # clusteringCount mismatch
        bad = r.cluster_size.copy(); bad[0] += 1
        orc.run(r.user, r.item, r.score, r.cl_user, r.cl_cluster, bad, 0.5, 100, 10)
    assert e.value.code == -4
    with pytest.raises(orc.OracleError) as e:      # duplicate rating
        orc.run(np.append(r.user, r.user[0]), np.append(r.item, r.item[0]), np.append(r.score, 1.0),
                r.cl_user, r.cl_cluster, r.cluster_size, 0.5, 100, 10)
    assert e.value.code == -3
    with pytest.raises(orc.OracleError) as e:      # rating of an unclustered user
        orc.run(np.append(r.user, 99), np.append(r.item, 1), np.append(r.score, 1.0),
                r.cl_user, r.cl_cluster, r.cluster_size, 0.5, 100, 10)
    assert e.value.code == -5
    with pytest.raises(orc.OracleError) as e:      # clustered user without ratings
        keep = r.user != 7
        orc.run(r.user[keep], r.item[keep], r.score[keep], r.cl_user, r.cl_cluster, r.cluster_size, 0.5, 100, 10)
    assert e.value.code == -2


def test_nonpositive_scores_are_ignored(golden_ratings):
    r = golden_ratings
    a = _run(r, 0.5, 100, 10)
    b = orc.run(np.append(r.user, [1, 2]), np.append(r.item, [100, 99]), np.append(r.score, [0.0, -3.0]),
                r.cl_user, r.cl_cluster, r.cluster_size, 0.5, 100, 10)
    assert np.array_equal(a["item"], b["item"]) and np.array_equal(a["score64"], b["score64"])


def test_synthetic_modes_agree_small():
    r = datagen.generate("tiny")
    a = _run(r, 0.1, r.n_items, 10, mode=orc.MODE_LITERAL, threads=1)
    b = _run(r, 0.1, r.n_items, 10, mode=orc.MODE_LITERAL_FAST)
    c = _run(r, 0.1, r.n_items, 10, mode=orc.MODE_GRAM)
    assert np.array_equal(a["score64"], b["score64"]) and np.array_equal(a["item"], b["item"])
    assert np.array_equal(a["item"], c["item"])


def test_cooccurrence_oracle_matches_numpy():
    r = datagen.generate("tiny")
    B = np.zeros((r.n_users + 1, r.n_items + 1), np.int64)
    B[r.user, r.item] = 1
    want = (B.T @ B).astype(np.int32)
    got = orc.cooccurrence(r.user, r.item, r.score, r.n_users + 1, r.n_items + 1)
    assert np.array_equal(got, want)


def test_generator_is_deterministic_and_exact():
    a, b = datagen.generate("small"), datagen.generate("small")
    assert a.sha256() == b.sha256()
    assert a.nnz == datagen.SHAPES["small"][2]
    assert len(np.unique(a.user.astype(np.int64) * 100000 + a.item)) == a.nnz
    assert a.cluster_size.min() >= 2 and a.cluster_size.sum() == a.n_users


def test_neighbour_list_mode_reproduces_the_goldens_when_the_list_is_the_cluster(golden, golden_ratings):
    # the oracle's neighbour mode (the `int[] neighbours` signature, AbstractRM2Reducer.java:321-323) is pinned by the
    # reference's own data: with N(u) = cluster(u) minus u it must give the 507 golden triples, and bit for bit what the
    # cluster mode gives (same K, same item universe, same ascending neighbour order)
    r = golden_ratings
    k = int(max(golden["clusteringCount"])) - 1
    nbr = -np.ones((r.n_users, k), np.int64)
    for q, u in enumerate(r.cl_user):
        mates = r.cl_user[(r.cl_cluster == r.cl_cluster[q]) & (r.cl_user != u)]
        nbr[q, :len(mates)] = mates
    out = orc.run_neighbours(r.user, r.item, r.score, r.cl_user, nbr, golden["lambda"], golden["numberOfItems"], 1000)
    gold = {(int(u), int(i)): s for u, i, s in golden["recommendations"]}
    assert len(out["user"]) == 507
    for u, i, s in zip(out["user"], out["item"], out["score32"]):
        assert abs(gold[(int(u), int(i))] - float(s)) <= golden["accuracy"]
    ref = orc.run(r.user, r.item, r.score, r.cl_user, r.cl_cluster, r.cluster_size, golden["lambda"], golden["numberOfItems"], 1000)
    key = lambda d: sorted(zip(d["user"].tolist(), d["item"].tolist(), d["score64"].tolist()))
    assert key(out) == key(ref)
