"""GPU suite, f3 (north_star part 1): RM2 scoring over EXPLICIT neighbour lists -- the `int[] neighbours`
argument of buildRecommendations (M/rm/AbstractRM2Reducer.java:321-323,342-346) -- fed by the engine's own kNN provider
(fy_knn_neighbours).  The reference only ever passes "the cluster minus u" (:215-216); with that list the call must
reproduce the cluster job (pinned by the goldens), beyond it the check is the oracle's neighbour mode
(oracle/rm2_oracle.py::run_neighbours, itself pinned by the goldens in tests/test_oracle_golden.py)."""
import numpy as np
import pytest

import filmyou_core_b200 as fy
from filmyou_core_b200 import datagen
from oracle import rm2_oracle as orc

from conftest import assert_parity, by_user

pytestmark = pytest.mark.gpu


def _gpu(r, users, nbr, lam, top_n):
    with fy.Rm2Engine(lam=lam, number_of_items=r.n_items, top_n=top_n) as eng:
        eng.set_ratings(r.user, r.item, r.score)
        eng.run_neighbours(users, nbr)
        out = eng.results()
        comp = fy.Rm2Engine.expand_compact(eng.results_compact())
        out["users_scored"] = eng.users_scored()
    assert all(np.array_equal(out[k], comp[k]) for k in ("user", "item", "score64"))
    return out


def test_cluster_minus_self_reproduces_the_cluster_job(golden, golden_ratings):
    r = golden_ratings
    k = int(max(golden["clusteringCount"])) - 1
    nbr = -np.ones((r.n_users, k), np.int32)
    for q, u in enumerate(r.cl_user):
        mates = r.cl_user[(r.cl_cluster == r.cl_cluster[q]) & (r.cl_user != u)]
        nbr[q, :len(mates)] = mates[::-1]                    # any order, any padding position
    got = _gpu(r, r.cl_user, nbr, 0.5, 1000)
    gold = {(int(u), int(i)): s for u, i, s in golden["recommendations"]}
    assert len(got["user"]) == 507
    for u, i, s in zip(got["user"], got["item"], got["score32"]):
        assert abs(gold[(int(u), int(i))] - float(s)) <= golden["accuracy"]       # the reference's own assertion
    with fy.Rm2Engine(lam=0.5, number_of_items=100, top_n=1000) as eng:
        eng.set_ratings(r.user, r.item, r.score)
        eng.set_clustering(r.cl_user, r.cl_cluster, r.cluster_size)
        eng.run()
        ref = by_user(eng.results())
    g = by_user(got)
    assert set(g) == set(ref)
    for u in ref:
        assert np.array_equal(g[u][0], ref[u][0]) and np.array_equal(g[u][1], ref[u][1])      # same ids, same score bits


@pytest.mark.parametrize("shape,k,n_sample", [("tiny", 7, None), ("ml-100k", 50, None), ("ml-1m", 100, 64)])
def test_knn_lists_feed_the_scoring_and_match_the_oracle(shape, k, n_sample):
    r = datagen.generate(shape)
    with fy.Rm2Engine(number_of_items=r.n_items) as eng:
        eng.set_ratings(r.user, r.item, r.score)
        nb, cnt, n, _ = eng.knn_neighbours(r.n_users + 1, r.n_items + 1, k)       # top-k co-rating users per user id
    users = r.cl_user if n_sample is None else np.sort(np.random.default_rng(9).choice(r.cl_user, n_sample, replace=False))
    lists = nb[users]
    assert (lists >= 0).any(axis=1).all()
    got = _gpu(r, users, lists, 0.1, 100)
    want = orc.run_neighbours(r.user, r.item, r.score, users, lists, 0.1, r.n_items, 100)
    worst = assert_parity(got, want, 1e-6, "%s kNN k=%d" % (shape, k))
    assert worst < 1e-9
    assert got["users_scored"] == want["users_scored"] == len(by_user(want))
    assert np.array_equal(got["cluster"], want["cluster"])                      # = position of the user in the call


def test_errors_and_padding():
    r = datagen.generate("tiny")
    with fy.Rm2Engine(lam=0.1, number_of_items=r.n_items, top_n=5) as eng:
        eng.set_ratings(r.user, r.item, r.score)
        with pytest.raises(fy.Rm2Error) as e:                                   # a neighbour nobody has heard of
            eng.run_neighbours([1], [[2, 9999]])
        assert e.value.code == -2
        eng.run_neighbours([3, 5], [[-1, 4, -1, 4, 3], [6, -1, -1, -1, -1]])    # padding, duplicates, self
        got = eng.results()
    want = orc.run_neighbours(r.user, r.item, r.score, [3, 5], [[4], [6]], 0.1, r.n_items, 5)
    assert_parity(got, want, 1e-6, "padding")
