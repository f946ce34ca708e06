"""Config 3's options around the co-occurrence count (filmyou_core_b200/baseline_job.py), checked against plain numpy.
The engine is replaced by a numpy stand-in with the same three calls, so the host logic is covered without a GPU; the
counting kernel itself is covered by tests/test_gpu_cooc.py.  Mahout 0.8 is not in the reference tree: parity unpinned."""
import numpy as np
import pytest

from filmyou_core_b200 import baseline_job as bj
from filmyou_core_b200 import datagen


class NumpyEngine:
    """set_ratings / cooc_counts / cooc_topk with the C ABI's semantics (binarised, self excluded, count desc + id asc)."""

    def set_ratings(self, user, item, score):
        self.u, self.i, self.s = np.asarray(user), np.asarray(item), np.asarray(score)

    def cooc_counts(self, n_user_ids, n_items, want_counts=True):
        B = np.zeros((n_user_ids, n_items), np.float64)
        pos = self.s > 0
        B[self.u[pos], self.i[pos]] = 1
        self.C = (B.T @ B).astype(np.int32)
        return (self.C.copy() if want_counts else None), 0.0

    def cooc_topk(self, n_items, k):
        n = self.C.shape[0]
        items = -np.ones((n, k), np.int32); counts = np.zeros((n, k), np.int32); cnt = np.zeros(n, np.int32)
        for i in range(n):
            row = self.C[i].astype(np.int64).copy()
            row[i] = 0
            nz = np.flatnonzero(row > 0)
            order = nz[np.lexsort((nz, -row[nz]))][:k]
            items[i, :len(order)] = order; counts[i, :len(order)] = row[order]; cnt[i] = len(order)
        return items, counts, cnt


def test_defaults_are_the_reference_defaults():
    # M/baselinerecommender/BaselineRecommenderJob.java:68-70,172
    job = bj.ItemSimilarityJob()
    assert (job.maxSimilaritiesPerItem, job.maxPrefsPerUserInItemSimilarity, job.minPrefsPerUser, job.threshold) == (100, 1000, 1, None)


def test_prepare_preferences_filters_and_samples():
    user = np.array([5, 5, 5, 5, 7, 7, 9, 9, 9, 9, 9, 9], np.int32)
    item = np.array([1, 2, 2, 3, 1, 4, 0, 1, 2, 3, 4, 5], np.int32)
    score = np.array([1, 2, 3, 0, 4, -1, 1, 1, 1, 1, 1, 1], np.float32)     # (5,2) twice, (5,3) and (7,4) not positive
    u, i, s, kept = bj.prepare_preferences(user, item, score, min_prefs_per_user=2, max_prefs_per_user=4, seed=3)
    assert kept == 2                                            # user 7 has one positive preference: dropped
    assert set(u.tolist()) == {5, 9}
    assert i[u == 5].tolist() == [1, 2]                         # de-duplicated, sorted
    assert (u == 9).sum() == 4 and set(i[u == 9].tolist()) <= {0, 1, 2, 3, 4, 5}     # sampled down to exactly 4
    u2, i2, _, _ = bj.prepare_preferences(user, item, score, 2, 4, seed=3)
    assert np.array_equal(u, u2) and np.array_equal(i, i2)      # seeded: reproducible
    with pytest.raises(ValueError):
        bj.prepare_preferences(user, item, score, 0, 4)
    eu, ei, es, ek = bj.prepare_preferences(user[:0], item[:0], score[:0])
    assert ek == 0 and eu.shape == (0,)


def test_apply_threshold_keeps_a_prefix():
    items = np.array([[3, 1, 2, -1], [0, 2, -1, -1]], np.int32)
    counts = np.array([[9, 4, 4, 0], [2, 1, 0, 0]], np.int32)
    n = np.array([3, 2], np.int32)
    it, ct, nn = bj.apply_threshold(items, counts, n, 4)
    assert nn.tolist() == [3, 0] and it.tolist() == [[3, 1, 2, -1], [-1, -1, -1, -1]] and ct.tolist() == [[9, 4, 4, 0], [0, 0, 0, 0]]
    it, ct, nn = bj.apply_threshold(items, counts, n, 5)
    assert nn.tolist() == [1, 0] and it[0].tolist() == [3, -1, -1, -1]
    it, ct, nn = bj.apply_threshold(items, counts, n, None)     # NO_THRESHOLD
    assert np.array_equal(it, items) and np.array_equal(nn, n)
    assert items[1, 0] == 0                                     # inputs untouched


@pytest.mark.parametrize("min_prefs,max_prefs,threshold", [(1, 1000, None), (25, 1000, None), (1, 30, None), (1, 1000, 3), (22, 40, 2)])
def test_job_equals_numpy_on_the_filtered_matrix(min_prefs, max_prefs, threshold):
    r = datagen.generate("small")
    n_items = r.n_items + 1
    job = bj.ItemSimilarityJob(maxSimilaritiesPerItem=10, maxPrefsPerUserInItemSimilarity=max_prefs, minPrefsPerUser=min_prefs,
                               threshold=threshold, seed=11)
    items, counts, n, info = job.run(NumpyEngine(), r.user, r.item, r.score, n_items)
    # independent restatement: dense 0/1 matrix of the users that survive, columns sampled by the same seeded rule
    u, i, s, kept = bj.prepare_preferences(r.user, r.item, r.score, min_prefs, max_prefs, seed=11)
    per_user = np.bincount(u, minlength=r.n_users + 1)
    raw = np.bincount(r.user[r.score > 0], minlength=r.n_users + 1)
    assert np.all(per_user[raw < min_prefs] == 0)
    assert np.all(per_user[raw >= min_prefs] == np.minimum(raw[raw >= min_prefs], max_prefs))
    assert info["users"] == int((raw >= min_prefs).sum()) == kept and info["preferences"] == len(u)
    B = np.zeros((r.n_users + 1, n_items), np.int64); B[u, i] = 1
    C = B.T @ B
    np.fill_diagonal(C, 0)
    for it in range(n_items):
        row = C[it]
        nz = np.flatnonzero(row >= (threshold if threshold is not None else 1))
        order = nz[np.lexsort((nz, -row[nz]))][:10]
        if threshold is not None:                               # the threshold applies to the top-10 list, as in the job
            top = np.flatnonzero(row > 0); top = top[np.lexsort((top, -row[top]))][:10]
            order = top[row[top] >= threshold]
        assert n[it] == len(order) and items[it, :n[it]].tolist() == order.tolist() and counts[it, :n[it]].tolist() == row[order].tolist()
        assert np.all(items[it, n[it]:] == -1) and np.all(counts[it, n[it]:] == 0)
