"""Host-side mirror of the item-similarity phase of the reference's baseline recommender (config 3).

M/baselinerecommender/BaselineRecommenderJob.java:140-172 declares the options, :214-224 hands `maxPrefsPerUserInItemSimilarity`
and `minPrefsPerUser` to the prepare job (BaselinePreparePreferenceMatrixJob.java:103-104,126-128) and :241-253 runs Mahout's
RowSimilarityJob(CooccurrenceCountSimilarity, maxSimilaritiesPerItem, excludeSelfSimilarity, threshold).  The counting itself
is the tcgen05 GEMM behind `fy_cooc_counts` / `fy_cooc_topk`; this module restates the three options around it, on the host:

  * minPrefsPerUser  -- users with fewer preferences are dropped from the similarity computation
                        (Mahout 0.8 ToUserVectorsReducer.MIN_PREFERENCES_PER_USER);
  * maxPrefsPerUserInItemSimilarity -- users with more preferences are sampled down to that many
                        (Mahout 0.8 ToItemVectorsMapper.SAMPLE_SIZE).  Mahout draws a random sample with an unseeded
                        generator, so no two runs of the reference agree on it either; here the sample is a seeded
                        choice, reproducible, and it is NOT Mahout's stream of random numbers;
  * threshold        -- item pairs whose count is below it are discarded (RowSimilarityJob --threshold; the reference's
                        default NO_THRESHOLD keeps every pair with a positive count).

Mahout 0.8 is a pom.xml dependency that is not vendored in the reference tree, so these semantics are restated from its
published behaviour and **parity is unpinned** (DESIGN.md 6); the tests check them against plain numpy.
"""
import numpy as np

DEFAULT_MAX_SIMILARITIES_PER_ITEM = 100    # BaselineRecommenderJob.java:68
DEFAULT_MAX_PREFS_PER_USER = 1000          # :69  (maxPrefsPerUserInItemSimilarity)
DEFAULT_MIN_PREFS_PER_USER = 1             # :70
NO_THRESHOLD = None                        # RowSimilarityJob.NO_THRESHOLD (:172)


def prepare_preferences(user, item, score, min_prefs_per_user=DEFAULT_MIN_PREFS_PER_USER,
                        max_prefs_per_user=DEFAULT_MAX_PREFS_PER_USER, seed=0):
    """The preference matrix the similarity phase sees: positive, de-duplicated (user, item) pairs of the users with at
    least `min_prefs_per_user` of them, users above `max_prefs_per_user` sampled down to exactly that many.
    Returns (user, item, score) int32/int32/float32 arrays sorted by (user, item), and the number of users kept
    (PreparePreferenceMatrixJob.NUM_USERS, read back at BaselineRecommenderJob.java:222-223)."""
    user = np.asarray(user, np.int32); item = np.asarray(item, np.int32); score = np.asarray(score, np.float32)
    if not (user.shape == item.shape == score.shape and user.ndim == 1):
        raise ValueError("user, item and score must be 1-D arrays of one length")
    if min_prefs_per_user < 1 or max_prefs_per_user < 1:
        raise ValueError("minPrefsPerUser and maxPrefsPerUser must be >= 1")
    keep = score > 0                                     # the count similarity only sees which pairs exist
    user, item, score = user[keep], item[keep], score[keep]
    order = np.lexsort((item, user))
    user, item, score = user[order], item[order], score[order]
    first = np.ones(user.shape[0], bool)
    first[1:] = (user[1:] != user[:-1]) | (item[1:] != item[:-1])      # a pair counts once
    user, item, score = user[first], item[first], score[first]
    if user.shape[0] == 0:
        return user, item, score, 0
    starts = np.flatnonzero(np.r_[True, user[1:] != user[:-1]])
    counts = np.diff(np.r_[starts, user.shape[0]])
    ok_user = counts >= min_prefs_per_user
    sel = np.repeat(ok_user, counts)
    rng = np.random.default_rng(seed)
    for s, c in zip(starts[ok_user & (counts > max_prefs_per_user)], counts[ok_user & (counts > max_prefs_per_user)]):
        drop = rng.choice(c, size=c - max_prefs_per_user, replace=False)
        sel[s + drop] = False
    return user[sel], item[sel], score[sel], int(ok_user.sum())


def apply_threshold(items, counts, n, threshold):
    """RowSimilarityJob --threshold on the (count desc, id asc) top-k lists of `fy_cooc_topk`: entries below the
    threshold are discarded (-1 / 0 padded), `n` shrinks accordingly.  `threshold` None = NO_THRESHOLD."""
    items = np.array(items, np.int32, copy=True); counts = np.array(counts, np.int32, copy=True); n = np.array(n, np.int32, copy=True)
    if threshold is None:
        return items, counts, n
    k = items.shape[1]
    live = (np.arange(k)[None, :] < n[:, None]) & (counts >= threshold)
    # lists are sorted by count descending, so the survivors are a prefix of each row
    n_new = live.sum(1).astype(np.int32)
    dead = np.arange(k)[None, :] >= n_new[:, None]
    items[dead] = -1
    counts[dead] = 0
    return items, counts, n_new


class ItemSimilarityJob:
    """`BaselineRecommenderJob`'s similarity phase with the reference's option names.

    run(engine, user, item, score, n_items) -> (similar_items[n_items, k], counts[n_items, k], n[n_items], info):
    `engine` is a `filmyou_core_b200.Rm2Engine` (any object with set_ratings / cooc_counts / cooc_topk)."""

    def __init__(self, maxSimilaritiesPerItem=DEFAULT_MAX_SIMILARITIES_PER_ITEM,
                 maxPrefsPerUserInItemSimilarity=DEFAULT_MAX_PREFS_PER_USER, minPrefsPerUser=DEFAULT_MIN_PREFS_PER_USER,
                 threshold=NO_THRESHOLD, seed=0):
        if maxSimilaritiesPerItem < 1:
            raise ValueError("maxSimilaritiesPerItem must be >= 1")
        self.maxSimilaritiesPerItem = int(maxSimilaritiesPerItem)
        self.maxPrefsPerUserInItemSimilarity = int(maxPrefsPerUserInItemSimilarity)
        self.minPrefsPerUser = int(minPrefsPerUser)
        self.threshold = threshold
        self.seed = seed

    def run(self, engine, user, item, score, n_items):
        u, i, s, n_users_kept = prepare_preferences(user, item, score, self.minPrefsPerUser,
                                                    self.maxPrefsPerUserInItemSimilarity, self.seed)
        if u.shape[0] == 0:
            raise ValueError("no preference left after the minPrefsPerUser filter")
        if int(i.max()) >= n_items or int(i.min()) < 0:
            raise ValueError("item id outside [0, n_items)")
        engine.set_ratings(u, i, s)
        _, ms = engine.cooc_counts(int(u.max()) + 1, int(n_items), want_counts=False)
        items, counts, n = engine.cooc_topk(int(n_items), self.maxSimilaritiesPerItem)     # excludeSelfSimilarity = true (:251)
        items, counts, n = apply_threshold(items, counts, n, self.threshold)
        return items, counts, n, {"users": n_users_kept, "preferences": int(u.shape[0]), "ms_gemm": ms}
