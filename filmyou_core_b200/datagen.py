"""Seeded synthetic rating matrices and clusterings of the shapes BASELINE.json names.

The reference ships no data sets (only the 30x100 golden matrix of T/testdata/RMTestData.java), so
every benchmark input is generated here, deterministically, from a seed:

  * user activity ~ lognormal(sigma=1), clipped to [20, n_items/2], rescaled to hit nnz exactly;
  * item popularity ~ Zipf(s=0.9) over a seeded permutation of the item ids;
  * each user's items drawn without replacement; ids are 1-based like the reference's fixtures
    (M/util/DataInitialization.java:168-174);
  * `clustering` = seeded uniform random assignment, every cluster non-empty with >= 2 users,
    in the content of the reference's two files `clustering` (user -> cluster) and
    `clusteringCount` (cluster -> size)  (M/common/AbstractByClusterMapper.java:57-66,
    M/rm/AbstractRM2Reducer.java:93-105).
"""
import hashlib
from dataclasses import dataclass

import numpy as np

# name -> (users, items, nnz, clusters, rating values, rating probabilities, seed)
_INT5 = (np.array([1, 2, 3, 4, 5], np.float32), np.array([.06, .11, .27, .34, .22]))
_HALF = (np.arange(1, 11, dtype=np.float32) / 2,
         np.array([.012, .034, .014, .072, .044, .215, .110, .278, .077, .144]))
SHAPES = {
    "ml-100k": (943, 1682, 100_000, 5, _INT5, 1001),
    "ml-1m": (6040, 3706, 1_000_209, 10, _INT5, 1002),
    "ml-20m": (138_493, 26_744, 20_000_263, 50, _HALF, 1004),
    "netflix": (480_189, 17_770, 100_480_507, 50, _INT5, 1005),
    # small shapes for tests
    "tiny": (60, 80, 1500, 4, _HALF, 7),
    "small": (300, 500, 12_000, 6, _HALF, 11),
}


@dataclass
class Ratings:
    name: str
    n_users: int
    n_items: int
    user: np.ndarray          # int32 [nnz], 1-based ids, arbitrary order
    item: np.ndarray          # int32 [nnz], 1-based ids
    score: np.ndarray         # float32 [nnz], > 0
    cl_user: np.ndarray       # int32 [n_users]   `clustering` keys
    cl_cluster: np.ndarray    # int32 [n_users]   `clustering` values (0-based)
    cluster_size: np.ndarray  # int32 [k]         `clusteringCount`
    seed: int

    @property
    def nnz(self):
        return int(self.user.shape[0])

    @property
    def n_clusters(self):
        return int(self.cluster_size.shape[0])

    def sha256(self):
        h = hashlib.sha256()
        for a in (self.user, self.item, self.score, self.cl_user, self.cl_cluster):
            h.update(np.ascontiguousarray(a).tobytes())
        return h.hexdigest()


def _activity(rng, n_users, n_items, nnz):
    lo, hi = min(20, max(1, n_items // 4)), max(1, n_items // 2)
    raw = rng.lognormal(mean=0.0, sigma=1.0, size=n_users)
    scale = nnz / raw.sum()
    for _ in range(60):                       # rescale under the clip until the total matches
        n = np.clip(raw * scale, lo, hi)
        scale *= nnz / n.sum()
    n = np.clip(np.floor(raw * scale), lo, hi).astype(np.int64)
    diff = int(nnz - n.sum())
    order = rng.permutation(n_users)
    k = 0
    while diff != 0:                           # distribute the rounding residue
        u = order[k % n_users]
        k += 1
        if diff > 0 and n[u] < hi:
            n[u] += 1
            diff -= 1
        elif diff < 0 and n[u] > lo:
            n[u] -= 1
            diff += 1
        if k > 100 * n_users + abs(diff) * 4 + 1000:
            raise ValueError("cannot reach nnz=%d with these bounds" % nnz)
    return n


def _sample_items(rng, need, cdf, n_items):
    """For every user u draw need[u] distinct item slots from the popularity cdf."""
    n_users = need.shape[0]
    need = need.copy()
    chosen = np.empty(0, np.int64)             # sorted keys user * n_items + slot
    users = np.arange(n_users, dtype=np.int64)
    rounds = 0
    while need.sum() > 0:
        rounds += 1
        draws = np.where(need > 0, (need * 3) // 2 + 8, 0)
        u = np.repeat(users, draws)
        slot = np.minimum(np.searchsorted(cdf, rng.random(u.shape[0]), side="right"), n_items - 1)
        key = u * n_items + slot
        if chosen.shape[0]:
            pos = np.minimum(np.searchsorted(chosen, key), chosen.shape[0] - 1)
            key = key[chosen[pos] != key]
        _, first = np.unique(key, return_index=True)
        key = key[np.sort(first)]              # distinct, still in draw order (grouped by user)
        ku = key // n_items
        cnt = np.bincount(ku, minlength=n_users)
        start = np.cumsum(cnt) - cnt
        rank = np.arange(key.shape[0]) - start[ku]
        key = key[rank < need[ku]]
        need -= np.bincount(key // n_items, minlength=n_users)
        chosen = np.sort(np.concatenate([chosen, key]))
        if rounds > 400:
            raise RuntimeError("item sampling did not converge")
    return chosen // n_items, chosen % n_items


def make_clustering(rng, n_users, k):
    if n_users < 2 * k:
        raise ValueError("need at least 2 users per cluster")
    perm = rng.permutation(n_users)
    cl = np.empty(n_users, np.int32)
    cl[perm[:2 * k]] = np.arange(2 * k) % k
    cl[perm[2 * k:]] = rng.integers(0, k, size=n_users - 2 * k)
    return cl, np.bincount(cl, minlength=k).astype(np.int32)


def generate(name, n_users=None, n_items=None, nnz=None, n_clusters=None, seed=None, shuffle=True):
    """Build the named shape (or a custom one when the sizes are given)."""
    if name in SHAPES:
        U, M, Z, K, (vals, probs), sd = SHAPES[name]
    else:
        U, M, Z, K, (vals, probs), sd = n_users, n_items, nnz, n_clusters, _HALF, 1
    U = n_users or U
    M = n_items or M
    Z = nnz or Z
    K = n_clusters or K
    sd = sd if seed is None else seed
    rng = np.random.default_rng(sd)
    n_per_user = _activity(rng, U, M, Z)
    w = 1.0 / np.arange(1, M + 1, dtype=np.float64) ** 0.9
    cdf = np.cumsum(w / w.sum())
    slot_to_item = rng.permutation(M).astype(np.int64)
    u, slot = _sample_items(rng, n_per_user, cdf, M)
    item = slot_to_item[slot]
    score = vals[rng.choice(vals.shape[0], size=u.shape[0], p=probs / probs.sum())]
    if shuffle:
        p = rng.permutation(u.shape[0])
        u, item, score = u[p], item[p], score[p]
    cl, csize = make_clustering(rng, U, K)
    return Ratings(name, U, M, (u + 1).astype(np.int32), (item + 1).astype(np.int32),
                   score.astype(np.float32), np.arange(1, U + 1, dtype=np.int32), cl, csize, sd)


def from_dense(A_item_by_user, clustering, cluster_size, name="dense"):
    """Ratings from a dense A[item][user] matrix, ids 1-based: the layout of the reference's
    fixture writer (M/util/DataInitialization.java:155-174)."""
    A = np.asarray(A_item_by_user, dtype=np.float64)
    it, us = np.nonzero(A > 0)
    return Ratings(name, A.shape[1], A.shape[0], (us + 1).astype(np.int32), (it + 1).astype(np.int32),
                   A[it, us].astype(np.float32), np.arange(1, A.shape[1] + 1, dtype=np.int32),
                   np.asarray(clustering, np.int32), np.asarray(cluster_size, np.int32), 0)
