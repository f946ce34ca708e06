import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run by the driver with -m gpu)")


@pytest.fixture(scope="session")
def golden():
    """The reference's own golden vectors (T/testdata/RMTestData.java, ClusteringTestData.java),
    extracted by tests/golden/make_golden.py."""
    with open(os.path.join(ROOT, "tests", "golden", "rm_test_data.json")) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def golden2():
    with open(os.path.join(ROOT, "tests", "golden", "rm_test_data2.json")) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def golden_ratings(golden):
    from filmyou_core_b200 import datagen
    k = golden["numberOfClusters"]                      # 10 configured, 5 real (RMTestData.java:27)
    csize = np.zeros(k, np.int32)
    csize[:len(golden["clusteringCount"])] = golden["clusteringCount"]
    return datagen.from_dense(golden["A_item_by_user"], golden["clustering"], csize, name="golden")


def by_user(res):
    """packed triples -> {user: (items[], scores[])} keeping emission order"""
    out = {}
    users = np.asarray(res["user"])
    if len(users) == 0:
        return out
    cut = np.flatnonzero(np.diff(users)) + 1
    starts = np.concatenate([[0], cut])
    ends = np.concatenate([cut, [len(users)]])
    for s, e in zip(starts, ends):
        u = int(users[s])
        assert u not in out, "user %d emitted in two groups" % u
        out[u] = (np.asarray(res["item"][s:e]), np.asarray(res["score64"][s:e]))
    return out


def assert_parity(got, want, rel=1e-6, what=""):
    """ids bit-exact and in the same order, scores within `rel` relative (north_star tolerance)."""
    g, w = by_user(got), by_user(want)
    assert set(g) == set(w), "%s: scored user sets differ: %d vs %d" % (what, len(g), len(w))
    worst = 0.0
    for u in w:
        gi, gs = g[u]
        wi, ws = w[u]
        assert len(gi) == len(wi), "%s: user %d emits %d items, oracle %d" % (what, u, len(gi), len(wi))
        if not np.array_equal(gi, wi):
            k = int(np.flatnonzero(gi != wi)[0])
            raise AssertionError("%s: user %d top-N differs at position %d: item %d (%.17g) vs oracle %d (%.17g)"
                                 % (what, u, k, gi[k], gs[k], wi[k], ws[k]))
        fin = np.isfinite(ws)
        assert np.array_equal(np.isfinite(gs), fin)
        assert np.array_equal(gs[~fin], ws[~fin])
        if fin.any():
            r = np.max(np.abs(gs[fin] - ws[fin]) / np.maximum(np.abs(ws[fin]), 1e-300))
            worst = max(worst, float(r))
    assert worst <= rel, "%s: worst relative score error %.3g > %.1g" % (what, worst, rel)
    return worst


def assert_parity_near_ties(got, want, rel=1e-6, tie_rel=1e-12, what=""):
    """Like assert_parity, for comparisons at full scale: the top-N ids must be the oracle's, in the oracle's order, EXCEPT
    that two items whose ORACLE scores differ by less than tie_rel relative may swap places (or straddle the N-th place):
    the engine and the oracle both round in the 1e-13 range, so such a pair has no defined order (the reference's own
    PriorityQueue leaves even exact ties unordered, M/util/IntDouble.java:31-34).  Returns (worst relative score error,
    number of users with such a swap)."""
    g, w = by_user(got), by_user(want)
    assert set(g) == set(w), "%s: scored user sets differ: %d vs %d" % (what, len(g), len(w))
    worst, swapped = 0.0, 0
    for u in w:
        gi, gs = g[u]
        wi, ws = w[u]
        assert len(gi) == len(wi), "%s: user %d emits %d items, oracle %d" % (what, u, len(gi), len(wi))
        r = np.max(np.abs(gs - ws) / np.maximum(np.abs(ws), 1e-300))     # position-wise: scores are sorted on both sides
        worst = max(worst, float(r))
        if np.array_equal(gi, wi):
            continue
        swapped += 1
        pos = {int(i): k for k, i in enumerate(wi)}
        for k in np.flatnonzero(gi != wi):
            k2 = pos.get(int(gi[k]))
            ref = ws[k2] if k2 is not None else ws[-1]                   # not in the oracle's list: must tie with its N-th place
            assert abs(ref - ws[k]) <= tie_rel * abs(ws[k]), \
                "%s: user %d position %d: item %d vs oracle %d is not a near-tie (oracle scores %.17g / %.17g)" % (what, u, k, gi[k], wi[k], ref, ws[k])
    assert worst <= rel, "%s: worst relative score error %.3g > %.1g" % (what, worst, rel)
    return worst, swapped
