"""ctypes binding of oracle/liboracle_rm2.so -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may
import this module.  The product package filmyou_core_b200 never does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

MODE_LITERAL, MODE_LITERAL_FAST, MODE_GRAM = 0, 1, 2

ERRORS = {
    -1: "bad argument", -2: "user without positive rating", -3: "duplicate rating",
    -4: "clusteringCount mismatch", -5: "rating of a user absent from clustering", -6: "out of memory",
}


class OracleError(RuntimeError):
    def __init__(self, code):
        super().__init__("oracle error %d: %s" % (code, ERRORS.get(code, "?")))
        self.code = code


class _Params(C.Structure):
    _fields_ = [("lambda_", C.c_double), ("number_of_items", C.c_int32), ("top_n", C.c_int32),
                ("filter_users", C.c_int32), ("mode", C.c_int32), ("threads", C.c_int32), ("cand_stride", C.c_int32)]


def build(force=False):
    so = os.path.join(_HERE, "liboracle_rm2.so")
    src = [os.path.join(_HERE, f) for f in ("rm2_oracle.c", "rm2_oracle.h")]
    if force or not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in src):
        subprocess.check_call(["make", "-C", _HERE, "-s", "-B", "liboracle_rm2.so"])
    return so


def lib():
    global _LIB
    if _LIB is None:
        L = C.CDLL(build())
        i32p, f32p, f64p = C.POINTER(C.c_int32), C.POINTER(C.c_float), C.POINTER(C.c_double)
        L.orc_rm2_stats.argtypes = [i32p, i32p, f32p, C.c_int64, i32p, C.c_int64, f64p, C.c_int32, f64p, f64p, f64p]
        L.orc_rm2_stats.restype = C.c_int
        L.orc_rm2_run.argtypes = [C.POINTER(_Params), i32p, i32p, f32p, C.c_int64, i32p, i32p, C.c_int64,
                                  i32p, C.c_int32, i32p, C.c_int64, C.POINTER(C.c_void_p)]
        L.orc_rm2_run.restype = C.c_int
        L.orc_rm2_run_ext.argtypes = [C.POINTER(_Params), i32p, i32p, f32p, C.c_int64, i32p, i32p, C.c_int64,
                                      i32p, C.c_int32, i32p, C.c_int64, f64p, C.c_int32, C.POINTER(C.c_void_p)]
        L.orc_rm2_run_ext.restype = C.c_int
        L.orc_result_count.argtypes = [C.c_void_p]
        L.orc_result_count.restype = C.c_int64
        L.orc_result_seconds.argtypes = [C.c_void_p]
        L.orc_result_seconds.restype = C.c_double
        L.orc_result_users_scored.argtypes = [C.c_void_p]
        L.orc_result_users_scored.restype = C.c_int64
        L.orc_result_copy.argtypes = [C.c_void_p, i32p, i32p, f64p, f32p, i32p]
        L.orc_result_copy.restype = None
        L.orc_result_free.argtypes = [C.c_void_p]
        L.orc_result_free.restype = None
        L.orc_cooccurrence.argtypes = [i32p, i32p, f32p, C.c_int64, C.c_int32, C.c_int32, i32p]
        L.orc_cooccurrence.restype = C.c_int
        _LIB = L
    return _LIB


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t))


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def stats(r_user, r_item, r_score, users):
    """RM2-1 / RM2-2: returns (user_sum[len(users)], item_sum[max_item+1], item_prob[max_item+1], total)."""
    r_user, r_item, users = _i32(r_user), _i32(r_item), _i32(users)
    r_score = np.ascontiguousarray(r_score, dtype=np.float32)
    max_item = int(r_item.max()) if len(r_item) else 0
    us = np.zeros(len(users), np.float64)
    isum = np.zeros(max_item + 1, np.float64)
    ip = np.zeros(max_item + 1, np.float64)
    tot = C.c_double(0)
    rc = lib().orc_rm2_stats(_p(r_user, C.c_int32), _p(r_item, C.c_int32), _p(r_score, C.c_float), len(r_user),
                             _p(users, C.c_int32), len(users), _p(us, C.c_double), max_item,
                             _p(isum, C.c_double), _p(ip, C.c_double), C.byref(tot))
    if rc:
        raise OracleError(rc)
    return us, isum, ip, tot.value


def run(r_user, r_item, r_score, cl_user, cl_cluster, cluster_size, lam, number_of_items, top_n,
        filter_users=0, mode=MODE_LITERAL_FAST, threads=0, only_users=None, cand_stride=1, ext_item_prob=None):
    """Whole RM2 job on the CPU.  Returns dict(user, item, score64, score32, cluster, seconds, users_scored)."""
    r_user, r_item = _i32(r_user), _i32(r_item)
    r_score = np.ascontiguousarray(r_score, dtype=np.float32)
    cl_user, cl_cluster, cluster_size = _i32(cl_user), _i32(cl_cluster), _i32(cluster_size)
    if threads <= 0:
        threads = os.cpu_count() or 1
    prm = _Params(float(lam), int(number_of_items), int(top_n), int(filter_users), int(mode), int(threads), int(cand_stride))
    only = _i32(only_users) if only_users is not None and len(only_users) else np.zeros(0, np.int32)
    h = C.c_void_p()
    ext = np.ascontiguousarray(ext_item_prob, dtype=np.float64) if ext_item_prob is not None else None
    rc = lib().orc_rm2_run_ext(C.byref(prm), _p(r_user, C.c_int32), _p(r_item, C.c_int32), _p(r_score, C.c_float),
                               len(r_user), _p(cl_user, C.c_int32), _p(cl_cluster, C.c_int32), len(cl_user),
                               _p(cluster_size, C.c_int32), len(cluster_size),
                               _p(only, C.c_int32), len(only),
                               _p(ext, C.c_double) if ext is not None else None, (len(ext) - 1) if ext is not None else -1,
                               C.byref(h))
    if rc:
        raise OracleError(rc)
    try:
        n = lib().orc_result_count(h)
        out = dict(user=np.zeros(n, np.int32), item=np.zeros(n, np.int32), score64=np.zeros(n, np.float64),
                   score32=np.zeros(n, np.float32), cluster=np.zeros(n, np.int32))
        lib().orc_result_copy(h, _p(out["user"], C.c_int32), _p(out["item"], C.c_int32),
                              _p(out["score64"], C.c_double), _p(out["score32"], C.c_float),
                              _p(out["cluster"], C.c_int32))
        out["seconds"] = lib().orc_result_seconds(h)
        out["users_scored"] = lib().orc_result_users_scored(h)
        out["threads"] = threads
    finally:
        lib().orc_result_free(h)
    return out


def run_neighbours(r_user, r_item, r_score, users, neighbours, lam, number_of_items, top_n,
                   mode=MODE_LITERAL_FAST, threads=0):
    """buildRecommendations over EXPLICIT neighbour lists (the `int[] neighbours` of AbstractRM2Reducer.java:321-323,342-346):
    users[q] is scored as the reducer would score it in a reduce() group made of users[q] and neighbours[q] (-1 = empty
    slot) -- K = |N(u)| + 1 (:143,:329), items = whatever the group rated (:164-174), neighbour sum over N(u) (:342-346),
    userSum / itemColl = the statistics jobs' global outputs.  Restated by literally building those groups (fresh user ids,
    members in ascending id) and running the cluster oracle on them with the global p(i|C) handed in; pinned by
    tests/test_oracle_golden.py: with N(u) = cluster(u) minus u it reproduces the reference's 507 golden triples.
    Returns the dict of run() with `user` mapped back and `cluster` = position of the user in `users`."""
    r_user, r_item = _i32(r_user), _i32(r_item)
    r_score = np.ascontiguousarray(r_score, dtype=np.float32)
    users = _i32(users)
    neighbours = np.asarray(neighbours, dtype=np.int64).reshape(len(users), -1)
    keep = r_score > 0
    order = np.lexsort((r_item[keep], r_user[keep]))                      # ascending (user, item): the canonical statistics order
    ru, ri, rs = r_user[keep][order], r_item[keep][order], r_score[keep][order]
    max_item = int(ri.max())
    usum = np.zeros(int(ru.max()) + 1, np.float64)
    isum = np.zeros(max_item + 1, np.float64)
    np.add.at(usum, ru, rs.astype(np.float64))                            # unbuffered, in order of appearance
    np.add.at(isum, ri, rs.astype(np.float64))
    total = float(np.sum(usum.astype(np.int64) * 100)) / 100.0            # (long) sum * OFFSET, DoubleSumAndCountReducer.java:41
    iprob = np.where(isum > 0, isum / total, 0.0)
    start = np.searchsorted(ru, np.arange(int(ru.max()) + 2))
    xu, xi, xs, vuser, vcluster, vreal, owner = [], [], [], [], [], [], []
    vid = 1
    for q, u in enumerate(users):
        members = np.unique(np.concatenate([[int(u)], neighbours[q][(neighbours[q] >= 0) & (neighbours[q] != u)]]))
        for m in members:
            a, b = start[m], start[m + 1]
            if b <= a:
                raise OracleError(-2)
            xu.append(np.full(b - a, vid, np.int32)); xi.append(ri[a:b]); xs.append(rs[a:b])
            vuser.append(vid); vcluster.append(q); vreal.append(int(m))
            if m == u:
                owner.append(vid)
            vid += 1
    vreal = np.array(vreal, np.int32)
    csize = np.bincount(np.array(vcluster), minlength=len(users)).astype(np.int32)
    out = run(np.concatenate(xu), np.concatenate(xi), np.concatenate(xs), vuser, vcluster, csize, lam, number_of_items, top_n,
              mode=mode, threads=threads, only_users=np.array(owner, np.int32), ext_item_prob=iprob)
    out["user"] = vreal[out["user"] - 1]
    return out


def cooccurrence(r_user, r_item, r_score, n_user_ids, n_items):
    r_user, r_item = _i32(r_user), _i32(r_item)
    r_score = np.ascontiguousarray(r_score, dtype=np.float32)
    Cm = np.zeros((n_items, n_items), np.int32)
    rc = lib().orc_cooccurrence(_p(r_user, C.c_int32), _p(r_item, C.c_int32), _p(r_score, C.c_float),
                                len(r_user), n_user_ids, n_items, _p(Cm, C.c_int32))
    if rc:
        raise OracleError(rc)
    return Cm
