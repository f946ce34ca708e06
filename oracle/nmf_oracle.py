"""ctypes binding of oracle/liboracle_nmf.so (nmf_oracle.c) -- TEST INFRASTRUCTURE ONLY.

Only tests/ may import this module.  The product package filmyou_core_b200 never does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

NMF, PPC = 0, 1
ERRORS = {-1: "bad argument", -2: "user without positive rating", -6: "out of memory", -10: "item without positive rating"}


class OracleError(RuntimeError):
    def __init__(self, code, bad_id=-1):
        super().__init__("nmf oracle error %d: %s (id %d)" % (code, ERRORS.get(code, "?"), bad_id))
        self.code = code
        self.bad_id = bad_id


def build(force=False):
    so = os.path.join(_HERE, "liboracle_nmf.so")
    src = os.path.join(_HERE, "nmf_oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(src) > os.path.getmtime(so):
        subprocess.check_call(["make", "-C", _HERE, "-s", "-B", "liboracle_nmf.so"])
    return so


def lib():
    global _LIB
    if _LIB is None:
        L = C.CDLL(build())
        i32p, f32p, f64p = C.POINTER(C.c_int32), C.POINTER(C.c_float), C.POINTER(C.c_double)
        L.orc_nmf_run.argtypes = [C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, i32p, i32p, f32p, C.c_int64,
                                  f64p, f64p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, i32p]
        L.orc_nmf_run.restype = C.c_int
        L.orc_cluster_assign.argtypes = [f64p, C.c_int32, C.c_int32, i32p, i32p]
        L.orc_cluster_assign.restype = C.c_int
        _LIB = L
    return _LIB


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t))


def run(mode, user, item, score, H, W, n_iter, id_base=1, apply_normalization=False, normalization_frequency=-1,
        combine_len=0, split_rows=0):
    """n_iter iterations of NMFDriver (mode NMF) / PPCDriver (mode PPC); returns new (H, W)."""
    user = np.ascontiguousarray(user, np.int32); item = np.ascontiguousarray(item, np.int32)
    score = np.ascontiguousarray(score, np.float32)
    H = np.array(H, np.float64, order="C"); W = np.array(W, np.float64, order="C")
    assert H.shape[1] == W.shape[1]
    bad = C.c_int32(-1)
    rc = lib().orc_nmf_run(mode, H.shape[0], W.shape[0], H.shape[1], id_base, _p(user, C.c_int32), _p(item, C.c_int32),
                           _p(score, C.c_float), len(user), _p(H, C.c_double), _p(W, C.c_double), n_iter,
                           int(bool(apply_normalization)), normalization_frequency, combine_len, split_rows, C.byref(bad))
    if rc != 0:
        raise OracleError(rc, bad.value)
    return H, W


def cluster_assign(H):
    """(clustering[user row], clusteringCount[k]) = arg-max assignment + counts."""
    H = np.ascontiguousarray(H, np.float64)
    cl = np.zeros(H.shape[0], np.int32); cnt = np.zeros(H.shape[1], np.int32)
    rc = lib().orc_cluster_assign(_p(H, C.c_double), H.shape[0], H.shape[1], _p(cl, C.c_int32), _p(cnt, C.c_int32))
    if rc != 0:
        raise OracleError(rc)
    return cl, cnt


def coo_from_dense(A):
    """A[item][user] (0 = none) -> (user ids, item ids, scores), 1-based, in the order
    DataInitialization.createIntPairFloatFile writes them (M/util/DataInitialization.java:167-175)."""
    A = np.asarray(A, np.float64)
    ii, jj = np.nonzero(A > 0)
    return (jj + 1).astype(np.int32), (ii + 1).astype(np.int32), A[ii, jj].astype(np.float32)
