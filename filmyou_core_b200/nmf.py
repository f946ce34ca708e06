"""ctypes binding of the NMF / PPC clustering entry points (C ABI: include/filmyou_nmf.h) and the host-side
mirror of the reference's drivers for that step:

    PPCDriver / NMFDriver        M/nmf/ppc/PPCDriver.java, M/nmf/NMFDriver.java  (AbstractNMFDriver.run :92-141)
    ClusterAssignmentJob         M/nmf/clustering/ClusterAssignmentJob.java:47-95
    CountClustersJob             M/nmf/clustering/CountClustersJob.java:41-80

(M/ = /root/reference/src/main/java/es/udc/fi/dc/irlab/).  No CPU path exists behind this module.
"""
import ctypes as C

import numpy as np

from .engine import NMF_EXPORTS, Rm2Error, _i32, _ptr, load_library

NMF, PPC = 0, 1


class NmfParams(C.Structure):
    _fields_ = [("mode", C.c_int32), ("number_of_users", C.c_int32), ("number_of_items", C.c_int32),
                ("number_of_clusters", C.c_int32), ("number_of_iterations", C.c_int32),
                ("normalization_frequency", C.c_int32), ("apply_normalization", C.c_int32), ("id_base", C.c_int32),
                ("combine_len", C.c_int32), ("split_rows", C.c_int32), ("device", C.c_int32)]


class NmfProfile(C.Structure):
    _fields_ = [("ms_index", C.c_double), ("ms_iterations", C.c_double), ("ms_per_iteration", C.c_double),
                ("join_bytes", C.c_double), ("ratings", C.c_int64), ("kernel_launches", C.c_int64),
                ("iterations", C.c_int32), ("graph_replays", C.c_int32)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


_BOUND = False


def _lib():
    global _BOUND
    L = load_library()
    if not _BOUND:
        vp, i32p, f32p, f64p = C.c_void_p, C.POINTER(C.c_int32), C.POINTER(C.c_float), C.POINTER(C.c_double)
        L.fy_nmf_default_params.argtypes = [C.POINTER(NmfParams)]
        L.fy_nmf_default_params.restype = None
        L.fy_nmf_create.argtypes = [C.POINTER(vp), C.POINTER(NmfParams)]
        L.fy_nmf_destroy.argtypes = [vp]
        L.fy_nmf_destroy.restype = None
        L.fy_nmf_last_error.argtypes = [vp]
        L.fy_nmf_last_error.restype = C.c_char_p
        L.fy_nmf_set_ratings.argtypes = [vp, i32p, i32p, f32p, C.c_int64]
        L.fy_nmf_set_factors.argtypes = [vp, f64p, f64p]
        L.fy_nmf_init_random.argtypes = [vp, C.c_uint64]
        L.fy_nmf_run.argtypes = [vp]
        L.fy_nmf_get_factors.argtypes = [vp, f64p, f64p]
        L.fy_nmf_cluster_assignment.argtypes = [vp, i32p, i32p]
        L.fy_nmf_get_profile.argtypes = [vp, C.POINTER(NmfProfile)]
        for name in NMF_EXPORTS:
            getattr(L, name)
        _BOUND = True
    return L


class NmfEngine:
    """One context = one GPU.  Thin, 1:1 over the C ABI."""

    def __init__(self, mode, number_of_users, number_of_items, number_of_clusters, number_of_iterations=10,
                 normalization_frequency=12, apply_normalization=False, id_base=1, combine_len=None, split_rows=None,
                 device=0):
        self._L = _lib()
        self._h = C.c_void_p()
        p = NmfParams()
        self._L.fy_nmf_default_params(C.byref(p))
        p.mode, p.number_of_users, p.number_of_items = int(mode), int(number_of_users), int(number_of_items)
        p.number_of_clusters, p.number_of_iterations = int(number_of_clusters), int(number_of_iterations)
        p.normalization_frequency, p.apply_normalization = int(normalization_frequency), int(bool(apply_normalization))
        p.id_base, p.device = int(id_base), int(device)
        if combine_len is not None:
            p.combine_len = int(combine_len)
        if split_rows is not None:
            p.split_rows = int(split_rows)
        rc = self._L.fy_nmf_create(C.byref(self._h), C.byref(p))
        if rc != 0:
            self._h = C.c_void_p()
            raise Rm2Error(rc, "fy_nmf_create failed (is a B200 visible?)")
        self.params = p

    def close(self):
        if self._h:
            self._L.fy_nmf_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def _check(self, rc):
        if rc != 0:
            raise Rm2Error(rc, self._L.fy_nmf_last_error(self._h).decode())

    def set_ratings(self, user, item, score):
        user, item = _i32(user), _i32(item)
        score = np.ascontiguousarray(score, dtype=np.float32)
        self._check(self._L.fy_nmf_set_ratings(self._h, _ptr(user, C.c_int32), _ptr(item, C.c_int32), _ptr(score, C.c_float), len(user)))

    def set_factors(self, H, W):
        p = self.params
        H = np.ascontiguousarray(H, dtype=np.float64); W = np.ascontiguousarray(W, dtype=np.float64)
        if H.shape != (p.number_of_users, p.number_of_clusters) or W.shape != (p.number_of_items, p.number_of_clusters):
            raise Rm2Error(-1, "H must be [numberOfUsers x k] and W [numberOfItems x k]")
        self._check(self._L.fy_nmf_set_factors(self._h, _ptr(H, C.c_double), _ptr(W, C.c_double)))

    def init_random(self, seed=0):
        self._check(self._L.fy_nmf_init_random(self._h, C.c_uint64(seed)))

    def run(self):
        self._check(self._L.fy_nmf_run(self._h))

    def factors(self):
        p = self.params
        H = np.empty((p.number_of_users, p.number_of_clusters), np.float64)
        W = np.empty((p.number_of_items, p.number_of_clusters), np.float64)
        self._check(self._L.fy_nmf_get_factors(self._h, _ptr(H, C.c_double), _ptr(W, C.c_double)))
        return H, W

    def cluster_assignment(self):
        """(clustering[row of H], clusteringCount[k]): what ClusterAssignmentJob + CountClustersJob write."""
        p = self.params
        cl = np.empty(p.number_of_users, np.int32); cnt = np.empty(p.number_of_clusters, np.int32)
        self._check(self._L.fy_nmf_cluster_assignment(self._h, _ptr(cl, C.c_int32), _ptr(cnt, C.c_int32)))
        return cl, cnt

    def run_files(self, input_dir, h_in=None, w_in=None, seed=0, h_out=None, w_out=None, clustering_out=None,
                  clustering_count_out=None):
        """PPCDriver / NMFDriver + ClusterAssignmentJob + CountClustersJob at the file level (SequenceFiles in and out)."""
        enc = lambda s: s.encode() if s else None
        rc = self._L.fy_nmf_run_files(self._h, C.byref(self.params), enc(input_dir), enc(h_in), enc(w_in), C.c_uint64(seed),
                                      enc(h_out), enc(w_out), enc(clustering_out), enc(clustering_count_out))
        if rc != 0:
            raise Rm2Error(rc, self._L.fy_seq_last_error().decode())

    def profile(self):
        p = NmfProfile()
        self._check(self._L.fy_nmf_get_profile(self._h, C.byref(p)))
        return p.as_dict()


def cluster_users(user, item, score, number_of_users, number_of_items, number_of_clusters, number_of_iterations=10,
                  mode=PPC, H=None, W=None, seed=0, device=0, **kw):
    """RMRecommenderDriver's clustering phase (M/rmrecommender/RMRecommenderDriver.java:169-186): PPCDriver, then
    ClusterAssignmentJob and CountClustersJob.  Returns (user ids, clustering, clusteringCount) ready for
    Rm2Engine.set_clustering."""
    with NmfEngine(mode, number_of_users, number_of_items, number_of_clusters, number_of_iterations, device=device, **kw) as eng:
        eng.set_ratings(user, item, score)
        if H is not None and W is not None:
            eng.set_factors(H, W)
        else:
            eng.init_random(seed)
        eng.run()
        cl, cnt = eng.cluster_assignment()
        ids = np.arange(number_of_users, dtype=np.int32) + eng.params.id_base
        return ids, cl, cnt


def sub_cluster_ids(cluster, arg_max, number_of_users, number_of_clusters):
    """ClusterAssignmentJob(true): k' = arg max(h_j) + cluster * ceil(numberOfUsers / numberOfClusters)
    (M/nmf/clustering/FindSubClusterMapper.java:52-77)."""
    stride = -(-int(number_of_users) // int(number_of_clusters))
    return np.asarray(cluster, np.int64) * stride + np.asarray(arg_max, np.int64)


def refine_clusters(user, item, score, ids, clustering, number_of_clusters, users_per_sub_cluster, number_of_iterations=10,
                    seed=0, device=0, **kw):
    """RMRecommenderDriver.clusterRefinement (M/rmrecommender/RMRecommenderDriver.java:217-266): for every cluster, renumber
    its users and the items they rated densely (SubClusterMappingJob), run PPC on that sub-matrix with
    ceil(users / usersPerSubCluster) columns from a random start, and give every user the id
    cluster * ceil(numberOfUsers / numberOfClusters) + arg-max.  Returns (clustering', clusteringCount', sub-clusters made);
    clusteringCount' is indexed by the new ids (0 for ids that do not occur).

    The reference then sets numberOfClusters to the NUMBER of sub-clusters made (:262), smaller than the largest id it
    has just written, so its RM2 job would index clusterSizes[] out of bounds (M/rm/AbstractRM2Reducer.java:93-105);
    here the count array simply covers every id."""
    user, item = np.asarray(user, np.int64), np.asarray(item, np.int64)
    score = np.asarray(score, np.float32)
    ids, clustering = np.asarray(ids, np.int64), np.asarray(clustering, np.int64)
    keep = score > 0
    user, item, score = user[keep], item[keep], score[keep]
    order = np.argsort(ids)
    cl_of_rating = clustering[order][np.searchsorted(ids[order], user)]
    new_cl = np.empty(len(ids), np.int64)
    made = 0
    for c in range(int(number_of_clusters)):
        members = np.sort(ids[clustering == c])
        if len(members) == 0:
            continue
        sel = cl_of_rating == c
        items_c, item_new = np.unique(item[sel], return_inverse=True)
        user_new = np.searchsorted(members, user[sel])
        sub = -(-len(members) // int(users_per_sub_cluster))
        made += sub
        with NmfEngine(PPC, len(members), len(items_c), sub, number_of_iterations, device=device, **kw) as eng:
            eng.set_ratings(user_new + 1, item_new + 1, score[sel])
            eng.init_random(seed + c)
            eng.run()
            arg_max, _ = eng.cluster_assignment()
        new_cl[np.searchsorted(ids[order], members)] = sub_cluster_ids(c, arg_max, len(ids), number_of_clusters)
    out = np.empty(len(ids), np.int64)
    out[order] = new_cl
    return out.astype(np.int32), np.bincount(out, minlength=int(out.max()) + 1).astype(np.int32), made
