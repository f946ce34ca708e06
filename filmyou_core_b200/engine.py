"""ctypes binding of libfilmyou_rm2.so (C ABI: include/filmyou_rm2.h).

The library is built in-tree by `build_library()` (called from __graft_entry__.build()) with
  nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo
and must be present to compute anything: there is no eager / CPU path behind this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_CSRC = os.path.join(_HERE, "csrc")
_SO = os.environ.get("FY_RM2_LIB") or os.path.join(_HERE, "libfilmyou_rm2.so")   # FY_RM2_LIB: A/B-test another build
_LIB = None

STATUS = {
    0: "FY_OK", -1: "FY_E_ARG", -2: "FY_E_USER_WITHOUT_RATING", -3: "FY_E_DUPLICATE_RATING",
    -4: "FY_E_CLUSTER_SIZE", -5: "FY_E_UNKNOWN_USER", -6: "FY_E_NOMEM", -7: "FY_E_CUDA",
    -8: "FY_E_STATE", -9: "FY_E_UNSUPPORTED", -10: "FY_E_ITEM_WITHOUT_RATING",
}

# every symbol include/filmyou_rm2.h declares
EXPORTS = [
    "fy_rm2_abi_version", "fy_rm2_default_params", "fy_rm2_create", "fy_rm2_destroy", "fy_rm2_last_error",
    "fy_rm2_set_stream", "fy_rm2_set_ratings", "fy_rm2_set_clustering", "fy_rm2_run", "fy_rm2_max_item",
    "fy_rm2_stats", "fy_rm2_result_count", "fy_rm2_users_scored", "fy_rm2_results", "fy_rm2_results_device", "fy_rm2_score_group",
    "fy_rm2_get_profile", "fy_cooc_counts", "fy_cooc_topk", "fy_knn_neighbours",
    "fy_rm2_result_row_count", "fy_rm2_result_rows", "fy_rm2_nccl_unique_id", "fy_rm2_comm_init", "fy_rm2_comm_destroy",
    "fy_rm2_shard_bounds", "fy_rm2_probe_plane_read", "fy_rm2_user_count", "fy_rm2_run_neighbours",
]
# include/filmyou_seqfile.h
SEQ_EXPORTS = [
    "fy_seq_last_error", "fy_free", "fy_seq_write_intpair_float", "fy_seq_write_int_int", "fy_seq_write_int_double",
    "fy_mapfile_write_int_double", "fy_seq_read_intpair_float", "fy_seq_read_int_int", "fy_seq_read_int_double",
    "fy_rm2_run_files", "fy_seq_write_int_vector", "fy_seq_read_int_vector", "fy_nmf_run_files",
]
# include/filmyou_nmf.h
NMF_EXPORTS = [
    "fy_nmf_default_params", "fy_nmf_create", "fy_nmf_destroy", "fy_nmf_last_error", "fy_nmf_set_ratings",
    "fy_nmf_set_factors", "fy_nmf_init_random", "fy_nmf_run", "fy_nmf_get_factors", "fy_nmf_cluster_assignment",
    "fy_nmf_get_profile",
]


class Rm2Error(RuntimeError):
    """Mirrors the reference's only error convention: the job fails (M/rm/RM2Job.java:265-268)."""

    def __init__(self, code, text=""):
        super().__init__("%s (%d): %s" % (STATUS.get(code, "?"), code, text))
        self.code = code


class Rm2Params(C.Structure):
    _fields_ = [("lambda_", C.c_double), ("number_of_items", C.c_int32), ("top_n", C.c_int32),
                ("filter_users", C.c_int32), ("device", C.c_int32), ("shard_rank", C.c_int32),
                ("shard_count", C.c_int32), ("tie_break", C.c_int32), ("score_mode", C.c_int32),
                ("n_gpus", C.c_int32), ("reserved", C.c_int32)]


class Rm2Profile(C.Structure):
    _fields_ = [("ms_total", C.c_double), ("ms_index", C.c_double), ("ms_gram", C.c_double),
                ("ms_score", C.c_double), ("ms_topn", C.c_double), ("log_terms", C.c_double),
                ("score_bytes", C.c_double), ("gram_bytes", C.c_double), ("users_scored", C.c_int64),
                ("kernel_launches", C.c_int64), ("clusters_touched", C.c_int32), ("score_launches", C.c_int32),
                ("ms_refine", C.c_double), ("bytes_per_term", C.c_double), ("exact_rerun", C.c_int32),
                ("score_kernel", C.c_int32), ("ms_gather", C.c_double)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


def library_path():
    return _SO


def sources():
    return sorted(os.path.join(_CSRC, f) for f in os.listdir(_CSRC) if f.endswith((".cu", ".cpp")))


def source_hash():
    """sha256 over every source and header the library is built from (file names + contents)."""
    import hashlib
    deps = sources() + sorted(os.path.join(_CSRC, f) for f in os.listdir(_CSRC) if f.endswith(".cuh"))
    deps += [os.path.join(_HERE, "..", "include", h) for h in ("filmyou_rm2.h", "filmyou_seqfile.h", "filmyou_nmf.h")]
    h = hashlib.sha256()
    for d in deps:
        h.update(os.path.basename(d).encode())
        with open(d, "rb") as f:
            h.update(f.read())
    return h.hexdigest()


def build_library(force=False, verbose=False):
    """nvcc cross-compiles for sm_100a without a GPU (about a minute).  The library is rebuilt unless a stamp file next
    to it records the hash of exactly these sources (mtimes do not survive a copy to the GPU box and prove nothing);
    says which of the two happened."""
    want = source_hash()
    stamp = _SO + ".srchash"
    have = open(stamp).read().strip() if os.path.exists(stamp) else ""
    if not force and os.path.exists(_SO) and have == want:
        if verbose:
            print("libfilmyou_rm2.so: reused (source hash %s matches the stamp)" % want[:16])
        return _SO
    cmd = ["nvcc", "-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
           "-Xcompiler", "-fPIC", "-shared", "-o", _SO] + sources() + ["-ldl"]
    if verbose:
        print(" ".join(cmd))
    subprocess.check_call(cmd)
    with open(stamp, "w") as f:
        f.write(want + "\n")
    if verbose:
        print("libfilmyou_rm2.so: compiled from source (source hash %s)" % want[:16])
    return _SO


def load_library():
    """Load libfilmyou_rm2.so; raises (loudly) when it has not been built."""
    global _LIB
    if _LIB is not None:
        return _LIB
    if not os.path.exists(_SO):
        raise Rm2Error(-7, "libfilmyou_rm2.so is missing: run __graft_entry__.build() (no CPU fallback exists)")
    L = C.CDLL(_SO)
    i32p, f32p, f64p = C.POINTER(C.c_int32), C.POINTER(C.c_float), C.POINTER(C.c_double)
    vp = C.c_void_p
    L.fy_rm2_abi_version.restype = C.c_int
    L.fy_rm2_default_params.argtypes = [C.POINTER(Rm2Params)]
    L.fy_rm2_default_params.restype = None
    L.fy_rm2_create.argtypes = [C.POINTER(vp), C.POINTER(Rm2Params)]
    L.fy_rm2_destroy.argtypes = [vp]
    L.fy_rm2_destroy.restype = None
    L.fy_rm2_last_error.argtypes = [vp]
    L.fy_rm2_last_error.restype = C.c_char_p
    L.fy_rm2_set_stream.argtypes = [vp, vp]
    L.fy_rm2_set_ratings.argtypes = [vp, i32p, i32p, f32p, C.c_int64]
    L.fy_rm2_set_clustering.argtypes = [vp, i32p, i32p, C.c_int64, i32p, C.c_int32]
    L.fy_rm2_run.argtypes = [vp]
    L.fy_rm2_max_item.argtypes = [vp]
    L.fy_rm2_max_item.restype = C.c_int32
    L.fy_rm2_stats.argtypes = [vp, f64p, f64p, f64p]
    L.fy_rm2_result_count.argtypes = [vp]
    L.fy_rm2_result_count.restype = C.c_int64
    L.fy_rm2_users_scored.argtypes = [vp]
    L.fy_rm2_users_scored.restype = C.c_int64
    L.fy_rm2_results.argtypes = [vp, i32p, i32p, f64p, f32p, i32p]
    L.fy_rm2_results_device.argtypes = [vp] + [C.POINTER(vp)] * 5
    L.fy_rm2_score_group.argtypes = [vp, C.c_int32, C.c_int32, C.c_int32, i32p, f64p, C.c_int32,
                                     i32p, i32p, f32p, C.c_int64, f64p, C.c_int32]
    L.fy_rm2_get_profile.argtypes = [vp, C.POINTER(Rm2Profile)]
    L.fy_rm2_result_row_count.argtypes = [vp]
    L.fy_rm2_result_row_count.restype = C.c_int64
    L.fy_rm2_result_rows.argtypes = [vp, i32p, i32p, i32p]
    L.fy_rm2_nccl_unique_id.argtypes = [vp]
    L.fy_rm2_comm_init.argtypes = [vp, vp, C.c_int32, C.c_int32]
    L.fy_rm2_comm_destroy.argtypes = [vp]
    L.fy_rm2_shard_bounds.argtypes = [vp, i32p]
    L.fy_rm2_run_neighbours.argtypes = [vp, i32p, i32p, C.c_int32, C.c_int64]
    L.fy_rm2_user_count.argtypes = [vp]
    L.fy_rm2_user_count.restype = C.c_int64
    L.fy_rm2_probe_plane_read.argtypes = [vp, C.c_int32, C.c_int32, C.c_int32, C.c_int32, f64p, f64p]
    L.fy_cooc_counts.argtypes = [vp, C.c_int32, C.c_int32, i32p, f64p]
    L.fy_cooc_topk.argtypes = [vp, C.c_int32, i32p, i32p, i32p]
    L.fy_knn_neighbours.argtypes = [vp, C.c_int32, C.c_int32, C.c_int32, i32p, i32p, i32p, f64p]
    L.fy_seq_last_error.restype = C.c_char_p
    L.fy_free.argtypes = [vp]
    L.fy_free.restype = None
    L.fy_seq_write_intpair_float.argtypes = [C.c_char_p, i32p, i32p, f32p, C.c_int64]
    L.fy_seq_write_int_int.argtypes = [C.c_char_p, i32p, i32p, C.c_int64]
    L.fy_seq_write_int_double.argtypes = [C.c_char_p, i32p, f64p, C.c_int64]
    L.fy_mapfile_write_int_double.argtypes = [C.c_char_p, i32p, f64p, C.c_int64]
    L.fy_seq_read_intpair_float.argtypes = [C.c_char_p, C.POINTER(i32p), C.POINTER(i32p), C.POINTER(f32p), C.POINTER(C.c_int64)]
    L.fy_seq_read_int_int.argtypes = [C.c_char_p, C.POINTER(i32p), C.POINTER(i32p), C.POINTER(C.c_int64)]
    L.fy_seq_read_int_double.argtypes = [C.c_char_p, C.POINTER(i32p), C.POINTER(f64p), C.POINTER(C.c_int64)]
    L.fy_rm2_run_files.argtypes = [vp, C.c_char_p, C.c_char_p, C.c_char_p, C.c_int32, C.c_char_p, C.c_char_p]
    L.fy_seq_write_int_vector.argtypes = [C.c_char_p, i32p, f64p, C.c_int64, C.c_int32]
    L.fy_seq_read_int_vector.argtypes = [C.c_char_p, C.POINTER(i32p), C.POINTER(f64p), C.POINTER(C.c_int64), C.POINTER(C.c_int32)]
    L.fy_nmf_run_files.argtypes = [vp, vp, C.c_char_p, C.c_char_p, C.c_char_p, C.c_uint64, C.c_char_p, C.c_char_p, C.c_char_p, C.c_char_p]
    for name in EXPORTS + SEQ_EXPORTS:
        getattr(L, name)
    _LIB = L
    return L


def _ptr(a, t):
    return a.ctypes.data_as(C.POINTER(t))


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


class _DeviceArray:
    def __init__(self, ptr, n, typestr):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": typestr, "data": (ptr, False), "version": 2,
                                         "strides": None}


class Rm2Engine:
    """One context = one GPU.  Thin, 1:1 over the C ABI."""

    def __init__(self, lam=0.1, number_of_items=0, top_n=1000, filter_users=0, device=0,
                 shard_rank=0, shard_count=1, score_mode=0, n_gpus=0):
        self._L = load_library()
        self._h = C.c_void_p()
        p = Rm2Params()
        self._L.fy_rm2_default_params(C.byref(p))
        p.lambda_, p.number_of_items, p.top_n, p.filter_users = float(lam), int(number_of_items), int(top_n), int(filter_users)
        p.device, p.shard_rank, p.shard_count = int(device), int(shard_rank), int(shard_count)
        p.score_mode = int(score_mode)
        p.n_gpus = int(n_gpus)
        rc = self._L.fy_rm2_create(C.byref(self._h), C.byref(p))
        if rc != 0:
            self._h = C.c_void_p()
            raise Rm2Error(rc, "fy_rm2_create failed (is a B200 visible?)")
        self.params = p
        self._n_users = 0

    def close(self):
        if self._h:
            self._L.fy_rm2_destroy(self._h)
            self._h = C.c_void_p()
            if hasattr(self._L, "fy_rm2_debug_violations"):           # a -DFY_BOUNDS_CHECK build (tools/build_checked.py)
                self._L.fy_rm2_debug_violations.restype = C.c_longlong
                v = self._L.fy_rm2_debug_violations()
                if v != 0:
                    raise Rm2Error(-7, "%d device-side bounds violations counted by the checked build" % v)

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
            raise Rm2Error(rc, self._L.fy_rm2_last_error(self._h).decode())

    def set_stream(self, cuda_stream_ptr):
        self._check(self._L.fy_rm2_set_stream(self._h, C.c_void_p(cuda_stream_ptr)))

    def set_ratings(self, user, item, score):
        user, item = _i32(user), _i32(item)
        score = np.ascontiguousarray(score, dtype=np.float32)
        self._check(self._L.fy_rm2_set_ratings(self._h, _ptr(user, C.c_int32), _ptr(item, C.c_int32),
                                                _ptr(score, C.c_float), len(user)))

    def set_clustering(self, user, cluster, cluster_size):
        user, cluster, cluster_size = _i32(user), _i32(cluster), _i32(cluster_size)
        self._n_users = len(user)
        self._check(self._L.fy_rm2_set_clustering(self._h, _ptr(user, C.c_int32), _ptr(cluster, C.c_int32),
                                                   len(user), _ptr(cluster_size, C.c_int32), len(cluster_size)))

    def run(self):
        self._check(self._L.fy_rm2_run(self._h))

    def stats(self):
        """(user_sum in set_clustering order, item_prob by item id, total)"""
        mi = self._L.fy_rm2_max_item(self._h)
        us = np.zeros(self._n_users, np.float64)
        ip = np.zeros(mi + 1, np.float64)
        tot = C.c_double(0)
        self._check(self._L.fy_rm2_stats(self._h, _ptr(us, C.c_double), _ptr(ip, C.c_double), C.byref(tot)))
        return us, ip, tot.value

    def result_count(self):
        return int(self._L.fy_rm2_result_count(self._h))

    def users_scored(self):
        return int(self._L.fy_rm2_users_scored(self._h))

    def results(self, out=None):
        """dict(user, item, score64, score32, cluster) of packed triples; `out` may hold preallocated
        (e.g. pinned) arrays of at least result_count() entries."""
        n = self.result_count()
        if n < 0:
            raise Rm2Error(-8, "no results")
        if out is None:
            out = dict(user=np.empty(n, np.int32), item=np.empty(n, np.int32), score64=np.empty(n, np.float64),
                       score32=np.empty(n, np.float32), cluster=np.empty(n, np.int32))
        self._check(self._L.fy_rm2_results(self._h, _ptr(out["user"], C.c_int32), _ptr(out["item"], C.c_int32),
                                            _ptr(out["score64"], C.c_double), _ptr(out["score32"], C.c_float),
                                            _ptr(out["cluster"], C.c_int32)))
        return {k: v[:n] for k, v in out.items()}

    def result_rows(self):
        """(user, cluster, count) per scored row, rows in the order of the packed triples."""
        n = int(self._L.fy_rm2_result_row_count(self._h))
        if n < 0:
            raise Rm2Error(-8, "no results")
        u, c, k = np.empty(n, np.int32), np.empty(n, np.int32), np.empty(n, np.int32)
        self._check(self._L.fy_rm2_result_rows(self._h, _ptr(u, C.c_int32), _ptr(c, C.c_int32), _ptr(k, C.c_int32)))
        return u, c, k

    def results_compact(self, out=None):
        """The 12-byte-per-triple read-back: dict(item, score64, row_user, row_cluster, row_count); `out` may hold
        preallocated (pinned) item / score64 arrays.  expand_compact() rebuilds the five packed arrays."""
        n = self.result_count()
        if n < 0:
            raise Rm2Error(-8, "no results")
        if out is None:
            out = dict(item=np.empty(n, np.int32), score64=np.empty(n, np.float64))
        self._check(self._L.fy_rm2_results(self._h, None, _ptr(out["item"], C.c_int32), _ptr(out["score64"], C.c_double), None, None))
        u, c, k = self.result_rows()
        return dict(item=out["item"][:n], score64=out["score64"][:n], row_user=u, row_cluster=c, row_count=k)

    @staticmethod
    def expand_compact(r):
        """(item, score64, rows) -> the packed triples fy_rm2_results returns (user, item, score64, score32, cluster)."""
        return dict(user=np.repeat(r["row_user"], r["row_count"]), item=r["item"], score64=r["score64"],
                    score32=r["score64"].astype(np.float32), cluster=np.repeat(r["row_cluster"], r["row_count"]))

    @staticmethod
    def nccl_unique_id():
        """128 bytes from ncclGetUniqueId (rank 0 calls this and hands the bytes to every rank)."""
        buf = (C.c_ubyte * 128)()
        rc = load_library().fy_rm2_nccl_unique_id(C.cast(buf, C.c_void_p))
        if rc != 0:
            raise Rm2Error(rc, "fy_rm2_nccl_unique_id failed (is libnccl.so.2 loadable?)")
        return bytes(buf)

    def comm_init(self, unique_id, world, rank):
        """Attach an NCCL communicator: fy_rm2_run then ends with the in-library exchange of the top-N blocks and
        results() on every rank holds the whole job."""
        buf = (C.c_ubyte * 128).from_buffer_copy(bytes(unique_id))
        self._check(self._L.fy_rm2_comm_init(self._h, C.cast(buf, C.c_void_p), int(world), int(rank)))

    def comm_destroy(self):
        self._check(self._L.fy_rm2_comm_destroy(self._h))

    def shard_bounds(self):
        b = np.zeros(int(self.params.shard_count if self.params.n_gpus <= 1 else self.params.n_gpus) + 1, np.int32)
        self._check(self._L.fy_rm2_shard_bounds(self._h, _ptr(b, C.c_int32)))
        return b

    def probe_plane_read(self, n_rows, n_users, rows_per_user, reps=5):
        """(GB/s, ms per launch) of the score kernel's access pattern without its arithmetic (roofline probe)."""
        g, ms = C.c_double(0), C.c_double(0)
        self._check(self._L.fy_rm2_probe_plane_read(self._h, int(n_rows), int(n_users), int(rows_per_user), int(reps),
                                                     C.byref(g), C.byref(ms)))
        return g.value, ms.value

    def results_device(self):
        """field -> object exposing __cuda_array_interface__ over the device-resident packed arrays
        (zero copy; wrap with torch.as_tensor(x, device="cuda")).  Valid until the next run()."""
        n = self.result_count()
        ptrs = [C.c_void_p() for _ in range(5)]
        self._check(self._L.fy_rm2_results_device(self._h, *[C.byref(p) for p in ptrs]))
        spec = (("user", "<i4"), ("item", "<i4"), ("score64", "<f8"), ("score32", "<f4"), ("cluster", "<i4"))
        return {name: _DeviceArray(p.value or 0, n, ts) for (name, ts), p in zip(spec, ptrs)}

    def score_group(self, cluster_id, split, n_splits, group_user, group_user_sum, r_user, r_item, r_score, item_prob):
        group_user, r_user, r_item = _i32(group_user), _i32(r_user), _i32(r_item)
        group_user_sum = np.ascontiguousarray(group_user_sum, dtype=np.float64)
        r_score = np.ascontiguousarray(r_score, dtype=np.float32)
        item_prob = np.ascontiguousarray(item_prob, dtype=np.float64)
        self._check(self._L.fy_rm2_score_group(
            self._h, int(cluster_id), int(split), int(n_splits), _ptr(group_user, C.c_int32),
            _ptr(group_user_sum, C.c_double), len(group_user), _ptr(r_user, C.c_int32), _ptr(r_item, C.c_int32),
            _ptr(r_score, C.c_float), len(r_user), _ptr(item_prob, C.c_double), len(item_prob) - 1))

    def run_neighbours(self, user, neighbour):
        """Score `user[q]` over the explicit neighbour list `neighbour[q]` ([n, k] user ids, -1 = empty slot)."""
        user = _i32(user)
        neighbour = np.ascontiguousarray(neighbour, dtype=np.int32).reshape(len(user), -1)
        self._check(self._L.fy_rm2_run_neighbours(self._h, _ptr(user, C.c_int32), _ptr(neighbour, C.c_int32),
                                                   neighbour.shape[1], len(user)))

    def run_files(self, input_dir, clustering_dir, clustering_count_dir, number_of_clusters, output_dir, rm2_dir=None):
        """RM2Job.run at the file level (SequenceFiles in, SequenceFile / MapFile out)."""
        rc = self._L.fy_rm2_run_files(self._h, input_dir.encode(), clustering_dir.encode(), clustering_count_dir.encode(),
                                      int(number_of_clusters), output_dir.encode(), rm2_dir.encode() if rm2_dir else None)
        if rc != 0:
            raise Rm2Error(rc, self._L.fy_seq_last_error().decode())

    def profile(self):
        p = Rm2Profile()
        self._check(self._L.fy_rm2_get_profile(self._h, C.byref(p)))
        return p.as_dict()

    def cooc_counts(self, n_user_ids, n_items, want_counts=True):
        out = np.zeros((n_items, n_items), np.int32) if want_counts else None
        ms = C.c_double(0)
        self._check(self._L.fy_cooc_counts(self._h, int(n_user_ids), int(n_items),
                                            _ptr(out, C.c_int32) if want_counts else None, C.byref(ms)))
        return out, ms.value

    def knn_neighbours(self, n_user_ids, n_items, k):
        """(neighbours[n_user_ids, k] (-1 padded), counts, n, ms_gemm): top-k co-rating users per user id."""
        nb = np.zeros((n_user_ids, k), np.int32)
        cnt = np.zeros((n_user_ids, k), np.int32)
        n = np.zeros(n_user_ids, np.int32)
        ms = C.c_double(0)
        self._check(self._L.fy_knn_neighbours(self._h, int(n_user_ids), int(n_items), int(k), _ptr(nb, C.c_int32),
                                               _ptr(cnt, C.c_int32), _ptr(n, C.c_int32), C.byref(ms)))
        return nb, cnt, n, ms.value

    def cooc_topk(self, n_items, k):
        items = np.zeros((n_items, k), np.int32)
        counts = np.zeros((n_items, k), np.int32)
        n = np.zeros(n_items, np.int32)
        self._check(self._L.fy_cooc_topk(self._h, int(k), _ptr(items, C.c_int32), _ptr(counts, C.c_int32), _ptr(n, C.c_int32)))
        return items, counts, n
