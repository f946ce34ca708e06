"""ctypes wrappers of the SequenceFile / MapFile readers and writers (include/filmyou_seqfile.h)."""
import ctypes as C

import numpy as np

from .engine import load_library, Rm2Error


def _chk(rc):
    if rc != 0:
        raise Rm2Error(rc, load_library().fy_seq_last_error().decode())


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t))


def write_intpair_float(path, first, second, value):
    first, second, value = _i32(first), _i32(second), np.ascontiguousarray(value, dtype=np.float32)
    _chk(load_library().fy_seq_write_intpair_float(path.encode(), _p(first, C.c_int32), _p(second, C.c_int32), _p(value, C.c_float), len(first)))


def write_int_int(path, key, value):
    key, value = _i32(key), _i32(value)
    _chk(load_library().fy_seq_write_int_int(path.encode(), _p(key, C.c_int32), _p(value, C.c_int32), len(key)))


def write_int_double(path, key, value):
    key, value = _i32(key), np.ascontiguousarray(value, dtype=np.float64)
    _chk(load_library().fy_seq_write_int_double(path.encode(), _p(key, C.c_int32), _p(value, C.c_double), len(key)))


def write_mapfile_int_double(directory, key, value):
    key, value = _i32(key), np.ascontiguousarray(value, dtype=np.float64)
    _chk(load_library().fy_mapfile_write_int_double(directory.encode(), _p(key, C.c_int32), _p(value, C.c_double), len(key)))


def _take(ptr, n, dtype):
    L = load_library()
    out = np.ctypeslib.as_array(ptr, shape=(max(n, 1),))[:n].astype(dtype, copy=True)
    L.fy_free(ptr)
    return out


def read_intpair_float(path):
    L = load_library()
    a, b, v, n = C.POINTER(C.c_int32)(), C.POINTER(C.c_int32)(), C.POINTER(C.c_float)(), C.c_int64(0)
    _chk(L.fy_seq_read_intpair_float(path.encode(), C.byref(a), C.byref(b), C.byref(v), C.byref(n)))
    return _take(a, n.value, np.int32), _take(b, n.value, np.int32), _take(v, n.value, np.float32)


def read_int_int(path):
    L = load_library()
    a, b, n = C.POINTER(C.c_int32)(), C.POINTER(C.c_int32)(), C.c_int64(0)
    _chk(L.fy_seq_read_int_int(path.encode(), C.byref(a), C.byref(b), C.byref(n)))
    return _take(a, n.value, np.int32), _take(b, n.value, np.int32)


def read_int_double(path):
    L = load_library()
    a, b, n = C.POINTER(C.c_int32)(), C.POINTER(C.c_double)(), C.c_int64(0)
    _chk(L.fy_seq_read_int_double(path.encode(), C.byref(a), C.byref(b), C.byref(n)))
    return _take(a, n.value, np.int32), _take(b, n.value, np.float64)


def write_int_vector(path, key, rows):
    """SequenceFile<IntWritable, VectorWritable> of dense vectors: the H / W files (DataInitialization.java:113-138)."""
    key, rows = _i32(key), np.ascontiguousarray(rows, dtype=np.float64)
    assert rows.ndim == 2 and rows.shape[0] == len(key)
    _chk(load_library().fy_seq_write_int_vector(path.encode(), _p(key, C.c_int32), _p(rows, C.c_double), len(key), rows.shape[1]))


def read_int_vector(path):
    """(keys, rows[n x cols]) in file order"""
    L = load_library()
    a, b, n, c = C.POINTER(C.c_int32)(), C.POINTER(C.c_double)(), C.c_int64(0), C.c_int32(0)
    _chk(L.fy_seq_read_int_vector(path.encode(), C.byref(a), C.byref(b), C.byref(n), C.byref(c)))
    keys = _take(a, n.value, np.int32)
    rows = _take(b, n.value * c.value, np.float64).reshape(n.value, c.value)
    return keys, rows
