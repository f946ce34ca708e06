"""CPU suite: Hadoop SequenceFile / MapFile readers and writers.  No serialized fixture
exists in the reference (its tests write the files at run time with Hadoop itself), so the byte layout is
checked against an image assembled here, independently, from the published format."""
import os
import struct

import numpy as np
import pytest

import filmyou_core_b200 as fy
from filmyou_core_b200 import seqfile


def _text(s):
    b = s.encode()
    assert len(b) < 128
    return bytes([len(b)]) + b


def _header(key_class, val_class, sync):
    return b"SEQ\x06" + _text(key_class) + _text(val_class) + b"\x00\x00" + struct.pack(">i", 0) + sync


def test_int_int_byte_image(tmp_path):
    # what DataInitialization.createIntIntFileParent(conf, {1,3,0}, dir, "clustering", 1) appends:
    # (IntWritable(i + start), IntWritable(data[i]))    M/util/DataInitialization.java:214-216
    p = str(tmp_path / "data")
    seqfile.write_int_int(p, [1, 2, 3], [1, 3, 0])
    raw = open(p, "rb").read()
    hdr_len = len(_header("org.apache.hadoop.io.IntWritable", "org.apache.hadoop.io.IntWritable", b"\x00" * 16))
    sync = raw[hdr_len - 16:hdr_len]
    want = _header("org.apache.hadoop.io.IntWritable", "org.apache.hadoop.io.IntWritable", sync)
    for k, v in ((1, 1), (2, 3), (3, 0)):
        want += struct.pack(">iiii", 8, 4, k, v)         # record length, key length, key, value (big-endian)
    assert raw == want


def test_intpair_float_byte_image(tmp_path):
    # (IntPairWritable(j + 1, i + 1), FloatWritable((float) data[i][j]))   M/util/DataInitialization.java:171-172
    p = str(tmp_path / "data")
    seqfile.write_intpair_float(p, [7, -2], [100, 5], [4.5, 1.0])
    raw = open(p, "rb").read()
    hdr_len = len(_header("org.apache.mahout.common.IntPairWritable", "org.apache.hadoop.io.FloatWritable", b"\x00" * 16))
    want = _header("org.apache.mahout.common.IntPairWritable", "org.apache.hadoop.io.FloatWritable", raw[hdr_len - 16:hdr_len])
    want += struct.pack(">iiiif", 12, 8, 7, 100, 4.5) + struct.pack(">iiiif", 12, 8, -2, 5, 1.0)
    assert raw == want


def test_round_trips_with_sync_markers(tmp_path):
    rng = np.random.default_rng(0)
    n = 5000                                              # 20 bytes/record -> a sync escape every 100 records
    u, i = rng.integers(-2**31, 2**31 - 1, n), rng.integers(0, 10**6, n)
    s = rng.random(n).astype(np.float32) * 5
    p = str(tmp_path / "ratings")
    seqfile.write_intpair_float(p, u, i, s)
    raw = open(p, "rb").read()
    assert raw.count(struct.pack(">i", -1) + raw[raw.index(b"\x00\x00\x00\x00\x00\x00") + 6:][:16]) >= 40
    ru, ri, rs = seqfile.read_intpair_float(p)
    assert np.array_equal(ru, u.astype(np.int32)) and np.array_equal(ri, i.astype(np.int32)) and np.array_equal(rs, s)
    d = rng.standard_normal(n)
    p2 = str(tmp_path / "usersum")
    seqfile.write_int_double(p2, i, d)
    rk, rd = seqfile.read_int_double(p2)
    assert np.array_equal(rk, i.astype(np.int32)) and np.array_equal(rd, d)


def test_directories_part_files_and_mapfile(tmp_path):
    d = tmp_path / "out"
    d.mkdir()
    seqfile.write_int_int(str(d / "part-r-00000"), [1, 2], [10, 20])
    seqfile.write_int_int(str(d / "part-r-00001"), [3], [30])
    (d / "_SUCCESS").write_bytes(b"")
    (d / ".part-r-00000.crc").write_bytes(b"junk")
    k, v = seqfile.read_int_int(str(d))
    assert k.tolist() == [1, 2, 3] and v.tolist() == [10, 20, 30]
    # MapFile: data + index, index entry every 128 keys pointing at the record's position in data
    keys = np.arange(5, 5 + 1000, dtype=np.int32)
    vals = np.linspace(0, 1, 1000)
    m = tmp_path / "itemColl" / "part-r-00000"
    seqfile.write_mapfile_int_double(str(m), keys, vals)
    rk, rv = seqfile.read_int_double(str(tmp_path / "itemColl"))       # directory of MapFiles -> their data files
    assert np.array_equal(rk, keys) and np.array_equal(rv, vals)
    raw_index = open(str(m / "index"), "rb").read()
    assert b"org.apache.hadoop.io.LongWritable" in raw_index
    data = open(str(m / "data"), "rb").read()
    body = raw_index[raw_index.index(b"\x00\x00\x00\x00\x00\x00") + 6 + 16:]
    entries = [struct.unpack(">iiiq", body[o:o + 20]) for o in range(0, len(body), 20)]
    assert [e[2] for e in entries] == keys[::128].tolist()
    for _, _, key, pos in entries:
        rec = data[pos:]
        if struct.unpack(">i", rec[:4])[0] == -1:          # the index may point at the sync escape before the record
            rec = rec[20:]
        assert struct.unpack(">iii", rec[:12]) == (12, 4, key)
    with pytest.raises(fy.Rm2Error):
        seqfile.write_mapfile_int_double(str(tmp_path / "bad"), [3, 2], [0.0, 1.0])


def test_rejects_what_it_cannot_read(tmp_path):
    p = str(tmp_path / "f")
    seqfile.write_int_int(p, [1], [2])
    raw = bytearray(open(p, "rb").read())
    with pytest.raises(fy.Rm2Error) as e:                  # wrong record types
        seqfile.read_int_double(p)
    assert e.value.code == -1
    flag = raw.index(b"IntWritable", raw.index(b"IntWritable") + 1) + len(b"IntWritable")
    comp = bytearray(raw); comp[flag] = 1                  # "compressed" header flag
    open(p + ".c", "wb").write(comp)
    with pytest.raises(fy.Rm2Error) as e:
        seqfile.read_int_int(p + ".c")
    assert e.value.code == -9
    open(p + ".t", "wb").write(raw[:-3])                   # truncated record
    with pytest.raises(fy.Rm2Error):
        seqfile.read_int_int(p + ".t")
    with pytest.raises(fy.Rm2Error):
        seqfile.read_int_int(str(tmp_path / "missing"))


def _uvarint(v):
    out = b""
    while v >= 0x80:
        out += bytes([(v & 0x7f) | 0x80]); v >>= 7
    return out + bytes([v])


def test_int_vector_byte_image_and_round_trip(tmp_path):
    # what DataInitialization.createDoubleMatrix(conf, data, dir, "H", 1) appends: (IntWritable(i), VectorWritable(DenseVector))
    # M/util/DataInitialization.java:127-131; Mahout 0.8 VectorWritable: flags 0x03 (dense | sequential), varint size, doubles
    p = str(tmp_path / "H")
    rows = np.array([[0.25, 0.75], [1e-12, 3.0]])
    seqfile.write_int_vector(p, [1, 2], rows)
    raw = open(p, "rb").read()
    hdr_len = len(_header("org.apache.hadoop.io.IntWritable", "org.apache.mahout.math.VectorWritable", b"\x00" * 16))
    want = _header("org.apache.hadoop.io.IntWritable", "org.apache.mahout.math.VectorWritable", raw[hdr_len - 16:hdr_len])
    for k, r in zip((1, 2), rows):
        val = b"\x03" + _uvarint(2) + struct.pack(">dd", *r)
        want += struct.pack(">iii", 4 + len(val), 4, k) + val
    assert raw == want
    rng = np.random.default_rng(3)
    big = rng.random((700, 200))                           # size 200 needs a two-byte varint; 1.6 KB records -> sync escapes
    keys = rng.permutation(700).astype(np.int32) + 1
    seqfile.write_int_vector(p + "2", keys, big)
    k2, r2 = seqfile.read_int_vector(p + "2")
    assert np.array_equal(k2, keys) and np.array_equal(r2, big)


def test_int_vector_reader_accepts_sparse_and_lax_vectors(tmp_path):
    hdr = _header("org.apache.hadoop.io.IntWritable", "org.apache.mahout.math.VectorWritable", bytes(range(16)))
    recs = []
    # SequentialAccessSparseVector (flags 0x02): varint nnz, then (index delta, double)
    recs.append((5, b"\x02" + _uvarint(6) + _uvarint(2) + _uvarint(1) + struct.pack(">d", 1.5) + _uvarint(3) + struct.pack(">d", -2.0)))
    # RandomAccessSparseVector (flags 0x00): plain indices
    recs.append((6, b"\x00" + _uvarint(6) + _uvarint(2) + _uvarint(5) + struct.pack(">d", 7.0) + _uvarint(0) + struct.pack(">d", 8.0)))
    # dense, lax precision (flags 0x09): floats
    recs.append((7, b"\x09" + _uvarint(6) + struct.pack(">6f", 1, 2, 3, 4, 5, 6)))
    body = b"".join(struct.pack(">iii", 4 + len(v), 4, k) + v for k, v in recs)
    p = str(tmp_path / "V")
    open(p, "wb").write(hdr + body)
    k, r = seqfile.read_int_vector(p)
    assert k.tolist() == [5, 6, 7]
    assert r.tolist() == [[0, 1.5, 0, 0, -2.0, 0], [8.0, 0, 0, 0, 0, 7.0], [1, 2, 3, 4, 5, 6]]
    open(p + "bad", "wb").write(hdr + struct.pack(">iii", 4 + 3, 4, 1) + b"\x13" + _uvarint(1) + b"\x00")
    with pytest.raises(fy.Rm2Error):
        seqfile.read_int_vector(p + "bad")
