// seqfile.cpp -- Hadoop SequenceFile / MapFile readers and writers for the record types on either side
// of the RM2 path, host only, so that the coarse seam is a drop-in at the
// file level without a JVM:
//   ratings in / recommendations out : SequenceFile<IntPairWritable(user,item), FloatWritable>
//        M/util/DataInitialization.java:155-174 (fixture writer), M/rm/RM2HDFSReducer.java:44-50 (sink)
//   clustering, clusteringCount      : SequenceFile<IntWritable, IntWritable>
//        M/util/DataInitialization.java:200-222, M/common/AbstractByClusterMapper.java:57-66
//   rm2/userSum                      : SequenceFile<IntWritable, DoubleWritable>     M/rm/RM2Job.java:138-142
//   rm2/itemColl                     : MapFile<IntWritable, DoubleWritable>          M/rm/RM2Job.java:190-196
//
// The byte formats are Hadoop 1.2.1's (SequenceFile version 6, uncompressed records) and Mahout 0.8's
// IntPairWritable (two big-endian ints); neither library is vendored under /root/reference and the
// reference holds no serialized fixture (its tests write them at run time), so the layout below is a
// restatement of the published formats -- PARITY UNPINNED at the byte level; tests check round trips
// and an independently assembled byte image.  What each piece restates (Apache sources, from memory of the
// published code; line numbers are those of the tagged releases and may be off by a few lines):
//   header + records : hadoop-1.2.1 src/core/org/apache/hadoop/io/SequenceFile.java -- VERSION = {'S','E','Q',6} (:~190),
//                      Writer.writeFileHeader (:~980: version, Text key class, Text value class, compression flags,
//                      Metadata, sync), Writer.append(Object, Object) (:~1010: checkAndWriteSync, int recordLength,
//                      int keyLength, key, value), SYNC_ESCAPE = -1, SYNC_HASH_SIZE = 16, SYNC_SIZE = 20,
//                      SYNC_INTERVAL = 100 * SYNC_SIZE (:~200), checkAndWriteSync (:~990: a marker once
//                      out.getPos() >= lastSyncPos + SYNC_INTERVAL)
//   Text / VInt      : hadoop-1.2.1 .../io/Text.java writeString / write (:~280: VInt length + UTF-8),
//                      .../io/WritableUtils.java writeVLong / readVLong / decodeVIntSize (:~260-330)
//   IntPairWritable  : mahout-core-0.8 org/apache/mahout/common/IntPairWritable.java -- a 2 * 4 byte array b[]; set() -> putInt
//                      stores each int most significant byte first (b[i] = (byte) (value >> j), j = 24, 16, 8, 0); the sign is
//                      handled by the raw comparator (compareInts), not by the stored bytes; write(DataOutput) = out.write(b)
//   VectorWritable   : mahout-core-0.8 org/apache/mahout/math/VectorWritable.java writeVector (:~100: flag byte
//                      FLAG_DENSE 0x01 | FLAG_SEQUENTIAL 0x02 | FLAG_NAMED 0x04 | FLAG_LAX_PRECISION 0x08, VInt size,
//                      dense: doubles; sparse: VInt nondefault count, then (VInt index [delta-coded if sequential], value))
//   MapFile          : hadoop-1.2.1 .../io/MapFile.java -- directory with DATA_FILE_NAME "data" + INDEX_FILE_NAME "index"
//                      (:~50), Writer.append (:~190: every indexInterval = 128-th key goes to the index with the data
//                      file position as a LongWritable)
// If a byte differs from what a real Hadoop writer emits, these are the places to compare.
//
//   header : 'S' 'E' 'Q' 6 | Text keyClass | Text valueClass | bool compressed | bool blockCompressed |
//            int32 metadataCount (+ Text pairs) | 16-byte sync marker
//   record : int32 recordLength | int32 keyLength | key bytes | value bytes           (all big-endian)
//   sync   : int32 -1 | 16-byte sync marker, written before a record once 2000 bytes have passed
//   Text   : Hadoop VInt length + UTF-8 bytes
#include "../../include/filmyou_rm2.h"
#include "../../include/filmyou_nmf.h"

#include <dirent.h>
#include <sys/stat.h>

#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

namespace {

const char* K_INTPAIR = "org.apache.mahout.common.IntPairWritable";
const char* K_INT = "org.apache.hadoop.io.IntWritable";
const char* K_FLOAT = "org.apache.hadoop.io.FloatWritable";
const char* K_DOUBLE = "org.apache.hadoop.io.DoubleWritable";
const char* K_LONG = "org.apache.hadoop.io.LongWritable";
const char* K_VECTOR = "org.apache.mahout.math.VectorWritable";
const int SYNC_INTERVAL = 100 * (4 + 16);      // SequenceFile.SYNC_INTERVAL

thread_local char g_err[512] = {0};
int fail(int code, const char* fmt, const char* a = "", const char* b = "") {
    snprintf(g_err, sizeof(g_err), fmt, a, b);
    return code;
}

void put_be32(std::vector<uint8_t>& o, uint32_t v) { for (int s = 24; s >= 0; s -= 8) o.push_back((uint8_t)(v >> s)); }
void put_be64(std::vector<uint8_t>& o, uint64_t v) { for (int s = 56; s >= 0; s -= 8) o.push_back((uint8_t)(v >> s)); }
uint32_t get_be32(const uint8_t* p) { return ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | p[3]; }
uint64_t get_be64(const uint8_t* p) { return ((uint64_t)get_be32(p) << 32) | get_be32(p + 4); }

// Hadoop WritableUtils.writeVLong for non-negative lengths
void put_vint(std::vector<uint8_t>& o, int64_t v) {
    if (v >= -112 && v <= 127) { o.push_back((uint8_t)v); return; }
    int len = -112;
    if (v < 0) { v ^= -1LL; len = -120; }
    int64_t tmp = v;
    while (tmp != 0) { tmp >>= 8; len--; }
    o.push_back((uint8_t)len);
    len = (len < -120) ? -(len + 120) : -(len + 112);
    for (int idx = len; idx != 0; idx--) o.push_back((uint8_t)((v >> ((idx - 1) * 8)) & 0xff));
}
bool get_vint(const uint8_t*& p, const uint8_t* end, int64_t& out) {
    if (p >= end) return false;
    const int8_t first = (int8_t)*p++;
    if (first >= -112) { out = first; return true; }
    const bool neg = first < -120;
    const int len = neg ? -(first + 120) : -(first + 112);
    if ((size_t)len > (size_t)(end - p)) return false;
    int64_t v = 0;
    for (int k = 0; k < len; k++) v = (v << 8) | *p++;
    out = neg ? (v ^ -1LL) : v;
    return true;
}
void put_text(std::vector<uint8_t>& o, const char* s) {
    const size_t n = strlen(s);
    put_vint(o, (int64_t)n);
    o.insert(o.end(), s, s + n);
}
bool get_text(const uint8_t*& p, const uint8_t* end, std::string& out) {
    int64_t n;
    if (!get_vint(p, end, n) || n < 0 || (uint64_t)n > (uint64_t)(end - p)) return false;
    out.assign((const char*)p, (size_t)n);
    p += n;
    return true;
}

struct Writer {
    std::vector<uint8_t> buf;
    uint8_t sync[16];
    size_t last_sync = 0;
    Writer(const char* key_class, const char* val_class, uint64_t seed) {
        // deterministic sync marker (Hadoop uses an MD5 of a UID and the time; any 16 bytes are valid)
        uint64_t x = seed * 0x9E3779B97F4A7C15ull + 0xD1B54A32D192ED03ull;
        for (int k = 0; k < 16; k++) { x ^= x >> 12; x ^= x << 25; x ^= x >> 27; sync[k] = (uint8_t)((x * 0x2545F4914F6CDD1Dull) >> 56); }
        buf.insert(buf.end(), {'S', 'E', 'Q', 6});
        put_text(buf, key_class);
        put_text(buf, val_class);
        buf.push_back(0);            // compressed
        buf.push_back(0);            // block compressed
        put_be32(buf, 0);            // metadata entries
        buf.insert(buf.end(), sync, sync + 16);
        last_sync = buf.size();
    }
    size_t pos() const { return buf.size(); }
    void append(const uint8_t* key, int klen, const uint8_t* val, int vlen) {
        if (buf.size() >= last_sync + SYNC_INTERVAL) {          // SequenceFile.Writer.checkAndWriteSync
            put_be32(buf, 0xffffffffu);
            buf.insert(buf.end(), sync, sync + 16);
            last_sync = buf.size();
        }
        put_be32(buf, (uint32_t)(klen + vlen));
        put_be32(buf, (uint32_t)klen);
        buf.insert(buf.end(), key, key + klen);
        buf.insert(buf.end(), val, val + vlen);
    }
    int flush(const std::string& path) {
        FILE* f = fopen(path.c_str(), "wb");
        if (!f) return fail(FY_E_ARG, "cannot create %s", path.c_str());
        const size_t w = fwrite(buf.data(), 1, buf.size(), f);
        fclose(f);
        return w == buf.size() ? FY_OK : fail(FY_E_ARG, "short write to %s", path.c_str());
    }
};

int read_file(const std::string& path, std::vector<uint8_t>& out) {
    FILE* f = fopen(path.c_str(), "rb");
    if (!f) return fail(FY_E_ARG, "cannot open %s", path.c_str());
    fseek(f, 0, SEEK_END);
    const long n = ftell(f);
    fseek(f, 0, SEEK_SET);
    out.resize((size_t)std::max(n, 0L));
    const size_t r = n > 0 ? fread(out.data(), 1, (size_t)n, f) : 0;
    fclose(f);
    return r == out.size() ? FY_OK : fail(FY_E_ARG, "short read from %s", path.c_str());
}

// a path is either one SequenceFile or a directory of them; files starting with '_' or '.' are skipped
// and a MapFile directory contributes its "data" file (M/util/HadoopUtils.java getSequenceReaders)
int list_files(const std::string& path, std::vector<std::string>& files) {
    struct stat st;
    if (stat(path.c_str(), &st) != 0) return fail(FY_E_ARG, "no such path %s", path.c_str());
    if (!S_ISDIR(st.st_mode)) { files.push_back(path); return FY_OK; }
    DIR* d = opendir(path.c_str());
    if (!d) return fail(FY_E_ARG, "cannot list %s", path.c_str());
    std::vector<std::string> names;
    while (dirent* e = readdir(d)) {
        const std::string n = e->d_name;
        if (n.empty() || n[0] == '_' || n[0] == '.') continue;
        names.push_back(n);
    }
    closedir(d);
    std::sort(names.begin(), names.end());
    for (const std::string& n : names) {
        const std::string p = path + "/" + n;
        if (stat(p.c_str(), &st) != 0) continue;
        if (S_ISDIR(st.st_mode)) {
            const std::string data = p + "/data";
            if (stat(data.c_str(), &st) == 0 && S_ISREG(st.st_mode)) files.push_back(data);
        } else if (n != "index") {
            files.push_back(p);
        }
    }
    return FY_OK;
}

// calls rec(key, klen, val, vlen) for every record of one file
template <class F>
int scan_file(const std::string& path, const char* key_class, const char* val_class, F&& rec) {
    std::vector<uint8_t> buf;
    int rc = read_file(path, buf);
    if (rc != FY_OK) return rc;
    const uint8_t* p = buf.data();
    const uint8_t* end = p + buf.size();
    if (buf.size() < 4 || p[0] != 'S' || p[1] != 'E' || p[2] != 'Q') return fail(FY_E_ARG, "%s is not a SequenceFile", path.c_str());
    if (p[3] != 6) return fail(FY_E_UNSUPPORTED, "%s: only SequenceFile version 6 is supported", path.c_str());
    p += 4;
    std::string kc, vc;
    if (!get_text(p, end, kc) || !get_text(p, end, vc)) return fail(FY_E_ARG, "%s: truncated header", path.c_str());
    if (kc != key_class || vc != val_class) return fail(FY_E_ARG, "%s: unexpected key/value classes (%s)", path.c_str(), (kc + ", " + vc).c_str());
    if ((size_t)(end - p) < 2) return fail(FY_E_ARG, "%s: truncated header", path.c_str());
    if (p[0] != 0 || p[1] != 0) return fail(FY_E_UNSUPPORTED, "%s: compressed SequenceFiles are not supported", path.c_str());
    p += 2;
    if ((size_t)(end - p) < 4) return fail(FY_E_ARG, "%s: truncated header", path.c_str());
    const uint32_t meta = get_be32(p); p += 4;
    for (uint32_t k = 0; k < meta; k++) { std::string a, b; if (!get_text(p, end, a) || !get_text(p, end, b)) return fail(FY_E_ARG, "%s: bad metadata", path.c_str()); }
    if ((size_t)(end - p) < 16) return fail(FY_E_ARG, "%s: truncated header", path.c_str());
    uint8_t sync[16];
    memcpy(sync, p, 16); p += 16;
    while (p < end) {
        if ((size_t)(end - p) < 4) return fail(FY_E_ARG, "%s: truncated record", path.c_str());
        const uint32_t len = get_be32(p); p += 4;
        if (len == 0xffffffffu) {                                 // sync escape
            if ((size_t)(end - p) < 16 || memcmp(p, sync, 16) != 0) return fail(FY_E_ARG, "%s: corrupt sync marker", path.c_str());
            p += 16;
            continue;
        }
        if ((size_t)(end - p) < 4) return fail(FY_E_ARG, "%s: truncated record", path.c_str());
        const uint32_t klen = get_be32(p); p += 4;
        if (klen > len || (uint64_t)len > (uint64_t)(end - p)) return fail(FY_E_ARG, "%s: corrupt record length", path.c_str());
        rc = rec(p, (int)klen, p + klen, (int)(len - klen));
        if (rc != FY_OK) return rc;
        p += len;
    }
    return FY_OK;
}

template <class T>
T* dup_vec(const std::vector<T>& v) {
    T* p = (T*)malloc(std::max<size_t>(v.size(), 1) * sizeof(T));
    if (p && !v.empty()) memcpy(p, v.data(), v.size() * sizeof(T));
    return p;
}

int mkdirs(const std::string& path) {
    std::string cur;
    for (size_t k = 0; k <= path.size(); k++) {
        if (k == path.size() || path[k] == '/') {
            if (!cur.empty() && mkdir(cur.c_str(), 0777) != 0) {
                struct stat st;
                if (stat(cur.c_str(), &st) != 0 || !S_ISDIR(st.st_mode)) return fail(FY_E_ARG, "cannot create directory %s", cur.c_str());
            }
        }
        if (k < path.size()) cur.push_back(path[k]);
    }
    return FY_OK;
}

}  // namespace

extern "C" const char* fy_seq_last_error(void) { return g_err; }
extern "C" void fy_free(void* p) { free(p); }

extern "C" int fy_seq_write_intpair_float(const char* path, const int32_t* first, const int32_t* second, const float* value, int64_t n) {
    if (!path || n < 0 || (n > 0 && (!first || !second || !value))) return fail(FY_E_ARG, "bad argument");
    Writer w(K_INTPAIR, K_FLOAT, (uint64_t)n + 17);
    for (int64_t k = 0; k < n; k++) {
        uint8_t key[8], val[4];
        uint32_t bits;
        memcpy(&bits, &value[k], 4);
        for (int s = 0; s < 4; s++) { key[s] = (uint8_t)((uint32_t)first[k] >> (24 - 8 * s)); key[4 + s] = (uint8_t)((uint32_t)second[k] >> (24 - 8 * s)); val[s] = (uint8_t)(bits >> (24 - 8 * s)); }
        w.append(key, 8, val, 4);
    }
    return w.flush(path);
}

extern "C" int fy_seq_write_int_int(const char* path, const int32_t* key, const int32_t* value, int64_t n) {
    if (!path || n < 0 || (n > 0 && (!key || !value))) return fail(FY_E_ARG, "bad argument");
    Writer w(K_INT, K_INT, (uint64_t)n + 29);
    for (int64_t k = 0; k < n; k++) {
        uint8_t kb[4], vb[4];
        for (int s = 0; s < 4; s++) { kb[s] = (uint8_t)((uint32_t)key[k] >> (24 - 8 * s)); vb[s] = (uint8_t)((uint32_t)value[k] >> (24 - 8 * s)); }
        w.append(kb, 4, vb, 4);
    }
    return w.flush(path);
}

extern "C" int fy_seq_write_int_double(const char* path, const int32_t* key, const double* value, int64_t n) {
    if (!path || n < 0 || (n > 0 && (!key || !value))) return fail(FY_E_ARG, "bad argument");
    Writer w(K_INT, K_DOUBLE, (uint64_t)n + 43);
    for (int64_t k = 0; k < n; k++) {
        uint8_t kb[4], vb[8];
        uint64_t bits;
        memcpy(&bits, &value[k], 8);
        for (int s = 0; s < 4; s++) kb[s] = (uint8_t)((uint32_t)key[k] >> (24 - 8 * s));
        for (int s = 0; s < 8; s++) vb[s] = (uint8_t)(bits >> (56 - 8 * s));
        w.append(kb, 4, vb, 8);
    }
    return w.flush(path);
}

// MapFile<IntWritable, DoubleWritable>: dir/data (sorted by key) + dir/index (every 128th key -> position)
extern "C" int fy_mapfile_write_int_double(const char* dir, const int32_t* key, const double* value, int64_t n) {
    if (!dir || n < 0 || (n > 0 && (!key || !value))) return fail(FY_E_ARG, "bad argument");
    for (int64_t k = 1; k < n; k++) if (key[k] <= key[k - 1]) return fail(FY_E_ARG, "MapFile keys must be strictly ascending");
    int rc = mkdirs(dir);
    if (rc != FY_OK) return rc;
    Writer data(K_INT, K_DOUBLE, (uint64_t)n + 59), index(K_INT, K_LONG, (uint64_t)n + 61);
    for (int64_t k = 0; k < n; k++) {
        uint8_t kb[4], vb[8];
        uint64_t bits;
        memcpy(&bits, &value[k], 8);
        for (int s = 0; s < 4; s++) kb[s] = (uint8_t)((uint32_t)key[k] >> (24 - 8 * s));
        for (int s = 0; s < 8; s++) vb[s] = (uint8_t)(bits >> (56 - 8 * s));
        // MapFile.Writer.append: index entry with the data position BEFORE the record (and its sync)
        if (k % 128 == 0) {
            // the sync escape, if any, is written by append(); the index must point at it
            std::vector<uint8_t> pb;
            put_be64(pb, (uint64_t)data.pos());
            index.append(kb, 4, pb.data(), 8);
        }
        data.append(kb, 4, vb, 8);
    }
    rc = data.flush(std::string(dir) + "/data");
    if (rc != FY_OK) return rc;
    return index.flush(std::string(dir) + "/index");
}

extern "C" int fy_seq_read_intpair_float(const char* path, int32_t** first, int32_t** second, float** value, int64_t* n) {
    if (!path || !first || !second || !value || !n) return fail(FY_E_ARG, "bad argument");
    std::vector<std::string> files;
    int rc = list_files(path, files);
    if (rc != FY_OK) return rc;
    std::vector<int32_t> a, b;
    std::vector<float> v;
    for (const std::string& f : files) {
        rc = scan_file(f, K_INTPAIR, K_FLOAT, [&](const uint8_t* k, int kl, const uint8_t* val, int vl) {
            if (kl != 8 || vl != 4) return fail(FY_E_ARG, "%s: record of unexpected size", f.c_str());
            a.push_back((int32_t)get_be32(k)); b.push_back((int32_t)get_be32(k + 4));
            const uint32_t bits = get_be32(val);
            float x; memcpy(&x, &bits, 4);
            v.push_back(x);
            return (int)FY_OK;
        });
        if (rc != FY_OK) return rc;
    }
    *first = dup_vec(a); *second = dup_vec(b); *value = dup_vec(v); *n = (int64_t)a.size();
    return (*first && *second && *value) ? FY_OK : fail(FY_E_NOMEM, "out of memory");
}

extern "C" int fy_seq_read_int_int(const char* path, int32_t** key, int32_t** value, int64_t* n) {
    if (!path || !key || !value || !n) return fail(FY_E_ARG, "bad argument");
    std::vector<std::string> files;
    int rc = list_files(path, files);
    if (rc != FY_OK) return rc;
    std::vector<int32_t> a, b;
    for (const std::string& f : files) {
        rc = scan_file(f, K_INT, K_INT, [&](const uint8_t* k, int kl, const uint8_t* val, int vl) {
            if (kl != 4 || vl != 4) return fail(FY_E_ARG, "%s: record of unexpected size", f.c_str());
            a.push_back((int32_t)get_be32(k)); b.push_back((int32_t)get_be32(val));
            return (int)FY_OK;
        });
        if (rc != FY_OK) return rc;
    }
    *key = dup_vec(a); *value = dup_vec(b); *n = (int64_t)a.size();
    return (*key && *value) ? FY_OK : fail(FY_E_NOMEM, "out of memory");
}

extern "C" int fy_seq_read_int_double(const char* path, int32_t** key, double** value, int64_t* n) {
    if (!path || !key || !value || !n) return fail(FY_E_ARG, "bad argument");
    std::vector<std::string> files;
    int rc = list_files(path, files);
    if (rc != FY_OK) return rc;
    std::vector<int32_t> a;
    std::vector<double> b;
    for (const std::string& f : files) {
        rc = scan_file(f, K_INT, K_DOUBLE, [&](const uint8_t* k, int kl, const uint8_t* val, int vl) {
            if (kl != 4 || vl != 8) return fail(FY_E_ARG, "%s: record of unexpected size", f.c_str());
            a.push_back((int32_t)get_be32(k));
            const uint64_t bits = get_be64(val);
            double x; memcpy(&x, &bits, 8);
            b.push_back(x);
            return (int)FY_OK;
        });
        if (rc != FY_OK) return rc;
    }
    *key = dup_vec(a); *value = dup_vec(b); *n = (int64_t)a.size();
    return (*key && *value) ? FY_OK : fail(FY_E_NOMEM, "out of memory");
}

// ---------------------------------------------------------------------------------------------
// RM2Job.run at the file level (M/rm/RM2Job.java:76-100): read the ratings directory and the two
// clustering files, run the engine, write <output>/part-r-00000, <rm2>/userSum/part-r-00000 and the
// MapFile <rm2>/itemColl/part-r-00000 exactly where the reference's three jobs leave them.
// ---------------------------------------------------------------------------------------------
extern "C" int fy_rm2_run_files(fy_rm2_ctx* ctx, const char* input_dir, const char* clustering_dir,
                                const char* clustering_count_dir, int32_t number_of_clusters,
                                const char* output_dir, const char* rm2_dir) {
    if (!ctx || !input_dir || !clustering_dir || !clustering_count_dir || !output_dir || number_of_clusters <= 0)
        return fail(FY_E_ARG, "bad argument");
    int32_t *ru = nullptr, *ri = nullptr, *cu = nullptr, *cc = nullptr, *sk = nullptr, *sv = nullptr;
    float* rs = nullptr;
    int64_t nnz = 0, nu = 0, nk = 0;
    int rc = fy_seq_read_intpair_float(input_dir, &ru, &ri, &rs, &nnz);
    if (rc == FY_OK) rc = fy_seq_read_int_int(clustering_dir, &cu, &cc, &nu);
    if (rc == FY_OK) rc = fy_seq_read_int_int(clustering_count_dir, &sk, &sv, &nk);
    std::vector<int32_t> csize((size_t)number_of_clusters, 0);       // clusterSizes = new int[numberOfClusters]
    if (rc == FY_OK)
        for (int64_t k = 0; k < nk; k++) {
            if (sk[k] < 0 || sk[k] >= number_of_clusters) { rc = fail(FY_E_ARG, "clusteringCount key outside [0, numberOfClusters)"); break; }
            csize[sk[k]] = sv[k];                                    // AbstractRM2Reducer.java:101-105
        }
    if (rc == FY_OK) { rc = fy_rm2_set_ratings(ctx, ru, ri, rs, nnz); if (rc != FY_OK) fail(rc, "%s", fy_rm2_last_error(ctx)); }
    if (rc == FY_OK) { rc = fy_rm2_set_clustering(ctx, cu, cc, nu, csize.data(), number_of_clusters); if (rc != FY_OK) fail(rc, "%s", fy_rm2_last_error(ctx)); }
    if (rc == FY_OK) { rc = fy_rm2_run(ctx); if (rc != FY_OK) fail(rc, "RM2-3 failed! %s", fy_rm2_last_error(ctx)); }
    if (rc == FY_OK) {
        const int64_t n = fy_rm2_result_count(ctx);
        std::vector<int32_t> u((size_t)n), i((size_t)n);
        std::vector<float> s((size_t)n);
        rc = fy_rm2_results(ctx, u.data(), i.data(), nullptr, s.data(), nullptr);
        if (rc == FY_OK) rc = mkdirs(output_dir);
        if (rc == FY_OK) rc = fy_seq_write_intpair_float((std::string(output_dir) + "/part-r-00000").c_str(), u.data(), i.data(), s.data(), n);
    }
    if (rc == FY_OK && rm2_dir) {
        std::vector<double> us((size_t)nu), ip((size_t)fy_rm2_max_item(ctx) + 1);
        double total = 0;
        rc = fy_rm2_stats(ctx, us.data(), ip.data(), &total);
        if (rc == FY_OK) rc = mkdirs(std::string(rm2_dir) + "/userSum");
        if (rc == FY_OK) {
            std::vector<int32_t> order((size_t)nu);
            for (int64_t k = 0; k < nu; k++) order[k] = (int32_t)k;
            std::sort(order.begin(), order.end(), [&](int32_t a, int32_t b) { return cu[a] < cu[b]; });
            std::vector<int32_t> ku((size_t)nu);
            std::vector<double> vu((size_t)nu);
            for (int64_t k = 0; k < nu; k++) { ku[k] = cu[order[k]]; vu[k] = us[order[k]]; }
            rc = fy_seq_write_int_double((std::string(rm2_dir) + "/userSum/part-r-00000").c_str(), ku.data(), vu.data(), nu);
        }
        if (rc == FY_OK) {
            std::vector<int32_t> ki;
            std::vector<double> vi;
            for (size_t it = 0; it < ip.size(); it++) if (ip[it] > 0) { ki.push_back((int32_t)it); vi.push_back(ip[it]); }
            rc = fy_mapfile_write_int_double((std::string(rm2_dir) + "/itemColl/part-r-00000").c_str(), ki.data(), vi.data(), (int64_t)ki.size());
        }
    }
    free(ru); free(ri); free(rs); free(cu); free(cc); free(sk); free(sv);
    return rc;
}

// ---------------------------------------------------------------------------------------------
// SequenceFile<IntWritable, VectorWritable>: the H and W factor matrices of the NMF / PPC step
// (M/util/DataInitialization.java:76-88,120-131 write them, M/nmf/common/AbstractJoinVectorMapper.java:66-76
// and every *ComputationReducer read them).  Mahout 0.8 VectorWritable (un-vendored dependency,
// org.apache.mahout:mahout-core:0.8), restated from its published format -- PARITY UNPINNED at the byte level:
//   byte flags  (0x01 dense | 0x02 sequential access | 0x04 named | 0x08 lax precision = floats)
//   unsigned varint size           (7 bits per byte, least significant group first, 0x80 = more)
//   dense : size doubles (big-endian) ;  sparse : varint nnz, then (varint index or index delta, double) pairs
//   named : DataOutput.writeUTF name after the elements
// The writer emits DenseVector records (flags 0x03), which is what DataInitialization and the reducers
// (Vector.times / plus on dense vectors) produce.
// ---------------------------------------------------------------------------------------------
namespace {
void put_uvarint(std::vector<uint8_t>& o, uint32_t v) {
    while (v >= 0x80) { o.push_back((uint8_t)((v & 0x7f) | 0x80)); v >>= 7; }
    o.push_back((uint8_t)v);
}
bool get_uvarint(const uint8_t*& p, const uint8_t* end, uint32_t& out) {
    uint32_t v = 0;
    for (int shift = 0; shift < 35; shift += 7) {
        if (p >= end) return false;
        const uint8_t b = *p++;
        v |= (uint32_t)(b & 0x7f) << shift;
        if (!(b & 0x80)) { out = v; return true; }
    }
    return false;
}
}  // namespace

extern "C" int fy_seq_write_int_vector(const char* path, const int32_t* key, const double* rows, int64_t n, int32_t cols) {
    if (!path || n < 0 || cols <= 0 || (n > 0 && (!key || !rows))) return fail(FY_E_ARG, "bad argument");
    Writer w(K_INT, K_VECTOR, (uint64_t)n * 31 + (uint64_t)cols);
    std::vector<uint8_t> val;
    for (int64_t k = 0; k < n; k++) {
        uint8_t kb[4];
        for (int s = 0; s < 4; s++) kb[s] = (uint8_t)((uint32_t)key[k] >> (24 - 8 * s));
        val.clear();
        val.push_back(0x03);                                  // FLAG_DENSE | FLAG_SEQUENTIAL
        put_uvarint(val, (uint32_t)cols);
        for (int32_t c = 0; c < cols; c++) {
            uint64_t bits;
            memcpy(&bits, &rows[(size_t)k * cols + c], 8);
            put_be64(val, bits);
        }
        w.append(kb, 4, val.data(), (int)val.size());
    }
    return w.flush(path);
}

// rows come back row-major [n x cols] in file order; every vector must have the same size
extern "C" int fy_seq_read_int_vector(const char* path, int32_t** key, double** rows, int64_t* n, int32_t* cols) {
    if (!path || !key || !rows || !n || !cols) return fail(FY_E_ARG, "bad argument");
    std::vector<std::string> files;
    int rc = list_files(path, files);
    if (rc != FY_OK) return rc;
    std::vector<int32_t> ks;
    std::vector<double> vs;
    int32_t width = -1;
    for (const std::string& f : files) {
        rc = scan_file(f, K_INT, K_VECTOR, [&](const uint8_t* k, int klen, const uint8_t* v, int vlen) -> int {
            if (klen != 4 || vlen < 2) return fail(FY_E_ARG, "%s: bad <IntWritable, VectorWritable> record", f.c_str());
            const uint8_t* p = v; const uint8_t* end = v + vlen;
            const uint8_t flags = *p++;
            if (flags >> 4) return fail(FY_E_ARG, "%s: unknown VectorWritable flags", f.c_str());
            const bool dense = flags & 1, sequential = flags & 2, lax = flags & 8;
            uint32_t size = 0;
            if (!get_uvarint(p, end, size)) return fail(FY_E_ARG, "%s: truncated VectorWritable", f.c_str());
            if (width < 0) width = (int32_t)size;
            if ((int32_t)size != width) return fail(FY_E_ARG, "%s: vectors of different sizes", f.c_str());
            const size_t base = vs.size();
            vs.resize(base + size, 0.0);
            auto element = [&](double& out) -> bool {
                if (lax) { if ((size_t)(end - p) < 4) return false; const uint32_t b = get_be32(p); p += 4; float x; memcpy(&x, &b, 4); out = x; }
                else { if ((size_t)(end - p) < 8) return false; const uint64_t b = get_be64(p); p += 8; memcpy(&out, &b, 8); }
                return true;
            };
            if (dense) {
                for (uint32_t c = 0; c < size; c++) if (!element(vs[base + c])) return fail(FY_E_ARG, "%s: truncated VectorWritable", f.c_str());
            } else {
                uint32_t nnz = 0, last = 0;
                if (!get_uvarint(p, end, nnz)) return fail(FY_E_ARG, "%s: truncated VectorWritable", f.c_str());
                for (uint32_t e = 0; e < nnz; e++) {
                    uint32_t idx = 0; double x = 0;
                    if (!get_uvarint(p, end, idx)) return fail(FY_E_ARG, "%s: truncated VectorWritable", f.c_str());
                    if (sequential) { idx += last; last = idx; }
                    if (idx >= size || !element(x)) return fail(FY_E_ARG, "%s: corrupt sparse VectorWritable", f.c_str());
                    vs[base + idx] = x;
                }
            }
            ks.push_back((int32_t)get_be32(k));
            return FY_OK;
        });
        if (rc != FY_OK) return rc;
    }
    *n = (int64_t)ks.size();
    *cols = width < 0 ? 0 : width;
    *key = dup_vec(ks); *rows = dup_vec(vs);
    if (!*key || !*rows) { free(*key); free(*rows); *key = nullptr; *rows = nullptr; return fail(FY_E_NOMEM, "out of memory"); }
    return FY_OK;
}

// ---------------------------------------------------------------------------------------------
// AbstractNMFDriver.run + ClusterAssignmentJob + CountClustersJob at the file level
// (M/nmf/AbstractNMFDriver.java:92-141, M/nmf/clustering/ClusterAssignmentJob.java:47-95,
//  M/nmf/clustering/CountClustersJob.java:41-80): read the ratings directory and the H / W files named by the
// "H" / "W" options (or start from createInitialMatrices when h_in is NULL), iterate on the GPU, and leave
// H, W, `clustering` and `clusteringCount` where the reference's jobs leave them.
// ---------------------------------------------------------------------------------------------
extern "C" int fy_nmf_run_files(fy_nmf_ctx* ctx, const fy_nmf_params* prm, const char* input_dir, const char* h_in, const char* w_in,
                                uint64_t seed, const char* h_out, const char* w_out, const char* clustering_out,
                                const char* clustering_count_out) {
    if (!ctx || !prm || !input_dir || (!h_in) != (!w_in)) return fail(FY_E_ARG, "bad argument");
    const int32_t U = prm->number_of_users, M = prm->number_of_items, k = prm->number_of_clusters, base = prm->id_base;
    int32_t *ru = nullptr, *ri = nullptr, *hk = nullptr, *wk = nullptr;
    float* rs = nullptr;
    double *hv = nullptr, *wv = nullptr;
    int64_t nnz = 0, hn = 0, wn = 0;
    int32_t hc = 0, wc = 0;
    auto engine_rc = [&](int rc, const char* what) { if (rc != FY_OK) fail(rc, "%s: %s", what, fy_nmf_last_error(ctx)); return rc; };
    int rc = fy_seq_read_intpair_float(input_dir, &ru, &ri, &rs, &nnz);
    if (rc == FY_OK) rc = engine_rc(fy_nmf_set_ratings(ctx, ru, ri, rs, nnz), "ratings");
    if (rc == FY_OK && h_in) {
        rc = fy_seq_read_int_vector(h_in, &hk, &hv, &hn, &hc);
        if (rc == FY_OK) rc = fy_seq_read_int_vector(w_in, &wk, &wv, &wn, &wc);
        if (rc == FY_OK && (hn != U || wn != M || hc != k || wc != k)) rc = fail(FY_E_ARG, "H / W do not have numberOfUsers / numberOfItems rows of numberOfClusters columns");
        if (rc == FY_OK) {
            std::vector<double> H((size_t)U * k), W((size_t)M * k);
            std::vector<char> seen_h((size_t)U, 0), seen_w((size_t)M, 0);
            for (int64_t r = 0; r < hn && rc == FY_OK; r++) {
                const int64_t row = (int64_t)hk[r] - base;
                if (row < 0 || row >= U || seen_h[row]) { rc = fail(FY_E_ARG, "H: row key outside the id range or repeated"); break; }
                seen_h[row] = 1;
                memcpy(&H[(size_t)row * k], &hv[(size_t)r * k], sizeof(double) * (size_t)k);
            }
            for (int64_t r = 0; r < wn && rc == FY_OK; r++) {
                const int64_t row = (int64_t)wk[r] - base;
                if (row < 0 || row >= M || seen_w[row]) { rc = fail(FY_E_ARG, "W: row key outside the id range or repeated"); break; }
                seen_w[row] = 1;
                memcpy(&W[(size_t)row * k], &wv[(size_t)r * k], sizeof(double) * (size_t)k);
            }
            if (rc == FY_OK) rc = engine_rc(fy_nmf_set_factors(ctx, H.data(), W.data()), "factors");
        }
    } else if (rc == FY_OK) {
        rc = engine_rc(fy_nmf_init_random(ctx, seed), "createInitialMatrices");
    }
    if (rc == FY_OK) rc = engine_rc(fy_nmf_run(ctx), "PPCJob failed!");
    if (rc == FY_OK && (h_out || w_out)) {
        std::vector<double> H((size_t)U * k), W((size_t)M * k);
        std::vector<int32_t> uk((size_t)U), ik((size_t)M);
        for (int32_t r = 0; r < U; r++) uk[r] = base + r;
        for (int32_t r = 0; r < M; r++) ik[r] = base + r;
        rc = engine_rc(fy_nmf_get_factors(ctx, H.data(), W.data()), "factors");
        if (rc == FY_OK && h_out) { rc = mkdirs(h_out); if (rc == FY_OK) rc = fy_seq_write_int_vector((std::string(h_out) + "/part-r-00000").c_str(), uk.data(), H.data(), U, k); }
        if (rc == FY_OK && w_out) { rc = mkdirs(w_out); if (rc == FY_OK) rc = fy_seq_write_int_vector((std::string(w_out) + "/part-m-00000").c_str(), ik.data(), W.data(), M, k); }
    }
    if (rc == FY_OK && clustering_out) {
        std::vector<int32_t> cl((size_t)U), cnt((size_t)k), uk((size_t)U);
        for (int32_t r = 0; r < U; r++) uk[r] = base + r;
        rc = engine_rc(fy_nmf_cluster_assignment(ctx, cl.data(), cnt.data()), "ClusterAssignmentJob failed!");
        if (rc == FY_OK) rc = mkdirs(clustering_out);
        if (rc == FY_OK) rc = fy_seq_write_int_int((std::string(clustering_out) + "/part-m-00000").c_str(), uk.data(), cl.data(), U);
        if (rc == FY_OK && clustering_count_out) {
            std::vector<int32_t> ck, cv;                      // CountReducer emits only the clusters that occur
            for (int32_t c = 0; c < k; c++) if (cnt[c] > 0) { ck.push_back(c); cv.push_back(cnt[c]); }
            rc = mkdirs(clustering_count_out);
            if (rc == FY_OK) rc = fy_seq_write_int_int((std::string(clustering_count_out) + "/part-r-00000").c_str(), ck.data(), cv.data(), (int64_t)ck.size());
        }
    }
    free(ru); free(ri); free(rs); free(hk); free(hv); free(wk); free(wv);
    return rc;
}
