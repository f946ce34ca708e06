/*
 * filmyou_seqfile.h -- C ABI of the Hadoop SequenceFile / MapFile readers and writers for the record
 * types on either side of the RM2 path.  Host only; part of
 * libfilmyou_rm2.so.  M/ = /root/reference/src/main/java/es/udc/fi/dc/irlab/.
 *
 *   fy_seq_*_intpair_float  ratings in / recommendations out, SequenceFile<IntPairWritable, FloatWritable>
 *                           (M/util/DataInitialization.java:155-174, M/rm/RM2HDFSReducer.java:44-50)
 *   fy_seq_*_int_int        clustering, clusteringCount, SequenceFile<IntWritable, IntWritable>
 *                           (M/util/DataInitialization.java:200-222, M/common/AbstractByClusterMapper.java:57-66)
 *   fy_seq_*_int_double     rm2/userSum, SequenceFile<IntWritable, DoubleWritable>  (M/rm/RM2Job.java:138-142)
 *   fy_mapfile_write_int_double   rm2/itemColl, MapFile<IntWritable, DoubleWritable> (M/rm/RM2Job.java:190-196,
 *                           M/util/MapFileOutputFormat.java:171-178); read it back with fy_seq_read_int_double
 *   fy_rm2_run_files        RM2Job.run at the file level (M/rm/RM2Job.java:76-100)
 *   fy_seq_*_int_vector     the H / W factor matrices, SequenceFile<IntWritable, VectorWritable> (Mahout 0.8)
 *   fy_nmf_run_files        PPCDriver / NMFDriver + ClusterAssignmentJob + CountClustersJob at the file level
 *
 * A `path` may be one file or a directory (files starting with '_' or '.' are skipped, a MapFile
 * sub-directory contributes its `data` file), like M/util/HadoopUtils.java getSequenceReaders.
 * Readers allocate the output arrays; release them with fy_free.  Formats: Hadoop 1.2.1 SequenceFile
 * version 6, uncompressed; Mahout 0.8 IntPairWritable = two big-endian ints.  PARITY UNPINNED at the
 * byte level (neither library is vendored; the reference holds no serialized fixture).
 */
#ifndef FILMYOU_SEQFILE_H
#define FILMYOU_SEQFILE_H

#include <stdint.h>
#include "filmyou_rm2.h"

#ifdef __cplusplus
extern "C" {
#endif

const char* fy_seq_last_error(void);
void fy_free(void* p);

int fy_seq_write_intpair_float(const char* path, const int32_t* first, const int32_t* second, const float* value, int64_t n);
int fy_seq_write_int_int(const char* path, const int32_t* key, const int32_t* value, int64_t n);
int fy_seq_write_int_double(const char* path, const int32_t* key, const double* value, int64_t n);
int fy_mapfile_write_int_double(const char* dir, const int32_t* key, const double* value, int64_t n);

int fy_seq_read_intpair_float(const char* path, int32_t** first, int32_t** second, float** value, int64_t* n);
int fy_seq_read_int_int(const char* path, int32_t** key, int32_t** value, int64_t* n);
int fy_seq_read_int_double(const char* path, int32_t** key, double** value, int64_t* n);

/* H / W factor matrices, SequenceFile<IntWritable, VectorWritable> (M/util/DataInitialization.java:76-88,120-131).
 * rows = [n x cols] row-major doubles in file order.  The writer emits Mahout 0.8 DenseVector records; the
 * reader also accepts sparse / lax-precision vectors. */
int fy_seq_write_int_vector(const char* path, const int32_t* key, const double* rows, int64_t n, int32_t cols);
int fy_seq_read_int_vector(const char* path, int32_t** key, double** rows, int64_t* n, int32_t* cols);

/* input_dir = mapred.input.dir, clustering_dir / clustering_count_dir = <directory>/<clustering>,
 * <directory>/<clusteringCount>, output_dir = mapred.output.dir, rm2_dir = <directory>/rm2 (may be NULL). */
int fy_rm2_run_files(fy_rm2_ctx* ctx, const char* input_dir, const char* clustering_dir,
                     const char* clustering_count_dir, int32_t number_of_clusters,
                     const char* output_dir, const char* rm2_dir);

/* AbstractNMFDriver.run + ClusterAssignmentJob + CountClustersJob at the file level (filmyou_nmf.h context; `prm`
 * = the parameters the context was created with).  h_in / w_in = the "H" / "W" options (both NULL: random
 * start from `seed`); any output path may be NULL.  Outputs: <h_out>/part-r-00000, <w_out>/part-m-00000,
 * <clustering_out>/part-m-00000, <clustering_count_out>/part-r-00000. */
struct fy_nmf_ctx;
struct fy_nmf_params;
int fy_nmf_run_files(struct fy_nmf_ctx* ctx, const struct fy_nmf_params* prm, const char* input_dir, const char* h_in,
                     const char* w_in, uint64_t seed, const char* h_out, const char* w_out, const char* clustering_out,
                     const char* clustering_count_out);

#ifdef __cplusplus
}
#endif
#endif
