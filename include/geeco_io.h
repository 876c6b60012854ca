/* geeco_io.h -- C-ABI of libgeeco_io.so: the host-side data formats either side of the train step.
 *
 * SURVEY.md section 8(f) rank 3 (native input pipeline) and rank 2 (checkpoint protocol).  Host code only
 * (g++ + zlib, no CUDA): it produces the pinned-host batches that geeco_b200.h's geeco_train_step consumes
 * and reads / writes the variable files that surround it.  Plain pointers and sizes; every function returns
 * 0 on success or a negative status, and geeco_io_last_error() describes the last failure of the calling
 * thread.  Handles are not thread-safe; different handles may be used from different threads (the Python
 * pipeline decodes one episode per worker thread, ctypes releases the GIL).
 *
 * What each group replaces in the reference (file:line into ogroth/geeco):
 *   TFRecord framing + ZLIB      tf.data.TFRecordDataset(compression_type='ZLIB')   src/data/geeco_gym.py:443-446
 *                                tf.python_io.TFRecordWriter                        src/data/data_recorder.py:150-156
 *   SequenceExample decode       tf.parse_single_sequence_example                   src/data/geeco_gym.py:298-301
 *                                (schema of PickAndPlaceEncodingV4.decode           src/data/geeco_gym.py:117-162)
 *   pixel bytes                  rgb recorded as uint8 cast to float (utils/tfrecord.py:75-76), `/= 255.0` at
 *                                geeco_gym.py:310 -- geeco_seqex_read_u8 recovers the bytes so the division
 *                                runs on the device (GEECO_FRAMES_U8 of geeco_b200.h)
 *   sliding windows              _window_v3                                         src/data/geeco_gym.py:614-631
 *   variable files               tf.train.Saver V2 bundle (`model.ckpt-N.index` + `.data-00000-of-00001`) that
 *                                tf.estimator writes and predictor.py:87-93 restores
 */
#ifndef GEECO_IO_H_
#define GEECO_IO_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum { GEECO_IO_OK = 0, GEECO_IO_ERR_ARG = -1, GEECO_IO_ERR_FILE = -2, GEECO_IO_ERR_FORMAT = -3,
       GEECO_IO_ERR_CRC = -4, GEECO_IO_ERR_MISSING = -5, GEECO_IO_ERR_SIZE = -6 };

enum { GEECO_IO_COMPRESSION_AUTO = -1, GEECO_IO_COMPRESSION_NONE = 0, GEECO_IO_COMPRESSION_ZLIB = 1,
       GEECO_IO_COMPRESSION_GZIP = 2 };

/* tf.train.Feature kinds (field numbers of the oneof in feature.proto) */
enum { GEECO_IO_KIND_NONE = 0, GEECO_IO_KIND_BYTES = 1, GEECO_IO_KIND_FLOAT = 2, GEECO_IO_KIND_INT64 = 3 };

enum { GEECO_IO_CONTEXT = 0, GEECO_IO_SEQUENCE = 1 };

const char* geeco_io_last_error(void);
int geeco_io_version(void);

/* ---- CRC-32C (Castagnoli) and TensorFlow's masking: ((crc >> 15) | (crc << 17)) + 0xa282ead8 ---- */
uint32_t geeco_io_crc32c(const void* data, size_t n);
uint32_t geeco_io_crc32c_extend(uint32_t crc, const void* data, size_t n);
uint32_t geeco_io_crc32c_mask(uint32_t crc);
uint32_t geeco_io_crc32c_unmask(uint32_t masked);

/* ---- TFRecord files: [u64 length][u32 masked crc of length][data][u32 masked crc of data], the whole file
 * optionally one zlib / gzip stream.  The reader inflates the file once and indexes its records. ---- */
typedef struct geeco_tfrecord geeco_tfrecord;
int geeco_tfrecord_open(const char* path, int compression, int verify_crc, geeco_tfrecord** out);
int64_t geeco_tfrecord_count(const geeco_tfrecord* f);
int geeco_tfrecord_get(const geeco_tfrecord* f, int64_t index, const uint8_t** data, uint64_t* len);
void geeco_tfrecord_close(geeco_tfrecord* f);
/* Writes `count` records (data[i], len[i]) as one file; compression as above (AUTO = by suffix). */
int geeco_tfrecord_write(const char* path, int compression, int64_t count, const uint8_t* const* data,
                         const uint64_t* len);

/* ---- tf.train.SequenceExample: an index over the serialized bytes (which must outlive the handle). ---- */
typedef struct geeco_seqex geeco_seqex;
int geeco_seqex_parse(const uint8_t* data, uint64_t len, geeco_seqex** out);
void geeco_seqex_free(geeco_seqex* s);
/* number of keys in the context (which = 0) or in feature_lists (which = 1), and the i-th key */
int64_t geeco_seqex_num_keys(const geeco_seqex* s, int which);
int geeco_seqex_key(const geeco_seqex* s, int which, int64_t i, const char** name, uint64_t* name_len);
/* kind, number of frames (1 for a context feature) and values per frame (-1 when frames differ in length:
 * tf.FixedLenSequenceFeature would reject such a list) */
int geeco_seqex_info(const geeco_seqex* s, int which, const char* name, int* kind, int64_t* frames,
                     int64_t* per_frame);
/* all frames of a float / int64 feature, concatenated in frame order; `capacity` in elements */
int geeco_seqex_read_f32(const geeco_seqex* s, int which, const char* name, float* dst, int64_t capacity);
int geeco_seqex_read_i64(const geeco_seqex* s, int which, const char* name, int64_t* dst, int64_t capacity);
/* float feature holding recorded pixel bytes: dst[i] = (uint8) value; *inexact counts the values that are not
 * integers in [0,255] (0 for data written by the reference's recorder) */
int geeco_seqex_read_u8(const geeco_seqex* s, int which, const char* name, uint8_t* dst, int64_t capacity,
                        int64_t* inexact);
/* value `j` of frame `frame` of a bytes feature (pointer into the serialized buffer) */
int geeco_seqex_bytes(const geeco_seqex* s, int which, const char* name, int64_t frame, int64_t j,
                      const uint8_t** data, uint64_t* len);

/* ---- sliding windows (_window_v3): dst[i] = src[w0+i .. w0+i+K) for i in [0, nwin); one frame =
 * frame_bytes bytes; frames w0 .. w0+nwin+K-2 must exist (checked against `frames`). ---- */
int geeco_io_window_gather(const void* src, int64_t frames, int64_t frame_bytes, int64_t K, int64_t w0,
                           int64_t nwin, void* dst);

/* ---- TF V2 checkpoint bundles (tensor_bundle): `<prefix>.index` is a leveldb-format table mapping "" to a
 * BundleHeaderProto and every variable name to a BundleEntryProto {dtype, shape, shard_id, offset, size,
 * crc32c}; `<prefix>.data-0000S-of-0000N` holds the raw little-endian tensors. ---- */
typedef struct geeco_bundle geeco_bundle;
enum { GEECO_IO_DT_FLOAT = 1, GEECO_IO_DT_INT32 = 3, GEECO_IO_DT_INT64 = 9 };   /* types.proto DataType */
int geeco_bundle_open(const char* prefix, geeco_bundle** out);
void geeco_bundle_close(geeco_bundle* b);
int64_t geeco_bundle_num_tensors(const geeco_bundle* b);
int geeco_bundle_name(const geeco_bundle* b, int64_t i, const char** name, uint64_t* name_len);
/* dtype (DataType enum), rank and dims (up to 8) and byte size of a stored tensor */
int geeco_bundle_info(const geeco_bundle* b, const char* name, int* dtype, int* ndim, int64_t* dims,
                      int64_t* nbytes);
/* copies the tensor's bytes into dst (capacity in bytes); verify_crc checks the entry's masked crc32c */
int geeco_bundle_read(const geeco_bundle* b, const char* name, void* dst, int64_t capacity, int verify_crc);
/* Writer: tensors are appended to one data shard in sorted-name order at finish. */
typedef struct geeco_bundle_writer geeco_bundle_writer;
int geeco_bundle_writer_create(const char* prefix, geeco_bundle_writer** out);
int geeco_bundle_writer_add(geeco_bundle_writer* w, const char* name, int dtype, int ndim, const int64_t* dims,
                            const void* data, int64_t nbytes);
int geeco_bundle_writer_finish(geeco_bundle_writer* w);   /* writes both files and frees the writer */
void geeco_bundle_writer_abort(geeco_bundle_writer* w);

#ifdef __cplusplus
}
#endif
#endif  /* GEECO_IO_H_ */
