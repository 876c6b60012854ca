// libgeeco_io.so -- host-side data formats either side of the GEECO train step (include/geeco_io.h).
//
// Three independent parts, all plain C++17 + zlib:
//   1. CRC-32C and TFRecord framing (reader inflates a ZLIB / GZIP file once and indexes its records);
//   2. a zero-copy index over a serialized tf.train.SequenceExample (the recorder's episode format,
//      reference src/data/data_recorder.py:37-59, schema src/data/geeco_gym.py:117-162) with bulk readers that
//      turn the float-encoded pixels back into the recorded bytes;
//   3. reader / writer of TF V2 checkpoint bundles (leveldb-format index table + raw data shard), the files
//      tf.estimator writes as model.ckpt-<step>.* and the reference's predictor restores (predictor.py:87-93).
// The wire formats are restated from their public specifications (protobuf encoding, leveldb table_format.md,
// tensor_bundle.proto, RFC 3720 appendix B.4 for CRC-32C, the snappy format description); TensorFlow's source
// is not part of the reference tree.
#include "../../include/geeco_io.h"

#include <zlib.h>

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <map>
#include <string>
#include <thread>
#include <utility>
#include <vector>

namespace {

thread_local std::string g_error;

int fail(int code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  g_error = buf;
  return code;
}

// ------------------------------------------------------------------------------------------------
// CRC-32C: slicing-by-8 tables built at load time; SSE4.2 crc32 instruction when the CPU has it
// ------------------------------------------------------------------------------------------------
struct CrcTables {
  uint32_t t[8][256];
  CrcTables() {
    for (uint32_t i = 0; i < 256; ++i) {
      uint32_t c = i;
      for (int k = 0; k < 8; ++k) c = (c & 1) ? (c >> 1) ^ 0x82F63B78u : c >> 1;
      t[0][i] = c;
    }
    for (uint32_t i = 0; i < 256; ++i)
      for (int s = 1; s < 8; ++s) t[s][i] = (t[s - 1][i] >> 8) ^ t[0][t[s - 1][i] & 0xff];
  }
};
const CrcTables g_crc;

uint32_t crc_sw(uint32_t crc, const uint8_t* p, size_t n) {
  crc = ~crc;
  while (n && (reinterpret_cast<uintptr_t>(p) & 7)) { crc = (crc >> 8) ^ g_crc.t[0][(crc ^ *p++) & 0xff]; --n; }
  while (n >= 8) {
    uint64_t w;
    memcpy(&w, p, 8);
    w ^= crc;
    crc = g_crc.t[7][w & 0xff] ^ g_crc.t[6][(w >> 8) & 0xff] ^ g_crc.t[5][(w >> 16) & 0xff] ^
          g_crc.t[4][(w >> 24) & 0xff] ^ g_crc.t[3][(w >> 32) & 0xff] ^ g_crc.t[2][(w >> 40) & 0xff] ^
          g_crc.t[1][(w >> 48) & 0xff] ^ g_crc.t[0][w >> 56];
    p += 8; n -= 8;
  }
  while (n--) crc = (crc >> 8) ^ g_crc.t[0][(crc ^ *p++) & 0xff];
  return ~crc;
}

#if defined(__x86_64__)
__attribute__((target("sse4.2"))) uint32_t crc_hw(uint32_t crc, const uint8_t* p, size_t n) {
  uint64_t c = static_cast<uint32_t>(~crc);
  while (n && (reinterpret_cast<uintptr_t>(p) & 7)) { c = __builtin_ia32_crc32qi(static_cast<uint32_t>(c), *p++); --n; }
  while (n >= 8) {
    uint64_t w;
    memcpy(&w, p, 8);
    c = __builtin_ia32_crc32di(c, w);
    p += 8; n -= 8;
  }
  while (n--) c = __builtin_ia32_crc32qi(static_cast<uint32_t>(c), *p++);
  return ~static_cast<uint32_t>(c);
}
const bool g_have_sse42 = __builtin_cpu_supports("sse4.2");
#else
const bool g_have_sse42 = false;
uint32_t crc_hw(uint32_t crc, const uint8_t* p, size_t n) { return crc_sw(crc, p, n); }
#endif

uint32_t crc_extend(uint32_t crc, const void* data, size_t n) {
  const uint8_t* p = static_cast<const uint8_t*>(data);
  return g_have_sse42 ? crc_hw(crc, p, n) : crc_sw(crc, p, n);
}
const uint32_t kMaskDelta = 0xa282ead8u;
inline uint32_t crc_mask(uint32_t c) { return ((c >> 15) | (c << 17)) + kMaskDelta; }
inline uint32_t crc_unmask(uint32_t m) { uint32_t r = m - kMaskDelta; return (r >> 17) | (r << 15); }

// ------------------------------------------------------------------------------------------------
// little helpers: files, fixed / varint coding
// ------------------------------------------------------------------------------------------------
int read_file(const std::string& path, std::vector<uint8_t>* out) {
  FILE* fp = fopen(path.c_str(), "rb");
  if (!fp) return fail(GEECO_IO_ERR_FILE, "cannot open %s", path.c_str());
  fseek(fp, 0, SEEK_END);
  long n = ftell(fp);
  fseek(fp, 0, SEEK_SET);
  if (n < 0) { fclose(fp); return fail(GEECO_IO_ERR_FILE, "cannot size %s", path.c_str()); }
  out->resize(static_cast<size_t>(n));
  size_t got = n ? fread(out->data(), 1, static_cast<size_t>(n), fp) : 0;
  fclose(fp);
  if (got != static_cast<size_t>(n)) return fail(GEECO_IO_ERR_FILE, "short read on %s", path.c_str());
  return GEECO_IO_OK;
}

int write_file(const std::string& path, const uint8_t* p, size_t n) {
  FILE* fp = fopen(path.c_str(), "wb");
  if (!fp) return fail(GEECO_IO_ERR_FILE, "cannot create %s", path.c_str());
  size_t put = n ? fwrite(p, 1, n, fp) : 0;
  int rc = fclose(fp);
  if (put != n || rc) return fail(GEECO_IO_ERR_FILE, "short write on %s", path.c_str());
  return GEECO_IO_OK;
}

inline uint32_t load32(const uint8_t* p) { uint32_t v; memcpy(&v, p, 4); return v; }
inline uint64_t load64(const uint8_t* p) { uint64_t v; memcpy(&v, p, 8); return v; }
inline void put32(std::vector<uint8_t>* b, uint32_t v) { uint8_t t[4]; memcpy(t, &v, 4); b->insert(b->end(), t, t + 4); }
inline void put64(std::vector<uint8_t>* b, uint64_t v) { uint8_t t[8]; memcpy(t, &v, 8); b->insert(b->end(), t, t + 8); }
inline void put_varint(std::vector<uint8_t>* b, uint64_t v) {
  while (v >= 0x80) { b->push_back(static_cast<uint8_t>(v) | 0x80); v >>= 7; }
  b->push_back(static_cast<uint8_t>(v));
}
inline void put_bytes(std::vector<uint8_t>* b, const void* p, size_t n) {
  const uint8_t* q = static_cast<const uint8_t*>(p);
  b->insert(b->end(), q, q + n);
}

// a cursor over protobuf / leveldb bytes; every getter returns false on truncation
struct Span {
  const uint8_t* p = nullptr;
  const uint8_t* e = nullptr;
  Span() {}
  Span(const uint8_t* b, size_t n) : p(b), e(b + n) {}
  size_t size() const { return static_cast<size_t>(e - p); }
  bool empty() const { return p >= e; }
  bool varint(uint64_t* v) {
    uint64_t r = 0;
    for (int shift = 0; shift < 64 && p < e; shift += 7) {
      uint8_t b = *p++;
      r |= static_cast<uint64_t>(b & 0x7f) << shift;
      if (!(b & 0x80)) { *v = r; return true; }
    }
    return false;
  }
  bool bytes(Span* out) {
    uint64_t n;
    if (!varint(&n) || n > size()) return false;
    *out = Span(p, static_cast<size_t>(n));
    p += n;
    return true;
  }
  bool skip(size_t n) { if (n > size()) return false; p += n; return true; }
  // reads a field header; wire types: 0 varint, 1 fixed64, 2 length-delimited, 5 fixed32
  bool tag(uint32_t* field, uint32_t* wire) {
    uint64_t t;
    if (!varint(&t)) return false;
    *field = static_cast<uint32_t>(t >> 3);
    *wire = static_cast<uint32_t>(t & 7);
    return true;
  }
  bool skip_value(uint32_t wire) {
    uint64_t v; Span s;
    switch (wire) {
      case 0: return varint(&v);
      case 1: return skip(8);
      case 2: return bytes(&s);
      case 5: return skip(4);
      default: return false;
    }
  }
};

// ------------------------------------------------------------------------------------------------
// zlib
// ------------------------------------------------------------------------------------------------
int inflate_all(const std::vector<uint8_t>& in, int window_bits, std::vector<uint8_t>* out, const char* what) {
  z_stream zs;
  memset(&zs, 0, sizeof zs);
  if (inflateInit2(&zs, window_bits) != Z_OK) return fail(GEECO_IO_ERR_FORMAT, "inflateInit2 failed");
  out->clear();
  out->resize(std::max<size_t>(in.size() * 4, 1 << 16));
  zs.next_in = const_cast<Bytef*>(in.data());
  size_t in_left = in.size(), produced = 0;
  int rc = Z_OK;
  while (true) {
    if (produced == out->size()) out->resize(out->size() * 2);
    // zlib's counters are 32-bit: feed and drain in bounded pieces
    uInt give = static_cast<uInt>(std::min<size_t>(in_left, 1u << 30));
    uInt room = static_cast<uInt>(std::min<size_t>(out->size() - produced, 1u << 30));
    zs.avail_in = give;
    zs.next_out = out->data() + produced;
    zs.avail_out = room;
    rc = inflate(&zs, Z_NO_FLUSH);
    in_left -= give - zs.avail_in;
    produced += room - zs.avail_out;
    if (rc == Z_STREAM_END) {
      if (in_left == 0) break;
      if (inflateReset(&zs) != Z_OK) { rc = Z_DATA_ERROR; break; }   // concatenated members
      continue;
    }
    if (rc == Z_BUF_ERROR && zs.avail_out == 0) continue;            // output full: grow
    if (rc != Z_OK) break;
    if (in_left == 0 && zs.avail_out != 0) { rc = Z_DATA_ERROR; break; }   // truncated stream
  }
  inflateEnd(&zs);
  if (rc != Z_STREAM_END) return fail(GEECO_IO_ERR_FORMAT, "%s: corrupt or truncated compressed stream (zlib %d)", what, rc);
  out->resize(produced);
  return GEECO_IO_OK;
}

int deflate_all(const std::vector<uint8_t>& in, int window_bits, std::vector<uint8_t>* out) {
  z_stream zs;
  memset(&zs, 0, sizeof zs);
  if (deflateInit2(&zs, Z_DEFAULT_COMPRESSION, Z_DEFLATED, window_bits, 8, Z_DEFAULT_STRATEGY) != Z_OK)
    return fail(GEECO_IO_ERR_FORMAT, "deflateInit2 failed");
  out->resize(deflateBound(&zs, static_cast<uLong>(in.size())) + 64);
  zs.next_in = const_cast<Bytef*>(in.data());
  zs.avail_in = static_cast<uInt>(in.size());
  zs.next_out = out->data();
  zs.avail_out = static_cast<uInt>(out->size());
  int rc = deflate(&zs, Z_FINISH);
  size_t n = zs.total_out;
  deflateEnd(&zs);
  if (rc != Z_STREAM_END) return fail(GEECO_IO_ERR_FORMAT, "deflate failed (%d)", rc);
  out->resize(n);
  return GEECO_IO_OK;
}

bool ends_with(const std::string& s, const char* suffix) {
  size_t n = strlen(suffix);
  return s.size() >= n && s.compare(s.size() - n, n, suffix) == 0;
}

int resolve_compression(const std::string& path, int compression) {
  if (compression != GEECO_IO_COMPRESSION_AUTO) return compression;
  if (ends_with(path, ".zlib")) return GEECO_IO_COMPRESSION_ZLIB;
  if (ends_with(path, ".gzip") || ends_with(path, ".gz")) return GEECO_IO_COMPRESSION_GZIP;
  return GEECO_IO_COMPRESSION_NONE;
}

}  // namespace

// ================================================================================================
// public: misc + crc
// ================================================================================================
extern "C" const char* geeco_io_last_error(void) { return g_error.c_str(); }
extern "C" int geeco_io_version(void) { return 1; }
extern "C" uint32_t geeco_io_crc32c(const void* data, size_t n) { return crc_extend(0, data, n); }
extern "C" uint32_t geeco_io_crc32c_extend(uint32_t crc, const void* data, size_t n) { return crc_extend(crc, data, n); }
extern "C" uint32_t geeco_io_crc32c_mask(uint32_t crc) { return crc_mask(crc); }
extern "C" uint32_t geeco_io_crc32c_unmask(uint32_t masked) { return crc_unmask(masked); }

// ================================================================================================
// TFRecord
// ================================================================================================
struct geeco_tfrecord {
  std::vector<uint8_t> raw;                                  // the (inflated) file
  std::vector<std::pair<uint64_t, uint64_t>> records;        // (offset, length) of each payload
};

extern "C" int geeco_tfrecord_open(const char* path, int compression, int verify_crc, geeco_tfrecord** out) {
  if (!path || !out) return fail(GEECO_IO_ERR_ARG, "geeco_tfrecord_open: null argument");
  *out = nullptr;
  std::string p(path);
  int comp = resolve_compression(p, compression);
  if (comp < 0 || comp > GEECO_IO_COMPRESSION_GZIP) return fail(GEECO_IO_ERR_ARG, "unknown compression %d", compression);
  auto* f = new geeco_tfrecord;
  int rc;
  if (comp == GEECO_IO_COMPRESSION_NONE) {
    rc = read_file(p, &f->raw);
  } else {
    std::vector<uint8_t> packed;
    rc = read_file(p, &packed);
    if (rc == GEECO_IO_OK) rc = inflate_all(packed, comp == GEECO_IO_COMPRESSION_ZLIB ? 15 : 15 + 16, &f->raw, path);
  }
  if (rc != GEECO_IO_OK) { delete f; return rc; }
  const uint8_t* b = f->raw.data();
  uint64_t pos = 0, total = f->raw.size();
  while (pos < total) {
    if (total - pos < 12) { delete f; return fail(GEECO_IO_ERR_FORMAT, "%s: truncated record header at byte %llu", path, (unsigned long long)pos); }
    uint64_t len = load64(b + pos);
    if (verify_crc && crc_mask(crc_extend(0, b + pos, 8)) != load32(b + pos + 8)) {
      delete f; return fail(GEECO_IO_ERR_CRC, "%s: corrupt length checksum at byte %llu", path, (unsigned long long)pos);
    }
    if (len > total - pos - 12 || total - pos - 12 - len < 4) {
      delete f; return fail(GEECO_IO_ERR_FORMAT, "%s: truncated record of %llu bytes at byte %llu", path, (unsigned long long)len, (unsigned long long)pos);
    }
    if (verify_crc && crc_mask(crc_extend(0, b + pos + 12, len)) != load32(b + pos + 12 + len)) {
      size_t at = f->records.size();
      delete f; return fail(GEECO_IO_ERR_CRC, "%s: corrupt data checksum in record %zu", path, at);
    }
    f->records.emplace_back(pos + 12, len);
    pos += 12 + len + 4;
  }
  *out = f;
  return GEECO_IO_OK;
}

extern "C" int64_t geeco_tfrecord_count(const geeco_tfrecord* f) { return f ? static_cast<int64_t>(f->records.size()) : 0; }

extern "C" int geeco_tfrecord_get(const geeco_tfrecord* f, int64_t index, const uint8_t** data, uint64_t* len) {
  if (!f || !data || !len) return fail(GEECO_IO_ERR_ARG, "geeco_tfrecord_get: null argument");
  if (index < 0 || index >= static_cast<int64_t>(f->records.size()))
    return fail(GEECO_IO_ERR_ARG, "record index %lld out of range [0, %zu)", (long long)index, f->records.size());
  *data = f->raw.data() + f->records[index].first;
  *len = f->records[index].second;
  return GEECO_IO_OK;
}

extern "C" void geeco_tfrecord_close(geeco_tfrecord* f) { delete f; }

extern "C" int geeco_tfrecord_write(const char* path, int compression, int64_t count, const uint8_t* const* data,
                                    const uint64_t* len) {
  if (!path || count < 0 || (count && (!data || !len))) return fail(GEECO_IO_ERR_ARG, "geeco_tfrecord_write: bad argument");
  std::string p(path);
  int comp = resolve_compression(p, compression);
  if (comp < 0 || comp > GEECO_IO_COMPRESSION_GZIP) return fail(GEECO_IO_ERR_ARG, "unknown compression %d", compression);
  std::vector<uint8_t> raw;
  size_t total = 0;
  for (int64_t i = 0; i < count; ++i) total += 16 + len[i];
  raw.reserve(total);
  for (int64_t i = 0; i < count; ++i) {
    size_t at = raw.size();
    put64(&raw, len[i]);
    put32(&raw, crc_mask(crc_extend(0, raw.data() + at, 8)));
    put_bytes(&raw, data[i], len[i]);
    put32(&raw, crc_mask(crc_extend(0, data[i], len[i])));
  }
  if (comp == GEECO_IO_COMPRESSION_NONE) return write_file(p, raw.data(), raw.size());
  if (raw.size() >= (1ull << 32)) return fail(GEECO_IO_ERR_SIZE, "compressed writer is limited to 4 GiB per file");
  std::vector<uint8_t> packed;
  int rc = deflate_all(raw, comp == GEECO_IO_COMPRESSION_ZLIB ? 15 : 15 + 16, &packed);
  if (rc != GEECO_IO_OK) return rc;
  return write_file(p, packed.data(), packed.size());
}

// ================================================================================================
// SequenceExample
//   SequenceExample { Features context = 1; FeatureLists feature_lists = 2; }
//   Features        { map<string, Feature> feature = 1; }          map entry: key = 1, value = 2
//   FeatureLists    { map<string, FeatureList> feature_list = 1; }
//   FeatureList     { repeated Feature feature = 1; }
//   Feature         { oneof kind { BytesList bytes_list = 1; FloatList float_list = 2; Int64List int64_list = 3; } }
//   BytesList / FloatList (packed) / Int64List (packed) { repeated ... value = 1; }
// ================================================================================================
struct geeco_seqex {
  // name -> serialized Feature messages (one per frame; exactly one for a context feature)
  std::map<std::string, std::vector<Span>> keys[2];
  std::vector<const std::string*> order[2];                  // iteration order = sorted names
};

namespace {

bool parse_map_entry(Span entry, std::string* key, Span* value) {
  uint32_t field, wire;
  *value = Span();
  key->clear();
  while (!entry.empty()) {
    if (!entry.tag(&field, &wire)) return false;
    if (field == 1 && wire == 2) {
      Span k;
      if (!entry.bytes(&k)) return false;
      key->assign(reinterpret_cast<const char*>(k.p), k.size());
    } else if (field == 2 && wire == 2) {
      if (!entry.bytes(value)) return false;
    } else if (!entry.skip_value(wire)) {
      return false;
    }
  }
  return true;
}

// the kind of a serialized Feature and the payload of its list message (last one wins, as for any oneof)
bool feature_kind(Span feat, int* kind, Span* list) {
  uint32_t field, wire;
  *kind = GEECO_IO_KIND_NONE;
  *list = Span();
  while (!feat.empty()) {
    if (!feat.tag(&field, &wire)) return false;
    if (field >= 1 && field <= 3 && wire == 2) {
      if (!feat.bytes(list)) return false;
      *kind = static_cast<int>(field);
    } else if (!feat.skip_value(wire)) {
      return false;
    }
  }
  return true;
}

// number of values in a BytesList / FloatList / Int64List payload
bool count_values(int kind, Span list, int64_t* n) {
  uint32_t field, wire;
  int64_t c = 0;
  while (!list.empty()) {
    if (!list.tag(&field, &wire)) return false;
    if (field != 1) { if (!list.skip_value(wire)) return false; continue; }
    if (kind == GEECO_IO_KIND_BYTES) {
      Span s;
      if (wire != 2 || !list.bytes(&s)) return false;
      ++c;
    } else if (kind == GEECO_IO_KIND_FLOAT) {
      if (wire == 2) { Span s; if (!list.bytes(&s) || (s.size() & 3)) return false; c += static_cast<int64_t>(s.size() / 4); }
      else if (wire == 5) { if (!list.skip(4)) return false; ++c; }
      else return false;
    } else {
      if (wire == 2) {
        Span s;
        if (!list.bytes(&s)) return false;
        for (const uint8_t* q = s.p; q < s.e; ++q) c += !(*q & 0x80);
        if (s.size() && (s.e[-1] & 0x80)) return false;
      } else if (wire == 0) { uint64_t v; if (!list.varint(&v)) return false; ++c; }
      else return false;
    }
  }
  *n = c;
  return true;
}

template <typename Sink>   // Sink(const float*, count) for packed runs, one element at a time otherwise
bool each_float_run(Span list, Sink&& sink) {
  uint32_t field, wire;
  while (!list.empty()) {
    if (!list.tag(&field, &wire)) return false;
    if (field != 1) { if (!list.skip_value(wire)) return false; continue; }
    if (wire == 2) { Span s; if (!list.bytes(&s) || (s.size() & 3)) return false; if (!sink(s.p, s.size() / 4)) return false; }
    else if (wire == 5) { if (list.size() < 4) return false; if (!sink(list.p, 1)) return false; list.skip(4); }
    else return false;
  }
  return true;
}

int find_key(const geeco_seqex* s, int which, const char* name, const std::vector<Span>** out) {
  if (!s || !name || which < 0 || which > 1) return fail(GEECO_IO_ERR_ARG, "bad SequenceExample query");
  auto it = s->keys[which].find(name);
  if (it == s->keys[which].end())
    return fail(GEECO_IO_ERR_MISSING, "%s feature '%s' is not in the SequenceExample", which ? "sequence" : "context", name);
  *out = &it->second;
  return GEECO_IO_OK;
}

}  // namespace

extern "C" int geeco_seqex_parse(const uint8_t* data, uint64_t len, geeco_seqex** out) {
  if (!out || (!data && len)) return fail(GEECO_IO_ERR_ARG, "geeco_seqex_parse: null argument");
  *out = nullptr;
  auto* s = new geeco_seqex;
  Span top(data, len);
  uint32_t field, wire;
  bool ok = true;
  while (ok && !top.empty()) {
    ok = top.tag(&field, &wire);
    if (!ok) break;
    if ((field == 1 || field == 2) && wire == 2) {
      Span body;
      ok = top.bytes(&body);
      const int which = field == 1 ? GEECO_IO_CONTEXT : GEECO_IO_SEQUENCE;
      while (ok && !body.empty()) {
        uint32_t f2, w2;
        ok = body.tag(&f2, &w2);
        if (!ok) break;
        if (f2 != 1 || w2 != 2) { ok = body.skip_value(w2); continue; }
        Span entry, value;
        std::string key;
        ok = body.bytes(&entry) && parse_map_entry(entry, &key, &value);
        if (!ok) break;
        std::vector<Span>& slot = s->keys[which][key];
        slot.clear();                                     // duplicate map keys: the last entry wins
        if (which == GEECO_IO_CONTEXT) {
          slot.push_back(value);
        } else {                                          // FeatureList: repeated Feature feature = 1
          while (ok && !value.empty()) {
            uint32_t f3, w3;
            ok = value.tag(&f3, &w3);
            if (!ok) break;
            if (f3 == 1 && w3 == 2) { Span feat; ok = value.bytes(&feat); if (ok) slot.push_back(feat); }
            else ok = value.skip_value(w3);
          }
        }
      }
    } else {
      ok = top.skip_value(wire);
    }
  }
  if (!ok) { delete s; return fail(GEECO_IO_ERR_FORMAT, "malformed SequenceExample (%llu bytes)", (unsigned long long)len); }
  for (int w = 0; w < 2; ++w)
    for (auto& kv : s->keys[w]) s->order[w].push_back(&kv.first);
  *out = s;
  return GEECO_IO_OK;
}

extern "C" void geeco_seqex_free(geeco_seqex* s) { delete s; }

extern "C" int64_t geeco_seqex_num_keys(const geeco_seqex* s, int which) {
  return (s && which >= 0 && which <= 1) ? static_cast<int64_t>(s->order[which].size()) : 0;
}

extern "C" int geeco_seqex_key(const geeco_seqex* s, int which, int64_t i, const char** name, uint64_t* name_len) {
  if (!s || which < 0 || which > 1 || !name || !name_len || i < 0 || i >= static_cast<int64_t>(s->order[which].size()))
    return fail(GEECO_IO_ERR_ARG, "geeco_seqex_key: bad argument");
  *name = s->order[which][i]->c_str();
  *name_len = s->order[which][i]->size();
  return GEECO_IO_OK;
}

extern "C" int geeco_seqex_info(const geeco_seqex* s, int which, const char* name, int* kind, int64_t* frames,
                                int64_t* per_frame) {
  const std::vector<Span>* feats = nullptr;
  if (int rc = find_key(s, which, name, &feats)) return rc;
  int k_all = GEECO_IO_KIND_NONE;
  int64_t per = -2;
  for (const Span& f : *feats) {
    int k; Span list; int64_t n = 0;
    if (!feature_kind(f, &k, &list) || (k != GEECO_IO_KIND_NONE && !count_values(k, list, &n)))
      return fail(GEECO_IO_ERR_FORMAT, "malformed Feature in '%s'", name);
    if (k != GEECO_IO_KIND_NONE) {
      if (k_all != GEECO_IO_KIND_NONE && k_all != k) return fail(GEECO_IO_ERR_FORMAT, "feature '%s' mixes value kinds", name);
      k_all = k;
    }
    per = (per == -2 || per == n) ? n : -1;
  }
  if (kind) *kind = k_all;
  if (frames) *frames = static_cast<int64_t>(feats->size());
  if (per_frame) *per_frame = per == -2 ? 0 : per;
  return GEECO_IO_OK;
}

extern "C" int geeco_seqex_read_f32(const geeco_seqex* s, int which, const char* name, float* dst, int64_t capacity) {
  const std::vector<Span>* feats = nullptr;
  if (int rc = find_key(s, which, name, &feats)) return rc;
  if (!dst && capacity) return fail(GEECO_IO_ERR_ARG, "null destination");
  int64_t at = 0;
  for (const Span& f : *feats) {
    int k; Span list;
    if (!feature_kind(f, &k, &list)) return fail(GEECO_IO_ERR_FORMAT, "malformed Feature in '%s'", name);
    if (k == GEECO_IO_KIND_NONE) continue;
    if (k != GEECO_IO_KIND_FLOAT) return fail(GEECO_IO_ERR_FORMAT, "feature '%s' is not a float list", name);
    bool room = true;
    bool ok = each_float_run(list, [&](const uint8_t* p, size_t n) {
      if (at + static_cast<int64_t>(n) > capacity) { room = false; return false; }
      memcpy(dst + at, p, n * 4);
      at += static_cast<int64_t>(n);
      return true;
    });
    if (!room) return fail(GEECO_IO_ERR_SIZE, "feature '%s' holds more than %lld values", name, (long long)capacity);
    if (!ok) return fail(GEECO_IO_ERR_FORMAT, "malformed FloatList in '%s'", name);
  }
  if (at != capacity) return fail(GEECO_IO_ERR_SIZE, "feature '%s' holds %lld values, expected %lld", name, (long long)at, (long long)capacity);
  return GEECO_IO_OK;
}

extern "C" int geeco_seqex_read_u8(const geeco_seqex* s, int which, const char* name, uint8_t* dst, int64_t capacity,
                                   int64_t* inexact) {
  const std::vector<Span>* feats = nullptr;
  if (int rc = find_key(s, which, name, &feats)) return rc;
  if (!dst && capacity) return fail(GEECO_IO_ERR_ARG, "null destination");
  int64_t at = 0, bad = 0;
  for (const Span& f : *feats) {
    int k; Span list;
    if (!feature_kind(f, &k, &list)) return fail(GEECO_IO_ERR_FORMAT, "malformed Feature in '%s'", name);
    if (k == GEECO_IO_KIND_NONE) continue;
    if (k != GEECO_IO_KIND_FLOAT) return fail(GEECO_IO_ERR_FORMAT, "feature '%s' is not a float list", name);
    bool room = true;
    bool ok = each_float_run(list, [&](const uint8_t* p, size_t n) {
      if (at + static_cast<int64_t>(n) > capacity) { room = false; return false; }
      uint8_t* d = dst + at;
      int64_t b = 0;
      // blocks of 4096 values with a 32-bit mismatch counter: the inner loop has no 64-bit lane and vectorises
      for (size_t i0 = 0; i0 < n; i0 += 4096) {
        const size_t i1 = i0 + 4096 < n ? i0 + 4096 : n;
        uint32_t miss = 0;
        for (size_t i = i0; i < i1; ++i) {
          float v;
          memcpy(&v, p + 4 * i, 4);                    // unaligned-safe load
          float c = v > 0.f ? v : 0.f;                  // NaN compares false: becomes 0 and counts as inexact
          c = c < 255.f ? c : 255.f;
          const uint8_t q = static_cast<uint8_t>(static_cast<int32_t>(c));
          d[i] = q;
          miss += (static_cast<float>(q) != v);
        }
        b += miss;
      }
      bad += b;
      at += static_cast<int64_t>(n);
      return true;
    });
    if (!room) return fail(GEECO_IO_ERR_SIZE, "feature '%s' holds more than %lld values", name, (long long)capacity);
    if (!ok) return fail(GEECO_IO_ERR_FORMAT, "malformed FloatList in '%s'", name);
  }
  if (at != capacity) return fail(GEECO_IO_ERR_SIZE, "feature '%s' holds %lld values, expected %lld", name, (long long)at, (long long)capacity);
  if (inexact) *inexact = bad;
  return GEECO_IO_OK;
}

extern "C" int geeco_seqex_read_i64(const geeco_seqex* s, int which, const char* name, int64_t* dst, int64_t capacity) {
  const std::vector<Span>* feats = nullptr;
  if (int rc = find_key(s, which, name, &feats)) return rc;
  if (!dst && capacity) return fail(GEECO_IO_ERR_ARG, "null destination");
  int64_t at = 0;
  for (const Span& f : *feats) {
    int k; Span list;
    if (!feature_kind(f, &k, &list)) return fail(GEECO_IO_ERR_FORMAT, "malformed Feature in '%s'", name);
    if (k == GEECO_IO_KIND_NONE) continue;
    if (k != GEECO_IO_KIND_INT64) return fail(GEECO_IO_ERR_FORMAT, "feature '%s' is not an int64 list", name);
    uint32_t field, wire;
    while (!list.empty()) {
      if (!list.tag(&field, &wire)) return fail(GEECO_IO_ERR_FORMAT, "malformed Int64List in '%s'", name);
      if (field != 1) { if (!list.skip_value(wire)) return fail(GEECO_IO_ERR_FORMAT, "malformed Int64List in '%s'", name); continue; }
      Span run;
      if (wire == 2) { if (!list.bytes(&run)) return fail(GEECO_IO_ERR_FORMAT, "malformed Int64List in '%s'", name); }
      else if (wire != 0) return fail(GEECO_IO_ERR_FORMAT, "malformed Int64List in '%s'", name);
      Span& from = wire == 2 ? run : list;          // packed run, or one unpacked varint
      do {
        uint64_t v;
        if (!from.varint(&v)) return fail(GEECO_IO_ERR_FORMAT, "malformed Int64List in '%s'", name);
        if (at >= capacity) return fail(GEECO_IO_ERR_SIZE, "feature '%s' holds more than %lld values", name, (long long)capacity);
        dst[at++] = static_cast<int64_t>(v);
      } while (wire == 2 && !run.empty());
    }
  }
  if (at != capacity) return fail(GEECO_IO_ERR_SIZE, "feature '%s' holds %lld values, expected %lld", name, (long long)at, (long long)capacity);
  return GEECO_IO_OK;
}

extern "C" int geeco_seqex_bytes(const geeco_seqex* s, int which, const char* name, int64_t frame, int64_t j,
                                 const uint8_t** data, uint64_t* len) {
  const std::vector<Span>* feats = nullptr;
  if (int rc = find_key(s, which, name, &feats)) return rc;
  if (!data || !len || frame < 0 || frame >= static_cast<int64_t>(feats->size()) || j < 0)
    return fail(GEECO_IO_ERR_ARG, "geeco_seqex_bytes: bad argument");
  int k; Span list;
  if (!feature_kind((*feats)[frame], &k, &list)) return fail(GEECO_IO_ERR_FORMAT, "malformed Feature in '%s'", name);
  if (k != GEECO_IO_KIND_BYTES) return fail(GEECO_IO_ERR_FORMAT, "feature '%s' is not a bytes list", name);
  uint32_t field, wire;
  int64_t seen = 0;
  while (!list.empty()) {
    if (!list.tag(&field, &wire)) return fail(GEECO_IO_ERR_FORMAT, "malformed BytesList in '%s'", name);
    if (field == 1 && wire == 2) {
      Span v;
      if (!list.bytes(&v)) return fail(GEECO_IO_ERR_FORMAT, "malformed BytesList in '%s'", name);
      if (seen++ == j) { *data = v.p; *len = v.size(); return GEECO_IO_OK; }
    } else if (!list.skip_value(wire)) {
      return fail(GEECO_IO_ERR_FORMAT, "malformed BytesList in '%s'", name);
    }
  }
  return fail(GEECO_IO_ERR_ARG, "feature '%s' frame %lld has %lld values, asked for %lld", name, (long long)frame, (long long)seen, (long long)j);
}

// ================================================================================================
// sliding windows
// ================================================================================================
extern "C" int geeco_io_window_gather(const void* src, int64_t frames, int64_t frame_bytes, int64_t K, int64_t w0,
                                      int64_t nwin, void* dst) {
  if (!src || !dst || frame_bytes <= 0 || K <= 0 || w0 < 0 || nwin < 0)
    return fail(GEECO_IO_ERR_ARG, "geeco_io_window_gather: bad argument");
  if (nwin && w0 + nwin + K - 1 > frames)
    return fail(GEECO_IO_ERR_SIZE, "windows %lld..%lld of %lld frames reach past frame %lld", (long long)w0,
                (long long)(w0 + nwin - 1), (long long)K, (long long)(frames - 1));
  const uint8_t* s = static_cast<const uint8_t*>(src);
  uint8_t* d = static_cast<uint8_t*>(dst);
  const size_t wb = static_cast<size_t>(K) * frame_bytes;
  auto copy_range = [=](int64_t lo, int64_t hi) {
    for (int64_t i = lo; i < hi; ++i) memcpy(d + i * wb, s + (w0 + i) * frame_bytes, wb);
  };
  // image windows (tens of MB per batch) are split over a few threads: one memcpy stream does not saturate the
  // host's memory bandwidth; small state vectors stay on the calling thread
  const size_t total = static_cast<size_t>(nwin) * wb;
  int nthreads = total >= (8u << 20) ? static_cast<int>(std::min<int64_t>(4, nwin)) : 1;
  if (nthreads <= 1) {
    copy_range(0, nwin);
    return GEECO_IO_OK;
  }
  std::vector<std::thread> pool;
  const int64_t per = (nwin + nthreads - 1) / nthreads;
  for (int t = 1; t < nthreads; ++t) {
    const int64_t lo = t * per, hi = std::min<int64_t>(nwin, lo + per);
    if (lo < hi) pool.emplace_back(copy_range, lo, hi);
  }
  copy_range(0, std::min<int64_t>(nwin, per));
  for (auto& th : pool) th.join();
  return GEECO_IO_OK;
}

// ================================================================================================
// TF V2 checkpoint bundles
// ================================================================================================
namespace {

const uint64_t kTableMagic = 0xdb4775248b80fb57ull;
const size_t kFooterSize = 48;        // two block handles padded to 40 bytes + 8 bytes of magic
const size_t kBlockTrailer = 5;       // 1 byte compression type + 4 bytes masked crc32c

struct BundleEntry {
  int dtype = 0;
  std::vector<int64_t> dims;
  bool unknown_rank = false;
  int32_t shard = 0;
  int64_t offset = 0, size = 0;
  uint32_t crc = 0;                   // masked
  bool sliced = false;
};

// snappy raw format: varint uncompressed length, then literal / copy elements
bool snappy_uncompress(Span in, std::vector<uint8_t>* out) {
  uint64_t n;
  if (!in.varint(&n) || n > (1ull << 32)) return false;
  out->clear();
  out->reserve(n);
  while (!in.empty()) {
    uint8_t tag = *in.p++;
    size_t len, off;
    switch (tag & 3) {
      case 0: {
        len = (tag >> 2) + 1;
        if (len > 60) {
          size_t extra = len - 60;
          if (in.size() < extra) return false;
          len = 0;
          for (size_t i = 0; i < extra; ++i) len |= static_cast<size_t>(in.p[i]) << (8 * i);
          len += 1;
          in.p += extra;
        }
        if (in.size() < len) return false;
        out->insert(out->end(), in.p, in.p + len);
        in.p += len;
        continue;
      }
      case 1:
        if (in.size() < 1) return false;
        len = ((tag >> 2) & 7) + 4;
        off = (static_cast<size_t>(tag >> 5) << 8) | *in.p++;
        break;
      case 2:
        if (in.size() < 2) return false;
        len = (tag >> 2) + 1;
        off = in.p[0] | (static_cast<size_t>(in.p[1]) << 8);
        in.p += 2;
        break;
      default:
        if (in.size() < 4) return false;
        len = (tag >> 2) + 1;
        off = load32(in.p);
        in.p += 4;
        break;
    }
    if (off == 0 || off > out->size()) return false;
    size_t from = out->size() - off;
    for (size_t i = 0; i < len; ++i) out->push_back((*out)[from + i]);   // overlapping copies repeat the pattern
  }
  return out->size() == n;
}

// reads the block a handle points at, checks its trailer, undoes snappy
int read_block(const std::vector<uint8_t>& file, uint64_t offset, uint64_t size, std::vector<uint8_t>* out, const char* what) {
  if (offset > file.size() || size > file.size() - offset || file.size() - offset - size < kBlockTrailer)
    return fail(GEECO_IO_ERR_FORMAT, "%s: block handle (%llu, %llu) outside the file", what, (unsigned long long)offset, (unsigned long long)size);
  const uint8_t* b = file.data() + offset;
  if (crc_unmask(load32(b + size + 1)) != crc_extend(0, b, size + 1))
    return fail(GEECO_IO_ERR_CRC, "%s: block checksum mismatch at offset %llu", what, (unsigned long long)offset);
  if (b[size] == 0) { out->assign(b, b + size); return GEECO_IO_OK; }
  if (b[size] == 1) {
    if (!snappy_uncompress(Span(b, size), out)) return fail(GEECO_IO_ERR_FORMAT, "%s: corrupt snappy block", what);
    return GEECO_IO_OK;
  }
  return fail(GEECO_IO_ERR_FORMAT, "%s: unknown block compression %d", what, (int)b[size]);
}

// walks the prefix-compressed entries of a block
template <typename Fn>
bool each_block_entry(const std::vector<uint8_t>& block, Fn&& fn) {
  if (block.size() < 4) return false;
  uint32_t nrestarts = load32(block.data() + block.size() - 4);
  if (static_cast<uint64_t>(nrestarts) * 4 + 4 > block.size()) return false;
  Span s(block.data(), block.size() - 4 - static_cast<size_t>(nrestarts) * 4);
  std::string key;
  while (!s.empty()) {
    uint64_t shared, non_shared, vlen;
    if (!s.varint(&shared) || !s.varint(&non_shared) || !s.varint(&vlen)) return false;
    if (shared > key.size() || non_shared > s.size() || vlen > s.size() - non_shared) return false;
    key.resize(shared);
    key.append(reinterpret_cast<const char*>(s.p), non_shared);
    s.p += non_shared;
    Span value(s.p, vlen);
    s.p += vlen;
    if (!fn(key, value)) return false;
  }
  return true;
}

bool parse_shape(Span s, BundleEntry* e) {
  uint32_t field, wire;
  while (!s.empty()) {
    if (!s.tag(&field, &wire)) return false;
    if (field == 2 && wire == 2) {                 // repeated Dim dim = 2 { int64 size = 1; string name = 2; }
      Span dim;
      if (!s.bytes(&dim)) return false;
      int64_t size = 0;
      while (!dim.empty()) {
        uint32_t f2, w2;
        if (!dim.tag(&f2, &w2)) return false;
        if (f2 == 1 && w2 == 0) { uint64_t v; if (!dim.varint(&v)) return false; size = static_cast<int64_t>(v); }
        else if (!dim.skip_value(w2)) return false;
      }
      e->dims.push_back(size);
    } else if (field == 3 && wire == 0) {
      uint64_t v; if (!s.varint(&v)) return false; e->unknown_rank = v != 0;
    } else if (!s.skip_value(wire)) {
      return false;
    }
  }
  return true;
}

bool parse_entry(Span s, BundleEntry* e) {
  uint32_t field, wire;
  while (!s.empty()) {
    if (!s.tag(&field, &wire)) return false;
    uint64_t v;
    if (field == 1 && wire == 0) { if (!s.varint(&v)) return false; e->dtype = static_cast<int>(v); }
    else if (field == 2 && wire == 2) { Span sh; if (!s.bytes(&sh) || !parse_shape(sh, e)) return false; }
    else if (field == 3 && wire == 0) { if (!s.varint(&v)) return false; e->shard = static_cast<int32_t>(v); }
    else if (field == 4 && wire == 0) { if (!s.varint(&v)) return false; e->offset = static_cast<int64_t>(v); }
    else if (field == 5 && wire == 0) { if (!s.varint(&v)) return false; e->size = static_cast<int64_t>(v); }
    else if (field == 6 && wire == 5) { if (s.size() < 4) return false; e->crc = load32(s.p); s.skip(4); }
    else if (field == 7 && wire == 2) { Span sl; if (!s.bytes(&sl)) return false; e->sliced = true; }
    else if (!s.skip_value(wire)) return false;
  }
  return true;
}

void put_tag(std::vector<uint8_t>* b, uint32_t field, uint32_t wire) { put_varint(b, (static_cast<uint64_t>(field) << 3) | wire); }

std::vector<uint8_t> encode_entry(const BundleEntry& e) {
  std::vector<uint8_t> shape;
  for (int64_t d : e.dims) {
    std::vector<uint8_t> dim;
    if (d != 0) { put_tag(&dim, 1, 0); put_varint(&dim, static_cast<uint64_t>(d)); }
    put_tag(&shape, 2, 2); put_varint(&shape, dim.size()); put_bytes(&shape, dim.data(), dim.size());
  }
  std::vector<uint8_t> out;                                     // proto3: default-valued scalars are omitted
  if (e.dtype) { put_tag(&out, 1, 0); put_varint(&out, static_cast<uint64_t>(e.dtype)); }
  put_tag(&out, 2, 2); put_varint(&out, shape.size()); put_bytes(&out, shape.data(), shape.size());
  if (e.shard) { put_tag(&out, 3, 0); put_varint(&out, static_cast<uint64_t>(e.shard)); }
  if (e.offset) { put_tag(&out, 4, 0); put_varint(&out, static_cast<uint64_t>(e.offset)); }
  if (e.size) { put_tag(&out, 5, 0); put_varint(&out, static_cast<uint64_t>(e.size)); }
  if (e.crc) { put_tag(&out, 6, 5); put32(&out, e.crc); }
  return out;
}

std::string shard_path(const std::string& prefix, int shard, int num_shards) {
  char buf[64];
  snprintf(buf, sizeof buf, ".data-%05d-of-%05d", shard, num_shards);
  return prefix + buf;
}

// builds one leveldb block: prefix compression with a restart point every 16 entries
struct BlockBuilder {
  std::vector<uint8_t> buf;
  std::vector<uint32_t> restarts{0};
  std::string last;
  int since_restart = 0;
  bool empty() const { return buf.empty(); }
  void add(const std::string& key, const uint8_t* value, size_t vlen) {
    size_t shared = 0;
    if (since_restart < 16) {
      size_t lim = std::min(last.size(), key.size());
      while (shared < lim && last[shared] == key[shared]) ++shared;
    } else {
      restarts.push_back(static_cast<uint32_t>(buf.size()));
      since_restart = 0;
    }
    put_varint(&buf, shared);
    put_varint(&buf, key.size() - shared);
    put_varint(&buf, vlen);
    put_bytes(&buf, key.data() + shared, key.size() - shared);
    put_bytes(&buf, value, vlen);
    last = key;
    ++since_restart;
  }
  std::vector<uint8_t> finish() {
    std::vector<uint8_t> out = buf;
    for (uint32_t r : restarts) put32(&out, r);
    put32(&out, static_cast<uint32_t>(restarts.size()));
    return out;
  }
};

// appends block + trailer to the file image, returns its handle encoding
std::vector<uint8_t> emit_block(std::vector<uint8_t>* file, const std::vector<uint8_t>& block) {
  std::vector<uint8_t> handle;
  put_varint(&handle, file->size());
  put_varint(&handle, block.size());
  size_t at = file->size();
  put_bytes(file, block.data(), block.size());
  file->push_back(0);                                          // no compression
  put32(file, crc_mask(crc_extend(0, file->data() + at, block.size() + 1)));
  return handle;
}

}  // namespace

struct geeco_bundle {
  std::string prefix;
  int num_shards = 1;
  std::map<std::string, BundleEntry> entries;
  std::vector<const std::string*> order;
};

extern "C" int geeco_bundle_open(const char* prefix, geeco_bundle** out) {
  if (!prefix || !out) return fail(GEECO_IO_ERR_ARG, "geeco_bundle_open: null argument");
  *out = nullptr;
  std::string index_path = std::string(prefix) + ".index";
  std::vector<uint8_t> file;
  int rc = read_file(index_path, &file);
  if (rc != GEECO_IO_OK) return rc;
  const char* what = index_path.c_str();
  if (file.size() < kFooterSize || load64(file.data() + file.size() - 8) != kTableMagic)
    return fail(GEECO_IO_ERR_FORMAT, "%s is not a checkpoint index table (bad magic)", what);
  Span footer(file.data() + file.size() - kFooterSize, kFooterSize - 8);
  uint64_t meta_off, meta_size, idx_off, idx_size;
  if (!footer.varint(&meta_off) || !footer.varint(&meta_size) || !footer.varint(&idx_off) || !footer.varint(&idx_size))
    return fail(GEECO_IO_ERR_FORMAT, "%s: corrupt table footer", what);
  std::vector<uint8_t> index_block;
  rc = read_block(file, idx_off, idx_size, &index_block, what);
  if (rc != GEECO_IO_OK) return rc;
  auto* b = new geeco_bundle;
  b->prefix = prefix;
  bool header_seen = false;
  int inner = GEECO_IO_OK;
  bool ok = each_block_entry(index_block, [&](const std::string&, Span handle) {
    uint64_t off, size;
    if (!handle.varint(&off) || !handle.varint(&size)) return false;
    std::vector<uint8_t> block;
    inner = read_block(file, off, size, &block, what);
    if (inner != GEECO_IO_OK) return false;
    return each_block_entry(block, [&](const std::string& key, Span value) {
      if (key.empty()) {                           // BundleHeaderProto { int32 num_shards = 1; endianness = 2; version = 3; }
        uint32_t field, wire;
        header_seen = true;
        while (!value.empty()) {
          if (!value.tag(&field, &wire)) return false;
          uint64_t v;
          if (field == 1 && wire == 0) { if (!value.varint(&v)) return false; b->num_shards = static_cast<int>(v); }
          else if (field == 2 && wire == 0) { if (!value.varint(&v)) return false; if (v != 0) { inner = fail(GEECO_IO_ERR_FORMAT, "%s: big-endian bundles are not supported", what); return false; } }
          else if (!value.skip_value(wire)) return false;
        }
        return true;
      }
      BundleEntry e;
      if (!parse_entry(value, &e)) return false;
      b->entries[key] = std::move(e);
      return true;
    });
  });
  if (!ok || !header_seen) {
    delete b;
    if (inner != GEECO_IO_OK) return inner;
    return fail(GEECO_IO_ERR_FORMAT, "%s: %s", what, ok ? "no bundle header entry" : "corrupt table block");
  }
  for (auto& kv : b->entries) b->order.push_back(&kv.first);
  *out = b;
  return GEECO_IO_OK;
}

extern "C" void geeco_bundle_close(geeco_bundle* b) { delete b; }
extern "C" int64_t geeco_bundle_num_tensors(const geeco_bundle* b) { return b ? static_cast<int64_t>(b->order.size()) : 0; }

extern "C" int geeco_bundle_name(const geeco_bundle* b, int64_t i, const char** name, uint64_t* name_len) {
  if (!b || !name || !name_len || i < 0 || i >= static_cast<int64_t>(b->order.size()))
    return fail(GEECO_IO_ERR_ARG, "geeco_bundle_name: bad argument");
  *name = b->order[i]->c_str();
  *name_len = b->order[i]->size();
  return GEECO_IO_OK;
}

extern "C" int geeco_bundle_info(const geeco_bundle* b, const char* name, int* dtype, int* ndim, int64_t* dims,
                                 int64_t* nbytes) {
  if (!b || !name) return fail(GEECO_IO_ERR_ARG, "geeco_bundle_info: null argument");
  auto it = b->entries.find(name);
  if (it == b->entries.end()) return fail(GEECO_IO_ERR_MISSING, "tensor '%s' is not in checkpoint %s", name, b->prefix.c_str());
  const BundleEntry& e = it->second;
  if (e.dims.size() > 8) return fail(GEECO_IO_ERR_SIZE, "tensor '%s' has rank %zu > 8", name, e.dims.size());
  if (dtype) *dtype = e.dtype;
  if (ndim) *ndim = static_cast<int>(e.dims.size());
  if (dims) for (size_t i = 0; i < e.dims.size(); ++i) dims[i] = e.dims[i];
  if (nbytes) *nbytes = e.size;
  return GEECO_IO_OK;
}

extern "C" int geeco_bundle_read(const geeco_bundle* b, const char* name, void* dst, int64_t capacity, int verify_crc) {
  if (!b || !name || (!dst && capacity)) return fail(GEECO_IO_ERR_ARG, "geeco_bundle_read: null argument");
  auto it = b->entries.find(name);
  if (it == b->entries.end()) return fail(GEECO_IO_ERR_MISSING, "tensor '%s' is not in checkpoint %s", name, b->prefix.c_str());
  const BundleEntry& e = it->second;
  if (e.sliced) return fail(GEECO_IO_ERR_FORMAT, "tensor '%s' is stored as slices (partitioned variable): not supported", name);
  if (e.size != capacity) return fail(GEECO_IO_ERR_SIZE, "tensor '%s' holds %lld bytes, destination has %lld", name, (long long)e.size, (long long)capacity);
  std::string path = shard_path(b->prefix, e.shard, b->num_shards);
  FILE* fp = fopen(path.c_str(), "rb");
  if (!fp) return fail(GEECO_IO_ERR_FILE, "cannot open %s", path.c_str());
  size_t got = 0;
  if (fseek(fp, static_cast<long>(e.offset), SEEK_SET) == 0 && e.size) got = fread(dst, 1, static_cast<size_t>(e.size), fp);
  fclose(fp);
  if (got != static_cast<size_t>(e.size)) return fail(GEECO_IO_ERR_FILE, "%s: short read of tensor '%s'", path.c_str(), name);
  if (verify_crc && crc_unmask(e.crc) != crc_extend(0, dst, static_cast<size_t>(e.size)))
    return fail(GEECO_IO_ERR_CRC, "tensor '%s': checksum mismatch in %s", name, path.c_str());
  return GEECO_IO_OK;
}

struct geeco_bundle_writer {
  std::string prefix;
  std::map<std::string, std::pair<BundleEntry, std::vector<uint8_t>>> tensors;
};

extern "C" int geeco_bundle_writer_create(const char* prefix, geeco_bundle_writer** out) {
  if (!prefix || !out) return fail(GEECO_IO_ERR_ARG, "geeco_bundle_writer_create: null argument");
  *out = new geeco_bundle_writer;
  (*out)->prefix = prefix;
  return GEECO_IO_OK;
}

extern "C" int geeco_bundle_writer_add(geeco_bundle_writer* w, const char* name, int dtype, int ndim, const int64_t* dims,
                                       const void* data, int64_t nbytes) {
  if (!w || !name || !*name || ndim < 0 || ndim > 8 || (ndim && !dims) || nbytes < 0 || (nbytes && !data))
    return fail(GEECO_IO_ERR_ARG, "geeco_bundle_writer_add: bad argument");
  int64_t width = dtype == GEECO_IO_DT_FLOAT || dtype == GEECO_IO_DT_INT32 ? 4 : dtype == GEECO_IO_DT_INT64 ? 8 : 0;
  if (!width) return fail(GEECO_IO_ERR_ARG, "tensor '%s': unsupported dtype %d", name, dtype);
  int64_t numel = 1;
  for (int i = 0; i < ndim; ++i) { if (dims[i] < 0) return fail(GEECO_IO_ERR_ARG, "tensor '%s': negative dimension", name); numel *= dims[i]; }
  if (numel * width != nbytes) return fail(GEECO_IO_ERR_SIZE, "tensor '%s': %lld bytes given, shape needs %lld", name, (long long)nbytes, (long long)(numel * width));
  if (w->tensors.count(name)) return fail(GEECO_IO_ERR_ARG, "tensor '%s' added twice", name);
  auto& slot = w->tensors[name];
  slot.first.dtype = dtype;
  slot.first.dims.assign(dims, dims + ndim);
  slot.first.size = nbytes;
  slot.second.assign(static_cast<const uint8_t*>(data), static_cast<const uint8_t*>(data) + nbytes);
  return GEECO_IO_OK;
}

extern "C" void geeco_bundle_writer_abort(geeco_bundle_writer* w) { delete w; }

extern "C" int geeco_bundle_writer_finish(geeco_bundle_writer* w) {
  if (!w) return fail(GEECO_IO_ERR_ARG, "geeco_bundle_writer_finish: null argument");
  // data shard: tensors back to back in name order
  std::vector<uint8_t> data;
  for (auto& kv : w->tensors) {
    BundleEntry& e = kv.second.first;
    e.offset = static_cast<int64_t>(data.size());
    e.crc = crc_mask(crc_extend(0, kv.second.second.data(), kv.second.second.size()));
    put_bytes(&data, kv.second.second.data(), kv.second.second.size());
  }
  int rc = write_file(shard_path(w->prefix, 0, 1), data.data(), data.size());
  if (rc != GEECO_IO_OK) { delete w; return rc; }
  // index table: "" -> header {num_shards: 1, version {producer: 1}}, then every entry; 256 KiB data blocks
  std::vector<uint8_t> file;
  BlockBuilder block, index;
  std::string last_key;
  auto flush = [&]() {
    if (block.empty()) return;
    std::vector<uint8_t> handle = emit_block(&file, block.finish());
    index.add(last_key, handle.data(), handle.size());
    block = BlockBuilder();
  };
  const uint8_t header[] = {0x08, 0x01, 0x1a, 0x02, 0x08, 0x01};
  block.add("", header, sizeof header);
  last_key = "";
  for (auto& kv : w->tensors) {
    std::vector<uint8_t> enc = encode_entry(kv.second.first);
    block.add(kv.first, enc.data(), enc.size());
    last_key = kv.first;
    if (block.buf.size() >= 262144) flush();
  }
  flush();
  BlockBuilder meta;
  std::vector<uint8_t> meta_handle = emit_block(&file, meta.finish());
  std::vector<uint8_t> index_handle = emit_block(&file, index.finish());
  std::vector<uint8_t> footer;
  put_bytes(&footer, meta_handle.data(), meta_handle.size());
  put_bytes(&footer, index_handle.data(), index_handle.size());
  footer.resize(kFooterSize - 8, 0);
  put64(&footer, kTableMagic);
  put_bytes(&file, footer.data(), footer.size());
  rc = write_file(w->prefix + ".index", file.data(), file.size());
  delete w;
  return rc;
}
