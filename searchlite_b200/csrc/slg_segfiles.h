// slg_segfiles.h — host-side readers of the reference's on-disk segment files.
//
// Nothing here touches the device: these functions turn the bytes the reference's writer produced
// into the tables the residency loader uploads.  Formats, all little-endian
// (paths relative to searchlite-core/src/):
//   seg_<id>.terms   index/terms.rs:10-25      u64 n | n x { varint len | bytes | u64 post offset } | crc32(entries)
//   seg_<id>.post    index/postings.rs:78-129  per-term posting lists at the offsets of .terms
//   seg_<id>.fast    index/fastfields.rs:409-424, 910-1134   "FFV1" | u32 n_fields | fields (HashMap order)
//   seg_<id>.meta    index/segment.rs:43-53    pretty-printed JSON: avg_field_lengths, vector_fields, doc ids
//   <field>.bin      index/segment.rs:1030-1053  "VCTR" vector store
//   MANIFEST.json    index/manifest.rs:14-47   segments[]: paths, doc_count, deleted_docs, checksums
// Checksums are crc32 (IEEE, crc32fast 1.5.0 — util/checksum.rs:3-7) of whole files.
#pragma once
#include <algorithm>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

namespace slgf {

// ---- crc32 (IEEE 802.3, reflected, poly 0xEDB88320), slice-by-8 ----
struct Crc32Tables {
  uint32_t t[8][256];
  Crc32Tables() {
    for (uint32_t i = 0; i < 256; i++) {
      uint32_t c = i;
      for (int k = 0; k < 8; k++) c = (c & 1u) ? (c >> 1) ^ 0xEDB88320u : c >> 1;
      t[0][i] = c;
    }
    for (uint32_t i = 0; i < 256; i++)
      for (int s = 1; s < 8; s++) t[s][i] = (t[s - 1][i] >> 8) ^ t[0][t[s - 1][i] & 0xFFu];
  }
};

inline uint32_t crc32(const uint8_t *p, size_t n) {
  static const Crc32Tables T;
  uint32_t c = 0xFFFFFFFFu;
  while (n >= 8) {
    uint32_t lo, hi;
    std::memcpy(&lo, p, 4);
    std::memcpy(&hi, p + 4, 4);
    lo ^= c;
    c = T.t[7][lo & 0xFF] ^ T.t[6][(lo >> 8) & 0xFF] ^ T.t[5][(lo >> 16) & 0xFF] ^ T.t[4][lo >> 24] ^ T.t[3][hi & 0xFF] ^
        T.t[2][(hi >> 8) & 0xFF] ^ T.t[1][(hi >> 16) & 0xFF] ^ T.t[0][hi >> 24];
    p += 8;
    n -= 8;
  }
  while (n--) c = (c >> 8) ^ T.t[0][(c ^ *p++) & 0xFFu];
  return c ^ 0xFFFFFFFFu;
}

// crc(A || B) from crc(A), crc(B) and |B|: multiply crc(A) by x^(8|B|) in GF(2)[x] / P by repeated squaring of the
// "advance one zero bit" operator (the crc32_combine construction of zlib), then xor crc(B).
inline uint32_t gf2_times(const uint32_t *mat, uint32_t vec) {
  uint32_t sum = 0;
  for (; vec; vec >>= 1, mat++)
    if (vec & 1u) sum ^= *mat;
  return sum;
}
inline void gf2_square(uint32_t *sq, const uint32_t *mat) {
  for (int n = 0; n < 32; n++) sq[n] = gf2_times(mat, mat[n]);
}
inline uint32_t crc32_combine(uint32_t crc1, uint32_t crc2, uint64_t len2) {
  if (len2 == 0) return crc1;
  uint32_t even[32], odd[32];
  odd[0] = 0xEDB88320u;  // one zero bit
  for (int n = 1; n < 32; n++) odd[n] = 1u << (n - 1);
  gf2_square(even, odd);  // two zero bits
  gf2_square(odd, even);  // four
  do {
    gf2_square(even, odd);  // first pass: one zero byte
    if (len2 & 1u) crc1 = gf2_times(even, crc1);
    len2 >>= 1;
    if (!len2) break;
    gf2_square(odd, even);
    if (len2 & 1u) crc1 = gf2_times(odd, crc1);
    len2 >>= 1;
  } while (len2);
  return crc1 ^ crc2;
}

// whole-file checksum on all host cores: the `.post` file of a 10 M-doc index is ~12 GB
inline uint32_t crc32_parallel(const uint8_t *p, size_t n, unsigned max_threads = 0) {
  const size_t kMinChunk = 4u << 20;
  unsigned hw = max_threads ? max_threads : std::thread::hardware_concurrency();
  if (hw == 0) hw = 1;
  const size_t parts = std::min<size_t>(std::min<size_t>(hw, 64), n / kMinChunk);
  if (parts <= 1) return crc32(p, n);
  std::vector<uint32_t> crc(parts);
  std::vector<size_t> lo(parts + 1);
  for (size_t i = 0; i <= parts; i++) lo[i] = n / parts * i;
  lo[parts] = n;
  std::vector<std::thread> th;
  for (size_t i = 1; i < parts; i++) th.emplace_back([&, i] { crc[i] = crc32(p + lo[i], lo[i + 1] - lo[i]); });
  crc[0] = crc32(p, lo[1]);
  for (auto &t : th) t.join();
  uint32_t c = crc[0];
  for (size_t i = 1; i < parts; i++) c = crc32_combine(c, crc[i], lo[i + 1] - lo[i]);
  return c;
}

// ---- LEB128 as util/varint.rs:22-35 (read_u64: no length limit other than the buffer) ----
inline bool read_varint_u64(const uint8_t *buf, size_t n, uint64_t &value, size_t &used) {
  uint32_t shift = 0;
  value = 0;
  for (size_t i = 0; i < n; i++) {
    const uint8_t b = buf[i];
    if (shift < 64) value |= (uint64_t)(b & 0x7F) << shift;
    if (!(b & 0x80)) {
      used = i + 1;
      return true;
    }
    shift += 7;
  }
  return false;
}

// ---- .terms ----
struct TermEntry {
  const char *key;  // not NUL-terminated, points into the file image
  uint32_t key_len;
  uint64_t offset;  // of the term's list in .post
};

// index/terms.rs:27-75 (read_terms): length check, crc over the entry bytes, then the entries.
inline bool parse_terms(const uint8_t *buf, uint64_t n, std::vector<TermEntry> &out, std::string &err) {
  if (n < 12) {
    err = "terms file is truncated";
    return false;
  }
  uint64_t count;
  std::memcpy(&count, buf, 8);
  const uint8_t *data = buf + 8;
  const uint64_t dn = n - 12;
  uint32_t expected;
  std::memcpy(&expected, buf + n - 4, 4);
  if (crc32_parallel(data, dn) != expected) {
    err = "terms file failed checksum validation";
    return false;
  }
  out.clear();
  out.reserve((size_t)std::min<uint64_t>(count, dn / 10 + 1));
  uint64_t cur = 0;
  for (uint64_t i = 0; i < count; i++) {
    uint64_t len;
    size_t used;
    if (!read_varint_u64(data + cur, dn - cur, len, used)) {
      err = "unterminated varint in terms file";
      return false;
    }
    cur += used;
    if (len > dn - cur) {
      err = "terms file ended unexpectedly while reading term";
      return false;
    }
    TermEntry e;
    e.key = reinterpret_cast<const char *>(data + cur);
    e.key_len = (uint32_t)len;
    cur += len;
    if (cur + 8 > dn) {
      err = "terms file ended unexpectedly while reading offset";
      return false;
    }
    std::memcpy(&e.offset, data + cur, 8);
    cur += 8;
    out.push_back(e);
  }
  return true;
}

// ---- .fast ----
struct FastColumn {
  std::string name;
  int type = -1;  // FieldType code, index/fastfields.rs:41-56: 0 I64, 1 F64, 2 Str, 3 I64List, 4 F64List, 5 StrList,
                  // 6 I64Nested, 7 F64Nested, 8 StrNested, 9 / 10 nested bookkeeping (walked over)
  uint32_t doc_len = 0;
  const uint8_t *presence = nullptr;  // I64 / F64: doc_len bytes
  const uint8_t *values = nullptr;    // I64 / F64: doc_len x 8 bytes (unaligned); Str: doc_len x u32 ords;
                                      // lists / nested: n_values x (8 | 4) bytes
  std::vector<std::string> dict;      // Str / StrList / StrNested
  // lists and nested columns: the values of doc d are values[list_offsets[d] .. list_offsets[d+1]).  A nested column
  // (doc -> objects -> values) is flattened here: "any value of any object" (fastfields.rs:510-527, 597-610) is
  // "any value of the doc" because the object offsets are running sums (write_field, fastfields.rs:939-965).
  std::vector<uint32_t> list_offsets;  // doc_len + 1 entries
  uint64_t n_values = 0;
  int kind() const { return type <= 2 ? type : (type <= 8 ? 3 + (type - 3) % 3 : -1); }  // 0..2 scalar, 3..5 list forms
};

struct Cursor {
  const uint8_t *p;
  uint64_t n, pos = 0;
  bool u32(uint32_t &v) {
    if (pos + 4 > n) return false;
    std::memcpy(&v, p + pos, 4);
    pos += 4;
    return true;
  }
  bool skip(uint64_t k) {
    if (k > n - pos) return false;
    pos += k;
    return true;
  }
  // u32 array of `count` entries; returns its last element (0 when empty) — the writer's running offsets
  bool offsets(uint64_t count, uint32_t &last, std::vector<uint32_t> *keep = nullptr) {
    last = 0;
    if (count == 0) return true;
    if (count > (n - pos) / 4) return false;
    std::memcpy(&last, p + pos + (count - 1) * 4, 4);
    if (keep) {
      keep->resize(count);
      std::memcpy(keep->data(), p + pos, count * 4);
    }
    pos += count * 4;
    return true;
  }
};

// read_fields, index/fastfields.rs:1166-1436.  Scalar columns are returned as views into the image;
// list / nested / bookkeeping columns are walked over (their sizes are data dependent) and reported
// with their type so that the caller can say what it ignored.
inline bool parse_fast(const uint8_t *buf, uint64_t n, std::vector<FastColumn> &out, std::string &err) {
  Cursor c{buf, n};
  if (n < 8 || std::memcmp(buf, "FFV1", 4) != 0) {
    err = "fast-field file lacks the FFV1 magic";
    return false;
  }
  c.pos = 4;
  uint32_t n_fields;
  if (!c.u32(n_fields)) return false;
  auto bad = [&](const char *what) {
    err = std::string("fast-field file ended unexpectedly in ") + what;
    return false;
  };
  for (uint32_t f = 0; f < n_fields; f++) {
    uint32_t name_len;
    if (!c.u32(name_len) || name_len > c.n - c.pos) return bad("a field name");
    FastColumn col;
    col.name.assign(reinterpret_cast<const char *>(buf + c.pos), name_len);
    c.pos += name_len;
    if (c.pos + 1 > c.n) return bad("a field type");
    col.type = buf[c.pos++];
    if (!c.u32(col.doc_len)) return bad("a column length");
    const uint64_t dl = col.doc_len;
    uint32_t last, last2;
    auto read_dict = [&]() {
      uint32_t dict_len;
      if (!c.u32(dict_len)) return false;
      for (uint32_t i = 0; i < dict_len; i++) {
        uint32_t bl;
        if (!c.u32(bl) || bl > c.n - c.pos) return false;
        col.dict.emplace_back(reinterpret_cast<const char *>(buf + c.pos), bl);
        c.pos += bl;
      }
      return true;
    };
    // running offsets: ascending and ending at `last` (every entry then indexes inside the values / the next table)
    auto monotone = [](const std::vector<uint32_t> &o, uint32_t last_v) {
      for (size_t i = 1; i < o.size(); i++)
        if (o[i] < o[i - 1]) return false;
      return o.empty() || o.back() == last_v;
    };
    switch (col.type) {
      case 0:
      case 1:
        col.presence = buf + c.pos;
        if (!c.skip(dl)) return bad("a presence vector");
        col.values = buf + c.pos;
        if (!c.skip(dl * 8)) return bad("a numeric column");
        break;
      case 2:
        if (!read_dict()) return bad("a keyword dictionary");
        col.values = buf + c.pos;
        if (!c.skip(dl * 4)) return bad("a keyword column");
        break;
      case 3:
      case 4:
      case 5: {
        if (col.type == 5 && !read_dict()) return bad("a keyword dictionary");
        if (!c.offsets(dl + 1, last, &col.list_offsets)) return bad("a list column's offsets");
        col.values = buf + c.pos;
        col.n_values = last;
        if (!c.skip((uint64_t)last * (col.type == 5 ? 4 : 8))) return bad("a list column");
        if (!monotone(col.list_offsets, last)) {
          err = "fast-field list column '" + col.name + "' has descending offsets";
          return false;
        }
        break;
      }
      case 6:
      case 7:
      case 8: {
        std::vector<uint32_t> doc_off, obj_off;
        if (col.type == 8 && !read_dict()) return bad("a keyword dictionary");
        if (!c.offsets(dl + 1, last, &doc_off) || !c.offsets((uint64_t)last + 1, last2, &obj_off)) return bad("a nested column's offsets");
        col.values = buf + c.pos;
        col.n_values = last2;
        if (!c.skip((uint64_t)last2 * (col.type == 8 ? 4 : 8))) return bad("a nested column");
        if (!monotone(doc_off, last) || !monotone(obj_off, last2)) {
          err = "fast-field nested column '" + col.name + "' has descending offsets";
          return false;
        }
        col.list_offsets.resize(doc_off.size());
        for (size_t d = 0; d < doc_off.size(); d++) col.list_offsets[d] = obj_off[doc_off[d]];
        break;
      }
      case 9:
        if (!c.skip(dl * 4)) return bad("a nested count column");
        break;
      case 10:
        if (!c.offsets(dl + 1, last) || !c.skip((uint64_t)last * 4)) return bad("a nested parent column");
        break;
      default:
        err = "unknown fast-field type " + std::to_string(col.type);
        return false;
    }
    out.push_back(std::move(col));
  }
  return true;
}

// ---- JSON (just enough for .meta and MANIFEST.json, both written by serde_json) ----
struct Json {
  const char *b = nullptr, *e = nullptr;  // the value's text, trimmed
  bool ok() const { return b != nullptr; }
  char kind() const { return ok() && b < e ? *b : '\0'; }  // '{' '[' '"' digit/- t f n
};

inline const char *json_ws(const char *p, const char *e) {
  while (p < e && (*p == ' ' || *p == '\n' || *p == '\r' || *p == '\t')) p++;
  return p;
}
inline const char *json_skip_string(const char *p, const char *e) {  // p at the opening quote
  for (p++; p < e; p++) {
    if (*p == '\\') p++;
    else if (*p == '"') return p + 1;
  }
  return nullptr;
}
inline const char *json_skip_value(const char *p, const char *e) {
  p = json_ws(p, e);
  if (p >= e) return nullptr;
  if (*p == '"') return json_skip_string(p, e);
  if (*p == '{' || *p == '[') {
    int depth = 0;
    while (p < e) {
      if (*p == '"') {
        p = json_skip_string(p, e);
        if (!p) return nullptr;
        continue;
      }
      if (*p == '{' || *p == '[') depth++;
      else if (*p == '}' || *p == ']') {
        depth--;
        if (depth == 0) return p + 1;
      }
      p++;
    }
    return nullptr;
  }
  while (p < e && *p != ',' && *p != '}' && *p != ']' && *p != ' ' && *p != '\n' && *p != '\r' && *p != '\t') p++;
  return p;
}
inline std::string json_unescape(const char *b, const char *e) {  // b..e inside the quotes
  std::string s;
  for (const char *p = b; p < e; p++) {
    if (*p != '\\' || p + 1 >= e) {
      s.push_back(*p);
      continue;
    }
    p++;
    switch (*p) {
      case 'n': s.push_back('\n'); break;
      case 't': s.push_back('\t'); break;
      case 'r': s.push_back('\r'); break;
      case 'b': s.push_back('\b'); break;
      case 'f': s.push_back('\f'); break;
      case 'u': {
        unsigned cp = 0;
        for (int i = 0; i < 4 && p + 1 < e; i++) {
          p++;
          cp = cp * 16 + (unsigned)(*p <= '9' ? *p - '0' : (*p | 32) - 'a' + 10);
        }
        if (cp < 0x80) s.push_back((char)cp);
        else if (cp < 0x800) {
          s.push_back((char)(0xC0 | (cp >> 6)));
          s.push_back((char)(0x80 | (cp & 0x3F)));
        } else {
          s.push_back((char)(0xE0 | (cp >> 12)));
          s.push_back((char)(0x80 | ((cp >> 6) & 0x3F)));
          s.push_back((char)(0x80 | (cp & 0x3F)));
        }
        break;
      }
      default: s.push_back(*p);
    }
  }
  return s;
}
// member `key` of an object value
inline Json json_get(Json obj, const char *key) {
  Json none;
  if (obj.kind() != '{') return none;
  const char *p = obj.b + 1, *e = obj.e;
  const size_t kl = std::strlen(key);
  for (;;) {
    p = json_ws(p, e);
    if (p >= e || *p == '}') return none;
    if (*p == ',') {
      p++;
      continue;
    }
    if (*p != '"') return none;
    const char *ks = p + 1;
    const char *ke = json_skip_string(p, e);
    if (!ke) return none;
    const std::string k = json_unescape(ks, ke - 1);
    p = json_ws(ke, e);
    if (p >= e || *p != ':') return none;
    p = json_ws(p + 1, e);
    const char *ve = json_skip_value(p, e);
    if (!ve) return none;
    if (k.size() == kl && std::memcmp(k.data(), key, kl) == 0) {
      Json v;
      v.b = p;
      v.e = ve;
      return v;
    }
    p = ve;
  }
}
// elements of an array value
inline bool json_elements(Json arr, std::vector<Json> &out) {
  if (arr.kind() != '[') return false;
  const char *p = arr.b + 1, *e = arr.e;
  for (;;) {
    p = json_ws(p, e);
    if (p >= e) return false;
    if (*p == ']') return true;
    if (*p == ',') {
      p++;
      continue;
    }
    const char *ve = json_skip_value(p, e);
    if (!ve) return false;
    Json v;
    v.b = p;
    v.e = ve;
    out.push_back(v);
    p = ve;
  }
}
inline std::string json_string(Json v) { return v.kind() == '"' ? json_unescape(v.b + 1, v.e - 1) : std::string(); }
inline double json_number(Json v, double dflt = 0.0) {
  if (!v.ok() || v.kind() == '"' || v.kind() == '{' || v.kind() == '[' || v.kind() == 'n') return dflt;
  return std::strtod(std::string(v.b, v.e).c_str(), nullptr);
}
inline Json json_root(const uint8_t *buf, uint64_t n) {
  Json v;
  const char *b = reinterpret_cast<const char *>(buf), *e = b + n;
  b = json_ws(b, e);
  const char *ve = json_skip_value(b, e);
  if (ve) {
    v.b = b;
    v.e = ve;
  }
  return v;
}

// ---- <field>.bin (read_vector_file, index/segment.rs:1056-1119) ----
struct VectorFile {
  uint32_t dim = 0, doc_count = 0, vector_count = 0;
  uint8_t metric = 0;               // 0 cosine, 1 l2 (segment.rs:981-995)
  const uint8_t *offsets = nullptr; // doc_count x u32
  const uint8_t *values = nullptr;  // vector_count x dim x f32
};
inline bool parse_vector_file(const uint8_t *buf, uint64_t n, VectorFile &out, std::string &err) {
  if (n < 24) {
    err = "vector file is truncated";
    return false;
  }
  uint32_t magic, version;
  std::memcpy(&magic, buf, 4);
  std::memcpy(&version, buf + 4, 4);
  if (magic != 0x56435452u) {
    err = "invalid vector file magic";
    return false;
  }
  if (version != 1) {
    err = "unsupported vector file version " + std::to_string(version);
    return false;
  }
  std::memcpy(&out.dim, buf + 8, 4);
  out.metric = buf[12];
  if (out.metric > 1) {
    err = "unknown vector metric code " + std::to_string(out.metric);
    return false;
  }
  std::memcpy(&out.doc_count, buf + 16, 4);
  std::memcpy(&out.vector_count, buf + 20, 4);
  // (checked: vector_count * dim * 4 can pass 2^64)
  const uint64_t head = 24 + (uint64_t)out.doc_count * 4;
  const uint64_t row_bytes = (uint64_t)out.dim * 4;
  if (head > n || (row_bytes && (uint64_t)out.vector_count > (n - head) / row_bytes)) {
    err = "vector file ended unexpectedly";
    return false;
  }
  out.offsets = buf + 24;
  out.values = buf + 24 + (uint64_t)out.doc_count * 4;
  return true;
}

}  // namespace slgf
