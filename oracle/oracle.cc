/*
 * oracle.cc — CPU restatement of searchlite's BM25 / WAND / BMW top-k path.
 *
 * TEST INFRASTRUCTURE ONLY (see searchlite_oracle.h).  Nothing under searchlite_b200/
 * links or loads this file.  Every function cites the reference lines it follows;
 * citations are relative to /root/reference/searchlite-core/src/.
 *
 * Build: see oracle/Makefile (g++ -O2 -ffp-contract=off -pthread).  -ffp-contract=off matters:
 * Rust never fuses a*b+c, so the f32 arithmetic below must not be contracted either.
 */
#include "searchlite_oracle.h"

#include <algorithm>
#include <cctype>
#include <cmath>
#include <cstring>
#include <limits>
#include <string>
#include <unordered_map>
#include <vector>

#include <atomic>
#include <thread>

namespace {

constexpr uint32_t DOCID_END = 0xFFFFFFFFu;          // query/wand.rs:12
constexpr uint32_t DEFAULT_BLOCK_SIZE = 128;         // index/postings.rs:11
constexpr uint32_t BLOCK_META_FLAG = 1u << 31;       // index/postings.rs:12

/* ---------------------------------------------------------------- arithmetic */

// query/bm25.rs:1-6
inline float bm25(float tf, float df, float doc_len, float avgdl, float docs, float k1, float b) {
  // f32::max returns the non-NaN operand (ln of a negative ratio, df > N + 0.5 after deletions): fmaxf
  float idf = fmaxf(logf((docs - df + 0.5f) / (df + 0.5f)), 0.0f) + 1.0f;
  float norm_dl = avgdl > 0.0f ? doc_len / avgdl : 1.0f;
  float denom = tf + k1 * (1.0f - b + b * norm_dl);
  return idf * (tf * (k1 + 1.0f)) / fmaxf(denom, 1e-6f);
}

// query/wand.rs:269-286
inline float score_tf(float tf, float df, float doc_len, float avgdl, float docs, float k1, float b,
                      float weight) {
  float norm_len = doc_len > 0.0f ? doc_len : std::max(avgdl, tf);
  float base = bm25(tf, df, norm_len, avgdl, docs, k1, b);
  return base * weight;
}

// query/wand.rs:289-303
inline float upper_bound_tf(float tf, float df, float doc_len, float avgdl, float docs, float k1, float b,
                            float weight) {
  if (tf <= 0.0f) return 0.0f;
  return score_tf(tf, df, doc_len, avgdl, docs, k1, b, weight);
}

// f32::total_cmp (used by RankedDoc::cmp query/wand.rs:30-36 and finalize_heap :918-926)
inline int32_t total_key(float f) {
  int32_t bits;
  std::memcpy(&bits, &f, 4);
  bits ^= (int32_t)(((uint32_t)(bits >> 31)) >> 1);
  return bits;
}
inline int total_cmp(float a, float b) {
  int32_t ka = total_key(a), kb = total_key(b);
  return ka < kb ? -1 : (ka > kb ? 1 : 0);
}

struct RankedDoc {  // query/wand.rs:16-20
  uint32_t doc_id;
  float score;
};
// RankedDoc::cmp query/wand.rs:30-36: score total_cmp, then SMALLER doc id is greater
inline int ranked_cmp(const RankedDoc &a, const RankedDoc &b) {
  int c = total_cmp(a.score, b.score);
  if (c != 0) return c;
  return b.doc_id < a.doc_id ? -1 : (b.doc_id > a.doc_id ? 1 : 0);
}

/* ------------------------------------------------ Rust std BinaryHeap, restated
 * The pop order of equal keys decides the float summation order in wand_loop, so the
 * container is restated (alloc::collections::binary_heap: push/sift_up, pop/
 * sift_down_to_bottom, From<Vec>/rebuild) rather than replaced by std::priority_queue. */
template <class T, class Le /* a <= b in the heap's Ord */>
struct RustBinaryHeap {
  std::vector<T> data;
  Le le;
  explicit RustBinaryHeap(Le l) : le(l) {}
  bool empty() const { return data.empty(); }
  size_t size() const { return data.size(); }
  T &peek() { return data[0]; }
  void sift_up(size_t start, size_t pos) {
    T elem = std::move(data[pos]);
    while (pos > start) {
      size_t parent = (pos - 1) / 2;
      if (le(elem, data[parent])) break;
      data[pos] = std::move(data[parent]);
      pos = parent;
    }
    data[pos] = std::move(elem);
  }
  void sift_down_range(size_t pos, size_t end) {
    T elem = std::move(data[pos]);
    size_t child = 2 * pos + 1;
    size_t lim = end >= 2 ? end - 2 : 0;
    while (child <= lim && end >= 2) {
      if (le(data[child], data[child + 1])) child += 1;
      if (le(data[child], elem)) {  // hole.element() >= hole.get(child)
        data[pos] = std::move(elem);
        return;
      }
      data[pos] = std::move(data[child]);
      pos = child;
      child = 2 * pos + 1;
    }
    if (child + 1 == end && !le(data[child], elem)) {  // element < child
      data[pos] = std::move(data[child]);
      pos = child;
    }
    data[pos] = std::move(elem);
  }
  void sift_down_to_bottom(size_t pos) {
    size_t end = data.size();
    size_t start = pos;
    T elem = std::move(data[pos]);
    size_t child = 2 * pos + 1;
    size_t lim = end >= 2 ? end - 2 : 0;
    while (child <= lim && end >= 2) {
      if (le(data[child], data[child + 1])) child += 1;
      data[pos] = std::move(data[child]);
      pos = child;
      child = 2 * pos + 1;
    }
    if (child + 1 == end) {
      data[pos] = std::move(data[child]);
      pos = child;
    }
    data[pos] = std::move(elem);
    sift_up(start, pos);
  }
  void push(T item) {
    size_t old_len = data.size();
    data.push_back(std::move(item));
    sift_up(0, old_len);
  }
  T pop() {
    T item = std::move(data.back());
    data.pop_back();
    if (!data.empty()) {
      std::swap(item, data[0]);
      sift_down_to_bottom(0);
    }
    return item;
  }
  void rebuild() {  // From<Vec<T>>
    size_t n = data.size() / 2;
    while (n > 0) {
      n -= 1;
      sift_down_range(n, data.size());
    }
  }
};

/* ---------------------------------------------------------------- varint / postings codec */

// util/varint.rs:5-15
inline void write_u32_var(uint32_t v32, std::vector<uint8_t> &out) {
  uint64_t v = v32;
  while (v >= 0x80) {
    out.push_back((uint8_t)((v & 0x7F) | 0x80));
    v >>= 7;
  }
  out.push_back((uint8_t)v);
}

// util/varint.rs:31-49 (byte at a time; error when shift passes 28)
inline size_t read_u32_var(const uint8_t *buf, size_t len, uint32_t *out) {
  uint32_t shift = 0, value = 0;
  size_t i = 0;
  for (;;) {
    if (i >= len) return 0;
    uint8_t b = buf[i++];
    value |= (uint32_t)(b & 0x7F) << shift;
    if ((b & 0x80) == 0) {
      *out = value;
      return i;
    }
    shift += 7;
    if (shift > 28) return 0;
  }
}

inline void put_u32(std::vector<uint8_t> &o, uint32_t v) {
  for (int i = 0; i < 4; i++) o.push_back((uint8_t)(v >> (8 * i)));
}
inline void put_f32(std::vector<uint8_t> &o, float f) {
  uint32_t v;
  std::memcpy(&v, &f, 4);
  put_u32(o, v);
}
inline uint32_t get_u32(const uint8_t *p) {
  return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24);
}
inline float get_f32(const uint8_t *p) {
  uint32_t v = get_u32(p);
  float f;
  std::memcpy(&f, &v, 4);
  return f;
}

// index/postings.rs:78-129 PostingsWriter::write_term
void encode_postings(const uint32_t *docs, const uint32_t *tfs, size_t n, bool keep_positions,
                     const uint32_t *pos_offsets, const uint32_t *positions, std::vector<uint8_t> &out) {
  put_u32(out, (uint32_t)n);
  out.push_back(keep_positions ? 1 : 0);
  uint32_t block_size = DEFAULT_BLOCK_SIZE;
  uint32_t block_count = (uint32_t)((n + block_size - 1) / block_size);
  put_u32(out, block_count > 0 ? (block_count | BLOCK_META_FLAG) : 0u);
  uint32_t max_doc_id = n ? docs[n - 1] : 0;
  float max_tf = 0.0f;
  for (size_t i = 0; i < n; i++) max_tf = std::max(max_tf, (float)tfs[i]);
  put_u32(out, max_doc_id);
  put_f32(out, max_tf);
  if (block_count > 0) {
    put_u32(out, block_size);
    for (size_t s = 0; s < n; s += block_size) put_u32(out, docs[std::min(n, s + block_size) - 1]);
    for (size_t s = 0; s < n; s += block_size) {
      float m = 0.0f;
      for (size_t i = s; i < std::min(n, s + block_size); i++) m = std::max(m, (float)tfs[i]);
      put_f32(out, m);
    }
  }
  for (size_t i = 0; i < n; i++) {
    write_u32_var(docs[i], out);  // absolute doc id, postings.rs:115
    write_u32_var(tfs[i], out);
    if (keep_positions) {
      uint32_t b = pos_offsets ? pos_offsets[i] : 0, e = pos_offsets ? pos_offsets[i + 1] : 0;
      write_u32_var(e - b, out);
      uint32_t prev = 0;
      for (uint32_t j = b; j < e; j++) {
        write_u32_var(positions[j] - prev, out);
        prev = positions[j];
      }
    }
  }
}

struct DecodedPostings {  // index/postings.rs:133-139 PostingsReader
  std::vector<uint32_t> docs, tfs;
  float max_tf = 0.0f;
  std::vector<uint32_t> block_max_doc_ids;
  std::vector<float> block_max_tfs;
  uint32_t block_size = DEFAULT_BLOCK_SIZE;
};

// index/postings.rs:142-212 PostingsReader::read_at.  Returns bytes consumed, 0 on error.
size_t decode_postings(const uint8_t *buf, size_t len, bool keep_positions, DecodedPostings &r) {
  size_t p = 0;
  if (len < 17) return 0;
  uint32_t doc_freq = get_u32(buf + p);
  p += 4;
  bool has_positions = buf[p] == 1 && keep_positions;
  bool stored_positions = buf[p] == 1;
  p += 1;
  uint32_t raw_block = get_u32(buf + p);
  p += 4;
  bool has_block_meta = (raw_block & BLOCK_META_FLAG) != 0;
  uint32_t block_count = raw_block & ~BLOCK_META_FLAG;
  uint32_t max_doc_id = get_u32(buf + p);
  p += 4;
  float max_tf = get_f32(buf + p);
  p += 4;
  r.block_size = DEFAULT_BLOCK_SIZE;
  r.block_max_doc_ids.clear();
  r.block_max_tfs.clear();
  if (has_block_meta && block_count > 0) {
    if (p + 4 + 8ull * block_count > len) return 0;
    r.block_size = get_u32(buf + p);
    p += 4;
    for (uint32_t i = 0; i < block_count; i++, p += 4) r.block_max_doc_ids.push_back(get_u32(buf + p));
    for (uint32_t i = 0; i < block_count; i++, p += 4) r.block_max_tfs.push_back(get_f32(buf + p));
  }
  r.docs.clear();
  r.tfs.clear();
  r.docs.reserve(doc_freq);
  r.tfs.reserve(doc_freq);
  for (uint32_t i = 0; i < doc_freq; i++) {
    uint32_t d, t;
    size_t c = read_u32_var(buf + p, len - p, &d);
    if (!c) return 0;
    p += c;
    c = read_u32_var(buf + p, len - p, &t);
    if (!c) return 0;
    p += c;
    // NB: the reference only consumes position bytes when has_positions (flag && keep_positions);
    // a reader opened with keep_positions=false on a positional file would mis-parse.  The
    // oracle skips them whenever they are stored, which is what a consistent open does.
    if (stored_positions) {
      uint32_t cnt;
      c = read_u32_var(buf + p, len - p, &cnt);
      if (!c) return 0;
      p += c;
      for (uint32_t j = 0; j < cnt; j++) {
        uint32_t dlt;
        c = read_u32_var(buf + p, len - p, &dlt);
        if (!c) return 0;
        p += c;
      }
    }
    (void)has_positions;
    r.docs.push_back(d);
    r.tfs.push_back(t);
  }
  if (r.block_max_doc_ids.empty()) {
    r.block_size = DEFAULT_BLOCK_SIZE;
    for (size_t s = 0; s < r.docs.size(); s += r.block_size) {
      size_t e = std::min(r.docs.size(), s + (size_t)r.block_size);
      float m = 0.0f;
      for (size_t i = s; i < e; i++) m = std::max(m, (float)r.tfs[i]);
      r.block_max_doc_ids.push_back(e > s ? r.docs[e - 1] : max_doc_id);
      r.block_max_tfs.push_back(m);
    }
  }
  float computed_max = 0.0f;
  for (float v : r.block_max_tfs) computed_max = std::max(computed_max, v);
  if (computed_max > max_tf) max_tf = computed_max;
  r.max_tf = max_tf;
  return p;
}

/* ---------------------------------------------------------------- index */

struct Column {
  int kind;  // 0 i64, 1 f64, 2 str; 3 i64 list, 4 f64 list, 5 str list (index/fastfields.rs: Column::I64List / F64List / StrList)
  std::vector<int64_t> i64;
  std::vector<double> f64;
  std::vector<uint8_t> present;
  std::vector<std::string> dict;
  std::vector<uint32_t> ords;
  std::vector<uint32_t> offsets;  // list kinds: doc_count + 1 running sums
  // doc_range, index/fastfields.rs:1136-1143
  bool doc_range(uint32_t doc, size_t &start, size_t &end) const {
    if (offsets.size() < (size_t)doc + 2) return false;
    start = offsets[doc];
    end = offsets[doc + 1];
    return true;
  }
};

}  // namespace

struct slo_index {
  uint32_t segment_ord = 0, doc_count = 0;
  float k1 = 0.9f, b = 0.4f;
  uint64_t n_terms = 0;
  const uint64_t *offsets = nullptr;
  const uint32_t *docs = nullptr;
  const uint32_t *tfs = nullptr;
  const uint64_t *pos_offsets = nullptr;  // BORROWED, one entry per posting + 1 (positions on)
  const uint32_t *positions = nullptr;
  bool has_lens = false;
  std::vector<float> lens;  // api/reader.rs:3604-3621
  float avgdl = 0.0f;       // index/segment.rs:946-957
  float min_doc_len = 1.0f; // query/wand.rs:110-121 (depends only on the segment)
  std::vector<uint8_t> deleted;
  uint32_t n_deleted = 0;
  std::vector<uint8_t> post_image;
  std::vector<uint64_t> post_off;
  std::vector<Column> columns;
};

namespace {

// View of one term's postings as the reference's PostingsReader would hold them.
struct PostingsView {
  const uint32_t *docs = nullptr;
  const uint32_t *tfs = nullptr;
  size_t len = 0;
  float max_tf = 0.0f;
  DecodedPostings owned;  // used in faithful_decode mode
  bool stored_blocks = false;
};

// query/wand.rs:65-85 ScoredTerm (+ group / flags for the matcher)
struct ScoredTerm {
  const PostingsView *postings;
  float weight, avgdl, docs, k1, b;
  uint32_t leaf;
  const std::vector<float> *doc_lengths;  // Option<Arc<Vec<f32>>>
  // query/wand.rs:77-84
  float doc_len(uint32_t doc_id) const {
    if (doc_lengths && doc_id < doc_lengths->size()) {
      float v = (*doc_lengths)[doc_id];
      if (v > 0.0f) return v;
    }
    return std::max(avgdl, 1.0f);
  }
};

// query/wand.rs:88-266 TermState
struct TermState {
  const uint32_t *pdocs;
  const uint32_t *ptfs;
  size_t plen;
  size_t idx = 0;
  float weight, df, avgdl, docs, k1, b;
  uint32_t leaf;
  float ub, min_doc_len;
  const std::vector<float> *doc_lengths;
  std::vector<uint32_t> block_max_doc_ids;
  std::vector<float> block_max_tfs;
  size_t block_size;

  // query/wand.rs:107-153 (+ build_block_meta :305-330)
  TermState(const ScoredTerm &term, size_t bsize, float seg_min_doc_len, bool scan_lengths) {
    pdocs = term.postings->docs;
    ptfs = term.postings->tfs;
    plen = term.postings->len;
    weight = term.weight;
    df = (float)plen;
    avgdl = term.avgdl;
    docs = term.docs;
    k1 = term.k1;
    b = term.b;
    leaf = term.leaf;
    doc_lengths = term.doc_lengths;
    block_size = std::max<size_t>(bsize, 1);
    if (term.postings->stored_blocks && block_size == term.postings->owned.block_size &&
        !term.postings->owned.block_max_doc_ids.empty()) {
      block_max_doc_ids = term.postings->owned.block_max_doc_ids;
      block_max_tfs = term.postings->owned.block_max_tfs;
    } else {
      for (size_t i = 0; i < plen; i += block_size) {
        size_t e = std::min(plen, i + block_size);
        float m = 0.0f;
        block_max_doc_ids.push_back(pdocs[e - 1]);
        for (size_t j = i; j < e; j++) m = std::max(m, (float)ptfs[j]);
        block_max_tfs.push_back(m);
      }
    }
    if (doc_lengths) {
      float mn;
      if (scan_lengths) {  // the reference rescans the whole vector per term per query, :111-116
        mn = std::numeric_limits<float>::infinity();
        for (float l : *doc_lengths)
          if (l > 0.0f) mn = std::min(mn, l);
        if (!std::isfinite(mn)) mn = std::max(avgdl, 1.0f);
      } else {
        mn = seg_min_doc_len;  // same value, hoisted
      }
      min_doc_len = mn;
    } else {
      min_doc_len = std::max(avgdl, 1.0f);
    }
    ub = upper_bound_tf(term.postings->max_tf, df, min_doc_len, avgdl, docs, k1, b, weight);
  }
  bool is_done() const { return idx >= plen; }
  uint32_t doc_id() const { return idx < plen ? pdocs[idx] : DOCID_END; }
  float doc_len(uint32_t d) const {
    if (doc_lengths && d < doc_lengths->size()) {
      float v = (*doc_lengths)[d];
      if (v > 0.0f) return v;
    }
    return std::max(avgdl, 1.0f);
  }
  float tf() const { return idx < plen ? (float)ptfs[idx] : 0.0f; }
  float score_current() const {
    return score_tf(tf(), df, doc_len(doc_id()), avgdl, docs, k1, b, weight);
  }
  size_t advance() {
    if (is_done()) return 0;
    idx += 1;
    return 1;
  }
  // query/wand.rs:205-232 galloping + partition_point
  size_t advance_to(uint32_t target) {
    if (is_done() || doc_id() >= target) return 0;
    size_t len = plen;
    size_t low = idx + 1;
    if (low >= len) {
      size_t delta = len - idx;
      idx = len;
      return delta;
    }
    size_t step = 1;
    while (low + step < len) {
      if (pdocs[low + step] >= target) break;
      step <<= 1;
    }
    size_t upper = std::min(low + step, len);
    size_t adv = std::lower_bound(pdocs + low, pdocs + upper, target) - (pdocs + low);
    size_t new_idx = std::min(low + adv, len);
    size_t delta = new_idx - idx;
    idx = new_idx;
    return delta;
  }
  // query/wand.rs:238-251 — bound of the block THE CURSOR IS IN
  float block_upper_bound() const {
    size_t bi = idx / block_size;
    float tfm = bi < block_max_tfs.size() ? block_max_tfs[bi] : 0.0f;
    return score_tf(tfm, df, min_doc_len, avgdl, docs, k1, b, weight);
  }
  float upper_bound() const { return ub; }
  // query/wand.rs:257-265
  size_t skip_to_block(uint32_t target) {
    size_t prev = idx;
    size_t bi = std::lower_bound(block_max_doc_ids.begin(), block_max_doc_ids.end(), target) -
                block_max_doc_ids.begin();
    size_t start = bi * block_size;
    if (start > idx) idx = std::min(start, plen);
    return idx - prev;
  }
};

struct RankedLe {  // heap of Reverse<RankedDoc>: Reverse(a) <= Reverse(b)  <=>  a >= b
  bool operator()(const RankedDoc &a, const RankedDoc &b) const { return ranked_cmp(a, b) >= 0; }
};
using TopHeap = RustBinaryHeap<RankedDoc, RankedLe>;

// query/wand.rs:905-916
void push_top_k(TopHeap &heap, RankedDoc doc, size_t k) {
  if (heap.size() < k) {
    heap.push(doc);
    return;
  }
  if (!heap.empty()) {
    if (ranked_cmp(doc, heap.peek()) > 0) {
      heap.pop();
      heap.push(doc);
    }
  }
}

// query/wand.rs:918-926
std::vector<RankedDoc> finalize_heap(TopHeap &heap) {
  std::vector<RankedDoc> out = heap.data;
  std::stable_sort(out.begin(), out.end(), [](const RankedDoc &a, const RankedDoc &b) {
    int c = total_cmp(b.score, a.score);
    if (c != 0) return c < 0;
    return a.doc_id < b.doc_id;
  });
  return out;
}

/* ---------------------------------------------------------------- filters & matcher */

// Unicode lowercase of one code point as Rust's char::to_lowercase gives it for the bicameral blocks a keyword field is
// likely to hold: Latin-1, Latin Extended-A / -B (the regular pairs) / Additional, Greek (+ tonos forms), Cyrillic (+ the
// extended pairs), Armenian, fullwidth Latin.  U+0130 (İ) lowers to two code points ("i̇"), the only multi-character mapping.
static void lower_cp(uint32_t c, std::u32string &out, bool final_sigma) {
  if (c < 0x80) {
    out.push_back(c >= 'A' && c <= 'Z' ? c + 32 : c);
    return;
  }
  if (c == 0x130) {
    out.push_back('i');
    out.push_back(0x307);
    return;
  }
  uint32_t r = c;
  if ((c >= 0xC0 && c <= 0xDE && c != 0xD7)) r = c + 0x20;
  else if ((c >= 0x100 && c <= 0x12F) || (c >= 0x132 && c <= 0x137) || (c >= 0x14A && c <= 0x177)) r = (c & 1u) ? c : c + 1;
  else if ((c >= 0x139 && c <= 0x148) || (c >= 0x179 && c <= 0x17E)) r = (c & 1u) ? c + 1 : c;
  else if (c == 0x178) r = 0xFF;
  else if (c == 0x1C4 || c == 0x1C5) r = 0x1C6;  // the DŽ / LJ / NJ / DZ digraphs: upper and title case lower to one letter
  else if (c == 0x1C7 || c == 0x1C8) r = 0x1C9;
  else if (c == 0x1CA || c == 0x1CB) r = 0x1CC;
  else if (c == 0x1F1 || c == 0x1F2) r = 0x1F3;
  else if ((c >= 0x1CD && c <= 0x1DC)) r = (c & 1u) ? c + 1 : c;
  else if ((c >= 0x1DE && c <= 0x1EF) || (c >= 0x1F8 && c <= 0x21F) || (c >= 0x222 && c <= 0x233)) r = (c & 1u) ? c : c + 1;
  else if (c == 0x386) r = 0x3AC;
  else if (c >= 0x388 && c <= 0x38A) r = c + 37;
  else if (c == 0x38C) r = 0x3CC;
  else if (c == 0x38E || c == 0x38F) r = c + 63;
  else if (c >= 0x391 && c <= 0x3AB && c != 0x3A2) r = (c == 0x3A3 && final_sigma) ? 0x3C2 : c + 0x20;
  else if (c >= 0x400 && c <= 0x40F) r = c + 0x50;
  else if (c >= 0x410 && c <= 0x42F) r = c + 0x20;
  else if ((c >= 0x460 && c <= 0x481) || (c >= 0x48A && c <= 0x4BF) || (c >= 0x4D0 && c <= 0x52F)) r = (c & 1u) ? c : c + 1;
  else if (c >= 0x4C1 && c <= 0x4CE) r = (c & 1u) ? c + 1 : c;
  else if (c == 0x4C0) r = 0x4CF;
  else if (c >= 0x531 && c <= 0x556) r = c + 0x30;
  else if (c >= 0x1E00 && c <= 0x1E95) r = (c & 1u) ? c : c + 1;
  else if (c >= 0x1EA0 && c <= 0x1EFF) r = (c & 1u) ? c : c + 1;
  else if (c >= 0xFF21 && c <= 0xFF3A) r = c + 0x20;
  out.push_back(r);
}
static bool is_cased_letter(uint32_t c) {  // enough of Unicode's "cased" for the final-sigma rule
  return (c >= 'A' && c <= 'Z') || (c >= 'a' && c <= 'z') || (c >= 0xC0 && c <= 0x24F && c != 0xD7 && c != 0xF7) ||
         (c >= 0x370 && c <= 0x3FF) || (c >= 0x400 && c <= 0x52F) || (c >= 0x531 && c <= 0x586) || (c >= 0x1E00 && c <= 0x1FFF);
}
static std::u32string utf8_lower(const std::string &s) {
  std::u32string cps;
  for (size_t i = 0; i < s.size();) {
    const unsigned char b = (unsigned char)s[i];
    uint32_t c = b;
    int n = 1;
    if (b >= 0xF0 && i + 3 < s.size()) {
      c = ((b & 7u) << 18) | (((unsigned char)s[i + 1] & 63u) << 12) | (((unsigned char)s[i + 2] & 63u) << 6) | ((unsigned char)s[i + 3] & 63u);
      n = 4;
    } else if (b >= 0xE0 && i + 2 < s.size()) {
      c = ((b & 15u) << 12) | (((unsigned char)s[i + 1] & 63u) << 6) | ((unsigned char)s[i + 2] & 63u);
      n = 3;
    } else if (b >= 0xC0 && i + 1 < s.size()) {
      c = ((b & 31u) << 6) | ((unsigned char)s[i + 1] & 63u);
      n = 2;
    }
    cps.push_back(c);
    i += n;
  }
  std::u32string out;
  for (size_t i = 0; i < cps.size(); i++) {
    // Final_Sigma: preceded by a cased letter and not followed by one (Unicode SpecialCasing, as Rust implements it)
    const bool fin = cps[i] == 0x3A3 && i > 0 && is_cased_letter(cps[i - 1]) && !(i + 1 < cps.size() && is_cased_letter(cps[i + 1]));
    lower_cp(cps[i], out, fin);
  }
  return out;
}

// index/fastfields.rs:475-481: eq_ignore_ascii_case when both sides are ASCII, else a.to_lowercase() == b.to_lowercase()
// (to_lowercase restated for the bicameral blocks above; tests/test_oracle_golden.py checks it against Python's str.lower())
bool case_insensitive_equals(const std::string &a, const std::string &b) {
  bool ascii = true;
  for (unsigned char ch : a) ascii = ascii && ch < 0x80;
  for (unsigned char ch : b) ascii = ascii && ch < 0x80;
  if (!ascii) return utf8_lower(a) == utf8_lower(b);
  if (a.size() != b.size()) return false;
  for (size_t i = 0; i < a.size(); i++)
    if (std::tolower((unsigned char)a[i]) != std::tolower((unsigned char)b[i])) return false;
  return true;
}

// query/filters.rs:84-149 over flat columns; index/fastfields.rs:490-657.  `pos` walks the prefix tree.
bool filter_eval(const slo_index *ix, uint32_t doc, const slo_filter_node_t *nodes, uint32_t n, uint32_t &pos,
                 const char *const *strings) {
  if (pos >= n) return false;
  const slo_filter_node_t &nd = nodes[pos++];
  const Column *col = (nd.column >= 0 && (size_t)nd.column < ix->columns.size()) ? &ix->columns[nd.column] : nullptr;
  switch (nd.op) {
    case SLO_F_KEYWORD_EQ:
    case SLO_F_KEYWORD_IN: {
      if (col && col->kind == 5) {  // matches_keyword / matches_keyword_in over Column::StrList: any value, fastfields.rs:497-509, 548-562
        size_t s0, e0;
        if (!col->doc_range(doc, s0, e0)) return false;
        for (size_t i = s0; i < e0 && i < col->ords.size(); i++) {
          uint32_t o = col->ords[i];
          if (o >= col->dict.size()) continue;
          for (uint32_t v = nd.value_begin; v < nd.value_end; v++)
            if (case_insensitive_equals(col->dict[o], strings[v])) return true;
        }
        return false;
      }
      if (!col || col->kind != 2 || doc >= col->ords.size()) return false;
      uint32_t o = col->ords[doc];
      if (o == 0xFFFFFFFFu || o >= col->dict.size()) return false;
      for (uint32_t v = nd.value_begin; v < nd.value_end; v++)
        if (case_insensitive_equals(col->dict[o], strings[v])) return true;
      return false;
    }
    case SLO_F_I64_RANGE: {
      if (col && col->kind == 3) {  // Column::I64List, fastfields.rs:602-609
        size_t s0, e0;
        if (!col->doc_range(doc, s0, e0)) return false;
        for (size_t i = s0; i < e0 && i < col->i64.size(); i++)
          if (col->i64[i] >= nd.i_min && col->i64[i] <= nd.i_max) return true;
        return false;
      }
      if (!col || col->kind != 0 || doc >= col->i64.size() || !col->present[doc]) return false;
      int64_t v = col->i64[doc];
      return v >= nd.i_min && v <= nd.i_max;
    }
    case SLO_F_F64_RANGE: {
      if (col && col->kind == 4) {  // Column::F64List, fastfields.rs:632-639
        size_t s0, e0;
        if (!col->doc_range(doc, s0, e0)) return false;
        for (size_t i = s0; i < e0 && i < col->f64.size(); i++)
          if (col->f64[i] >= nd.f_min && col->f64[i] <= nd.f_max) return true;
        return false;
      }
      if (!col || col->kind != 1 || doc >= col->f64.size() || !col->present[doc]) return false;
      double v = col->f64[doc];
      return v >= nd.f_min && v <= nd.f_max;
    }
    case SLO_F_AND: {
      bool ok = true;
      for (uint32_t c = 0; c < nd.n_children; c++) ok = filter_eval(ix, doc, nodes, n, pos, strings) && ok;
      return ok;
    }
    case SLO_F_OR: {
      bool any = false;
      for (uint32_t c = 0; c < nd.n_children; c++) any = filter_eval(ix, doc, nodes, n, pos, strings) || any;
      return any;
    }
    case SLO_F_NOT: {
      bool v = filter_eval(ix, doc, nodes, n, pos, strings);
      return !v;
    }
  }
  return false;
}

struct Matcher {  // flat Bool / QueryString matcher, api/reader.rs:1485-1582
  const slo_query_t *q;
  std::vector<std::vector<const PostingsView *>> group_lists;  // build_term_doc_lists api/reader.rs:1722
  bool group_matches(uint32_t g, uint32_t doc) const {
    for (const PostingsView *pv : group_lists[g])
      if (std::binary_search(pv->docs, pv->docs + pv->len, doc)) return true;
    return false;
  }
  bool matches(uint32_t doc) const {
    if (q->n_groups == 0) return true;  // plain OR of the scored terms: every scored doc is in some list
    uint32_t should_total = 0, should_hit = 0;
    for (uint32_t g = 0; g < q->n_groups; g++) {
      uint8_t role = q->group_role[g];
      if (role == SLO_ROLE_MUST) {
        if (!group_matches(g, doc)) return false;
      } else if (role == SLO_ROLE_MUST_NOT) {
        if (group_matches(g, doc)) return false;
      }
    }
    for (uint32_t g = 0; g < q->n_groups; g++)
      if (q->group_role[g] == SLO_ROLE_SHOULD) {
        should_total++;
        if (group_matches(g, doc)) should_hit++;
      }
    (void)should_total;
    return should_hit >= q->min_should;
  }
};

struct SearchCtx {
  const slo_index *ix;
  const slo_query_t *q;
  const Matcher *matcher;
  const slo_filter_node_t *filter;
  uint32_t n_filter;
  const char *const *strings;
  slo_stats_t *stats;
  // accept closure api/reader.rs:3009-3036 (the collector branch is not on this path)
  bool accept(uint32_t doc, float score) const {
    if (doc < ix->deleted.size() && ix->deleted[doc]) return false;
    if (!matcher->matches(doc)) return false;
    if (filter && n_filter) {
      uint32_t pos = 0;
      if (!filter_eval(ix, doc, filter, n_filter, pos, strings)) return false;
    }
    if (q->has_cursor) {
      // key.cmp(cur) for the `_score` desc plan, query/sort.rs:80-93: score desc (total_cmp), segment_ord asc, doc_id asc
      int ord = total_cmp(q->cursor_score, score);
      if (ord == 0) ord = ix->segment_ord < q->cursor_segment_ord ? -1 : (ix->segment_ord > q->cursor_segment_ord ? 1 : 0);
      if (ord == 0) ord = doc < q->cursor_doc_id ? -1 : (doc > q->cursor_doc_id ? 1 : 0);
      if (ord <= 0) {  // at or before the cursor: rejected (:3021-3027)
        if (ord == 0 && stats) stats->saw_cursor = 1;
        return false;
      }
    }
    if (stats) stats->total_matches += 1;
    return true;
  }
};

/* ---------------------------------------------------------------- execution */

// ScorePlan::evaluate, query/planner.rs:133-164, on the postfix form of the ScoreExpr tree.  Without
// plan nodes the plan is Sum of all leaves (QueryString, query/planner.rs:354-360).
float evaluate_plan(const slo_query_t *q, const std::vector<float> &leaves) {
  if (!q->n_plan_nodes) {
    float score = 0.0f;
    for (float l : leaves) score += l;  // ScoreExpr::Sum, query/planner.rs:135
    return score;
  }
  std::vector<float> st;
  for (uint32_t i = 0; i < q->n_plan_nodes; i++) {
    const slo_plan_node_t &n = q->plan[i];
    if (n.op == SLO_PLAN_LEAF) {
      st.push_back(n.arg < leaves.size() ? leaves[n.arg] : 0.0f);  // leaves.get(idx).unwrap_or(0.0), :134
      continue;
    }
    size_t c = std::min<size_t>(n.arg, st.size());
    size_t base = st.size() - c;
    float r;
    if (n.op == SLO_PLAN_SUM) {
      r = 0.0f;
      for (size_t j = base; j < st.size(); j++) r += st[j];
    } else {  // DisMax, :136-151
      if (c == 0) {
        r = 0.0f;
      } else {
        float mx = -INFINITY, sum = 0.0f;
        for (size_t j = base; j < st.size(); j++) {
          mx = std::fmax(mx, st[j]);
          sum += st[j];
        }
        r = mx + n.tie_breaker * (sum - mx);
      }
    }
    st.resize(base);
    st.push_back(r);
  }
  return st.empty() ? 0.0f : st.back();
}

// query/wand.rs:459-566 brute_force (hash-map accumulation; leaf buffers when a ScorePlan is given)
std::vector<RankedDoc> brute_force(const std::vector<ScoredTerm> &terms, size_t k, uint32_t leaf_count,
                                   const SearchCtx &cx) {
  TopHeap heap{RankedLe{}};
  if (leaf_count > 0) {
    std::unordered_map<uint32_t, std::vector<float>> scores;
    for (const ScoredTerm &term : terms) {
      float df = (float)term.postings->len;
      if (cx.stats) cx.stats->postings_advanced += term.postings->len;
      for (size_t i = 0; i < term.postings->len; i++) {
        uint32_t d = term.postings->docs[i];
        float s = score_tf((float)term.postings->tfs[i], df, term.doc_len(d), term.avgdl, term.docs, term.k1,
                           term.b, term.weight);
        auto it = scores.find(d);
        if (it == scores.end()) it = scores.emplace(d, std::vector<float>(leaf_count, 0.0f)).first;
        it->second[term.leaf] += s;
      }
    }
    if (cx.stats) {
      cx.stats->scored_docs += scores.size();
      cx.stats->candidates_examined += scores.size();
    }
    for (auto &kv : scores) {
      float score = evaluate_plan(cx.q, kv.second);  // plan.evaluate(&leaves), query/wand.rs:506
      if (!cx.accept(kv.first, score)) continue;
      push_top_k(heap, RankedDoc{kv.first, score}, k);
    }
    return finalize_heap(heap);
  }
  std::unordered_map<uint32_t, float> scores;
  for (const ScoredTerm &term : terms) {
    float df = (float)term.postings->len;
    if (cx.stats) cx.stats->postings_advanced += term.postings->len;
    for (size_t i = 0; i < term.postings->len; i++) {
      uint32_t d = term.postings->docs[i];
      float s = score_tf((float)term.postings->tfs[i], df, term.doc_len(d), term.avgdl, term.docs, term.k1, term.b,
                         term.weight);
      scores[d] += s;
    }
  }
  if (cx.stats) {
    cx.stats->scored_docs += scores.size();
    cx.stats->candidates_examined += scores.size();
  }
  for (auto &kv : scores) {
    if (!cx.accept(kv.first, kv.second)) continue;
    push_top_k(heap, RankedDoc{kv.first, kv.second}, k);
  }
  return finalize_heap(heap);
}

// Same result as brute_force with score_plan None, but a dense accumulator + touched flags: the
// "fair" CPU port.  Float adds happen in the same (term) order as above.
std::vector<RankedDoc> brute_force_dense(const std::vector<ScoredTerm> &terms, size_t k, const SearchCtx &cx,
                                         std::vector<float> &acc, std::vector<uint8_t> &touched) {
  TopHeap heap{RankedLe{}};
  uint32_t n = cx.ix->doc_count;
  if (acc.size() < n) acc.assign(n, 0.0f);
  if (touched.size() < n) touched.assign(n, 0);
  uint64_t scored = 0;
  for (const ScoredTerm &term : terms) {
    float df = (float)term.postings->len;
    if (cx.stats) cx.stats->postings_advanced += term.postings->len;
    for (size_t i = 0; i < term.postings->len; i++) {
      uint32_t d = term.postings->docs[i];
      float s = score_tf((float)term.postings->tfs[i], df, term.doc_len(d), term.avgdl, term.docs, term.k1, term.b,
                         term.weight);
      if (!touched[d]) {
        touched[d] = 1;
        acc[d] = 0.0f;
        scored++;
      }
      acc[d] += s;
    }
  }
  if (cx.stats) {
    cx.stats->scored_docs += scored;
    cx.stats->candidates_examined += scored;
  }
  for (const ScoredTerm &term : terms)
    for (size_t i = 0; i < term.postings->len; i++) {
      uint32_t d = term.postings->docs[i];
      if (!touched[d]) continue;
      touched[d] = 0;
      float s = acc[d];
      if (!cx.accept(d, s)) continue;
      push_top_k(heap, RankedDoc{d, s}, k);
    }
  return finalize_heap(heap);
}

struct TermLe {  // TermWrapper::cmp query/wand.rs:693-698: other.doc_id().cmp(self.doc_id()); a <= b
  bool operator()(const TermState *a, const TermState *b) const { return b->doc_id() <= a->doc_id(); }
};

// query/wand.rs:659-903 wand_loop.  use_block_bounds=true is the reference's `bmw`, restated
// faithfully INCLUDING the terminate-on-no-pivot at :770-778 that makes it inexact.
std::vector<RankedDoc> wand_loop(std::vector<TermState> &states, size_t k, bool use_block_bounds,
                                 uint32_t leaf_count, const SearchCtx &cx) {
  TopHeap heap{RankedLe{}};
  RustBinaryHeap<TermState *, TermLe> queue{TermLe{}};
  for (TermState &t : states)
    if (!t.is_done()) queue.data.push_back(&t);
  queue.rebuild();
  std::vector<float> leaf_scores(leaf_count, 0.0f);
  std::vector<TermState *> pending;
  bool rank_hits = k > 0;
  for (;;) {
    if (queue.empty()) break;
    if (queue.peek()->is_done()) {
      queue.pop();
      continue;
    }
    float heap_threshold = (rank_hits && heap.size() >= k) ? heap.peek().score : 0.0f;
    float pivot_threshold = heap_threshold;  // no collector on this path (:725-729)
    bool have_pivot = false;
    size_t p_idx = 0;
    float acc = 0.0f;
    while (!queue.empty()) {
      TermState *t = queue.pop();
      float bound = use_block_bounds ? t->block_upper_bound() : t->upper_bound();
      pending.push_back(t);
      if (!std::isfinite(bound)) continue;
      acc += bound;
      if (acc >= pivot_threshold) {
        have_pivot = true;
        p_idx = pending.size() - 1;
        break;
      }
    }
    if (!have_pivot) {
      for (TermState *t : pending)
        if (!t->is_done()) queue.push(t);
      pending.clear();
      break;
    }
    uint32_t pivot_doc = pending[p_idx]->doc_id();
    uint32_t smallest_doc = pending[0]->doc_id();
    if (pivot_doc == smallest_doc) {
      uint32_t doc_id = pivot_doc;
      while (!queue.empty() && queue.peek()->doc_id() == doc_id) pending.push_back(queue.pop());
      float score_sum = 0.0f;
      for (TermState *t : pending) {
        if (t->doc_id() != doc_id) continue;
        float c = t->score_current();
        score_sum += c;
        if (leaf_count) leaf_scores[t->leaf] += c;
        size_t moved = t->advance();
        if (cx.stats) cx.stats->postings_advanced += moved;
      }
      if (cx.stats) {
        cx.stats->candidates_examined += 1;
        cx.stats->scored_docs += 1;
      }
      float score = score_sum;
      if (leaf_count) {
        score = evaluate_plan(cx.q, leaf_scores);
        std::fill(leaf_scores.begin(), leaf_scores.end(), 0.0f);
      }
      if (cx.accept(doc_id, score)) {
        if (rank_hits && (heap.size() < k || score > heap_threshold)) push_top_k(heap, RankedDoc{doc_id, score}, k);
      }
    } else {
      for (size_t i = 0; i < p_idx; i++) {
        TermState *t = pending[i];
        if (use_block_bounds) {
          size_t moved = t->skip_to_block(pivot_doc);
          if (cx.stats) cx.stats->postings_advanced += moved;
        }
        size_t moved = t->advance_to(pivot_doc);
        if (cx.stats) cx.stats->postings_advanced += moved;
      }
    }
    for (TermState *t : pending)
      if (!t->is_done()) queue.push(t);
    pending.clear();
  }
  return finalize_heap(heap);
}

struct Scratch {
  std::vector<float> acc;
  std::vector<uint8_t> touched;
};

int32_t search_one(const slo_index *ix, const slo_query_t *q, uint32_t k, int exec, uint32_t block_size,
                   const slo_filter_node_t *filter, uint32_t n_filter, const char *const *strings, int faithful,
                   slo_hit_t *out, slo_stats_t *stats, Scratch &scratch) {
  if (stats) std::memset(stats, 0, sizeof(*stats));
  if (!ix || !q || !ix->offsets) return -1;
  // per-query doc-length vector (api/reader.rs:3604-3621): the reference rebuilds it per query
  std::vector<float> faithful_lens;
  const std::vector<float> *lens = ix->has_lens ? &ix->lens : nullptr;
  if (faithful && ix->has_lens) {
    faithful_lens.reserve(ix->doc_count);
    for (uint32_t d = 0; d < ix->doc_count; d++) faithful_lens.push_back(ix->lens[d]);
    lens = &faithful_lens;
  }
  float docs = (float)(ix->doc_count - ix->n_deleted);  // live_docs, index/segment.rs:1365-1370
  std::vector<PostingsView> views(q->n_terms);
  std::vector<char> present(q->n_terms, 0);
  for (uint32_t i = 0; i < q->n_terms; i++) {
    uint32_t t = q->terms[i].term_id;
    if (t == 0xFFFFFFFFu || t >= ix->n_terms) continue;  // seg.postings(key) == None
    PostingsView &pv = views[i];
    if (faithful) {
      if (ix->post_off.empty()) return -2;
      size_t off = ix->post_off[t], end = ix->post_off[t + 1];
      if (!decode_postings(ix->post_image.data() + off, end - off, false, pv.owned)) return -3;
      pv.docs = pv.owned.docs.data();
      pv.tfs = pv.owned.tfs.data();
      pv.len = pv.owned.docs.size();
      pv.max_tf = pv.owned.max_tf;
      pv.stored_blocks = true;
    } else {
      pv.docs = ix->docs + ix->offsets[t];
      pv.tfs = ix->tfs + ix->offsets[t];
      pv.len = ix->offsets[t + 1] - ix->offsets[t];
      float m = 0.0f;
      for (size_t j = 0; j < pv.len; j++) m = std::max(m, (float)pv.tfs[j]);
      pv.max_tf = m;
    }
    present[i] = 1;
  }
  Matcher matcher;
  matcher.q = q;
  matcher.group_lists.resize(q->n_groups);
  for (uint32_t i = 0; i < q->n_terms; i++)
    if (present[i] && q->terms[i].group < q->n_groups) matcher.group_lists[q->terms[i].group].push_back(&views[i]);
  std::vector<ScoredTerm> terms;
  for (uint32_t i = 0; i < q->n_terms; i++) {
    if (!present[i] || !(q->terms[i].flags & SLO_TERM_SCORED)) continue;
    ScoredTerm st;
    st.postings = &views[i];
    st.weight = q->terms[i].weight;
    st.avgdl = ix->avgdl;
    st.docs = docs;
    st.k1 = ix->k1;
    st.b = ix->b;
    st.leaf = q->terms[i].leaf;
    st.doc_lengths = lens;
    terms.push_back(std::move(st));
  }
  if (terms.empty() || k == 0) return 0;  // query/wand.rs:413-416, api/reader.rs:3003-3005
  SearchCtx cx{ix, q, &matcher, filter, n_filter, strings, stats};
  std::vector<RankedDoc> ranked;
  if (exec == SLO_EXEC_BM25) {
    ranked = brute_force(terms, k, q->leaf_count, cx);
  } else if (exec == SLO_EXEC_BM25_DENSE) {
    ranked = brute_force_dense(terms, k, cx, scratch.acc, scratch.touched);
  } else {
    size_t bsize = std::max<uint32_t>(block_size ? block_size : DEFAULT_BLOCK_SIZE, 1);
    std::vector<TermState> states;
    states.reserve(terms.size());
    for (const ScoredTerm &st : terms)
      if (st.postings->len > 0) states.emplace_back(st, bsize, ix->min_doc_len, faithful != 0);
    ranked = wand_loop(states, k, exec == SLO_EXEC_BMW, q->leaf_count, cx);
  }
  int32_t n = (int32_t)ranked.size();
  for (int32_t i = 0; i < n; i++) out[i] = slo_hit_t{ix->segment_ord, ranked[i].doc_id, ranked[i].score};
  return n;
}

}  // namespace

/* ================================================================ C ABI */

extern "C" {

float slo_bm25(float tf, float df, float doc_len, float avgdl, float docs, float k1, float b) {
  return bm25(tf, df, doc_len, avgdl, docs, k1, b);
}
float slo_score_tf(float tf, float df, float doc_len, float avgdl, float docs, float k1, float b, float weight) {
  return score_tf(tf, df, doc_len, avgdl, docs, k1, b, weight);
}
float slo_upper_bound_tf(float tf, float df, float doc_len, float avgdl, float docs, float k1, float b,
                         float weight) {
  return upper_bound_tf(tf, df, doc_len, avgdl, docs, k1, b, weight);
}
float slo_idf(float df, float docs) { return fmaxf(logf((docs - df + 0.5f) / (df + 0.5f)), 0.0f) + 1.0f; }

size_t slo_varint_write_u32(uint32_t v, uint8_t *out) {
  std::vector<uint8_t> tmp;
  write_u32_var(v, tmp);
  std::memcpy(out, tmp.data(), tmp.size());
  return tmp.size();
}
size_t slo_varint_read_u32(const uint8_t *buf, size_t len, uint32_t *out) { return read_u32_var(buf, len, out); }

size_t slo_postings_encode(const uint32_t *docs, const uint32_t *tfs, size_t n, int keep_positions,
                           const uint32_t *pos_offsets, const uint32_t *positions, uint8_t *out, size_t cap) {
  std::vector<uint8_t> tmp;
  encode_postings(docs, tfs, n, keep_positions != 0, pos_offsets, positions, tmp);
  if (out && cap >= tmp.size()) std::memcpy(out, tmp.data(), tmp.size());
  return tmp.size();
}
int slo_postings_peek_df(const uint8_t *buf, size_t len, uint32_t *df, uint32_t *block_count) {
  if (len < 9) return -1;
  *df = get_u32(buf);
  *block_count = get_u32(buf + 5) & ~BLOCK_META_FLAG;
  return 0;
}
int slo_postings_decode(const uint8_t *buf, size_t len, int keep_positions, uint32_t *docs, uint32_t *tfs,
                        float *max_tf, uint32_t *block_size, uint32_t *blk_max_doc, float *blk_max_tf,
                        uint32_t *n_blocks_out, size_t *consumed) {
  DecodedPostings r;
  size_t c = decode_postings(buf, len, keep_positions != 0, r);
  if (!c) return -1;
  std::memcpy(docs, r.docs.data(), r.docs.size() * 4);
  std::memcpy(tfs, r.tfs.data(), r.tfs.size() * 4);
  if (max_tf) *max_tf = r.max_tf;
  if (block_size) *block_size = r.block_size;
  if (blk_max_doc) std::memcpy(blk_max_doc, r.block_max_doc_ids.data(), r.block_max_doc_ids.size() * 4);
  if (blk_max_tf) std::memcpy(blk_max_tf, r.block_max_tfs.data(), r.block_max_tfs.size() * 4);
  if (n_blocks_out) *n_blocks_out = (uint32_t)r.block_max_doc_ids.size();
  if (consumed) *consumed = c;
  return 0;
}

slo_index_t *slo_index_new(uint32_t segment_ord, uint32_t doc_count, float k1, float b) {
  slo_index *ix = new slo_index();
  ix->segment_ord = segment_ord;
  ix->doc_count = doc_count;
  ix->k1 = k1;
  ix->b = b;
  ix->deleted.assign(doc_count, 0);
  return ix;
}
void slo_index_free(slo_index_t *ix) { delete ix; }

int slo_index_set_postings(slo_index_t *ix, uint64_t n_terms, const uint64_t *offsets, const uint32_t *docs,
                           const uint32_t *tfs) {
  ix->n_terms = n_terms;
  ix->offsets = offsets;
  ix->docs = docs;
  ix->tfs = tfs;
  return 0;
}

int slo_index_set_field_lengths(slo_index_t *ix, const int64_t *lens, const uint8_t *present,
                                uint64_t total_tokens) {
  ix->lens.resize(ix->doc_count);
  for (uint32_t d = 0; d < ix->doc_count; d++) {
    // i64_value(..).unwrap_or(0) as f32, api/reader.rs:3614-3616
    int64_t v = (present && !present[d]) ? 0 : lens[d];
    ix->lens[d] = (float)v;
  }
  ix->has_lens = true;
  // compute_avg_lengths index/segment.rs:946-957: sum as f32 / total_docs as f32
  ix->avgdl = ix->doc_count == 0 ? 0.0f : (float)total_tokens / (float)(uint64_t)ix->doc_count;
  float mn = std::numeric_limits<float>::infinity();
  for (float l : ix->lens)
    if (l > 0.0f) mn = std::min(mn, l);
  ix->min_doc_len = std::isfinite(mn) ? mn : std::max(ix->avgdl, 1.0f);
  return 0;
}

int slo_index_set_deleted(slo_index_t *ix, const uint32_t *docs, uint32_t n) {
  std::fill(ix->deleted.begin(), ix->deleted.end(), 0);
  ix->n_deleted = 0;
  for (uint32_t i = 0; i < n; i++)
    if (docs[i] < ix->doc_count && !ix->deleted[docs[i]]) {
      ix->deleted[docs[i]] = 1;
      ix->n_deleted++;
    }
  return 0;
}

int slo_index_set_positions(slo_index_t *ix, const uint64_t *pos_offsets, const uint32_t *positions) {
  ix->pos_offsets = pos_offsets;
  ix->positions = positions;
  return 0;
}

int slo_index_build_post_image(slo_index_t *ix) {
  ix->post_image.clear();
  ix->post_off.assign(ix->n_terms + 1, 0);
  for (uint64_t t = 0; t < ix->n_terms; t++) {
    ix->post_off[t] = ix->post_image.size();
    uint64_t o = ix->offsets[t], e = ix->offsets[t + 1];
    if (ix->pos_offsets) {  // positions on, as every reference test / bench / FFI index is written
      std::vector<uint32_t> rel(e - o + 1);
      const uint64_t base = ix->pos_offsets[o];
      for (uint64_t i = o; i <= e; i++) rel[i - o] = (uint32_t)(ix->pos_offsets[i] - base);
      encode_postings(ix->docs + o, ix->tfs + o, e - o, true, rel.data(), ix->positions + base, ix->post_image);
    } else {
      encode_postings(ix->docs + o, ix->tfs + o, e - o, false, nullptr, nullptr, ix->post_image);
    }
  }
  ix->post_off[ix->n_terms] = ix->post_image.size();
  return 0;
}
uint64_t slo_index_post_image_size(const slo_index_t *ix) { return ix->post_image.size(); }
const uint8_t *slo_index_post_image(const slo_index_t *ix) { return ix->post_image.data(); }
const uint64_t *slo_index_post_offsets(const slo_index_t *ix) { return ix->post_off.data(); }

float slo_index_avgdl(const slo_index_t *ix) { return ix->avgdl; }
float slo_index_live_docs(const slo_index_t *ix) { return (float)(ix->doc_count - ix->n_deleted); }
float slo_index_min_doc_len(const slo_index_t *ix) { return ix->min_doc_len; }

int32_t slo_index_add_i64_column(slo_index_t *ix, const int64_t *values, const uint8_t *present) {
  Column c;
  c.kind = 0;
  c.i64.assign(values, values + ix->doc_count);
  if (present) c.present.assign(present, present + ix->doc_count);
  else c.present.assign(ix->doc_count, 1);
  ix->columns.push_back(std::move(c));
  return (int32_t)ix->columns.size() - 1;
}
int32_t slo_index_add_f64_column(slo_index_t *ix, const double *values, const uint8_t *present) {
  Column c;
  c.kind = 1;
  c.f64.assign(values, values + ix->doc_count);
  if (present) c.present.assign(present, present + ix->doc_count);
  else c.present.assign(ix->doc_count, 1);
  ix->columns.push_back(std::move(c));
  return (int32_t)ix->columns.size() - 1;
}
int32_t slo_index_add_str_column(slo_index_t *ix, const char *const *dict, uint32_t n_dict, const uint32_t *ords) {
  Column c;
  c.kind = 2;
  for (uint32_t i = 0; i < n_dict; i++) c.dict.emplace_back(dict[i]);
  c.ords.assign(ords, ords + ix->doc_count);
  ix->columns.push_back(std::move(c));
  return (int32_t)ix->columns.size() - 1;
}

int32_t slo_index_add_i64_list_column(slo_index_t *ix, const uint32_t *offsets, const int64_t *values) {
  Column c;
  c.kind = 3;
  c.offsets.assign(offsets, offsets + ix->doc_count + 1);
  c.i64.assign(values, values + c.offsets.back());
  ix->columns.push_back(std::move(c));
  return (int32_t)ix->columns.size() - 1;
}
int32_t slo_index_add_f64_list_column(slo_index_t *ix, const uint32_t *offsets, const double *values) {
  Column c;
  c.kind = 4;
  c.offsets.assign(offsets, offsets + ix->doc_count + 1);
  c.f64.assign(values, values + c.offsets.back());
  ix->columns.push_back(std::move(c));
  return (int32_t)ix->columns.size() - 1;
}
int32_t slo_index_add_str_list_column(slo_index_t *ix, const char *const *dict, uint32_t n_dict, const uint32_t *offsets,
                                      const uint32_t *ords) {
  Column c;
  c.kind = 5;
  for (uint32_t i = 0; i < n_dict; i++) c.dict.emplace_back(dict[i]);
  c.offsets.assign(offsets, offsets + ix->doc_count + 1);
  c.ords.assign(ords, ords + c.offsets.back());
  ix->columns.push_back(std::move(c));
  return (int32_t)ix->columns.size() - 1;
}

int32_t slo_search(const slo_index_t *ix, const slo_query_t *q, uint32_t k, int exec, uint32_t block_size,
                   const slo_filter_node_t *filter, uint32_t n_filter_nodes, const char *const *strings,
                   int faithful_decode, slo_hit_t *out_hits, slo_stats_t *stats) {
  Scratch scratch;
  return search_one(ix, q, k, exec, block_size, filter, n_filter_nodes, strings, faithful_decode, out_hits, stats,
                    scratch);
}

int32_t slo_search_batch(const slo_index_t *ix, const slo_query_t *qs, uint32_t n_queries, uint32_t k, int exec,
                         uint32_t block_size, const slo_filter_node_t *filter, uint32_t n_filter_nodes,
                         const char *const *strings, int faithful_decode, int threads, slo_hit_t *out_hits,
                         uint32_t *out_counts, slo_stats_t *out_stats) {
  std::atomic<int32_t> rc{0};
  std::atomic<int64_t> next{0};
  if (threads < 1) threads = 1;
  auto worker = [&]() {
    Scratch scratch;
    for (;;) {
      int64_t i = next.fetch_add(1);
      if (i >= (int64_t)n_queries) break;
      int32_t n = search_one(ix, &qs[i], k, exec, block_size, filter, n_filter_nodes, strings, faithful_decode,
                             out_hits + (size_t)i * k, out_stats ? &out_stats[i] : nullptr, scratch);
      if (n < 0) {
        rc.store(n);
        out_counts[i] = 0;
      } else {
        out_counts[i] = (uint32_t)n;
      }
    }
  };
  if (threads == 1) {
    worker();
  } else {
    std::vector<std::thread> pool;
    for (int t = 0; t < threads; t++) pool.emplace_back(worker);
    for (auto &th : pool) th.join();
  }
  return rc.load();
}

int slo_max_threads(void) {
  unsigned n = std::thread::hardware_concurrency();
  return n ? (int)n : 1;
}

// api/reader.rs:2777 hits.sort_by(SortKey) with query/sort.rs:80-136 for the `_score` desc plan:
// score desc by total_cmp, then segment_ord asc, then doc_id asc; truncate to limit (:2838-2851).
float slo_plan_evaluate(const slo_plan_node_t *plan, uint32_t n_nodes, const float *leaves, uint32_t n_leaves) {
  slo_query_t q{};
  q.plan = plan;
  q.n_plan_nodes = n_nodes;
  return evaluate_plan(&q, std::vector<float>(leaves, leaves + n_leaves));
}

uint32_t slo_merge_hits(const slo_hit_t *hits, uint32_t n, uint32_t limit, slo_hit_t *out) {
  std::vector<slo_hit_t> v(hits, hits + n);
  std::stable_sort(v.begin(), v.end(), [](const slo_hit_t &a, const slo_hit_t &b) {
    int c = total_cmp(b.score, a.score);
    if (c != 0) return c < 0;
    if (a.segment_ord != b.segment_ord) return a.segment_ord < b.segment_ord;
    return a.doc_id < b.doc_id;
  });
  uint32_t m = std::min(n, limit);
  std::memcpy(out, v.data(), (size_t)m * sizeof(slo_hit_t));
  return m;
}

int slo_filter_bitmap(const slo_index_t *ix, const slo_filter_node_t *filter, uint32_t n_filter_nodes,
                      const char *const *strings, uint32_t *bitmap_out) {
  uint32_t words = (ix->doc_count + 31) / 32;
  std::memset(bitmap_out, 0, (size_t)words * 4);
  for (uint32_t d = 0; d < ix->doc_count; d++) {
    uint32_t pos = 0;
    if (filter_eval(ix, d, filter, n_filter_nodes, pos, strings)) bitmap_out[d >> 5] |= 1u << (d & 31);
  }
  return 0;
}

// ---- phrases ----
// query/phrase.rs:22-39: the recursive search, kept depth first and in the reference's iteration order.
static bool phrase_search(const uint32_t *const *positions, const uint32_t *counts, uint32_t n, uint32_t idx, uint32_t prev,
                          int32_t remaining) {
  if (idx >= n) return true;
  for (uint32_t i = 0; i < counts[idx]; i++) {
    const uint32_t pos = positions[idx][i];
    if (pos <= prev) continue;
    const uint32_t prev1 = prev == 0xFFFFFFFFu ? prev : prev + 1;           // prev.saturating_add(1)
    const int32_t gap = (int32_t)(pos > prev1 ? pos - prev1 : 0u);           // pos.saturating_sub(..) as i32
    if (gap > remaining) break;  // positions are sorted; no later entry will shrink the gap
    if (phrase_search(positions, counts, n, idx + 1, pos, remaining - gap)) return true;
  }
  return false;
}
// query/phrase.rs:4-48 from the point where the doc's position lists are collected (:15-47)
int slo_matches_phrase_positions(uint32_t n_terms, const uint32_t *const *positions, const uint32_t *counts, uint32_t slop) {
  if (n_terms == 0) return 1;                       // :5-7
  for (uint32_t j = 0; j < n_terms; j++)
    if (counts[j] == 0) return 0;                   // :16-18
  if (n_terms == 1) return 1;                       // :19-21
  for (uint32_t i = 0; i < counts[0]; i++)          // :41-46
    if (phrase_search(positions, counts, n_terms, 1, positions[0][i], (int32_t)slop)) return 1;
  return 0;
}
// matches_phrase for every doc of a segment held as CSR postings + positions (one variant, one term per
// phrase position, api/reader.rs:1584-1597 + 1681-1712); a phrase term the segment lacks => no doc matches
int slo_phrase_bitmap(uint32_t doc_count, uint64_t n_terms, const uint64_t *term_offsets, const uint32_t *docs,
                      const uint64_t *pos_offsets, const uint32_t *positions, const uint32_t *phrase_terms, uint32_t n_phrase,
                      uint32_t slop, uint32_t *bitmap_out) {
  const uint32_t words = (doc_count + 31) / 32;
  std::memset(bitmap_out, 0, (size_t)words * 4);
  for (uint32_t j = 0; j < n_phrase; j++)
    if (phrase_terms[j] >= n_terms || term_offsets[phrase_terms[j] + 1] == term_offsets[phrase_terms[j]]) return 0;
  std::vector<const uint32_t *> lists(n_phrase);
  std::vector<uint32_t> counts(n_phrase);
  for (uint32_t d = 0; d < doc_count; d++) {
    bool all = true;
    for (uint32_t j = 0; j < n_phrase && all; j++) {
      const uint64_t lo = term_offsets[phrase_terms[j]], hi = term_offsets[phrase_terms[j] + 1];
      const uint32_t *it = std::lower_bound(docs + lo, docs + hi, d);  // term_posts.iter().find(|p| p.doc_id == doc_id), :10
      if (it == docs + hi || *it != d) {
        all = false;
        break;
      }
      const uint64_t pi = (uint64_t)(it - docs);
      lists[j] = positions + pos_offsets[pi];
      counts[j] = (uint32_t)(pos_offsets[pi + 1] - pos_offsets[pi]);
    }
    if (all && slo_matches_phrase_positions(n_phrase, lists.data(), counts.data(), slop)) bitmap_out[d >> 5] |= 1u << (d & 31);
  }
  return 0;
}

// vectors/mod.rs:74-81
void slo_normalize_in_place(float *v, size_t dim) {
  float s = 0.0f;
  for (size_t i = 0; i < dim; i++) s += v[i] * v[i];
  float norm = sqrtf(s);
  if (norm > 0.0f)
    for (size_t i = 0; i < dim; i++) v[i] /= norm;
}
// vectors/mod.rs:107-120 (and l2_distance :98-105)
float slo_metric_similarity(int metric, const float *a, const float *b, size_t dim) {
  if (metric == SLO_METRIC_COSINE) {
    float dot = 0.0f;
    for (size_t i = 0; i < dim; i++) dot += a[i] * b[i];
    return std::isnan(dot) ? 0.0f : dot;
  }
  float sum = 0.0f;
  for (size_t i = 0; i < dim; i++) {
    float d = a[i] - b[i];
    sum += d * d;
  }
  return -sqrtf(sum);
}
// vectors/mod.rs:122-129
float slo_blend_scores(float bm25_s, float vector_score, float alpha, int higher_is_better) {
  float vec_component = higher_is_better ? vector_score : -vector_score;
  return alpha * bm25_s + (1.0f - alpha) * vec_component;
}
// api/reader.rs:218-254, one clause (denominator = 1)
float slo_hybrid_score(float bm25_score, int has_vec, float vec_score, float alpha, int metric) {
  float vs = has_vec ? vec_score : (metric == SLO_METRIC_COSINE ? -1.0f : -std::numeric_limits<float>::max());
  float blended;
  if (alpha >= 1.0f) blended = bm25_score;
  else if (alpha <= 0.0f) blended = vs;
  else blended = slo_blend_scores(bm25_score, vs, alpha, 1);
  float blended_sum = 0.0f;
  blended_sum += blended;
  return blended_sum / 1.0f;
}

// api/reader.rs:226-254 for any number of clauses: vec_scores[c] is the clause's similarity already multiplied by its
// boost (api/reader.rs:2421) when has_vec[c], else ignored.  Returns the final score; *vector_sum / *has_vector as the
// tuple's second and third members.
float slo_hybrid_score_clauses(float bm25_score, uint32_t n_clauses, const int *has_vec, const float *vec_scores, const float *alpha,
                               const int *metric, float *vector_sum_out, int *has_vector_out) {
  float blended_sum = 0.0f, vector_sum = 0.0f;
  int has_vector = 0;
  for (uint32_t c = 0; c < n_clauses; c++) {
    if (has_vec[c]) {
      vector_sum += vec_scores[c];
      has_vector = 1;
    }
    const float vec_score = has_vec[c] ? vec_scores[c] : (metric[c] == SLO_METRIC_COSINE ? -1.0f : -std::numeric_limits<float>::max());
    float blended;
    if (alpha[c] >= 1.0f) blended = bm25_score;
    else if (alpha[c] <= 0.0f) blended = vec_score;
    else blended = slo_blend_scores(bm25_score, vec_score, alpha[c], 1);
    blended_sum += blended;
  }
  const float denom = (float)(n_clauses > 1 ? n_clauses : 1);
  if (vector_sum_out) *vector_sum_out = vector_sum;
  if (has_vector_out) *has_vector_out = has_vector;
  return blended_sum / denom;
}
// the engine's bf16 storage option (not in the reference): round-to-nearest-even to the upper 16 bits, in place
void slo_round_bf16(float *v, size_t n) {
  for (size_t i = 0; i < n; i++) {
    uint32_t b;
    std::memcpy(&b, &v[i], 4);
    if ((b & 0x7F800000u) == 0x7F800000u && (b & 0x007FFFFFu)) b |= 0x00400000u;  // NaN stays NaN
    else b += 0x7FFFu + ((b >> 16) & 1u);
    b &= 0xFFFF0000u;
    std::memcpy(&v[i], &b, 4);
  }
}

// The hybrid step of IndexReader::search for the BM25 candidate set (api/reader.rs:2752-2773 -> merge_vector_hits
// :2477-2537), exact instead of HNSW-approximate similarities: every candidate's vector is looked up
// (VectorStore::vector, vectors/mod.rs:63-71), scored per clause (metric_similarity x boost, api/reader.rs:2421), blended
// (compute_hybrid_score), vector-less candidates of an all-vector plan dropped (:2474-2476), and the rest ordered by
// (score desc, segment_ord asc, doc_id asc).  One vector store per segment_ord listed in seg_ords.
int slo_rerank_batch(uint32_t n_queries, uint32_t stride, const slo_hit_t *cands, const uint32_t *counts, uint32_t n_segs,
                     const uint32_t *seg_ords, const uint32_t *seg_doc_counts, const uint32_t *const *seg_offsets,
                     const float *const *seg_values, const uint64_t *seg_rows, uint32_t dim, uint32_t n_clauses,
                     const float *const *clause_qv, const float *alpha, const float *boost, const int *metric, slo_hit_t *out_hits,
                     uint32_t *out_counts, float *out_vs, int threads) {
  bool all_vector_only = true;
  for (uint32_t c = 0; c < n_clauses; c++) all_vector_only = all_vector_only && alpha[c] <= 0.0f;
  auto work = [&](uint32_t q0, uint32_t q1) {
    std::vector<int> has(n_clauses);
    std::vector<float> vs(n_clauses);
    struct Row {
      slo_hit_t h;
      float v;
    };
    std::vector<Row> rows;
    for (uint32_t q = q0; q < q1; q++) {
      rows.clear();
      const uint32_t n = std::min(counts[q], stride);
      for (uint32_t i = 0; i < n; i++) {
        slo_hit_t h = cands[(size_t)q * stride + i];
        const float *row = nullptr;
        for (uint32_t s = 0; s < n_segs; s++) {
          if (seg_ords[s] != h.segment_ord || h.doc_id >= seg_doc_counts[s]) continue;
          const uint32_t o = seg_offsets[s][h.doc_id];
          if (o != 0xFFFFFFFFu && (uint64_t)o < seg_rows[s]) row = seg_values[s] + (size_t)o * dim;
        }
        for (uint32_t c = 0; c < n_clauses; c++) {
          has[c] = row != nullptr;
          vs[c] = row ? slo_metric_similarity(metric[c], clause_qv[c] + (size_t)q * dim, row, dim) * boost[c] : 0.0f;
        }
        float vsum = 0.0f;
        int hv = 0;
        h.score = slo_hybrid_score_clauses(h.score, n_clauses, has.data(), vs.data(), alpha, metric, &vsum, &hv);
        if (all_vector_only && !hv) continue;
        rows.push_back({h, hv ? vsum : 0.0f});
      }
      std::stable_sort(rows.begin(), rows.end(), [](const Row &a, const Row &b) {
        int c = total_cmp(b.h.score, a.h.score);
        if (c != 0) return c < 0;
        if (a.h.segment_ord != b.h.segment_ord) return a.h.segment_ord < b.h.segment_ord;
        return a.h.doc_id < b.h.doc_id;
      });
      for (uint32_t i = 0; i < stride; i++) {
        slo_hit_t h{};
        h.segment_ord = 0xFFFFFFFFu;
        h.doc_id = 0xFFFFFFFFu;
        float v = 0.0f;
        if (i < rows.size()) {
          h = rows[i].h;
          v = rows[i].v;
        }
        out_hits[(size_t)q * stride + i] = h;
        if (out_vs) out_vs[(size_t)q * stride + i] = v;
      }
      out_counts[q] = (uint32_t)rows.size();
    }
  };
  const uint32_t nt = (uint32_t)std::max(1, std::min<int>(threads, (int)n_queries));
  if (nt <= 1) {
    work(0, n_queries);
  } else {
    std::vector<std::thread> pool;
    for (uint32_t t = 0; t < nt; t++)
      pool.emplace_back(work, (uint32_t)((uint64_t)n_queries * t / nt), (uint32_t)((uint64_t)n_queries * (t + 1) / nt));
    for (auto &th : pool) th.join();
  }
  return 0;
}

}  // extern "C"
