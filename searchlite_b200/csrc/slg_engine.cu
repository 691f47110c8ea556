// slg_engine.cu — host side of libsearchlite_gpu.so: residency, batch preparation, kernel
// launches and the C ABI declared in include/searchlite_gpu.h.
//
// Mirrors, for the hot path only:
//   SegmentReader::open / live_docs / avg_field_length   searchlite-core/src/index/segment.rs:1239,1344,1365
//   IndexReader::search_segment                           src/api/reader.rs:2908-3128
//   execute_top_k_with_stats_and_mode_internal            src/query/wand.rs:398-456
//   hits.sort_by(SortKey)                                 src/api/reader.rs:2777
// There is no CPU fallback anywhere in this file: every search runs the CUDA kernels or fails.
#include "../../include/searchlite_gpu.h"

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <cub/cub.cuh>
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>
#include <limits>
#include <memory>
#include <string>
#include <unordered_map>
#include <vector>

#include "slg_filter.cuh"
#include "slg_kernels.cuh"
#include "slg_phrase.cuh"
#include "slg_postimage.cuh"
#include "slg_segfiles.h"
#include "slg_rerank.cuh"
#include "slg_sweep_kernel.cuh"
#include "slg_warp_kernel.cuh"

using namespace slg;

namespace {

thread_local std::string g_open_error;

// Per-batch buffers come from the device's stream-ordered pool (cudaMallocAsync): a prepare/free pair
// per batch then costs microseconds instead of a cudaMalloc/cudaFree round trip per buffer.  The scope
// guard names the stream; everything else (segment residency) uses plain cudaMalloc.
thread_local cudaStream_t g_pool_stream = nullptr;
struct PoolScope {
  cudaStream_t prev;
  explicit PoolScope(cudaStream_t s) : prev(g_pool_stream) { g_pool_stream = s; }
  ~PoolScope() { g_pool_stream = prev; }
};

struct DevBuf {
  void *p = nullptr;
  size_t bytes = 0;
  cudaStream_t pool = nullptr;  // non-null: allocated with cudaMallocAsync on this stream
  bool borrowed = false;        // a view into another DevBuf's allocation: never freed here
  DevBuf() = default;
  DevBuf(const DevBuf &) = delete;
  DevBuf &operator=(const DevBuf &) = delete;
  DevBuf(DevBuf &&o) noexcept : p(o.p), bytes(o.bytes), pool(o.pool), borrowed(o.borrowed) { o.p = nullptr; o.bytes = 0; }
  DevBuf &operator=(DevBuf &&o) noexcept {
    if (this != &o) {
      release();
      p = o.p;
      bytes = o.bytes;
      pool = o.pool;
      borrowed = o.borrowed;
      o.p = nullptr;
      o.bytes = 0;
    }
    return *this;
  }
  ~DevBuf() { release(); }
  void view(void *ptr, size_t n) {  // borrow [ptr, ptr + n) from a slab that outlives this view
    release();
    p = ptr;
    bytes = n;
    borrowed = true;
  }
  void release() {
    if (p && !borrowed) {
      if (pool) cudaFreeAsync(p, pool);
      else cudaFree(p);
    }
    p = nullptr;
    bytes = 0;
    borrowed = false;
  }
  cudaError_t alloc(size_t n) {
    release();
    if (n == 0) n = 16;
    borrowed = false;
    pool = g_pool_stream;
    cudaError_t e = pool ? cudaMallocAsync(&p, n, pool) : cudaMalloc(&p, n);
    if (e == cudaSuccess) bytes = n;
    else p = nullptr;
    return e;
  }
  template <class T>
  T *as() const { return reinterpret_cast<T *>(p); }
};

struct CastU64 {
  __host__ __device__ uint64_t operator()(uint32_t v) const { return (uint64_t)v; }
};

struct Column {
  int kind = -1;  // 0 i64, 1 f64, 2 str; -1 = the segment lacks this column (predicates on it are false)
  DevBuf values; // i64 / f64 / u32 ords
  DevBuf present;
  std::vector<std::string> dict;
};

struct Vectors {
  uint32_t dim = 0;
  uint64_t n_rows = 0;
  bool bf16 = false;
  DevBuf offsets;  // u32[doc_count]
  DevBuf values;   // f32 or bf16 [n_rows][dim]
};

struct Segment {
  uint32_t ord = 0, doc_count = 0;
  uint64_t n_terms = 0, n_postings = 0, n_post_padded = 0;
  uint32_t n_blocks = 0, n_deleted = 0;
  float k1 = 0.9f, b = 0.4f, avgdl = 0, live_docs = 0, min_doc_len = 1;
  std::vector<uint32_t> h_df;  // host copy (query ordering, validation)
  DevBuf post_doc, post_tf, term_start, term_df, term_idf, term_max_tf, term_wide, tf_wide, term_blk, blk_max_doc,
      blk_max_tf, nk, live_bits, post_score, post_pair, cols, term_col, col_tmax;
  uint32_t n_cols = 0, tmax_stride = 0;
  uint64_t col_stride = 0;
  std::vector<int32_t> h_term_col;  // host copy (tests, introspection); empty = no columns
  SegmentDev dev{};
  std::vector<Column> columns;
  std::vector<DevBuf> filter_bits;  // per filter id (owning, or a view into one of filter_slabs)
  std::vector<std::shared_ptr<DevBuf>> filter_slabs;  // per filter id: the slab a view borrows from (or null)
  std::vector<uint64_t> h_start;    // host copy of term_start
  DevBuf filter_ptrs;               // device array of pointers into filter_bits
  Vectors vec;
  // term positions (index/postings.rs:117-125), kept for phrase matching: positions of padded posting slot i are
  // pos[pos_begin[i] .. pos_begin[i+1])
  DevBuf pos_begin, pos;
  uint64_t n_positions = 0;
  bool has_positions = false;
  bool avgdl_given = false;  // avgdl comes from the segment's .meta file instead of total_tokens / doc_count
  // further text fields of a handle that scores several ("title:..." next to "body:..."): field 0 is the one the
  // load call passes directly; these are set before finish_segment, which consumes the device copies
  struct ExtraField {
    DevBuf d_lens, d_present;
    float avgdl = 0.0f;
  };
  std::vector<ExtraField> extra_fields;
  std::vector<uint8_t> h_term_field;      // per term, empty = single field
  std::vector<float> f_avgdl, f_min_len;  // per field (index 0 = avgdl / min_doc_len)
  DevBuf term_field;
  size_t resident() const {
    return post_doc.bytes + post_tf.bytes + term_start.bytes + term_df.bytes + term_idf.bytes + term_max_tf.bytes +
           term_wide.bytes + tf_wide.bytes + term_blk.bytes + blk_max_doc.bytes + blk_max_tf.bytes + nk.bytes + term_field.bytes +
           live_bits.bytes + post_score.bytes + post_pair.bytes + cols.bytes + term_col.bytes + col_tmax.bytes +
           pos_begin.bytes + pos.bytes;
  }
};

struct FilterProg {
  std::vector<slg_filter_node_t> nodes;
  std::vector<std::string> strings;
};

}  // namespace

struct slg_index {
  int device = 0;
  int n_sm = 148;
  size_t smem_optin = 0;
  cudaStream_t stream = nullptr;
  cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
  std::vector<std::unique_ptr<Segment>> segs;
  std::vector<FilterProg> filters;
  std::string err;
  slg_counters_t ctr{};
  void *pinned = nullptr;        // host staging buffer kept between batches (one batch at a time uses it)
  size_t pinned_bytes = 0;
  bool pinned_busy = false;
  uint32_t tile_docs = 16384;
  uint32_t ctas_per_sm = 0;  // 0 = as many as shared memory allows
  uint32_t sub_docs = 2048;  // warp kernel: docs per warp-private accumulator
  uint32_t kernel_choice = 0;  // 0 auto, 1 CTA-per-item kernel, 2 warp-per-item kernel, 3 register-tile kernel
  bool staging = true;         // use the resident per-posting scores (seg.post_score) where a kernel can
  // residency options (slg_set_option), applied to segments loaded afterwards
  bool resident_scores = true;   // build seg.post_score at load
  uint32_t dense_den = 8;        // a term gets a dense column when df * dense_den >= doc_count; 0 = no columns
  uint32_t dense_min_df = 256;   // ... and df >= this
  uint64_t max_column_bytes = 24ull << 30;
  uint32_t reg_tile_v = 8;       // sweep kernel: 128 * V docs per tile (4 or 8)
  uint64_t sweep_min_postings = 0;  // sweep: a query without column terms and sum(df) below this goes to the warp kernel (0 = doc_count / 64)
  uint32_t seed_docs = 16384;    // sweep: docs of the seed pass
  uint32_t part_tiles = 0;       // sweep: tiles per unit of work (0 = automatic)
  uint32_t maxscore_pct = 35;    // pruned warp kernel: non-essential bounds may sum to this % of the k-th score (0 = tile skip only)
  uint32_t heavy_kernel = 0;     // column front end: 0 = warp kernel summing column terms from their columns, 1 = tile-sweep kernel
  bool keep_positions = true;    // keep term positions resident when a posting image carries them (SegmentReader keep_positions)
  // term space of segments loaded from the reference's files: "field:token" key -> term id, in order of first appearance
  std::unordered_map<std::string, uint32_t> term_ids;
  std::string term_field;        // the text field(s) those keys belong to, as named at load ("body" or "title,body")
  std::vector<uint8_t> term_field_of;  // term id -> index of its field in that list
  // fast-field columns by name (handles are indices into every segment's `columns`)
  std::vector<std::string> column_names;
  Segment *find(uint32_t ord) {
    for (auto &s : segs)
      if (s->ord == ord) return s.get();
    return nullptr;
  }
};

struct slg_batch {
  slg_index *ix = nullptr;
  uint32_t Q = 0, k = 0, cap = 0, U = 0, T = 0;
  slg_exec_t exec = SLG_EXEC_BM25;
  bool matcher = false;
  bool want_stats = false;
  uint64_t posting_count = 0;
  // packed inputs
  std::vector<unsigned char> h_pack;
  DevBuf d_pack;
  size_t off_ut_term = 0, off_q_term_off = 0, off_qt_uterm = 0, off_qt_weight = 0, off_qt_group = 0, off_qt_flags = 0,
         off_q_order = 0, off_q_must = 0, off_q_not = 0, off_q_should = 0, off_q_min = 0, off_q_filter = 0;
  std::vector<uint32_t> h_ut_term, h_qt_uterm, h_q_term_off;
  // state + outputs
  DevBuf ut_rng, ut_tile_ub, thr_key, topk_count, lock, topk_keys, work_counter, stats;
  DevBuf qterms, qheads;  // warp / register kernel term tables
  bool staged = false;
  uint32_t max_terms = 0;
  bool use_warp = false, use_reg = false;
  bool has_cursor = false;   // some query carries a search-after cursor
  std::vector<uint8_t> h_has_cursor;
  DevBuf cursor_bounds;      // u64 [n_segs][Q]: exclusive upper key bound per segment
  DevBuf cursor_saw;         // u32 [Q]
  uint32_t n_cursor_segs = 0;
  bool has_plan = false;     // some query carries a ScorePlan: CTA-per-item kernel with per-leaf accumulator planes
  uint32_t max_leaves = 1;
  size_t off_qt_leaf = 0, off_q_leaves = 0, off_q_plan_off = 0, off_plan_nodes = 0;
  uint32_t plan_docs = 0, reg_v = 8;
  uint32_t n_heavy = 0, n_light = 0;  // sweep: slots [0, n_heavy) of q_order are swept, the rest go to the warp kernel
  uint32_t n_rows = 0, n_light_u = 0; // rows of the sweep's range table; unique terms of the light queries
  uint32_t sweep_tiles_max = 0, sub_tiles_max = 0;
  DevBuf d_u_row, d_row_u, d_light_u; // [U] row or ~0; [n_rows] unique term; [n_light_u] unique term
  DevBuf sw_sstat, sw_weights, sw_ubw, sw_rng, sw_records, d_chunk_cols;  // d_chunk_cols: [S][n_chunks][kSweepStage]
  uint32_t n_chunks = 0;
  bool any_weight = false;             // some scored term has weight != 1
  bool warp_cols = false;              // the warp kernel sums column terms from their dense columns
  DevBuf seg_hits, seg_counts;  // [S][Q][k], [S][Q]
  DevBuf out_hits, out_counts;  // merged (aliases seg buffers when S == 1)
  uint32_t n_segs_run = 0;
  void *pinned = nullptr;
  size_t pinned_bytes = 0;
  bool pinned_from_index = false;
  ~slg_batch() {
    if (pinned_from_index) ix->pinned_busy = false;
    else if (pinned) {
      if (!ix->pinned_busy && pinned_bytes > ix->pinned_bytes) {  // keep the larger buffer for the next batch
        if (ix->pinned) cudaFreeHost(ix->pinned);
        ix->pinned = pinned;
        ix->pinned_bytes = pinned_bytes;
      } else {
        cudaFreeHost(pinned);
      }
    }
  }
};

namespace {

int32_t fail(slg_index *ix, int32_t code, const char *fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  if (ix) ix->err = buf;
  else g_open_error = buf;
  return code;
}

#define SLG_CUDA(ix, call)                                                                         \
  do {                                                                                             \
    cudaError_t e__ = (call);                                                                      \
    if (e__ != cudaSuccess)                                                                        \
      return fail((ix), e__ == cudaErrorMemoryAllocation ? SLG_ERR_OOM : SLG_ERR_CUDA, "%s: %s (%s:%d)", #call, \
                  cudaGetErrorString(e__), __FILE__, __LINE__);                                    \
  } while (0)

inline void count_launch(slg_index *ix, uint64_t n = 1) { ix->ctr.kernel_launches += n; }

// idf exactly as query/bm25.rs:2 with docs = live docs (api/reader.rs:2985) and df = list length
// f32::max returns the non-NaN operand (ln of a negative ratio when df > N + 0.5 after deletions): fmaxf
inline float host_idf(float df, float docs) { return fmaxf(logf((docs - df + 0.5f) / (df + 0.5f)), 0.0f) + 1.0f; }
inline float host_nk(float dl, float avgdl, float k1, float b) {
  volatile float norm = avgdl > 0.0f ? dl / avgdl : 1.0f;
  volatile float bn = b * norm;
  volatile float omb = 1.0f - b;
  volatile float s = omb + bn;
  volatile float r = k1 * s;
  return r;
}

size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// SLG_LOAD_TRACE=1: wall-clock time of every residency stage on stderr (the stream is synchronised at each mark)
struct StageTimer {
  bool on;
  cudaStream_t st;
  std::chrono::steady_clock::time_point t0;
  explicit StageTimer(cudaStream_t s) : on(getenv("SLG_LOAD_TRACE") != nullptr), st(s), t0(std::chrono::steady_clock::now()) {}
  void mark(const char *what) {
    if (!on) return;
    cudaStreamSynchronize(st);
    const auto t1 = std::chrono::steady_clock::now();
    fprintf(stderr, "[slg load] %-28s %8.1f ms\n", what, std::chrono::duration<double, std::milli>(t1 - t0).count());
    t0 = t1;
  }
};

// --------------------------------------------------------------------------------------------
// residency
int32_t finish_segment(slg_index *ix, std::unique_ptr<Segment> seg, const int64_t *d_lens, const uint8_t *d_present,
                       uint64_t total_tokens, const uint32_t *deleted, uint32_t n_deleted) {
  cudaStream_t st = ix->stream;
  Segment *s = seg.get();
  // compute_avg_lengths index/segment.rs:946-957
  if (!s->avgdl_given) s->avgdl = s->doc_count == 0 ? 0.0f : (float)total_tokens / (float)(uint64_t)s->doc_count;
  // deleted docs -> live bitmap; live_docs index/segment.rs:1365-1370
  uint32_t words = (s->doc_count + 31) / 32;
  std::vector<uint32_t> live(words, 0xFFFFFFFFu);
  if (s->doc_count & 31) live[words - 1] = (1u << (s->doc_count & 31)) - 1u;
  uint32_t nd = 0;
  for (uint32_t i = 0; i < n_deleted; i++) {
    uint32_t d = deleted[i];
    if (d < s->doc_count && (live[d >> 5] >> (d & 31) & 1u)) {
      live[d >> 5] &= ~(1u << (d & 31));
      nd++;
    }
  }
  s->n_deleted = nd;
  s->live_docs = (float)(s->doc_count - nd);
  SLG_CUDA(ix, s->live_bits.alloc((size_t)std::max(words, 1u) * 4));
  if (words) SLG_CUDA(ix, cudaMemcpyAsync(s->live_bits.p, live.data(), (size_t)words * 4, cudaMemcpyHostToDevice, st));
  // idf per term on the host (glibc logf, like the reference's f32::ln)
  std::vector<float> idf(s->n_terms);
  for (uint64_t t = 0; t < s->n_terms; t++) idf[t] = host_idf((float)s->h_df[t], s->live_docs);
  SLG_CUDA(ix, s->term_idf.alloc(s->n_terms * 4));
  if (s->n_terms) SLG_CUDA(ix, cudaMemcpyAsync(s->term_idf.p, idf.data(), s->n_terms * 4, cudaMemcpyHostToDevice, st));
  // norms: one [doc_count] vector per scored text field
  const uint32_t n_fields = 1 + (uint32_t)s->extra_fields.size();
  if (n_fields > kMaxFields) return fail(ix, SLG_ERR_UNSUPPORTED, "more than %u text fields in one handle", kMaxFields);
  const size_t nk_stride = s->doc_count;
  SLG_CUDA(ix, s->nk.alloc(std::max<size_t>(nk_stride * n_fields, 1) * 4));
  s->f_avgdl.assign(n_fields, 0.0f);
  s->f_min_len.assign(n_fields, 1.0f);
  DevBuf minbits;
  SLG_CUDA(ix, minbits.alloc(4));
  for (uint32_t f = 0; f < n_fields; f++) {
    const float avgdl_f = f == 0 ? s->avgdl : s->extra_fields[f - 1].avgdl;
    const int64_t *lens_f = f == 0 ? d_lens : s->extra_fields[f - 1].d_lens.as<int64_t>();
    const uint8_t *pres_f = f == 0 ? d_present : s->extra_fields[f - 1].d_present.as<uint8_t>();
    uint32_t inf_bits = 0x7F800000u;
    SLG_CUDA(ix, cudaMemcpyAsync(minbits.p, &inf_bits, 4, cudaMemcpyHostToDevice, st));
    if (s->doc_count) {
      slg_norms_kernel<<<(s->doc_count + 255) / 256, 256, 0, st>>>(lens_f, pres_f, s->doc_count, avgdl_f, s->k1, s->b,
                                                                     s->nk.as<float>() + nk_stride * f, minbits.as<uint32_t>());
      count_launch(ix);
    }
    uint32_t got = 0;
    SLG_CUDA(ix, cudaMemcpyAsync(&got, minbits.p, 4, cudaMemcpyDeviceToHost, st));
    SLG_CUDA(ix, cudaStreamSynchronize(st));
    float mn;
    std::memcpy(&mn, &got, 4);
    s->f_avgdl[f] = avgdl_f;
    s->f_min_len[f] = std::isfinite(mn) ? mn : std::max(avgdl_f, 1.0f);  // query/wand.rs:117-121
  }
  s->min_doc_len = s->f_min_len[0];
  s->extra_fields.clear();  // the device copies of the length columns are no longer needed
  if (n_fields > 1) {
    if (s->h_term_field.size() != s->n_terms) return fail(ix, SLG_ERR_INVALID, "multi-field segment without a term -> field table");
    SLG_CUDA(ix, s->term_field.alloc(std::max<uint64_t>(s->n_terms, 1)));
    if (s->n_terms) SLG_CUDA(ix, cudaMemcpyAsync(s->term_field.p, s->h_term_field.data(), s->n_terms, cudaMemcpyHostToDevice, st));
  }

  SegmentDev &d = s->dev;
  d.post_doc = s->post_doc.as<uint32_t>();
  d.post_tf = s->post_tf.as<uint8_t>();
  d.term_start = s->term_start.as<uint64_t>();
  d.term_df = s->term_df.as<uint32_t>();
  d.term_idf = s->term_idf.as<float>();
  d.term_max_tf = s->term_max_tf.as<float>();
  d.term_wide = s->term_wide.as<uint64_t>();
  d.tf_wide = s->tf_wide.as<uint32_t>();
  d.term_blk = s->term_blk.as<uint32_t>();
  d.blk_max_doc = s->blk_max_doc.as<uint32_t>();
  d.blk_max_tf = s->blk_max_tf.as<float>();
  d.nk = s->nk.as<float>();
  d.live_bits = s->live_bits.as<uint32_t>();
  d.post_score = nullptr;
  d.cols = nullptr;
  d.term_col = nullptr;
  d.col_tmax = nullptr;
  d.tmax_stride = 0;
  d.col_stride = 0;
  d.n_terms = s->n_terms;
  d.doc_count = s->doc_count;
  d.k1p1 = s->k1 + 1.0f;
  d.min_nk = host_nk(s->min_doc_len, s->avgdl, s->k1, s->b);
  d.term_field = n_fields > 1 ? s->term_field.as<uint8_t>() : nullptr;
  for (uint32_t f = 0; f < kMaxFields; f++)
    d.min_nk_f[f] = f < n_fields ? host_nk(s->f_min_len[f], s->f_avgdl[f], s->k1, s->b) : d.min_nk;

  // resident unit-weight scores: score_tf(tf, df, doc_len, ...) of every posting, once
  if (ix->resident_scores && s->n_blocks) {
    SLG_CUDA(ix, s->post_score.alloc(s->n_post_padded * 4));
    SLG_CUDA(ix, cudaMemsetAsync(s->post_score.p, 0, s->n_post_padded * 4, st));
    slg_score_postings_kernel<<<s->n_blocks, 128, 0, st>>>(d, s->n_blocks, s->post_score.as<float>());
    count_launch(ix);
    SLG_CUDA(ix, cudaGetLastError());
    d.post_score = s->post_score.as<float>();
    if (s->n_post_padded < (1ull << 32)) {  // the sweep kernel's posting stream (32-bit posting indices)
      SLG_CUDA(ix, s->post_pair.alloc(s->n_post_padded * 8));
      slg_pair_postings_kernel<<<ix->n_sm * 8, 256, 0, st>>>(s->post_doc.as<uint32_t>(), s->post_score.as<float>(), s->n_post_padded,
                                                             s->post_pair.as<uint2>());
      count_launch(ix);
      SLG_CUDA(ix, cudaGetLastError());
    }
    // dense columns for the high-df terms, largest df first until the byte budget is spent
    if (ix->dense_den && s->doc_count) {
      const uint64_t stride = align_up((uint64_t)s->doc_count, 4096) + 4096;  // a whole staged block past the end stays in bounds and zero
      std::vector<uint32_t> cand;
      for (uint64_t t = 0; t < s->n_terms; t++) {
        const uint64_t df = s->h_df[t];
        if (df >= ix->dense_min_df && df * ix->dense_den >= s->doc_count) cand.push_back((uint32_t)t);
      }
      std::sort(cand.begin(), cand.end(), [&](uint32_t a, uint32_t b2) { return s->h_df[a] != s->h_df[b2] ? s->h_df[a] > s->h_df[b2] : a < b2; });
      const uint64_t max_cols = std::min<uint64_t>(65535, ix->max_column_bytes / (stride * 4));  // 16-bit column ids in the sweep records
      if (cand.size() > max_cols) cand.resize(max_cols);
      if (!cand.empty()) {
        std::vector<int32_t> tcol(s->n_terms, -1);
        for (size_t c = 0; c < cand.size(); c++) tcol[cand[c]] = (int32_t)c;
        DevBuf d_terms;
        SLG_CUDA(ix, d_terms.alloc(cand.size() * 4));
        SLG_CUDA(ix, cudaMemcpyAsync(d_terms.p, cand.data(), cand.size() * 4, cudaMemcpyHostToDevice, st));
        SLG_CUDA(ix, s->term_col.alloc(s->n_terms * 4));
        SLG_CUDA(ix, cudaMemcpyAsync(s->term_col.p, tcol.data(), s->n_terms * 4, cudaMemcpyHostToDevice, st));
        SLG_CUDA(ix, s->cols.alloc(cand.size() * stride * 4));
        SLG_CUDA(ix, cudaMemsetAsync(s->cols.p, 0, cand.size() * stride * 4, st));
        s->n_cols = (uint32_t)cand.size();
        s->col_stride = stride;
        d.col_stride = stride;
        for (size_t c0 = 0; c0 < cand.size(); c0 += 32768) {  // gridDim.y limit
          const uint32_t nc = (uint32_t)std::min<size_t>(32768, cand.size() - c0);
          slg_fill_columns_kernel<<<dim3(256, nc), 256, 0, st>>>(d, d_terms.as<uint32_t>() + c0, nc,
                                                                  s->cols.as<float>() + c0 * stride);
          count_launch(ix);
        }
        SLG_CUDA(ix, cudaGetLastError());
        // exact per-512-doc maxima of every column: the tile bounds of the pruned modes
        s->tmax_stride = (uint32_t)(stride / 512);
        SLG_CUDA(ix, s->col_tmax.alloc(cand.size() * s->tmax_stride * 4));
        {
          const uint64_t warps = (uint64_t)cand.size() * s->tmax_stride;
          slg_column_tmax_kernel<<<(unsigned)((warps * 32 + 255) / 256), 256, 0, st>>>(s->cols.as<float>(), stride, (uint32_t)cand.size(),
                                                                                     s->tmax_stride, s->col_tmax.as<float>());
          count_launch(ix);
        }
        SLG_CUDA(ix, cudaGetLastError());
        SLG_CUDA(ix, cudaStreamSynchronize(st));  // host vectors go out of scope
        s->h_term_col = std::move(tcol);
        d.cols = s->cols.as<float>();
        d.term_col = s->term_col.as<int32_t>();
        d.col_tmax = s->col_tmax.as<float>();
        d.tmax_stride = s->tmax_stride;
      }
    }
  }
  ix->ctr.resident_bytes += s->resident();
  // replace a segment with the same ordinal
  for (auto &old : ix->segs)
    if (old->ord == s->ord) {
      ix->ctr.resident_bytes -= old->resident();
      old = std::move(seg);
      return SLG_OK;
    }
  ix->segs.push_back(std::move(seg));
  std::sort(ix->segs.begin(), ix->segs.end(), [](const auto &a, const auto &b2) { return a->ord < b2->ord; });
  return SLG_OK;
}

// layout tables shared by both load paths: df -> padded starts, block starts
int32_t build_layout(slg_index *ix, Segment *s) {
  cudaStream_t st = ix->stream;
  std::vector<uint64_t> start(s->n_terms + 1);
  std::vector<uint32_t> blk(s->n_terms + 1);
  uint64_t pos = 0, nb = 0, np = 0;
  for (uint64_t t = 0; t < s->n_terms; t++) {
    start[t] = pos;
    blk[t] = (uint32_t)nb;
    uint32_t df = s->h_df[t];
    np += df;
    pos += align_up(df, kTermAlign);
    nb += (df + kBlock - 1) / kBlock;
    if (nb > 0xFFFFFFF0ull) return fail(ix, SLG_ERR_UNSUPPORTED, "segment has too many posting blocks");
  }
  start[s->n_terms] = pos;
  blk[s->n_terms] = (uint32_t)nb;
  s->n_postings = np;
  s->n_post_padded = pos + 1024;  // tail slack so that vector loads past a list never leave the buffer
  s->n_blocks = (uint32_t)nb;
  SLG_CUDA(ix, s->term_start.alloc((s->n_terms + 1) * 8));
  SLG_CUDA(ix, s->term_blk.alloc((s->n_terms + 1) * 4));
  SLG_CUDA(ix, s->term_df.alloc(std::max<uint64_t>(s->n_terms, 1) * 4));
  SLG_CUDA(ix, cudaMemcpyAsync(s->term_start.p, start.data(), (s->n_terms + 1) * 8, cudaMemcpyHostToDevice, st));
  SLG_CUDA(ix, cudaMemcpyAsync(s->term_blk.p, blk.data(), (s->n_terms + 1) * 4, cudaMemcpyHostToDevice, st));
  if (s->n_terms) SLG_CUDA(ix, cudaMemcpyAsync(s->term_df.p, s->h_df.data(), s->n_terms * 4, cudaMemcpyHostToDevice, st));
  SLG_CUDA(ix, s->post_doc.alloc(s->n_post_padded * 4));
  SLG_CUDA(ix, s->post_tf.alloc(s->n_post_padded));
  SLG_CUDA(ix, cudaMemsetAsync(s->post_doc.p, 0xFF, s->n_post_padded * 4, st));
  SLG_CUDA(ix, cudaMemsetAsync(s->post_tf.p, 0, s->n_post_padded, st));
  SLG_CUDA(ix, s->blk_max_doc.alloc((size_t)std::max(s->n_blocks, 1u) * 4));
  SLG_CUDA(ix, s->blk_max_tf.alloc((size_t)std::max(s->n_blocks, 1u) * 4));
  SLG_CUDA(ix, s->term_max_tf.alloc(std::max<uint64_t>(s->n_terms, 1) * 4));
  SLG_CUDA(ix, s->term_wide.alloc(std::max<uint64_t>(s->n_terms, 1) * 8));
  SLG_CUDA(ix, cudaMemsetAsync(s->term_wide.p, 0xFF, std::max<uint64_t>(s->n_terms, 1) * 8, st));
  SLG_CUDA(ix, cudaStreamSynchronize(st));  // host vectors go out of scope
  s->h_start = std::move(start);
  return SLG_OK;
}

// after post_tf / blk_max_tf are filled: per-term max tf and the wide-tf side table
int32_t build_wide(slg_index *ix, Segment *s, const uint64_t *d_csr_off, const uint32_t *d_csr_tfs) {
  cudaStream_t st = ix->stream;
  if (!s->n_terms) return SLG_OK;
  slg_term_max_tf_kernel<<<(unsigned)((s->n_terms + 255) / 256), 256, 0, st>>>(s->term_blk.as<uint32_t>(), s->blk_max_tf.as<float>(),
                                                                              s->n_terms, s->term_max_tf.as<float>());
  count_launch(ix);
  std::vector<float> mtf(s->n_terms);
  SLG_CUDA(ix, cudaMemcpyAsync(mtf.data(), s->term_max_tf.p, s->n_terms * 4, cudaMemcpyDeviceToHost, st));
  SLG_CUDA(ix, cudaStreamSynchronize(st));
  std::vector<uint32_t> wide_terms;
  std::vector<uint64_t> wide_off;
  uint64_t wpos = 0;
  for (uint64_t t = 0; t < s->n_terms; t++)
    if (mtf[t] >= 255.0f) {
      wide_terms.push_back((uint32_t)t);
      wide_off.push_back(wpos);
      wpos += s->h_df[t];
    }
  if (wide_terms.empty()) return SLG_OK;
  if (!d_csr_tfs) return fail(ix, SLG_ERR_UNSUPPORTED, "term frequency >= 255 needs the CSR load path");
  SLG_CUDA(ix, s->tf_wide.alloc(wpos * 4));
  DevBuf d_wt, d_wo;
  SLG_CUDA(ix, d_wt.alloc(wide_terms.size() * 4));
  SLG_CUDA(ix, d_wo.alloc(wide_off.size() * 8));
  SLG_CUDA(ix, cudaMemcpyAsync(d_wt.p, wide_terms.data(), wide_terms.size() * 4, cudaMemcpyHostToDevice, st));
  SLG_CUDA(ix, cudaMemcpyAsync(d_wo.p, wide_off.data(), wide_off.size() * 8, cudaMemcpyHostToDevice, st));
  // term_wide[t] = offset such that tf_wide[term_wide[t] + i] is posting i of the term
  std::vector<uint64_t> tw(s->n_terms, ~0ull);
  for (size_t w = 0; w < wide_terms.size(); w++) tw[wide_terms[w]] = wide_off[w];
  SLG_CUDA(ix, cudaMemcpyAsync(s->term_wide.p, tw.data(), s->n_terms * 8, cudaMemcpyHostToDevice, st));
  dim3 grid(64, (unsigned)wide_terms.size());
  slg_wide_tf_kernel<<<grid, 256, 0, st>>>(d_csr_off, d_csr_tfs, d_wt.as<uint32_t>(), d_wo.as<uint64_t>(),
                                           (uint32_t)wide_terms.size(), s->tf_wide.as<uint32_t>());
  count_launch(ix);
  SLG_CUDA(ix, cudaStreamSynchronize(st));
  return SLG_OK;
}

template <class T>
int32_t to_device(slg_index *ix, const T *src, size_t n, int space, DevBuf &tmp, const T **out) {
  if (!src || n == 0) {
    *out = nullptr;
    return SLG_OK;
  }
  if (space == SLG_MEM_DEVICE) {
    *out = src;
    return SLG_OK;
  }
  SLG_CUDA(ix, tmp.alloc(n * sizeof(T)));
  SLG_CUDA(ix, cudaMemcpyAsync(tmp.p, src, n * sizeof(T), cudaMemcpyHostToDevice, ix->stream));
  *out = tmp.as<T>();
  return SLG_OK;
}

// --------------------------------------------------------------------------------------------
// search
int32_t select_smem(slg_index *ix, uint32_t tile_docs, uint32_t cap, bool matcher, size_t *out, uint32_t planes = 1) {
  size_t smem = (size_t)tile_docs * 4 * planes + (size_t)cap * 8 + (matcher ? tile_docs : 0);
  if (smem + 1024 > ix->smem_optin)
    return fail(ix, SLG_ERR_UNSUPPORTED, "tile of %u docs with k buffer %u needs %zu B shared memory", tile_docs, cap, smem);
  *out = smem;
  return SLG_OK;
}

template <bool M, bool P, bool S, bool PL = false>
int32_t launch_score_t(slg_index *ix, const SegmentDev &sd, const BatchDev &bd, size_t smem, int grid) {
  auto kern = slg_score_tiles_kernel<M, P, S, PL>;
  SLG_CUDA(ix, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  kern<<<grid, kThreads, smem, ix->stream>>>(sd, bd);
  SLG_CUDA(ix, cudaGetLastError());
  return SLG_OK;
}

int32_t launch_score(slg_index *ix, bool matcher, bool prune, bool stats, const SegmentDev &sd, const BatchDev &bd,
                     size_t smem, int grid, bool plan = false) {
  int sel = (matcher ? 4 : 0) | (prune ? 2 : 0) | (stats ? 1 : 0);
  if (plan) {  // ScorePlan batches: the matcher form serves both (plain OR queries carry an empty group program)
    switch (sel & 3) {
      case 0: return matcher ? launch_score_t<true, false, false, true>(ix, sd, bd, smem, grid)
                             : launch_score_t<false, false, false, true>(ix, sd, bd, smem, grid);
      case 1: return matcher ? launch_score_t<true, false, true, true>(ix, sd, bd, smem, grid)
                             : launch_score_t<false, false, true, true>(ix, sd, bd, smem, grid);
      case 2: return matcher ? launch_score_t<true, true, false, true>(ix, sd, bd, smem, grid)
                             : launch_score_t<false, true, false, true>(ix, sd, bd, smem, grid);
      default: return matcher ? launch_score_t<true, true, true, true>(ix, sd, bd, smem, grid)
                              : launch_score_t<false, true, true, true>(ix, sd, bd, smem, grid);
    }
  }
  switch (sel) {
    case 0: return launch_score_t<false, false, false>(ix, sd, bd, smem, grid);
    case 1: return launch_score_t<false, false, true>(ix, sd, bd, smem, grid);
    case 2: return launch_score_t<false, true, false>(ix, sd, bd, smem, grid);
    case 3: return launch_score_t<false, true, true>(ix, sd, bd, smem, grid);
    case 4: return launch_score_t<true, false, false>(ix, sd, bd, smem, grid);
    case 5: return launch_score_t<true, false, true>(ix, sd, bd, smem, grid);
    case 6: return launch_score_t<true, true, false>(ix, sd, bd, smem, grid);
    default: return launch_score_t<true, true, true>(ix, sd, bd, smem, grid);
  }
}

template <bool M, bool P, bool S, bool G, bool C = false, bool PL = false>
int32_t launch_warp_t(slg_index *ix, const SegmentDev &sd, const WarpBatchDev &wb, size_t smem, int grid) {
  auto kern = slg_score_warp_kernel<M, P, S, G, C, PL>;
  SLG_CUDA(ix, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  kern<<<grid, kThreads, smem, ix->stream>>>(sd, wb);
  SLG_CUDA(ix, cudaGetLastError());
  return SLG_OK;
}

int32_t launch_warp(slg_index *ix, bool matcher, bool prune, bool stats, bool staged, bool cols, const SegmentDev &sd,
                    const WarpBatchDev &wb, size_t smem, int grid, bool plan = false) {
  int sel = (matcher ? 4 : 0) | (prune ? 2 : 0) | (stats ? 1 : 0);
  if (plan) {  // ScorePlans: resident scores for plain OR queries, in-place scoring + group masks otherwise
    const bool st = staged && !matcher;
    switch (sel & 3) {
      case 0: return st ? launch_warp_t<false, false, false, true, false, true>(ix, sd, wb, smem, grid)
                        : launch_warp_t<true, false, false, false, false, true>(ix, sd, wb, smem, grid);
      case 1: return st ? launch_warp_t<false, false, true, true, false, true>(ix, sd, wb, smem, grid)
                        : launch_warp_t<true, false, true, false, false, true>(ix, sd, wb, smem, grid);
      case 2: return st ? launch_warp_t<false, true, false, true, false, true>(ix, sd, wb, smem, grid)
                        : launch_warp_t<true, true, false, false, false, true>(ix, sd, wb, smem, grid);
      default: return st ? launch_warp_t<false, true, true, true, false, true>(ix, sd, wb, smem, grid)
                         : launch_warp_t<true, true, true, false, false, true>(ix, sd, wb, smem, grid);
    }
  }
  if (staged && !matcher && cols) {
    switch (sel) {
      case 0: return launch_warp_t<false, false, false, true, true>(ix, sd, wb, smem, grid);
      case 1: return launch_warp_t<false, false, true, true, true>(ix, sd, wb, smem, grid);
      case 2: return launch_warp_t<false, true, false, true, true>(ix, sd, wb, smem, grid);
      default: return launch_warp_t<false, true, true, true, true>(ix, sd, wb, smem, grid);
    }
  }
  if (staged && !matcher) {
    switch (sel) {
      case 0: return launch_warp_t<false, false, false, true>(ix, sd, wb, smem, grid);
      case 1: return launch_warp_t<false, false, true, true>(ix, sd, wb, smem, grid);
      case 2: return launch_warp_t<false, true, false, true>(ix, sd, wb, smem, grid);
      default: return launch_warp_t<false, true, true, true>(ix, sd, wb, smem, grid);
    }
  }
  switch (sel) {
    case 0: return launch_warp_t<false, false, false, false>(ix, sd, wb, smem, grid);
    case 1: return launch_warp_t<false, false, true, false>(ix, sd, wb, smem, grid);
    case 2: return launch_warp_t<false, true, false, false>(ix, sd, wb, smem, grid);
    case 3: return launch_warp_t<false, true, true, false>(ix, sd, wb, smem, grid);
    case 4: return launch_warp_t<true, false, false, false>(ix, sd, wb, smem, grid);
    case 5: return launch_warp_t<true, false, true, false>(ix, sd, wb, smem, grid);
    case 6: return launch_warp_t<true, true, false, false>(ix, sd, wb, smem, grid);
    default: return launch_warp_t<true, true, true, false>(ix, sd, wb, smem, grid);
  }
}

template <int V, bool P, bool S, bool W>
int32_t launch_sweep_t(slg_index *ix, const SegmentDev &sd, const SweepDev &sw, int grid) {
  auto kern = slg_score_sweep_kernel<V, P, S, W>;
  const size_t smem = sweep_smem_bytes<V>();
  if (smem + 1024 > ix->smem_optin) return fail(ix, SLG_ERR_UNSUPPORTED, "sweep kernel needs %zu B shared memory", smem);
  SLG_CUDA(ix, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  kern<<<grid, kSweepThreads, smem, ix->stream>>>(sd, sw);
  SLG_CUDA(ix, cudaGetLastError());
  return SLG_OK;
}

template <int V>
int32_t launch_sweep_v(slg_index *ix, bool prune, bool stats, bool weights, const SegmentDev &sd, const SweepDev &sw, int grid) {
  switch ((prune ? 4 : 0) | (stats ? 2 : 0) | (weights ? 1 : 0)) {
    case 0: return launch_sweep_t<V, false, false, false>(ix, sd, sw, grid);
    case 1: return launch_sweep_t<V, false, false, true>(ix, sd, sw, grid);
    case 2: return launch_sweep_t<V, false, true, false>(ix, sd, sw, grid);
    case 3: return launch_sweep_t<V, false, true, true>(ix, sd, sw, grid);
    case 4: return launch_sweep_t<V, true, false, false>(ix, sd, sw, grid);
    case 5: return launch_sweep_t<V, true, false, true>(ix, sd, sw, grid);
    case 6: return launch_sweep_t<V, true, true, false>(ix, sd, sw, grid);
    default: return launch_sweep_t<V, true, true, true>(ix, sd, sw, grid);
  }
}

int32_t launch_sweep(slg_index *ix, uint32_t v, bool prune, bool stats, bool weights, const SegmentDev &sd, const SweepDev &sw, int grid) {
  switch (v) {
    case 4: return launch_sweep_v<4>(ix, prune, stats, weights, sd, sw, grid);
    default: return launch_sweep_v<8>(ix, prune, stats, weights, sd, sw, grid);
  }
}

}  // namespace

/* ================================================================================================ C ABI */
extern "C" {

const char *slg_version(void) { return "searchlite-b200 0.1 (sm_100a)"; }

const char *slg_last_error(const slg_index_t *ix) { return ix ? ix->err.c_str() : g_open_error.c_str(); }

int32_t slg_open(int32_t device, slg_index_t **out) {
  if (!out) return fail(nullptr, SLG_ERR_INVALID, "out is NULL");
  *out = nullptr;
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n == 0)
    return fail(nullptr, SLG_ERR_NO_DEVICE, "no CUDA device (%s); this engine has no CPU fallback",
                e != cudaSuccess ? cudaGetErrorString(e) : "device count 0");
  if (device < 0 || device >= n) return fail(nullptr, SLG_ERR_INVALID, "device %d out of range (have %d)", device, n);
  e = cudaSetDevice(device);
  if (e != cudaSuccess) return fail(nullptr, SLG_ERR_CUDA, "cudaSetDevice: %s", cudaGetErrorString(e));
  cudaDeviceProp prop;
  e = cudaGetDeviceProperties(&prop, device);
  if (e != cudaSuccess) return fail(nullptr, SLG_ERR_CUDA, "cudaGetDeviceProperties: %s", cudaGetErrorString(e));
  if (prop.major < 10)
    return fail(nullptr, SLG_ERR_NO_DEVICE, "device %d is sm_%d%d; this library is built for sm_100a only", device, prop.major,
                prop.minor);
  slg_index *ix = new slg_index();
  ix->device = device;
  ix->n_sm = prop.multiProcessorCount;
  ix->smem_optin = prop.sharedMemPerBlockOptin;
  {
    // keep freed per-batch buffers in the pool instead of returning them to the driver
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
      uint64_t keep = ~0ull;
      cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
    }
  }
  e = cudaStreamCreateWithFlags(&ix->stream, cudaStreamNonBlocking);
  for (int i = 0; i < 4 && e == cudaSuccess; i++) e = cudaEventCreate(&ix->ev[i]);
  if (e != cudaSuccess) {
    int32_t rc = fail(nullptr, SLG_ERR_CUDA, "stream/event creation: %s", cudaGetErrorString(e));
    delete ix;
    return rc;
  }
  *out = ix;
  return SLG_OK;
}

int32_t slg_close(slg_index_t *ix) {
  if (!ix) return SLG_OK;
  cudaSetDevice(ix->device);
  cudaStreamSynchronize(ix->stream);
  if (ix->pinned) cudaFreeHost(ix->pinned);
  ix->segs.clear();
  for (auto &ev : ix->ev)
    if (ev) cudaEventDestroy(ev);
  if (ix->stream) cudaStreamDestroy(ix->stream);
  delete ix;
  return SLG_OK;
}

int32_t slg_configure(slg_index_t *ix, uint32_t tile_docs, uint32_t ctas_per_sm, uint32_t sub_docs, uint32_t kernel_choice) {
  if (!ix) return SLG_ERR_INVALID;
  if (sub_docs) {
    if (sub_docs % 128 || sub_docs > 8192) return fail(ix, SLG_ERR_INVALID, "sub_docs must be a multiple of 128 <= 8192");
    ix->sub_docs = sub_docs;
  }
  if ((kernel_choice & 0xFF) > 3) return fail(ix, SLG_ERR_INVALID, "kernel_choice must be 0..3 (+256 = score postings in place)");
  ix->kernel_choice = kernel_choice & 0xFF;
  ix->staging = !(kernel_choice & 256u);
  if (tile_docs) {
    if (tile_docs % 1024 || tile_docs > 49152) return fail(ix, SLG_ERR_INVALID, "tile_docs must be a multiple of 1024 <= 49152");
    ix->tile_docs = tile_docs;
  }
  ix->ctas_per_sm = ctas_per_sm;
  return SLG_OK;
}

int32_t slg_set_option(slg_index_t *ix, const char *name, uint64_t value) {
  if (!ix || !name) return SLG_ERR_INVALID;
  const std::string n(name);
  if (n == "resident_scores") ix->resident_scores = value != 0;
  else if (n == "dense_den") ix->dense_den = (uint32_t)value;
  else if (n == "dense_min_df") ix->dense_min_df = (uint32_t)value;
  else if (n == "max_column_bytes") ix->max_column_bytes = value;
  else if (n == "reg_tile_v") {
    if (value == 4)  // known issue: the 4-wide variant faulted intermittently on the B200 boxes, ~1 batch in 40
      return fail(ix, SLG_ERR_UNSUPPORTED, "reg_tile_v 4 is disabled: intermittent illegal memory access in the 4-wide tile-sweep variant");
    if (value != 8) return fail(ix, SLG_ERR_INVALID, "reg_tile_v must be 8");
    ix->reg_tile_v = (uint32_t)value;
  } else if (n == "sweep_min_postings") ix->sweep_min_postings = value;
  else if (n == "seed_docs") ix->seed_docs = (uint32_t)value;
  else if (n == "part_tiles") ix->part_tiles = (uint32_t)value;
  else if (n == "maxscore_pct") ix->maxscore_pct = (uint32_t)std::min<uint64_t>(value, 100);
  else if (n == "heavy_kernel") {
    if (value > 1) return fail(ix, SLG_ERR_INVALID, "heavy_kernel must be 0 (warp kernel + columns) or 1 (tile sweep)");
    ix->heavy_kernel = (uint32_t)value;
  }
  else if (n == "keep_positions") ix->keep_positions = value != 0;
  else return fail(ix, SLG_ERR_INVALID, "unknown option '%s'", name);
  return SLG_OK;
}

int32_t slg_load_segment(slg_index_t *ix, const slg_segment_view_t *v, float k1, float b) {
  if (!ix || !v) return SLG_ERR_INVALID;
  if (!v->term_offsets || (v->n_terms && (!v->post_docs || !v->post_tfs)))
    return fail(ix, SLG_ERR_INVALID, "segment view lacks postings");
  if (v->doc_count && !v->field_lengths) return fail(ix, SLG_ERR_INVALID, "segment view lacks field lengths");
  if (!ix->term_field.empty())
    return fail(ix, SLG_ERR_INVALID, "this handle's term ids come from the files of field '%s' (slg_term_lookup); caller-assigned ids need their own handle", ix->term_field.c_str());
  SLG_CUDA(ix, cudaSetDevice(ix->device));
  cudaStream_t st = ix->stream;
  auto seg = std::make_unique<Segment>();
  Segment *s = seg.get();
  s->ord = v->segment_ord;
  s->doc_count = v->doc_count;
  s->n_terms = v->n_terms;
  s->k1 = k1;
  s->b = b;
  // CSR offsets on the host (df, idf and the padded layout are host work: 1 M terms)
  std::vector<uint64_t> off(v->n_terms + 1);
  if (v->memory_space == SLG_MEM_DEVICE) {
    SLG_CUDA(ix, cudaMemcpyAsync(off.data(), v->term_offsets, (v->n_terms + 1) * 8, cudaMemcpyDeviceToHost, st));
    SLG_CUDA(ix, cudaStreamSynchronize(st));
  } else {
    std::memcpy(off.data(), v->term_offsets, (v->n_terms + 1) * 8);
  }
  s->h_df.resize(v->n_terms);
  for (uint64_t t = 0; t < v->n_terms; t++) {
    uint64_t df = off[t + 1] - off[t];
    if (off[t + 1] < off[t] || df > 0xFFFFFFFFull) return fail(ix, SLG_ERR_INVALID, "term_offsets not monotone at term %llu", (unsigned long long)t);
    s->h_df[t] = (uint32_t)df;
  }
  int32_t rc = build_layout(ix, s);
  if (rc) return rc;
  const uint64_t n_post = off[v->n_terms] - off[0];
  DevBuf t_off, t_docs, t_tfs, t_lens, t_pres;
  const uint64_t *d_off;
  const uint32_t *d_docs, *d_tfs;
  const int64_t *d_lens;
  const uint8_t *d_pres;
  if ((rc = to_device(ix, v->term_offsets, (size_t)v->n_terms + 1, v->memory_space, t_off, &d_off))) return rc;
  if ((rc = to_device(ix, v->post_docs ? v->post_docs + 0 : nullptr, (size_t)off[v->n_terms], v->memory_space, t_docs, &d_docs))) return rc;
  if ((rc = to_device(ix, v->post_tfs, (size_t)off[v->n_terms], v->memory_space, t_tfs, &d_tfs))) return rc;
  if ((rc = to_device(ix, v->field_lengths, (size_t)v->doc_count, v->memory_space, t_lens, &d_lens))) return rc;
  if ((rc = to_device(ix, v->field_length_present, (size_t)v->doc_count, v->memory_space, t_pres, &d_pres))) return rc;
  (void)n_post;
  if (s->n_blocks) {
    slg_transcode_csr_kernel<<<s->n_blocks, 128, 0, st>>>(d_off, d_docs, d_tfs, s->n_terms, s->term_start.as<uint64_t>(),
                                                          s->term_blk.as<uint32_t>(), s->n_blocks, s->post_doc.as<uint32_t>(),
                                                          s->post_tf.as<uint8_t>(), s->blk_max_doc.as<uint32_t>(),
                                                          s->blk_max_tf.as<float>());
    count_launch(ix);
    SLG_CUDA(ix, cudaGetLastError());
  }
  if ((rc = build_wide(ix, s, d_off, d_tfs))) return rc;
  rc = finish_segment(ix, std::move(seg), d_lens, d_pres, v->total_tokens, v->deleted_docs, v->n_deleted);
  if (rc) return rc;
  SLG_CUDA(ix, cudaStreamSynchronize(st));
  return SLG_OK;
}

namespace {

// exclusive prefix sum of per-slot position counts -> pos_begin[n_post_padded + 1]; allocates pos
int32_t scan_positions(slg_index *ix, Segment *s, DevBuf &npos) {
  cudaStream_t st = ix->stream;
  const uint64_t n = s->n_post_padded + 1;  // npos holds n entries, the last one zero
  if (n > 0x7FFFFFFFull) return fail(ix, SLG_ERR_UNSUPPORTED, "segment has too many posting slots for the position index");
  SLG_CUDA(ix, s->pos_begin.alloc(n * 8));
  size_t tmp_bytes = 0;
  auto in = cub::TransformInputIterator<uint64_t, CastU64, const uint32_t *>(npos.as<uint32_t>(), CastU64());
  SLG_CUDA(ix, cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, in, s->pos_begin.as<uint64_t>(), (int)n, st));
  DevBuf tmp;
  SLG_CUDA(ix, tmp.alloc(tmp_bytes));
  SLG_CUDA(ix, cub::DeviceScan::ExclusiveSum(tmp.p, tmp_bytes, in, s->pos_begin.as<uint64_t>(), (int)n, st));
  count_launch(ix, 2);
  uint64_t total = 0;
  SLG_CUDA(ix, cudaMemcpyAsync(&total, s->pos_begin.as<uint64_t>() + (n - 1), 8, cudaMemcpyDeviceToHost, st));
  SLG_CUDA(ix, cudaStreamSynchronize(st));
  s->n_positions = total;
  SLG_CUDA(ix, s->pos.alloc(std::max<uint64_t>(total, 1) * 4));
  return SLG_OK;
}

// Residency from a `.post` image.  begin[t] = offset of term t's list (UINT64_MAX: the segment lacks the
// term), end[t] = an offset the list does not reach past.  avgdl: the .meta value or nullptr (derive it).
struct FieldInput {  // a further text field: its `_len:` column (host) and its avgdl
  const int64_t *lens;
  const uint8_t *present;
  float avgdl;
};

int32_t load_post_image(slg_index *ix, const slg_segment_view_t *v, const uint8_t *post_image, uint64_t post_image_bytes,
                        const uint64_t *begin, const uint64_t *end, const float *avgdl, float k1, float b,
                        const std::vector<FieldInput> *more_fields = nullptr, const std::vector<uint8_t> *term_field = nullptr) {
  SLG_CUDA(ix, cudaSetDevice(ix->device));
  cudaStream_t st = ix->stream;
  auto seg = std::make_unique<Segment>();
  Segment *s = seg.get();
  s->ord = v->segment_ord;
  s->doc_count = v->doc_count;
  s->n_terms = v->n_terms;
  s->k1 = k1;
  s->b = b;
  if (avgdl) {
    s->avgdl = *avgdl;
    s->avgdl_given = true;
  }
  // parse the fixed part of every list header on the host (index/postings.rs:142-168)
  std::vector<PostTermHeader> hdr(v->n_terms);
  s->h_df.resize(v->n_terms);
  bool any_positions = false;
  for (uint64_t t = 0; t < v->n_terms; t++) {
    const uint64_t o = begin[t];
    if (o == ~0ull) {  // seg.postings(key) == None
      hdr[t] = PostTermHeader{0, 0, 0, 0};
      s->h_df[t] = 0;
      continue;
    }
    if (o + 17 > post_image_bytes) return fail(ix, SLG_ERR_INVALID, "posting header of term %llu is out of bounds", (unsigned long long)t);
    const uint8_t *p = post_image + o;
    uint32_t df, raw_block;
    std::memcpy(&df, p, 4);
    std::memcpy(&raw_block, p + 5, 4);
    uint32_t bc = raw_block & 0x7FFFFFFFu;
    bool has_meta = (raw_block >> 31) != 0;
    uint64_t payload = o + 17 + ((has_meta && bc > 0) ? 4 + 8ull * bc : 0);
    if (payload > post_image_bytes || end[t] > post_image_bytes || end[t] < payload)
      return fail(ix, SLG_ERR_INVALID, "block table of term %llu is out of bounds", (unsigned long long)t);
    hdr[t].payload = payload;
    hdr[t].end = end[t];
    hdr[t].df = df;
    hdr[t].has_positions = p[4] == 1;
    any_positions |= hdr[t].has_positions && df > 0;
    s->h_df[t] = df;
  }
  StageTimer tm(st);
  tm.mark("list headers (host)");
  int32_t rc = build_layout(ix, s);
  if (rc) return rc;
  tm.mark("layout tables");
  const bool keep_pos = any_positions && ix->keep_positions;
  DevBuf d_img, d_hdr, t_lens, t_pres, d_npos, d_posbyte;
  SLG_CUDA(ix, d_img.alloc(post_image_bytes + 16));
  SLG_CUDA(ix, cudaMemcpyAsync(d_img.p, post_image, post_image_bytes, cudaMemcpyHostToDevice, st));
  SLG_CUDA(ix, d_hdr.alloc(std::max<uint64_t>(v->n_terms, 1) * sizeof(PostTermHeader)));
  if (v->n_terms) SLG_CUDA(ix, cudaMemcpyAsync(d_hdr.p, hdr.data(), v->n_terms * sizeof(PostTermHeader), cudaMemcpyHostToDevice, st));
  if (keep_pos) {
    SLG_CUDA(ix, d_npos.alloc((s->n_post_padded + 1) * 4));
    SLG_CUDA(ix, d_posbyte.alloc((s->n_post_padded + 1) * 4));
    SLG_CUDA(ix, cudaMemsetAsync(d_npos.p, 0, (s->n_post_padded + 1) * 4, st));
  }
  DevBuf d_err;
  SLG_CUDA(ix, d_err.alloc(4));
  SLG_CUDA(ix, cudaMemsetAsync(d_err.p, 0, 4, st));
  tm.mark("image -> device");
  if (v->n_terms) {
    slg_decode_post_image_kernel<<<(unsigned)((v->n_terms + 3) / 4), 128, 0, st>>>(
        d_img.as<uint8_t>(), post_image_bytes, d_hdr.as<PostTermHeader>(), v->n_terms, s->term_start.as<uint64_t>(),
        s->term_blk.as<uint32_t>(), s->post_doc.as<uint32_t>(), s->post_tf.as<uint8_t>(), s->blk_max_doc.as<uint32_t>(),
        s->blk_max_tf.as<float>(), keep_pos ? d_npos.as<uint32_t>() : nullptr, keep_pos ? d_posbyte.as<uint32_t>() : nullptr,
        d_err.as<uint32_t>());
    count_launch(ix);
    SLG_CUDA(ix, cudaGetLastError());
  }
  uint32_t derr = 0;
  SLG_CUDA(ix, cudaMemcpyAsync(&derr, d_err.p, 4, cudaMemcpyDeviceToHost, st));
  SLG_CUDA(ix, cudaStreamSynchronize(st));
  if (derr == 1) return fail(ix, SLG_ERR_INVALID, "malformed varint in the posting image");
  if (derr == 2) return fail(ix, SLG_ERR_UNSUPPORTED, "term frequency >= 255 in a posting image (use the CSR load path)");
  if (derr == 3) return fail(ix, SLG_ERR_UNSUPPORTED, "a posting list with positions is longer than 4 GiB");
  tm.mark("decode docs + tfs");
  if (keep_pos) {
    if ((rc = scan_positions(ix, s, d_npos))) return rc;
    if (s->n_blocks) {
      slg_decode_positions_kernel<<<s->n_blocks, 128, 0, st>>>(d_img.as<uint8_t>(), d_hdr.as<PostTermHeader>(), s->n_terms,
                                                               s->term_start.as<uint64_t>(), s->term_blk.as<uint32_t>(),
                                                               s->term_df.as<uint32_t>(), s->n_blocks, d_posbyte.as<uint32_t>(),
                                                               s->pos_begin.as<uint64_t>(), s->pos.as<uint32_t>(), d_err.as<uint32_t>());
      count_launch(ix);
      SLG_CUDA(ix, cudaGetLastError());
      SLG_CUDA(ix, cudaMemcpyAsync(&derr, d_err.p, 4, cudaMemcpyDeviceToHost, st));
      SLG_CUDA(ix, cudaStreamSynchronize(st));
      if (derr) return fail(ix, SLG_ERR_INVALID, "malformed position varint in the posting image");
    }
    s->has_positions = true;
    tm.mark("decode positions");
  }
  if ((rc = build_wide(ix, s, nullptr, nullptr))) return rc;
  const int64_t *d_lens;
  const uint8_t *d_pres;
  if ((rc = to_device(ix, v->field_lengths, (size_t)v->doc_count, SLG_MEM_HOST, t_lens, &d_lens))) return rc;
  if ((rc = to_device(ix, v->field_length_present, (size_t)v->doc_count, SLG_MEM_HOST, t_pres, &d_pres))) return rc;
  if (more_fields && !more_fields->empty()) {
    for (const FieldInput &fi : *more_fields) {
      Segment::ExtraField ef;
      ef.avgdl = fi.avgdl;
      SLG_CUDA(ix, ef.d_lens.alloc(std::max<size_t>(v->doc_count, 1) * 8));
      SLG_CUDA(ix, ef.d_present.alloc(std::max<size_t>(v->doc_count, 1)));
      if (v->doc_count) {
        SLG_CUDA(ix, cudaMemcpyAsync(ef.d_lens.p, fi.lens, (size_t)v->doc_count * 8, cudaMemcpyHostToDevice, st));
        SLG_CUDA(ix, cudaMemcpyAsync(ef.d_present.p, fi.present, v->doc_count, cudaMemcpyHostToDevice, st));
      }
      s->extra_fields.push_back(std::move(ef));
    }
    if (term_field) s->h_term_field = *term_field;
  }
  rc = finish_segment(ix, std::move(seg), d_lens, d_pres, v->total_tokens, v->deleted_docs, v->n_deleted);
  if (rc) return rc;
  SLG_CUDA(ix, cudaStreamSynchronize(st));
  tm.mark("norms, scores, columns");
  return SLG_OK;
}

}  // namespace

int32_t slg_load_segment_post_image(slg_index_t *ix, const slg_segment_view_t *v, const uint8_t *post_image,
                                    uint64_t post_image_bytes, const uint64_t *term_post_offsets, float k1, float b) {
  if (!ix || !v || !post_image || !term_post_offsets) return SLG_ERR_INVALID;
  if (v->doc_count && !v->field_lengths) return fail(ix, SLG_ERR_INVALID, "segment view lacks field lengths");
  if (v->memory_space != SLG_MEM_HOST) return fail(ix, SLG_ERR_INVALID, "post image loads take host memory");
  return load_post_image(ix, v, post_image, post_image_bytes, term_post_offsets, term_post_offsets + 1, nullptr, k1, b);
}

/* ---- residency from the reference's segment files (SegmentReader::open, index/segment.rs:1239-1330) ---- */
namespace {

struct ParsedSegmentFiles {
  std::vector<slgf::TermEntry> terms;      // every key of .terms
  std::vector<slgf::FastColumn> fast;      // every column of .fast
  std::vector<std::string> fields;         // the scored text fields, in the caller's order
  std::vector<const slgf::FastColumn *> len_col;  // `_len:<field>` per field (or null)
  std::vector<float> avgdl;                // per field
  std::vector<int32_t> term_field;         // per entry of `terms`: index into fields or -1
  uint64_t n_field_terms = 0;
  uint64_t df_sum = 0;
  bool any_positions = false;
};

// "body" or "title,body": the text fields a handle scores
std::vector<std::string> split_fields(const char *spec) {
  std::vector<std::string> out;
  std::string cur;
  for (const char *p = spec;; p++) {
    if (*p == ',' || *p == '\0') {
      if (!cur.empty()) out.push_back(cur);
      cur.clear();
      if (!*p) break;
    } else {
      cur.push_back(*p);
    }
  }
  return out;
}

// verify_checksums + read_terms + FastFieldsReader::open + the .meta fields the search path reads
bool parse_segment_files(const slg_segment_files_t *f, const char *field_spec, ParsedSegmentFiles &out, std::string &err) {
  if (!f->terms || !f->post || !f->fast || !f->meta) {
    err = "segment files: terms, post, fast and meta images are all required";
    return false;
  }
  out.fields = split_fields(field_spec);
  if (out.fields.empty()) {
    err = "no text field named";
    return false;
  }
  if (out.fields.size() > kMaxFields) {
    err = "more than " + std::to_string(kMaxFields) + " text fields in one handle";
    return false;
  }
  if (f->checksums) {  // SegmentMeta.checksums, index/segment.rs:1140-1200
    const char *label[4] = {"terms", "postings", "fast", "meta"};
    const uint8_t *img[4] = {f->terms, f->post, f->fast, f->meta};
    const uint64_t len[4] = {f->terms_bytes, f->post_bytes, f->fast_bytes, f->meta_bytes};
    for (int i = 0; i < 4; i++) {
      const uint32_t actual = slgf::crc32_parallel(img[i], len[i]);
      if (actual != f->checksums[i]) {
        err = std::string("segment failed checksum for ") + label[i] + " (expected " + std::to_string(f->checksums[i]) +
              ", found " + std::to_string(actual) + ")";
        return false;
      }
    }
  }
  if (!slgf::parse_terms(f->terms, f->terms_bytes, out.terms, err)) return false;
  if (!slgf::parse_fast(f->fast, f->fast_bytes, out.fast, err)) return false;
  slgf::Json root = slgf::json_root(f->meta, f->meta_bytes);
  if (root.kind() != '{') {
    err = "segment meta is not a JSON object";
    return false;
  }
  for (auto &c : out.fast)
    if (c.type <= 2 && c.doc_len != f->doc_count) {
      err = "fast-field column '" + c.name + "' has " + std::to_string(c.doc_len) + " rows, the segment " + std::to_string(f->doc_count) + " docs";
      return false;
    }
  const slgf::Json avgs = slgf::json_get(root, "avg_field_lengths");
  for (auto &field : out.fields) {
    const std::string len_key = "_len:" + field;  // doc_length_key, index/fastfields.rs:1162-1164
    const slgf::FastColumn *lc = nullptr;
    for (auto &c : out.fast)
      if (c.name == len_key && c.type == 0) lc = &c;
    out.len_col.push_back(lc);
    // SegmentReader::avg_field_length, index/segment.rs:1344-1351: missing field => 0.0; serde_json reads an f32 as f64 -> f32
    out.avgdl.push_back((float)slgf::json_number(slgf::json_get(avgs, field.c_str()), 0.0));
  }
  out.term_field.assign(out.terms.size(), -1);
  for (size_t i = 0; i < out.terms.size(); i++) {
    const slgf::TermEntry &t = out.terms[i];
    if (t.offset + 17 > f->post_bytes) {
      err = "a term's posting offset lies outside the posting file";
      return false;
    }
    // keys are "<field>:<token>" (index/segment.rs:675-679); field names hold no ':' in the reference's schemas,
    // so the field is the text before the first colon
    const char *colon = static_cast<const char *>(std::memchr(t.key, ':', t.key_len));
    if (!colon) continue;
    const size_t fl = (size_t)(colon - t.key);
    for (size_t fi = 0; fi < out.fields.size(); fi++)
      if (out.fields[fi].size() == fl && std::memcmp(t.key, out.fields[fi].data(), fl) == 0) {
        out.term_field[i] = (int32_t)fi;
        out.n_field_terms++;
        uint32_t df;
        std::memcpy(&df, f->post + t.offset, 4);
        out.df_sum += df;
        out.any_positions |= f->post[t.offset + 4] == 1;
        break;
      }
  }
  return true;
}

}  // namespace

int32_t slg_inspect_segment_files(const slg_segment_files_t *files, const char *field, slg_segment_info_t *out, char *err,
                                  uint64_t err_cap) {
  if (!files || !field || !out) return SLG_ERR_INVALID;
  ParsedSegmentFiles ps;
  std::string e;
  if (!parse_segment_files(files, field, ps, e)) {
    if (err && err_cap) snprintf(err, (size_t)err_cap, "%s", e.c_str());
    return SLG_ERR_INVALID;
  }
  std::memset(out, 0, sizeof(*out));
  out->n_terms_total = ps.terms.size();
  out->n_terms_field = ps.n_field_terms;
  out->n_postings = ps.df_sum;
  out->avgdl = ps.avgdl[0];
  out->has_positions = ps.any_positions;
  out->has_length_column = ps.len_col[0] != nullptr;
  out->n_fast_columns = (uint32_t)ps.fast.size();
  for (auto &c : ps.fast) out->n_scalar_columns += c.type <= 2;
  out->crc_terms = slgf::crc32_parallel(files->terms, files->terms_bytes);
  out->crc_postings = slgf::crc32_parallel(files->post, files->post_bytes);
  out->crc_fast = slgf::crc32_parallel(files->fast, files->fast_bytes);
  out->crc_meta = slgf::crc32_parallel(files->meta, files->meta_bytes);
  return SLG_OK;
}

int32_t slg_load_segment_files(slg_index_t *ix, const slg_segment_files_t *f, const char *field, float k1, float b) {
  if (!ix || !f || !field) return SLG_ERR_INVALID;
  if (!ix->term_field.empty() && ix->term_field != field)
    return fail(ix, SLG_ERR_UNSUPPORTED, "this handle scores field(s) '%s'; every segment of a handle names the same field list", ix->term_field.c_str());
  if (ix->term_field.empty() && !ix->segs.empty())
    return fail(ix, SLG_ERR_INVALID, "this handle holds segments loaded with caller-assigned term ids; file segments need their own handle");
  ParsedSegmentFiles ps;
  std::string e;
  StageTimer tm(ix->stream);
  if (!parse_segment_files(f, field, ps, e)) return fail(ix, SLG_ERR_INVALID, "%s", e.c_str());
  tm.mark("crc32 + parse files (host)");
  // every list's end: the next list's offset in file order (lists of all fields share the file)
  std::vector<uint64_t> all_off;
  all_off.reserve(ps.terms.size() + 1);
  for (auto &t : ps.terms) all_off.push_back(t.offset);
  all_off.push_back(f->post_bytes);
  std::sort(all_off.begin(), all_off.end());
  // the handle's term space grows by the keys this segment adds
  struct Mine {
    uint32_t id;
    uint64_t offset;
    uint8_t field;
  };
  std::vector<Mine> mine;
  mine.reserve((size_t)ps.n_field_terms);
  for (size_t i = 0; i < ps.terms.size(); i++) {
    if (ps.term_field[i] < 0) continue;
    const slgf::TermEntry &t = ps.terms[i];
    auto it = ix->term_ids.emplace(std::string(t.key, t.key_len), (uint32_t)ix->term_ids.size()).first;
    if (it->second >= ix->term_field_of.size()) ix->term_field_of.resize(it->second + 1, 0);
    ix->term_field_of[it->second] = (uint8_t)ps.term_field[i];
    mine.push_back(Mine{it->second, t.offset, (uint8_t)ps.term_field[i]});
  }
  ix->term_field = field;
  const uint64_t n_terms = ix->term_ids.size();
  std::vector<uint64_t> begin(n_terms, ~0ull), end(n_terms, 0);
  for (auto &m : mine) {
    begin[m.id] = m.offset;
    end[m.id] = *std::upper_bound(all_off.begin(), all_off.end(), m.offset);
  }
  // `_len:<field>` per field (field_lengths_for, api/reader.rs:3604-3621: absent column or value => 0)
  const size_t n_fields = ps.fields.size();
  std::vector<std::vector<int64_t>> lens(n_fields, std::vector<int64_t>(f->doc_count, 0));
  std::vector<std::vector<uint8_t>> pres(n_fields, std::vector<uint8_t>(f->doc_count, 0));
  uint64_t total = 0;
  for (size_t fi = 0; fi < n_fields; fi++) {
    if (!ps.len_col[fi]) continue;
    std::memcpy(lens[fi].data(), ps.len_col[fi]->values, (size_t)f->doc_count * 8);
    std::memcpy(pres[fi].data(), ps.len_col[fi]->presence, f->doc_count);
    if (fi == 0)
      for (uint32_t d = 0; d < f->doc_count; d++)
        if (pres[0][d] && lens[0][d] > 0) total += (uint64_t)lens[0][d];
  }
  tm.mark("term space + lengths (host)");
  std::vector<FieldInput> more;
  for (size_t fi = 1; fi < n_fields; fi++) more.push_back(FieldInput{lens[fi].data(), pres[fi].data(), ps.avgdl[fi]});
  std::vector<uint8_t> term_field(ix->term_field_of.begin(), ix->term_field_of.end());
  term_field.resize(n_terms, 0);
  slg_segment_view_t v{};
  v.segment_ord = f->segment_ord;
  v.doc_count = f->doc_count;
  v.n_terms = n_terms;
  v.field_lengths = lens[0].data();
  v.field_length_present = pres[0].data();
  v.total_tokens = total;
  v.deleted_docs = f->deleted_docs;
  v.n_deleted = f->n_deleted;
  v.memory_space = SLG_MEM_HOST;
  int32_t rc = load_post_image(ix, &v, f->post, f->post_bytes, begin.data(), end.data(), &ps.avgdl[0], k1, b, &more, &term_field);
  if (rc) {
    if (ix->segs.empty()) {  // nothing loaded yet: a failed first load leaves no term space behind
      ix->term_ids.clear();
      ix->term_field_of.clear();
      ix->term_field.clear();
    }
    return rc;
  }
  // scalar fast-field columns, by name (the file's field order is HashMap order, index/fastfields.rs:414)
  Segment *s = ix->find(f->segment_ord);
  for (auto &c : ps.fast) {
    if (c.type > 2 || c.name.compare(0, 5, "_len:") == 0) continue;
    size_t h = 0;
    while (h < ix->column_names.size() && ix->column_names[h] != c.name) h++;
    if (h == ix->column_names.size()) ix->column_names.push_back(c.name);
    if (s->columns.size() <= h) s->columns.resize(h + 1);
    Column col;
    col.kind = c.type;
    const size_t elem = c.type == 2 ? 4 : 8;
    SLG_CUDA(ix, col.values.alloc(std::max<size_t>((size_t)s->doc_count * elem, 1)));
    SLG_CUDA(ix, cudaMemcpy(col.values.p, c.values, (size_t)s->doc_count * elem, cudaMemcpyHostToDevice));
    if (c.type != 2) {
      SLG_CUDA(ix, col.present.alloc(std::max<size_t>(s->doc_count, 1)));
      SLG_CUDA(ix, cudaMemcpy(col.present.p, c.presence, s->doc_count, cudaMemcpyHostToDevice));
    }
    col.dict = c.dict;
    s->columns[h] = std::move(col);
  }
  return SLG_OK;
}

int32_t slg_load_vector_file(slg_index_t *ix, uint32_t segment_ord, const uint8_t *bytes, uint64_t n_bytes, int32_t store_bf16,
                             int32_t *metric_out) {
  if (!ix || !bytes) return SLG_ERR_INVALID;
  Segment *s = ix->find(segment_ord);
  if (!s) return fail(ix, SLG_ERR_INVALID, "no segment %u", segment_ord);
  slgf::VectorFile vf;
  std::string e;
  if (!slgf::parse_vector_file(bytes, n_bytes, vf, e)) return fail(ix, SLG_ERR_INVALID, "%s", e.c_str());
  if (vf.doc_count != s->doc_count)
    return fail(ix, SLG_ERR_INVALID, "vector doc count mismatch: expected %u, found %u", s->doc_count, vf.doc_count);
  if (metric_out) *metric_out = vf.metric;
  // slg_load_vectors only copies from these pointers (no host dereference), so the image needs no alignment and no staging
  // copy — the rows of a 100 M-doc shard are 19 GB
  return slg_load_vectors(ix, segment_ord, vf.dim, reinterpret_cast<const uint32_t *>(vf.offsets),
                          reinterpret_cast<const float *>(vf.values), vf.vector_count, store_bf16);
}

namespace {
// read-only mapping of a file: the `.post` of a 10 M-doc index is ~12 GB, so no private copy is made —
// crc32 and the host-to-device copy read the page cache directly
struct MappedFile {
  const uint8_t *p = nullptr;
  size_t n = 0;
  MappedFile() = default;
  MappedFile(const MappedFile &) = delete;
  MappedFile &operator=(const MappedFile &) = delete;
  ~MappedFile() {
    if (p && n) munmap(const_cast<uint8_t *>(p), n);
  }
  bool open(const std::string &path) {
    const int fd = ::open(path.c_str(), O_RDONLY);
    if (fd < 0) return false;
    struct stat st;
    if (fstat(fd, &st) != 0) {
      ::close(fd);
      return false;
    }
    n = (size_t)st.st_size;
    if (n) {
      void *m = mmap(nullptr, n, PROT_READ, MAP_PRIVATE, fd, 0);
      if (m == MAP_FAILED) {
        ::close(fd);
        n = 0;
        return false;
      }
      madvise(m, n, MADV_SEQUENTIAL);
      p = static_cast<const uint8_t *>(m);
    } else {
      static const uint8_t empty = 0;
      p = &empty;
    }
    ::close(fd);
    return true;
  }
  const uint8_t *data() const { return p; }
  size_t size() const { return n; }
};
// the manifest stores `root.join(name)` strings (index/directory.rs:16-46); the index may have moved since
std::string in_dir(const std::string &dir, const std::string &stored) {
  size_t slash = stored.find_last_of('/');
  return dir + "/" + (slash == std::string::npos ? stored : stored.substr(slash + 1));
}
}  // namespace

int32_t slg_load_index_dir(slg_index_t *ix, const char *dir, const char *field, float k1, float b, const char *vector_field,
                           int32_t store_bf16, uint32_t *n_segments_out) {
  return slg_load_index_dir_shard(ix, dir, field, k1, b, vector_field, store_bf16, 0, 1, n_segments_out);
}

int32_t slg_load_index_dir_shard(slg_index_t *ix, const char *dir, const char *field, float k1, float b, const char *vector_field,
                                 int32_t store_bf16, uint32_t shard_rank, uint32_t shard_world, uint32_t *n_segments_out) {
  if (!ix || !dir || !field) return SLG_ERR_INVALID;
  if (shard_world == 0 || shard_rank >= shard_world) return fail(ix, SLG_ERR_INVALID, "shard %u of %u", shard_rank, shard_world);
  MappedFile man;
  const std::string d(dir);
  if (!man.open(d + "/MANIFEST.json")) return fail(ix, SLG_ERR_INVALID, "cannot read %s/MANIFEST.json", dir);
  slgf::Json root = slgf::json_root(man.data(), man.size());
  std::vector<slgf::Json> segs;
  if (!slgf::json_elements(slgf::json_get(root, "segments"), segs)) return fail(ix, SLG_ERR_INVALID, "manifest has no segments array");
  uint32_t ord = 0, loaded = 0;
  for (auto &sm : segs) {  // IndexReader::open keeps manifest order; segment_ord is that index (api/reader.rs:2670)
    if (ord % shard_world != shard_rank) {  // another GPU's segment (segment == shard)
      ord++;
      continue;
    }
    slgf::Json paths = slgf::json_get(sm, "paths");
    MappedFile terms, post, fast, meta;
    const char *names[4] = {"terms", "postings", "fast", "meta"};
    MappedFile *bufs[4] = {&terms, &post, &fast, &meta};
    uint32_t crcs[4];
    bool have_crc = true;
    slgf::Json sums = slgf::json_get(sm, "checksums");
    for (int i = 0; i < 4; i++) {
      const std::string stored = slgf::json_string(slgf::json_get(paths, names[i]));
      if (stored.empty()) return fail(ix, SLG_ERR_INVALID, "segment %u: manifest lacks paths.%s", ord, names[i]);
      if (!bufs[i]->open(in_dir(d, stored))) return fail(ix, SLG_ERR_INVALID, "cannot read %s", in_dir(d, stored).c_str());
      slgf::Json c = slgf::json_get(sums, names[i]);
      if (c.ok()) crcs[i] = (uint32_t)slgf::json_number(c);
      else have_crc = false;
    }
    std::vector<uint32_t> deleted;
    std::vector<slgf::Json> del;
    slgf::json_elements(slgf::json_get(sm, "deleted_docs"), del);
    for (auto &x : del) deleted.push_back((uint32_t)slgf::json_number(x));
    slg_segment_files_t f{};
    f.segment_ord = ord;
    f.doc_count = (uint32_t)slgf::json_number(slgf::json_get(sm, "doc_count"));
    f.terms = terms.data();
    f.terms_bytes = terms.size();
    f.post = post.data();
    f.post_bytes = post.size();
    f.fast = fast.data();
    f.fast_bytes = fast.size();
    f.meta = meta.data();
    f.meta_bytes = meta.size();
    f.deleted_docs = deleted.data();
    f.n_deleted = (uint32_t)deleted.size();
    f.checksums = have_crc ? crcs : nullptr;
    int32_t rc = slg_load_segment_files(ix, &f, field, k1, b);
    if (rc) return rc;
    if (vector_field && *vector_field) {
      const std::string vdir = slgf::json_string(slgf::json_get(paths, "vector_dir"));
      if (vdir.empty()) return fail(ix, SLG_ERR_INVALID, "segment missing vector directory path");  // segment.rs:969-972
      MappedFile vb;
      const std::string vp = in_dir(d, vdir) + "/" + vector_field + ".bin";
      if (!vb.open(vp)) return fail(ix, SLG_ERR_INVALID, "cannot read %s", vp.c_str());
      if ((rc = slg_load_vector_file(ix, ord, vb.data(), vb.size(), store_bf16, nullptr))) return rc;
    }
    ord++;
    loaded++;
  }
  if (n_segments_out) *n_segments_out = loaded;
  return SLG_OK;
}

int32_t slg_term_lookup(const slg_index_t *ix, const char *key, uint32_t *term_id) {
  if (!ix || !key || !term_id) return SLG_ERR_INVALID;
  auto it = ix->term_ids.find(key);
  *term_id = it == ix->term_ids.end() ? 0xFFFFFFFFu : it->second;
  return SLG_OK;
}

int32_t slg_column_lookup(const slg_index_t *ix, const char *name) {
  if (!ix || !name) return SLG_ERR_INVALID;
  for (size_t h = 0; h < ix->column_names.size(); h++)
    if (ix->column_names[h] == name) return (int32_t)h;
  return -1;
}

/* ---- term positions handed over as CSR (the positions of PostingEntry, index/postings.rs:14-19) ---- */
int32_t slg_load_positions(slg_index_t *ix, uint32_t segment_ord, const uint64_t *term_offsets, const uint64_t *position_offsets,
                           const uint32_t *positions, int32_t memory_space) {
  if (!ix || !term_offsets || !position_offsets) return SLG_ERR_INVALID;
  SLG_CUDA(ix, cudaSetDevice(ix->device));
  cudaStream_t st = ix->stream;
  Segment *s = ix->find(segment_ord);
  if (!s) return fail(ix, SLG_ERR_INVALID, "no segment %u", segment_ord);
  const bool dev = memory_space == SLG_MEM_DEVICE;
  std::vector<uint64_t> off(s->n_terms + 1);
  if (dev) SLG_CUDA(ix, cudaMemcpy(off.data(), term_offsets, (s->n_terms + 1) * 8, cudaMemcpyDeviceToHost));
  else std::memcpy(off.data(), term_offsets, (s->n_terms + 1) * 8);
  for (uint64_t t = 0; t < s->n_terms; t++)
    if (off[t + 1] - off[t] != s->h_df[t])
      return fail(ix, SLG_ERR_INVALID, "term offsets do not match the loaded postings at term %llu", (unsigned long long)t);
  const uint64_t n_post = off[s->n_terms];
  uint64_t n_pos = 0;
  if (dev) SLG_CUDA(ix, cudaMemcpy(&n_pos, position_offsets + n_post, 8, cudaMemcpyDeviceToHost));
  else n_pos = position_offsets[n_post];
  if (n_pos && !positions) return SLG_ERR_INVALID;
  ix->ctr.resident_bytes -= s->pos_begin.bytes + s->pos.bytes;
  s->has_positions = false;
  DevBuf t_off, t_poff, t_pos, d_npos;
  const uint64_t *d_off, *d_poff;
  const uint32_t *d_pos;
  int32_t rc;
  if ((rc = to_device(ix, term_offsets, (size_t)s->n_terms + 1, memory_space, t_off, &d_off))) return rc;
  if ((rc = to_device(ix, position_offsets, (size_t)n_post + 1, memory_space, t_poff, &d_poff))) return rc;
  if ((rc = to_device(ix, positions, (size_t)n_pos, memory_space, t_pos, &d_pos))) return rc;
  SLG_CUDA(ix, d_npos.alloc((s->n_post_padded + 1) * 4));
  SLG_CUDA(ix, cudaMemsetAsync(d_npos.p, 0, (s->n_post_padded + 1) * 4, st));
  if (s->n_blocks) {
    slg_csr_position_counts_kernel<<<s->n_blocks, 128, 0, st>>>(d_off, d_poff, s->n_terms, s->term_start.as<uint64_t>(),
                                                                s->term_blk.as<uint32_t>(), s->n_blocks, d_npos.as<uint32_t>());
    count_launch(ix);
  }
  if ((rc = scan_positions(ix, s, d_npos))) return rc;
  if (s->n_positions != n_pos) return fail(ix, SLG_ERR_INVALID, "position offsets are inconsistent");
  if (s->n_blocks) {
    slg_csr_position_copy_kernel<<<s->n_blocks, 128, 0, st>>>(d_off, d_poff, d_pos, s->n_terms, s->term_start.as<uint64_t>(), s->term_blk.as<uint32_t>(),
                                                              s->n_blocks, s->pos_begin.as<uint64_t>(), s->pos.as<uint32_t>());
    count_launch(ix);
  }
  SLG_CUDA(ix, cudaGetLastError());
  SLG_CUDA(ix, cudaStreamSynchronize(st));
  s->has_positions = true;
  ix->ctr.resident_bytes += s->pos_begin.bytes + s->pos.bytes;
  return SLG_OK;
}

int32_t slg_segment_stats(const slg_index_t *ixc, uint32_t segment_ord, float *avgdl, float *live_docs, float *min_doc_len,
                          uint64_t *n_postings) {
  slg_index *ix = const_cast<slg_index *>(ixc);
  if (!ix) return SLG_ERR_INVALID;
  Segment *s = ix->find(segment_ord);
  if (!s) return fail(ix, SLG_ERR_INVALID, "no segment %u", segment_ord);
  if (avgdl) *avgdl = s->avgdl;
  if (live_docs) *live_docs = s->live_docs;
  if (min_doc_len) *min_doc_len = s->min_doc_len;
  if (n_postings) *n_postings = s->n_postings;
  return SLG_OK;
}

int32_t slg_field_stats(const slg_index_t *ixc, uint32_t segment_ord, uint32_t field_index, float *avgdl, float *min_doc_len) {
  slg_index *ix = const_cast<slg_index *>(ixc);
  if (!ix) return SLG_ERR_INVALID;
  Segment *s = ix->find(segment_ord);
  if (!s) return fail(ix, SLG_ERR_INVALID, "no segment %u", segment_ord);
  if (field_index >= s->f_avgdl.size()) return fail(ix, SLG_ERR_INVALID, "segment %u scores %zu field(s)", segment_ord, s->f_avgdl.size());
  if (avgdl) *avgdl = s->f_avgdl[field_index];
  if (min_doc_len) *min_doc_len = s->f_min_len[field_index];
  return SLG_OK;
}

int32_t slg_term_has_column(const slg_index_t *ixc, uint32_t segment_ord, uint32_t term_id) {
  slg_index *ix = const_cast<slg_index *>(ixc);
  if (!ix) return SLG_ERR_INVALID;
  Segment *s = ix->find(segment_ord);
  if (!s) return fail(ix, SLG_ERR_INVALID, "no segment %u", segment_ord);
  return term_id < s->h_term_col.size() && s->h_term_col[term_id] >= 0 ? 1 : 0;
}

/* ---- fast-field columns + filters ---- */
static int32_t add_column(slg_index *ix, uint32_t segment_ord, int kind, const void *values, size_t elem,
                          const uint8_t *present, const char *const *dict, uint32_t n_dict) {
  if (!ix || !values) return SLG_ERR_INVALID;
  SLG_CUDA(ix, cudaSetDevice(ix->device));
  Segment *s = ix->find(segment_ord);
  if (!s) return fail(ix, SLG_ERR_INVALID, "no segment %u", segment_ord);
  Column c;
  c.kind = kind;
  SLG_CUDA(ix, c.values.alloc((size_t)s->doc_count * elem));
  SLG_CUDA(ix, cudaMemcpy(c.values.p, values, (size_t)s->doc_count * elem, cudaMemcpyHostToDevice));
  if (kind != 2) {
    SLG_CUDA(ix, c.present.alloc(std::max<size_t>(s->doc_count, 1)));
    if (present) SLG_CUDA(ix, cudaMemcpy(c.present.p, present, s->doc_count, cudaMemcpyHostToDevice));
    else SLG_CUDA(ix, cudaMemset(c.present.p, 1, std::max<size_t>(s->doc_count, 1)));
  }
  for (uint32_t i = 0; i < n_dict; i++) c.dict.emplace_back(dict[i]);
  s->columns.push_back(std::move(c));
  return (int32_t)s->columns.size() - 1;
}

int32_t slg_add_i64_column(slg_index_t *ix, uint32_t segment_ord, const int64_t *values, const uint8_t *present) {
  return add_column(ix, segment_ord, 0, values, 8, present, nullptr, 0);
}
int32_t slg_add_f64_column(slg_index_t *ix, uint32_t segment_ord, const double *values, const uint8_t *present) {
  return add_column(ix, segment_ord, 1, values, 8, present, nullptr, 0);
}
int32_t slg_add_str_column(slg_index_t *ix, uint32_t segment_ord, const char *const *dict, uint32_t n_dict,
                           const uint32_t *ords) {
  if (n_dict && !dict) return SLG_ERR_INVALID;
  return add_column(ix, segment_ord, 2, ords, 4, nullptr, dict, n_dict);
}

// index/fastfields.rs:475-481, ASCII path
static bool ci_equals(const std::string &a, const std::string &b) {
  if (a.size() != b.size()) return false;
  for (size_t i = 0; i < a.size(); i++) {
    unsigned char x = a[i], y = b[i];
    if (x >= 'A' && x <= 'Z') x += 32;
    if (y >= 'A' && y <= 'Z') y += 32;
    if (x != y) return false;
  }
  return true;
}

// A filter id names one bitmap per loaded segment (root filters, phrases and their combinations alike).
static int32_t register_filter(slg_index *ix, FilterProg fp, std::vector<DevBuf> &per_seg) {
  const size_t id = ix->filters.size();
  for (size_t si = 0; si < ix->segs.size(); si++) {
    Segment *s = ix->segs[si].get();
    s->filter_bits.resize(id + 1);
    s->filter_slabs.resize(id + 1);
    s->filter_bits[id] = std::move(per_seg[si]);
    std::vector<const uint32_t *> ptrs;
    for (auto &fb : s->filter_bits) ptrs.push_back(fb.as<uint32_t>());
    if (s->filter_ptrs.bytes < ptrs.size() * sizeof(void *)) {
      SLG_CUDA(ix, cudaStreamSynchronize(ix->stream));  // a running batch may still read the old table
      SLG_CUDA(ix, s->filter_ptrs.alloc(std::max<size_t>(64, ptrs.size() * 2) * sizeof(void *)));
    }
    SLG_CUDA(ix, cudaMemcpy(s->filter_ptrs.p, ptrs.data(), ptrs.size() * sizeof(void *), cudaMemcpyHostToDevice));
  }
  ix->filters.push_back(std::move(fp));
  return (int32_t)id;
}

static int32_t compile_filter_for_segment(slg_index *ix, Segment *s, const FilterProg &fp, DevBuf &bits_out) {
  // Resolve keyword predicates to dictionary-ordinal sets on the host (string compares happen once
  // per dictionary entry instead of once per doc, index/fastfields.rs:490-530), then evaluate the
  // program for every doc on the device.
  std::vector<FilterNodeDev> nodes(fp.nodes.size());
  std::vector<uint32_t> ordset;  // concatenated bitsets over dictionaries
  for (size_t i = 0; i < fp.nodes.size(); i++) {
    const slg_filter_node_t &n = fp.nodes[i];
    FilterNodeDev &d = nodes[i];
    d.op = n.op;
    d.n_children = n.n_children;
    d.i_min = n.i_min;
    d.i_max = n.i_max;
    d.f_min = n.f_min;
    d.f_max = n.f_max;
    d.values = nullptr;
    d.present = nullptr;
    d.set_off = 0;
    d.set_words = 0;
    const Column *c = (n.column >= 0 && (size_t)n.column < s->columns.size()) ? &s->columns[n.column] : nullptr;
    bool leaf = n.op <= SLG_F_F64_RANGE;
    if (!leaf) continue;
    int want = (n.op == SLG_F_I64_RANGE) ? 0 : (n.op == SLG_F_F64_RANGE ? 1 : 2);
    if (!c || c->kind != want) {
      d.op = FOP_FALSE;  // unknown field or wrong column type: predicate is false (fastfields.rs `_ => false`)
      continue;
    }
    d.values = c->values.p;
    d.present = c->present.as<uint8_t>();
    if (want == 2) {
      uint32_t words = ((uint32_t)c->dict.size() + 31) / 32;
      d.set_off = (uint32_t)ordset.size();
      d.set_words = words;
      ordset.resize(ordset.size() + words, 0u);
      if (n.value_end < n.value_begin || n.value_end > fp.strings.size()) return fail(ix, SLG_ERR_INVALID, "filter value range out of bounds");
      for (uint32_t o = 0; o < c->dict.size(); o++)
        for (uint32_t vi = n.value_begin; vi < n.value_end; vi++)
          if (ci_equals(c->dict[o], fp.strings[vi])) {
            ordset[d.set_off + (o >> 5)] |= 1u << (o & 31);
            break;
          }
    }
  }
  uint32_t words = (s->doc_count + 31) / 32;
  SLG_CUDA(ix, bits_out.alloc(std::max<size_t>(words, 1) * 4));
  DevBuf d_nodes, d_set;
  SLG_CUDA(ix, d_nodes.alloc(nodes.size() * sizeof(FilterNodeDev)));
  SLG_CUDA(ix, cudaMemcpyAsync(d_nodes.p, nodes.data(), nodes.size() * sizeof(FilterNodeDev), cudaMemcpyHostToDevice, ix->stream));
  SLG_CUDA(ix, d_set.alloc(std::max<size_t>(ordset.size(), 1) * 4));
  if (!ordset.empty()) SLG_CUDA(ix, cudaMemcpyAsync(d_set.p, ordset.data(), ordset.size() * 4, cudaMemcpyHostToDevice, ix->stream));
  if (words) {
    slg_filter_bitmap_kernel<<<(s->doc_count + 255) / 256, 256, 0, ix->stream>>>(d_nodes.as<FilterNodeDev>(), (uint32_t)nodes.size(),
                                                                                 d_set.as<uint32_t>(), s->doc_count,
                                                                                 bits_out.as<uint32_t>());
    count_launch(ix);
    SLG_CUDA(ix, cudaGetLastError());
  }
  SLG_CUDA(ix, cudaStreamSynchronize(ix->stream));
  return SLG_OK;
}

int32_t slg_filter_compile(slg_index_t *ix, const slg_filter_node_t *nodes, uint32_t n_nodes, const char *const *strings) {
  if (!ix || !nodes || !n_nodes) return SLG_ERR_INVALID;
  SLG_CUDA(ix, cudaSetDevice(ix->device));
  if (n_nodes > kMaxFilterNodes) return fail(ix, SLG_ERR_UNSUPPORTED, "filter has more than %u nodes", kMaxFilterNodes);
  FilterProg fp;
  fp.nodes.assign(nodes, nodes + n_nodes);
  uint32_t max_val = 0;
  for (auto &n : fp.nodes) {
    if (n.op > SLG_F_NOT) return fail(ix, SLG_ERR_INVALID, "unknown filter op %u", n.op);
    if (n.op <= SLG_F_KEYWORD_IN) max_val = std::max(max_val, n.value_end);
  }
  if (max_val && !strings) return fail(ix, SLG_ERR_INVALID, "keyword filter without strings");
  for (uint32_t i = 0; i < max_val; i++) fp.strings.emplace_back(strings[i]);
  // validate the prefix encoding
  {
    uint32_t pos = 0;
    std::vector<uint32_t> pending{1};
    while (!pending.empty()) {
      if (pending.back() == 0) {
        pending.pop_back();
        continue;
      }
      pending.back()--;
      if (pos >= n_nodes) return fail(ix, SLG_ERR_INVALID, "filter program is truncated");
      const auto &n = fp.nodes[pos++];
      if (n.op == SLG_F_NOT && n.n_children != 1) return fail(ix, SLG_ERR_INVALID, "Not takes one child");
      if (n.op >= SLG_F_AND) pending.push_back(n.n_children);
      if (pending.size() > kMaxFilterDepth) return fail(ix, SLG_ERR_UNSUPPORTED, "filter nesting deeper than %u", kMaxFilterDepth);
    }
    if (pos != n_nodes) return fail(ix, SLG_ERR_INVALID, "filter program has trailing nodes");
  }
  std::vector<DevBuf> per_seg(ix->segs.size());
  for (size_t si = 0; si < ix->segs.size(); si++) {
    int32_t rc = compile_filter_for_segment(ix, ix->segs[si].get(), fp, per_seg[si]);
    if (rc) return rc;
  }
  return register_filter(ix, std::move(fp), per_seg);
}

int32_t slg_filter_bitmap(slg_index_t *ix, int32_t filter_id, uint32_t segment_ord, uint32_t *bitmap_out) {
  if (!ix || !bitmap_out) return SLG_ERR_INVALID;
  SLG_CUDA(ix, cudaSetDevice(ix->device));
  Segment *s = ix->find(segment_ord);
  if (!s) return fail(ix, SLG_ERR_INVALID, "no segment %u", segment_ord);
  if (filter_id < 0 || (size_t)filter_id >= s->filter_bits.size() || !s->filter_bits[filter_id].p)
    return fail(ix, SLG_ERR_INVALID, "filter %d is not compiled for segment %u", filter_id, segment_ord);
  uint32_t words = (s->doc_count + 31) / 32;
  SLG_CUDA(ix, cudaMemcpy(bitmap_out, s->filter_bits[filter_id].p, (size_t)words * 4, cudaMemcpyDeviceToHost));
  return SLG_OK;
}

// n consecutive filter ids whose bitmaps are rows of one slab per segment
static int32_t register_slabs(slg_index *ix, std::vector<std::shared_ptr<DevBuf>> &slabs, const std::vector<uint64_t> &stride,
                              uint32_t n, int32_t *out_ids) {
  const size_t id0 = ix->filters.size();
  for (size_t si = 0; si < ix->segs.size(); si++) {
    Segment *s = ix->segs[si].get();
    s->filter_bits.resize(id0 + n);
    s->filter_slabs.resize(id0 + n);
    for (uint32_t i = 0; i < n; i++) {
      s->filter_bits[id0 + i].view(slabs[si]->as<uint32_t>() + (uint64_t)i * stride[si], stride[si] * 4);
      s->filter_slabs[id0 + i] = slabs[si];
    }
    std::vector<const uint32_t *> ptrs;
    for (auto &fb : s->filter_bits) ptrs.push_back(fb.as<uint32_t>());
    if (s->filter_ptrs.bytes < ptrs.size() * sizeof(void *)) {
      SLG_CUDA(ix, cudaStreamSynchronize(ix->stream));  // a running batch may still read the old table
      SLG_CUDA(ix, s->filter_ptrs.alloc(std::max<size_t>(64, ptrs.size() * 2) * sizeof(void *)));
    }
    SLG_CUDA(ix, cudaMemcpy(s->filter_ptrs.p, ptrs.data(), ptrs.size() * sizeof(void *), cudaMemcpyHostToDevice));
  }
  for (uint32_t i = 0; i < n; i++) {
    ix->filters.push_back(FilterProg{});
    out_ids[i] = (int32_t)(id0 + i);
  }
  return SLG_OK;
}

/* ---- phrases (query/phrase.rs:4-48) and bitmap algebra ---- */
int32_t slg_phrase_compile_batch(slg_index_t *ix, const uint32_t *term_ids, const uint32_t *phrase_offsets, const uint32_t *slops,
                                 uint32_t n_phrases, int32_t *out_ids) {
  if (!ix || !term_ids || !phrase_offsets || !n_phrases || !out_ids) return SLG_ERR_INVALID;
  if (ix->segs.empty()) return fail(ix, SLG_ERR_INVALID, "no segment loaded");
  for (uint32_t i = 0; i < n_phrases; i++) {
    const uint32_t n = phrase_offsets[i + 1] - phrase_offsets[i];
    if (phrase_offsets[i + 1] <= phrase_offsets[i]) return fail(ix, SLG_ERR_INVALID, "phrase %u has no terms", i);
    if (n > kMaxPhraseTerms) return fail(ix, SLG_ERR_UNSUPPORTED, "phrase %u has more than %u terms", i, kMaxPhraseTerms);
  }
  SLG_CUDA(ix, cudaSetDevice(ix->device));
  cudaStream_t st = ix->stream;
  // per segment: one slab holding every phrase's bitmap, one table of phrase descriptors, one launch
  std::vector<std::shared_ptr<DevBuf>> slabs(ix->segs.size());
  std::vector<uint64_t> stride(ix->segs.size());
  for (size_t si = 0; si < ix->segs.size(); si++) {
    Segment *s = ix->segs[si].get();
    if (!s->has_positions) return fail(ix, SLG_ERR_INVALID, "segment %u holds no term positions (index without positions, or keep_positions = 0)", s->ord);
    const uint64_t words = align_up(std::max<uint64_t>((s->doc_count + 31) / 32, 1), 32);  // 128-byte rows
    stride[si] = words;
    slabs[si] = std::make_shared<DevBuf>();
    SLG_CUDA(ix, slabs[si]->alloc(words * 4 * n_phrases));
    SLG_CUDA(ix, cudaMemsetAsync(slabs[si]->p, 0, words * 4 * n_phrases, st));
    std::vector<PhraseDev> ph(n_phrases);
    std::vector<uint32_t> blk_off(n_phrases + 1, 0);
    for (uint32_t i = 0; i < n_phrases; i++) {
      PhraseDev &p = ph[i];
      std::memset(&p, 0, sizeof(p));
      const uint32_t *t = term_ids + phrase_offsets[i];
      p.n = phrase_offsets[i + 1] - phrase_offsets[i];
      p.slop = slops ? slops[i] : 0;
      bool absent = false;
      for (uint32_t j = 0; j < p.n; j++) {
        if (t[j] == 0xFFFFFFFFu || t[j] >= s->n_terms || s->h_df[t[j]] == 0) {  // api/reader.rs:1690-1697: no postings => no variant => no match
          absent = true;
          break;
        }
        p.df[j] = s->h_df[t[j]];
        p.start[j] = s->h_start[t[j]];
        if (p.df[j] < p.df[p.driver]) p.driver = j;
      }
      const uint64_t blocks = absent ? 0 : ((uint64_t)p.df[p.driver] + 255) / 256;
      if (absent) p.n = 0;
      if (blk_off[i] + blocks > 0x7FFFFFFFull) return fail(ix, SLG_ERR_UNSUPPORTED, "phrase batch is too large for one launch");
      blk_off[i + 1] = blk_off[i] + (uint32_t)blocks;
    }
    if (blk_off[n_phrases]) {
      DevBuf d_ph, d_off;
      SLG_CUDA(ix, d_ph.alloc(ph.size() * sizeof(PhraseDev)));
      SLG_CUDA(ix, d_off.alloc(blk_off.size() * 4));
      SLG_CUDA(ix, cudaMemcpyAsync(d_ph.p, ph.data(), ph.size() * sizeof(PhraseDev), cudaMemcpyHostToDevice, st));
      SLG_CUDA(ix, cudaMemcpyAsync(d_off.p, blk_off.data(), blk_off.size() * 4, cudaMemcpyHostToDevice, st));
      slg_phrase_bitmap_kernel<<<blk_off[n_phrases], 256, 0, st>>>(d_ph.as<PhraseDev>(), d_off.as<uint32_t>(), n_phrases,
                                                                 s->post_doc.as<uint32_t>(), s->pos_begin.as<uint64_t>(),
                                                                 s->pos.as<uint32_t>(), s->doc_count, slabs[si]->as<uint32_t>(), words);
      count_launch(ix);
      SLG_CUDA(ix, cudaGetLastError());
      SLG_CUDA(ix, cudaStreamSynchronize(st));  // the host tables go out of scope
    }
  }
  SLG_CUDA(ix, cudaStreamSynchronize(st));
  return register_slabs(ix, slabs, stride, n_phrases, out_ids);
}

int32_t slg_phrase_compile(slg_index_t *ix, const uint32_t *term_ids, uint32_t n_terms, uint32_t slop) {
  if (!ix || !term_ids || !n_terms) return SLG_ERR_INVALID;
  const uint32_t off[2] = {0, n_terms};
  int32_t id = -1;
  const int32_t rc = slg_phrase_compile_batch(ix, term_ids, off, &slop, 1, &id);
  return rc ? rc : id;
}

int32_t slg_filter_combine(slg_index_t *ix, uint32_t op, int32_t a, int32_t b) {
  if (!ix) return SLG_ERR_INVALID;
  if (op > SLG_COMBINE_AND_NOT) return fail(ix, SLG_ERR_INVALID, "unknown combine op %u", op);
  SLG_CUDA(ix, cudaSetDevice(ix->device));
  cudaStream_t st = ix->stream;
  std::vector<DevBuf> per_seg(ix->segs.size());
  for (size_t si = 0; si < ix->segs.size(); si++) {
    Segment *s = ix->segs[si].get();
    for (int32_t f : {a, b})
      if (f < 0 || (size_t)f >= s->filter_bits.size() || !s->filter_bits[f].p)
        return fail(ix, SLG_ERR_INVALID, "filter %d is not compiled for segment %u", f, s->ord);
    const uint32_t words = (s->doc_count + 31) / 32;
    SLG_CUDA(ix, per_seg[si].alloc(std::max<size_t>(words, 1) * 4));
    if (words) {
      slg_bitmap_combine_kernel<<<(words + 255) / 256, 256, 0, st>>>(s->filter_bits[a].as<uint32_t>(), s->filter_bits[b].as<uint32_t>(),
                                                                    words, op, per_seg[si].as<uint32_t>());
      count_launch(ix);
    }
  }
  SLG_CUDA(ix, cudaGetLastError());
  SLG_CUDA(ix, cudaStreamSynchronize(st));
  return register_filter(ix, FilterProg{}, per_seg);
}

int32_t slg_filter_combine_batch(slg_index_t *ix, uint32_t op, const int32_t *a, const int32_t *b, uint32_t n, int32_t *out_ids) {
  if (!ix || !a || !b || !n || !out_ids) return SLG_ERR_INVALID;
  if (op > SLG_COMBINE_AND_NOT) return fail(ix, SLG_ERR_INVALID, "unknown combine op %u", op);
  if (n > 65535) return fail(ix, SLG_ERR_UNSUPPORTED, "at most 65535 combinations per call");
  if (ix->segs.empty()) return fail(ix, SLG_ERR_INVALID, "no segment loaded");
  SLG_CUDA(ix, cudaSetDevice(ix->device));
  cudaStream_t st = ix->stream;
  std::vector<std::shared_ptr<DevBuf>> slabs(ix->segs.size());
  std::vector<uint64_t> stride(ix->segs.size());
  for (size_t si = 0; si < ix->segs.size(); si++) {
    Segment *s = ix->segs[si].get();
    std::vector<const uint32_t *> pa(n), pb(n);
    for (uint32_t i = 0; i < n; i++) {
      for (int32_t f : {a[i], b[i]})
        if (f < 0 || (size_t)f >= s->filter_bits.size() || !s->filter_bits[f].p)
          return fail(ix, SLG_ERR_INVALID, "filter %d is freed or not compiled for segment %u", f, s->ord);
      pa[i] = s->filter_bits[a[i]].as<uint32_t>();
      pb[i] = s->filter_bits[b[i]].as<uint32_t>();
    }
    const uint32_t words = (s->doc_count + 31) / 32;
    stride[si] = align_up(std::max<uint64_t>(words, 1), 32);
    slabs[si] = std::make_shared<DevBuf>();
    SLG_CUDA(ix, slabs[si]->alloc(stride[si] * 4 * n));
    SLG_CUDA(ix, cudaMemsetAsync(slabs[si]->p, 0, stride[si] * 4 * n, st));
    if (words) {
      DevBuf d_a, d_b;
      SLG_CUDA(ix, d_a.alloc(n * sizeof(void *)));
      SLG_CUDA(ix, d_b.alloc(n * sizeof(void *)));
      SLG_CUDA(ix, cudaMemcpyAsync(d_a.p, pa.data(), n * sizeof(void *), cudaMemcpyHostToDevice, st));
      SLG_CUDA(ix, cudaMemcpyAsync(d_b.p, pb.data(), n * sizeof(void *), cudaMemcpyHostToDevice, st));
      slg_bitmap_combine_batch_kernel<<<dim3((words + 255) / 256, n), 256, 0, st>>>(
          d_a.as<const uint32_t *>(), d_b.as<const uint32_t *>(), words, op, slabs[si]->as<uint32_t>(), stride[si]);
      count_launch(ix);
      SLG_CUDA(ix, cudaGetLastError());
      SLG_CUDA(ix, cudaStreamSynchronize(st));  // the host tables go out of scope
    }
  }
  return register_slabs(ix, slabs, stride, n, out_ids);
}

int32_t slg_filter_free(slg_index_t *ix, int32_t filter_id) {
  if (!ix) return SLG_ERR_INVALID;
  if (filter_id < 0 || (size_t)filter_id >= ix->filters.size()) return fail(ix, SLG_ERR_INVALID, "no filter %d", filter_id);
  SLG_CUDA(ix, cudaSetDevice(ix->device));
  SLG_CUDA(ix, cudaStreamSynchronize(ix->stream));
  for (auto &s : ix->segs)
    if ((size_t)filter_id < s->filter_bits.size()) {
      s->filter_bits[filter_id].release();
      if ((size_t)filter_id < s->filter_slabs.size()) s->filter_slabs[filter_id].reset();  // the slab goes with its last view
    }
  return SLG_OK;
}

/* ---- batched search ---- */
int32_t slg_batch_prepare(slg_index_t *ix, const slg_query_t *queries, uint32_t n_queries, uint32_t k, slg_exec_t exec,
                          uint32_t bmw_block_size, slg_batch_t **out) {
  if (!ix || !out) return SLG_ERR_INVALID;
  *out = nullptr;
  if (!queries || n_queries == 0) return fail(ix, SLG_ERR_INVALID, "empty query batch");
  if (k == 0) return fail(ix, SLG_ERR_INVALID, "k must be > 0 (the reference bails on limit == 0, api/reader.rs:2540)");
  if (k > SLG_MAX_K) return fail(ix, SLG_ERR_UNSUPPORTED, "k = %u exceeds the built maximum %u", k, SLG_MAX_K);
  if (exec != SLG_EXEC_BM25 && exec != SLG_EXEC_WAND && exec != SLG_EXEC_BMW) return fail(ix, SLG_ERR_INVALID, "unknown execution strategy");
  if (ix->segs.empty()) return fail(ix, SLG_ERR_INVALID, "no segment loaded");
  (void)bmw_block_size;  // bounds are taken over the stored 128-posting blocks; any block size gives the same (exact) result
  SLG_CUDA(ix, cudaSetDevice(ix->device));
  PoolScope pool_scope(ix->stream);
  auto bt = std::make_unique<slg_batch>();
  bt->ix = ix;
  bt->Q = n_queries;
  bt->k = k;
  bt->exec = exec;
  bt->cap = std::max(1024u, 1u << (32 - __builtin_clz(4 * k - 1)));
  const Segment *s0 = ix->segs[0].get();
  uint64_t n_terms_space = 0;
  for (auto &s : ix->segs) n_terms_space = std::max(n_terms_space, s->n_terms);

  std::unordered_map<uint32_t, uint32_t> umap;
  std::vector<uint32_t> &ut = bt->h_ut_term;
  std::vector<uint32_t> &q_off = bt->h_q_term_off;
  std::vector<uint32_t> &qt_u = bt->h_qt_uterm;
  std::vector<float> qt_w;
  std::vector<uint8_t> qt_g, qt_f, q_must(n_queries, 0), q_not(n_queries, 0), q_should(n_queries, 0), q_min(n_queries, 0);
  std::vector<int32_t> q_filter(n_queries, -1);
  std::vector<uint64_t> q_cost(n_queries, 0);
  std::vector<uint8_t> qt_leaf, q_leaves(n_queries, 0);
  std::vector<uint32_t> q_plan_off(n_queries + 1, 0);
  std::vector<PlanNodeDev> plan_nodes;
  q_off.assign(n_queries + 1, 0);
  bool matcher = false;
  for (uint32_t qi = 0; qi < n_queries; qi++) {
    const slg_query_t &q = queries[qi];
    if (q.n_terms && !q.terms) return fail(ix, SLG_ERR_INVALID, "query %u has no terms pointer", qi);
    if (q.n_groups > 8) return fail(ix, SLG_ERR_UNSUPPORTED, "query %u has %u term groups; this build supports 8", qi, q.n_groups);
    if (q.n_groups && !q.group_role) return fail(ix, SLG_ERR_INVALID, "query %u has no group roles", qi);
    if (q.filter_id >= (int32_t)ix->filters.size()) return fail(ix, SLG_ERR_INVALID, "query %u names unknown filter %d", qi, q.filter_id);
    q_filter[qi] = q.filter_id < 0 ? -1 : q.filter_id;
    if (q.n_plan_nodes) {
      // ScorePlan: a well-formed postfix program over leaves 0..leaf_count-1 (query/planner.rs:113-164)
      if (!q.plan) return fail(ix, SLG_ERR_INVALID, "query %u has no plan pointer", qi);
      if (q.leaf_count == 0 || q.leaf_count > SLG_MAX_PLAN_LEAVES)
        return fail(ix, SLG_ERR_UNSUPPORTED, "query %u: a plan needs 1..%u leaves, got %u", qi, SLG_MAX_PLAN_LEAVES, q.leaf_count);
      if (q.n_plan_nodes > SLG_MAX_PLAN_NODES)
        return fail(ix, SLG_ERR_UNSUPPORTED, "query %u: plan of %u nodes; the maximum is %u", qi, q.n_plan_nodes, SLG_MAX_PLAN_NODES);
      uint32_t depth = 0;
      for (uint32_t n = 0; n < q.n_plan_nodes; n++) {
        const slg_plan_node_t &pn = q.plan[n];
        if (pn.op == SLG_PLAN_LEAF) {
          if (pn.arg >= q.leaf_count) return fail(ix, SLG_ERR_INVALID, "query %u plan node %u: leaf %u out of range", qi, n, pn.arg);
          depth++;
        } else if (pn.op == SLG_PLAN_SUM || pn.op == SLG_PLAN_DISMAX) {
          if (pn.arg > depth) return fail(ix, SLG_ERR_INVALID, "query %u plan node %u: %u children but %u values", qi, n, pn.arg, depth);
          if (pn.op == SLG_PLAN_DISMAX && !(pn.tie_breaker >= 0.0f && pn.tie_breaker <= 1.0f))
            return fail(ix, SLG_ERR_INVALID, "query %u plan node %u: tie_breaker must lie in [0, 1]", qi, n);  // planner.rs:850-858
          depth = depth - pn.arg + 1;
        } else {
          return fail(ix, SLG_ERR_INVALID, "query %u plan node %u: unknown op %u", qi, n, pn.op);
        }
        plan_nodes.push_back(PlanNodeDev{pn.op, pn.arg, pn.tie_breaker});
      }
      if (depth != 1) return fail(ix, SLG_ERR_INVALID, "query %u: the plan leaves %u values instead of one", qi, depth);
      q_leaves[qi] = (uint8_t)q.leaf_count;
      bt->has_plan = true;
      bt->max_leaves = std::max(bt->max_leaves, q.leaf_count);
    }
    q_plan_off[qi + 1] = (uint32_t)plan_nodes.size();
    if (q.has_cursor) {
      if (!(q.cursor.score >= 0.0f) || !std::isfinite(q.cursor.score))
        return fail(ix, SLG_ERR_INVALID, "query %u: the cursor score must be finite and >= 0", qi);
      bt->has_cursor = true;
    }
    if (q.filter_id >= 0)
      for (auto &sg : ix->segs)
        if ((size_t)q.filter_id >= sg->filter_bits.size() || !sg->filter_bits[q.filter_id].p)
          return fail(ix, SLG_ERR_INVALID, "query %u names filter %d, which is freed or not compiled for segment %u", qi, q.filter_id, sg->ord);
    uint32_t kept = 0;
    bool need_mask = q.n_groups > 0;
    for (uint32_t t = 0; t < q.n_terms; t++) {
      const slg_term_t &tm = q.terms[t];
      if (tm.term_id == 0xFFFFFFFFu || tm.term_id >= n_terms_space) continue;  // seg.postings(key) == None
      bool scored = tm.flags & SLG_TERM_SCORED;
      if (scored && !(tm.weight > 0.0f && std::isfinite(tm.weight)))
        return fail(ix, SLG_ERR_UNSUPPORTED, "query %u term %u: weight must be finite and > 0", qi, t);
      if (q.n_groups && tm.group >= q.n_groups) return fail(ix, SLG_ERR_INVALID, "query %u term %u: group out of range", qi, t);
      if (scored && q.n_plan_nodes && tm.leaf >= q.leaf_count)
        return fail(ix, SLG_ERR_INVALID, "query %u term %u: leaf %u but the plan has %u leaves", qi, t, tm.leaf, q.leaf_count);  // wand.rs:489-494
      if (!scored) need_mask = true;
      auto it = umap.find(tm.term_id);
      uint32_t u;
      if (it == umap.end()) {
        u = (uint32_t)ut.size();
        umap.emplace(tm.term_id, u);
        ut.push_back(tm.term_id);
      } else {
        u = it->second;
      }
      qt_u.push_back(u);
      qt_w.push_back(tm.weight);
      qt_g.push_back((uint8_t)(q.n_groups ? tm.group : 0));
      qt_f.push_back(scored ? 1 : 0);
      qt_leaf.push_back((uint8_t)(scored && q.n_plan_nodes ? tm.leaf : 0));
      if (scored && tm.term_id < s0->n_terms) q_cost[qi] += s0->h_df[tm.term_id];
      kept++;
    }
    if (kept > SLG_MAX_QUERY_TERMS) return fail(ix, SLG_ERR_UNSUPPORTED, "query %u has %u terms; the maximum is %u", qi, kept, SLG_MAX_QUERY_TERMS);
    q_off[qi + 1] = q_off[qi] + kept;
    bt->max_terms = std::max(bt->max_terms, kept);
    if (need_mask) {
      matcher = true;
      if (q.n_groups == 0) {  // non-scored terms without groups: plain OR over group 0
        q_should[qi] = 1;
        q_min[qi] = 1;
      }
      for (uint32_t g = 0; g < q.n_groups; g++) {
        uint8_t bit = (uint8_t)(1u << g);
        if (q.group_role[g] == SLG_ROLE_MUST) q_must[qi] |= bit;
        else if (q.group_role[g] == SLG_ROLE_MUST_NOT) q_not[qi] |= bit;
        else q_should[qi] |= bit;
      }
      if (q.n_groups) q_min[qi] = (uint8_t)std::min<uint32_t>(q.min_should, 255);
    }
    bt->posting_count += q_cost[qi];
  }
  if (bt->has_cursor) {
    if (ix->kernel_choice == 3 && ix->heavy_kernel == 1)
      return fail(ix, SLG_ERR_UNSUPPORTED, "cursors are not handled by the tile-sweep kernel (heavy_kernel 1)");
    // key.cmp(cursor) (query/sort.rs:80-93: score desc, segment_ord asc, doc_id asc) folded into one exclusive bound on
    // the 64-bit keys of each segment: same segment (score, ~doc); an earlier segment loses ties; a later one wins them
    const uint32_t n_segs = (uint32_t)ix->segs.size();
    std::vector<unsigned long long> bounds((size_t)n_segs * n_queries, ~0ull);
    bt->h_has_cursor.assign(n_queries, 0);
    for (uint32_t qi = 0; qi < n_queries; qi++) {
      const slg_query_t &q = queries[qi];
      if (!q.has_cursor) continue;
      bt->h_has_cursor[qi] = 1;
      uint32_t sb;
      std::memcpy(&sb, &q.cursor.score, 4);
      for (uint32_t si = 0; si < n_segs; si++) {
        const uint32_t ord = ix->segs[si]->ord;
        unsigned long long b;
        if (ord == q.cursor.segment_ord) b = ((unsigned long long)sb << 32) | (unsigned long long)(0xFFFFFFFFu - q.cursor.doc_id);
        else if (ord < q.cursor.segment_ord) b = (unsigned long long)sb << 32;
        else b = ((unsigned long long)sb + 1ull) << 32;
        bounds[(size_t)si * n_queries + qi] = b;
      }
    }
    bt->n_cursor_segs = n_segs;
    SLG_CUDA(ix, bt->cursor_bounds.alloc(bounds.size() * 8));
    SLG_CUDA(ix, bt->cursor_saw.alloc((size_t)n_queries * 4));
    SLG_CUDA(ix, cudaMemcpyAsync(bt->cursor_bounds.p, bounds.data(), bounds.size() * 8, cudaMemcpyHostToDevice, ix->stream));
    SLG_CUDA(ix, cudaStreamSynchronize(ix->stream));  // bounds is a local
  }
  bt->matcher = matcher;
  bt->U = (uint32_t)ut.size();
  bt->T = (uint32_t)qt_u.size();
  // kernel selection
  const bool small = k <= kWarpMaxK && bt->max_terms <= kWarpMaxTerms;
  bool all_scores = ix->staging;
  for (auto &s : ix->segs) all_scores = all_scores && (s->post_score.p != nullptr || s->n_blocks == 0);
  bool sweepable = all_scores;  // the sweep addresses postings with 32-bit indices
  for (auto &s : ix->segs) sweepable = sweepable && (s->post_pair.p != nullptr || s->n_blocks == 0);
  // tile-sweep kernel: plain OR queries (no matcher), small k, few terms, resident scores
  // column front end: plain OR queries (no matcher), small k, few terms, resident scores and columns
  if (bt->has_plan && ix->kernel_choice == 3)
    return fail(ix, SLG_ERR_UNSUPPORTED, "ScorePlan queries run on the warp or CTA-per-item kernel (kernel_choice 0, 1 or 2)");
  bt->use_reg = ix->kernel_choice == 3 || (ix->kernel_choice == 0 && small && !matcher && sweepable && !bt->has_plan);
  if (bt->use_reg && !(small && !matcher && sweepable))
    return fail(ix, SLG_ERR_UNSUPPORTED,
                "the sweep kernel handles plain OR queries, k <= %u, <= %u terms per query, resident scores, < 2^32 postings per segment",
                kWarpMaxK, kWarpMaxTerms);
  bt->use_warp = !bt->use_reg && (ix->kernel_choice == 2 || (ix->kernel_choice == 0 && small));
  if (bt->use_warp && !small)
    return fail(ix, SLG_ERR_UNSUPPORTED, "the warp kernel handles k <= %u and <= %u terms per query", kWarpMaxK, kWarpMaxTerms);

  // processing order inside a tile: most expensive queries first.  Sweep kernel: the "heavy" queries
  // (a column term in any segment, or enough postings that most tiles hold some) come first and are
  // swept; the "light" rest is scored posting-driven by the warp kernel.
  std::vector<uint32_t> order(n_queries);
  for (uint32_t i = 0; i < n_queries; i++) order[i] = i;
  std::vector<uint8_t> heavy(n_queries, 0);
  if (bt->use_reg) {
    std::vector<uint8_t> u_col(bt->U, 0);
    for (uint32_t u = 0; u < bt->U; u++)
      for (auto &sg : ix->segs)
        if (ut[u] < sg->h_term_col.size() && sg->h_term_col[ut[u]] >= 0) u_col[u] = 1;
    uint64_t max_docs = 0;
    for (auto &sg : ix->segs) max_docs = std::max<uint64_t>(max_docs, sg->doc_count);
    bt->warp_cols = ix->sub_docs <= 4096;
    for (uint32_t qi = 0; qi < n_queries && ix->heavy_kernel == 1 && !bt->has_cursor; qi++) {
      bool h = q_cost[qi] >= (ix->sweep_min_postings ? ix->sweep_min_postings : std::max<uint64_t>(1, max_docs / 64));
      for (uint32_t t = q_off[qi]; t < q_off[qi + 1] && !h; t++) h = u_col[qt_u[t]] != 0;
      heavy[qi] = h;
      bt->n_heavy += h;
    }
    bt->n_light = n_queries - bt->n_heavy;
  }
  // swept queries are ordered by their first column term (segment 0): the warps of a CTA walk
  // neighbouring slots at the same time and share that column's tile slice through L1
  std::vector<uint32_t> q_col(n_queries, 0xFFFFFFFFu);
  if (bt->use_reg && !ix->segs.empty()) {
    const Segment *sg0 = ix->segs[0].get();
    for (uint32_t qi = 0; qi < n_queries; qi++)
      for (uint32_t t = q_off[qi]; t < q_off[qi + 1]; t++) {
        const uint32_t term = ut[qt_u[t]];
        if (term < sg0->h_term_col.size() && sg0->h_term_col[term] >= 0) {
          q_col[qi] = (uint32_t)sg0->h_term_col[term];
          break;
        }
      }
  }
  std::stable_sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b2) {
    if (heavy[a] != heavy[b2]) return heavy[a] > heavy[b2];
    if (q_col[a] != q_col[b2]) return q_col[a] < q_col[b2];
    return q_cost[a] > q_cost[b2];
  });

  // pack all inputs into one buffer: one H2D copy per batch
  size_t pos = 0;
  auto place = [&](size_t bytes) {
    size_t o = pos;
    pos = align_up(pos + std::max<size_t>(bytes, 4), 256);
    return o;
  };
  bt->off_ut_term = place((size_t)bt->U * 4);
  bt->off_q_term_off = place((size_t)(n_queries + 1) * 4);
  bt->off_qt_uterm = place((size_t)bt->T * 4);
  bt->off_qt_weight = place((size_t)bt->T * 4);
  bt->off_qt_group = place(bt->T);
  bt->off_qt_flags = place(bt->T);
  bt->off_q_order = place((size_t)n_queries * 4);
  bt->off_q_must = place(n_queries);
  bt->off_q_not = place(n_queries);
  bt->off_q_should = place(n_queries);
  bt->off_q_min = place(n_queries);
  bt->off_q_filter = place((size_t)n_queries * 4);
  if (bt->has_plan) {
    bt->off_qt_leaf = place(bt->T);
    bt->off_q_leaves = place(n_queries);
    bt->off_q_plan_off = place((size_t)(n_queries + 1) * 4);
    bt->off_plan_nodes = place(plan_nodes.size() * sizeof(PlanNodeDev));
  }
  bt->h_pack.assign(pos, 0);
  unsigned char *hp = bt->h_pack.data();
  auto put = [&](size_t off, const void *src, size_t bytes) {
    if (bytes) std::memcpy(hp + off, src, bytes);
  };
  put(bt->off_ut_term, ut.data(), (size_t)bt->U * 4);
  put(bt->off_q_term_off, q_off.data(), (size_t)(n_queries + 1) * 4);
  put(bt->off_qt_uterm, qt_u.data(), (size_t)bt->T * 4);
  put(bt->off_qt_weight, qt_w.data(), (size_t)bt->T * 4);
  put(bt->off_qt_group, qt_g.data(), bt->T);
  put(bt->off_qt_flags, qt_f.data(), bt->T);
  put(bt->off_q_order, order.data(), (size_t)n_queries * 4);
  put(bt->off_q_must, q_must.data(), n_queries);
  put(bt->off_q_not, q_not.data(), n_queries);
  put(bt->off_q_should, q_should.data(), n_queries);
  put(bt->off_q_min, q_min.data(), n_queries);
  put(bt->off_q_filter, q_filter.data(), (size_t)n_queries * 4);
  if (bt->has_plan) {
    put(bt->off_qt_leaf, qt_leaf.data(), bt->T);
    put(bt->off_q_leaves, q_leaves.data(), n_queries);
    put(bt->off_q_plan_off, q_plan_off.data(), (size_t)(n_queries + 1) * 4);
    put(bt->off_plan_nodes, plan_nodes.data(), plan_nodes.size() * sizeof(PlanNodeDev));
  }
  SLG_CUDA(ix, bt->d_pack.alloc(pos));
  SLG_CUDA(ix, cudaMemcpyAsync(bt->d_pack.p, hp, pos, cudaMemcpyHostToDevice, ix->stream));
  ix->ctr.last_h2d_bytes = pos;

  bt->plan_docs = bt->use_reg ? ix->sub_docs : (bt->use_warp ? ix->sub_docs : ix->tile_docs);
  if (bt->has_plan) {
    // one accumulator plane per leaf: shrink the doc tile so that the planes together stay near the configured tile
    uint32_t planes = 1;
    while (planes < bt->max_leaves) planes <<= 1;
    bt->plan_docs = bt->use_warp ? std::max(512u, (ix->sub_docs / planes) & ~127u) : std::max(1024u, (ix->tile_docs / planes) & ~1023u);
  }
  const uint32_t plan_docs = bt->plan_docs;
  uint32_t max_tiles = 0;
  for (auto &s : ix->segs) max_tiles = std::max(max_tiles, (s->doc_count + plan_docs - 1) / plan_docs);
  max_tiles = std::max(max_tiles, 1u);
  const uint32_t n_warp_slots = bt->use_reg ? bt->n_light : n_queries;
  if (bt->use_warp || (bt->use_reg && bt->n_light)) {
    SLG_CUDA(ix, bt->qterms.alloc((size_t)n_warp_slots * kWarpMaxTerms * sizeof(QTerm)));
    SLG_CUDA(ix, bt->qheads.alloc((size_t)n_warp_slots * sizeof(QHead)));
    // the (doc, score) stream form of the warp kernel: no matcher, resident scores
    bt->staged = all_scores && !matcher && bt->U > 0;
  }
  if (bt->use_reg) {
    bt->reg_v = ix->reg_tile_v;
    const size_t nseg = ix->segs.size();
    // rows of the sweep's range table: the unique terms of the heavy queries; the light queries'
    // unique terms get rows of the warp kernel's table
    std::vector<uint32_t> u_row(std::max(bt->U, 1u), 0xFFFFFFFFu), row_u, light_u;
    std::vector<uint8_t> u_light(bt->U, 0);
    std::vector<uint32_t> inst(bt->U, 0);
    for (uint32_t qi = 0; qi < n_queries; qi++)
      for (uint32_t t = q_off[qi]; t < q_off[qi + 1]; t++) {
        const uint32_t u = qt_u[t];
        if (heavy[qi]) {
          inst[u]++;
          if (u_row[u] == 0xFFFFFFFFu) {
            u_row[u] = (uint32_t)row_u.size();
            row_u.push_back(u);
          }
        } else if (!u_light[u]) {
          u_light[u] = 1;
          light_u.push_back(u);
        }
      }
    bt->n_rows = (uint32_t)row_u.size();
    bt->n_light_u = (uint32_t)light_u.size();
    for (auto &sg : ix->segs) {
      const uint32_t tile = 128u * bt->reg_v;
      bt->sweep_tiles_max = std::max(bt->sweep_tiles_max, std::max(1u, (sg->doc_count + tile - 1) / tile));
    }
    // per segment and chunk of kSweepChunk swept slots: the columns staged in shared memory — the ones
    // the chunk's queries name most
    bt->n_chunks = (bt->n_heavy + kSweepChunk - 1) / kSweepChunk;
    std::vector<uint32_t> chunk_cols(std::max<size_t>(1, nseg * bt->n_chunks * kSweepStage), 0xFFFFFFFFu);
    for (size_t si = 0; si < nseg; si++) {
      const Segment *sg = ix->segs[si].get();
      if (sg->h_term_col.empty()) continue;
      for (uint32_t ch = 0; ch < bt->n_chunks; ch++) {
        std::vector<std::pair<uint32_t, uint32_t>> cnt;  // (column, uses)
        for (uint32_t sl = ch * kSweepChunk; sl < std::min(bt->n_heavy, (ch + 1) * kSweepChunk); sl++) {
          const uint32_t qi = order[sl];
          for (uint32_t t = q_off[qi]; t < q_off[qi + 1]; t++) {
            const uint32_t term = ut[qt_u[t]];
            if (term >= sg->h_term_col.size() || sg->h_term_col[term] < 0) continue;
            const uint32_t col = (uint32_t)sg->h_term_col[term];
            auto it = std::find_if(cnt.begin(), cnt.end(), [&](const auto &p2) { return p2.first == col; });
            if (it == cnt.end()) cnt.emplace_back(col, 1u);
            else it->second++;
          }
        }
        std::stable_sort(cnt.begin(), cnt.end(), [](const auto &a, const auto &b2) { return a.second > b2.second; });
        for (size_t i = 0; i < cnt.size() && i < kSweepStage; i++) chunk_cols[(si * bt->n_chunks + ch) * kSweepStage + i] = cnt[i].first;
      }
    }
    auto upload = [&](DevBuf &d, const void *src, size_t bytes) -> cudaError_t {
      cudaError_t e = d.alloc(bytes);
      if (e != cudaSuccess || !bytes) return e;
      ix->ctr.last_h2d_bytes += bytes;
      return cudaMemcpyAsync(d.p, src, bytes, cudaMemcpyHostToDevice, ix->stream);
    };
    SLG_CUDA(ix, upload(bt->d_chunk_cols, chunk_cols.data(), chunk_cols.size() * 4));
    SLG_CUDA(ix, upload(bt->d_u_row, u_row.data(), u_row.size() * 4));
    SLG_CUDA(ix, upload(bt->d_row_u, row_u.data(), row_u.size() * 4));
    SLG_CUDA(ix, upload(bt->d_light_u, light_u.data(), light_u.size() * 4));
    SLG_CUDA(ix, cudaStreamSynchronize(ix->stream));  // host vectors go out of scope
    if (bt->n_heavy) {
      for (float w : qt_w) bt->any_weight = bt->any_weight || w != 1.0f;
      SLG_CUDA(ix, bt->sw_sstat.alloc((size_t)bt->n_heavy * kSweepSlotWords * 4));
      SLG_CUDA(ix, bt->sw_weights.alloc((size_t)bt->n_heavy * 8 * 4));
      SLG_CUDA(ix, bt->sw_ubw.alloc((size_t)bt->n_heavy * 8 * 4));
      SLG_CUDA(ix, bt->sw_records.alloc((size_t)bt->sweep_tiles_max * std::min(bt->n_heavy, kSweepMaxSlots) * kSweepRecWords * 4));
      SLG_CUDA(ix, bt->sw_rng.alloc((size_t)std::max(bt->n_rows, 1u) * (bt->sweep_tiles_max + 1) * 4));
    }
  }
  size_t S = ix->segs.size();
  bt->sub_tiles_max = max_tiles;
  if (!bt->use_reg || bt->n_light) {
    SLG_CUDA(ix, bt->ut_rng.alloc((size_t)std::max(bt->U, 1u) * (max_tiles + 1) * 4));
    if (exec != SLG_EXEC_BM25) SLG_CUDA(ix, bt->ut_tile_ub.alloc((size_t)std::max(bt->U, 1u) * max_tiles * 4));
  }
  SLG_CUDA(ix, bt->thr_key.alloc((size_t)n_queries * 8));
  SLG_CUDA(ix, bt->topk_count.alloc((size_t)n_queries * 4));
  SLG_CUDA(ix, bt->lock.alloc((size_t)n_queries * 4));
  SLG_CUDA(ix, bt->topk_keys.alloc((size_t)n_queries * k * 8));
  SLG_CUDA(ix, bt->work_counter.alloc(64 * 4));
  SLG_CUDA(ix, bt->stats.alloc((size_t)n_queries * 5 * 8));  // [Q][4] counters, then [Q] accepted docs
  SLG_CUDA(ix, bt->seg_hits.alloc(S * n_queries * k * sizeof(HitDev)));
  SLG_CUDA(ix, bt->seg_counts.alloc(S * n_queries * 4));
  if (S > 1) {
    SLG_CUDA(ix, bt->out_hits.alloc((size_t)n_queries * k * sizeof(HitDev)));
    SLG_CUDA(ix, bt->out_counts.alloc((size_t)n_queries * 4));
  }
  bt->pinned_bytes = (size_t)n_queries * k * sizeof(slg_hit_t) + (size_t)n_queries * 4 + (size_t)n_queries * 40;
  if (!ix->pinned_busy && ix->pinned && ix->pinned_bytes >= bt->pinned_bytes) {
    bt->pinned = ix->pinned;
    bt->pinned_from_index = true;
    ix->pinned_busy = true;
  } else {
    SLG_CUDA(ix, cudaMallocHost(&bt->pinned, bt->pinned_bytes));
  }
  SLG_CUDA(ix, cudaStreamSynchronize(ix->stream));
  *out = bt.release();
  return SLG_OK;
}

int32_t slg_batch_run(slg_batch_t *bt, int32_t sync) {
  if (!bt) return SLG_ERR_INVALID;
  slg_index *ix = bt->ix;
  SLG_CUDA(ix, cudaSetDevice(ix->device));
  cudaStream_t st = ix->stream;
  const bool prune = bt->exec != SLG_EXEC_BM25;
  const uint32_t Q = bt->Q, k = bt->k;
  unsigned char *dp = bt->d_pack.as<unsigned char>();
  size_t smem = 0;
  int32_t rc = bt->has_plan ? select_smem(ix, bt->plan_docs, bt->cap, bt->matcher, &smem, bt->max_leaves)
                            : select_smem(ix, ix->tile_docs, bt->cap, bt->matcher, &smem);
  if (rc) return rc;
  uint32_t per_sm = (uint32_t)std::max<size_t>(1, (ix->smem_optin + 1024) / (smem + 1024));
  per_sm = std::min(per_sm, 8u);
  if (ix->ctas_per_sm) per_sm = std::min(per_sm, ix->ctas_per_sm);
  SLG_CUDA(ix, cudaEventRecord(ix->ev[0], st));
  SLG_CUDA(ix, cudaMemsetAsync(bt->stats.p, 0, (size_t)Q * 40, st));
  if (bt->has_cursor) {
    if (bt->n_cursor_segs != ix->segs.size()) return fail(ix, SLG_ERR_INVALID, "a segment was loaded after the batch with cursors was prepared");
    SLG_CUDA(ix, cudaMemsetAsync(bt->cursor_saw.p, 0, (size_t)Q * 4, st));
  }
  uint32_t si = 0;
  for (auto &sp : ix->segs) {
    Segment *s = sp.get();
    BatchDev bd{};
    bd.ut_term = reinterpret_cast<const uint32_t *>(dp + bt->off_ut_term);
    bd.ut_rng = bt->ut_rng.as<uint32_t>();
    bd.ut_tile_ub = bt->ut_tile_ub.as<float>();
    bd.q_term_off = reinterpret_cast<const uint32_t *>(dp + bt->off_q_term_off);
    bd.qt_uterm = reinterpret_cast<const uint32_t *>(dp + bt->off_qt_uterm);
    bd.qt_weight = reinterpret_cast<const float *>(dp + bt->off_qt_weight);
    bd.qt_group = dp + bt->off_qt_group;
    bd.qt_flags = dp + bt->off_qt_flags;
    bd.q_order = reinterpret_cast<const uint32_t *>(dp + bt->off_q_order);
    bd.q_must = dp + bt->off_q_must;
    bd.q_not = dp + bt->off_q_not;
    bd.q_should = dp + bt->off_q_should;
    bd.q_min_should = dp + bt->off_q_min;
    bd.q_filter = reinterpret_cast<const int32_t *>(dp + bt->off_q_filter);
    bd.filter_bits = reinterpret_cast<const uint32_t *const *>(s->filter_ptrs.p);
    if (bt->has_plan) {
      bd.qt_leaf = dp + bt->off_qt_leaf;
      bd.q_leaves = dp + bt->off_q_leaves;
      bd.q_plan_off = reinterpret_cast<const uint32_t *>(dp + bt->off_q_plan_off);
      bd.plan_nodes = reinterpret_cast<const PlanNodeDev *>(dp + bt->off_plan_nodes);
    }
    bd.max_leaves = bt->max_leaves;
    if (bt->has_cursor) {
      bd.q_cursor = bt->cursor_bounds.as<unsigned long long>() + (size_t)si * Q;
      bd.q_saw = bt->cursor_saw.as<uint32_t>();
    }
    bd.n_queries = Q;
    bd.n_uterms = bt->U;
    bd.k = k;
    bd.cap = bt->cap;
    const uint32_t plan_docs = bt->plan_docs;
    bd.tile_docs = plan_docs;
    bd.n_tiles = std::max(1u, (s->doc_count + plan_docs - 1) / plan_docs);
    bd.thr_key = bt->thr_key.as<unsigned long long>();
    bd.topk_count = bt->topk_count.as<uint32_t>();
    bd.lock = bt->lock.as<uint32_t>();
    bd.topk_keys = bt->topk_keys.as<unsigned long long>();
    bd.work_counter = bt->work_counter.as<uint32_t>();
    bd.stats = bt->stats.as<unsigned long long>();
    bd.match_count = bd.stats + (size_t)Q * 4;
    // queries that name a filter need its bitmap on every segment
    if (!ix->filters.empty() && s->filter_bits.size() < ix->filters.size())
      return fail(ix, SLG_ERR_INVALID, "segment %u was loaded after its filters were compiled", s->ord);

    slg_fill_u64_kernel<<<(Q + 255) / 256, 256, 0, st>>>(bd.thr_key, kThrInit, Q);
    count_launch(ix);
    SLG_CUDA(ix, cudaMemsetAsync(bd.topk_count, 0, (size_t)Q * 4, st));
    SLG_CUDA(ix, cudaMemsetAsync(bd.lock, 0, (size_t)Q * 4, st));
    SLG_CUDA(ix, cudaMemsetAsync(bd.work_counter, 0, 64 * 4, st));
    if (bt->U && s->doc_count) {
      // ---- plans and per-segment query tables ----
      BatchDev bw = bd;  // the view of the warp kernel: all queries, or the light ones behind the heavy slots
      uint32_t n_warp_rows = bt->U;
      const uint32_t *warp_rows = nullptr;
      if (bt->use_reg && bt->n_heavy) {  // only the light queries' terms behind the swept slots
        bw.q_order = bd.q_order + bt->n_heavy;
        bw.n_queries = bt->n_light;
        n_warp_rows = bt->n_light_u;
        warp_rows = bt->d_light_u.as<uint32_t>();
      }
      const bool run_warp_side = !bt->use_reg || bt->n_light;
      if (run_warp_side && n_warp_rows) {
        uint64_t n = (uint64_t)n_warp_rows * (bd.n_tiles + 1);
        uint64_t n2 = (uint64_t)n_warp_rows * bd.n_tiles;
        if (warp_rows) {
          slg_plan_ranges_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(s->dev, bd, warp_rows, n_warp_rows);
        } else {
          // every unique term: short lists are walked once, long lists take one binary search per boundary
          slg_sweep_plan_kernel<<<dim3(n_warp_rows, 8), 256, 0, st>>>(s->dev, bd.ut_term, nullptr, n_warp_rows, bd.tile_docs, bd.n_tiles,
                                                                     true, bd.ut_rng);
        }
        count_launch(ix);
        if (prune) {
          slg_plan_bounds_kernel<<<(unsigned)((n2 + 255) / 256), 256, 0, st>>>(s->dev, bd, warp_rows, n_warp_rows);
          count_launch(ix);
        }
      }
      if ((bt->use_warp || (bt->use_reg && bt->n_light)) && bw.n_queries) {
        slg_build_qterms_kernel<<<(bw.n_queries + 127) / 128, 128, 0, st>>>(s->dev, bw, bt->qterms.as<QTerm>(), bt->qheads.as<QHead>(),
                                                                            bt->use_reg && bt->warp_cols && bt->staged);
        count_launch(ix);
      }
      const uint32_t sw_tile = 128u * bt->reg_v;
      const uint32_t sw_tiles = std::max(1u, (s->doc_count + sw_tile - 1) / sw_tile);
      if (bt->use_reg && bt->n_heavy) {
        if (bt->n_rows) {
          slg_sweep_plan_kernel<<<dim3(bt->n_rows, 8), 256, 0, st>>>(s->dev, bd.ut_term, bt->d_row_u.as<uint32_t>(), bt->n_rows, sw_tile,
                                                                      sw_tiles, bt->want_stats, bt->sw_rng.as<uint32_t>());
          count_launch(ix);
        }
        slg_build_sweep_kernel<<<(bt->n_heavy + 127) / 128, 128, 0, st>>>(
            s->dev, bd, bt->n_heavy, bt->d_u_row.as<uint32_t>(), bt->d_chunk_cols.as<uint32_t>() + (size_t)si * bt->n_chunks * kSweepStage,
            bt->sw_sstat.as<uint4>(), bt->sw_weights.as<float>(), bt->sw_ubw.as<float>());
        count_launch(ix);
      }
      SLG_CUDA(ix, cudaGetLastError());
      SLG_CUDA(ix, cudaEventRecord(ix->ev[2], st));
      // ---- scoring ----
      if (bt->use_reg && bt->n_heavy) {
        // seed pass over the first tiles (gives every query a threshold), then the rest in ranges of tiles
        const uint32_t seed_tiles = std::min(sw_tiles, std::max(1u, std::min(sw_tiles / 4, ix->seed_docs / sw_tile)));
        uint32_t launch_id = 0;
        for (uint32_t c0 = 0; c0 < bt->n_heavy; c0 += kSweepMaxSlots) {
          if (launch_id + 3 >= 64) return fail(ix, SLG_ERR_UNSUPPORTED, "more than %u swept queries in one batch", 30 * kSweepMaxSlots);
          SweepDev sw{};
          sw.sstat = bt->sw_sstat.as<uint4>() + (size_t)c0 * (kSweepSlotWords / 4);
          sw.weights = bt->sw_weights.as<float>() + (size_t)c0 * 8;
          sw.ubw = bt->sw_ubw.as<float>() + (size_t)c0 * 8;
          sw.records = bt->sw_records.as<uint32_t>();
          sw.rng = bt->sw_rng.as<uint32_t>();
          sw.post_pair = s->post_pair.as<uint2>();
          sw.col_tmax = s->col_tmax.as<float>();
          sw.chunk_cols = bt->d_chunk_cols.as<uint32_t>() + ((size_t)si * bt->n_chunks + c0 / kSweepChunk) * kSweepStage;
          sw.filter_bits = bd.filter_bits;
          sw.n_slots = std::min(kSweepMaxSlots, bt->n_heavy - c0);
          sw.n_chunks = (sw.n_slots + kSweepChunk - 1) / kSweepChunk;
          sw.k = k;
          sw.n_tiles = sw_tiles;
          sw.tmax_stride = s->tmax_stride;
          sw.thr_key = bd.thr_key;
          sw.topk_count = bd.topk_count;
          sw.lock = bd.lock;
          sw.topk_keys = bd.topk_keys;
          sw.stats = bd.stats;
          {
            const dim3 rgrid((sw.n_slots + 127) / 128, (sw_tiles + kSweepTileGroup - 1) / kSweepTileGroup);
            if (prune) slg_sweep_records_kernel<true><<<rgrid, 128, 0, st>>>(sw, bt->reg_v, bt->want_stats);
            else slg_sweep_records_kernel<false><<<rgrid, 128, 0, st>>>(sw, bt->reg_v, bt->want_stats);
            SLG_CUDA(ix, cudaGetLastError());
            count_launch(ix);
          }
          sw.work_counter = bd.work_counter + 1 + launch_id++;
          sw.tile_begin = 0;
          sw.tile_end = seed_tiles;
          sw.part_tiles = seed_tiles;
          rc = launch_sweep(ix, bt->reg_v, prune, bt->want_stats, bt->any_weight, s->dev, sw, (int)std::min<uint32_t>((uint32_t)ix->n_sm, sw.n_chunks));
          if (rc) return rc;
          count_launch(ix);
          if (seed_tiles < sw_tiles) {
            sw.work_counter = bd.work_counter + 1 + launch_id++;
            sw.tile_begin = seed_tiles;
            sw.tile_end = sw_tiles;
            // enough units for an even finish: about 24 per SM
            const uint32_t want_parts = std::max(1u, (24u * (uint32_t)ix->n_sm + sw.n_chunks - 1) / sw.n_chunks);
            sw.part_tiles = ix->part_tiles ? ix->part_tiles : std::max(32u, (sw_tiles - seed_tiles + want_parts - 1) / want_parts);
            sw.part_tiles = (sw.part_tiles + 31u) & ~31u;  // whole staged blocks
            const uint64_t units = (uint64_t)sw.n_chunks * ((sw_tiles - seed_tiles + sw.part_tiles - 1) / sw.part_tiles);
            rc = launch_sweep(ix, bt->reg_v, prune, bt->want_stats, bt->any_weight, s->dev, sw, (int)std::min<uint64_t>((uint64_t)ix->n_sm, units));
            if (rc) return rc;
            count_launch(ix);
          }
        }
        ix->ctr.score_launches++;
      }
      if ((bt->use_warp || (bt->use_reg && bt->n_light)) && bw.n_queries) {
        WarpBatchDev wb{};
        wb.qterms = bt->qterms.as<QTerm>();
        wb.qheads = bt->qheads.as<QHead>();
        wb.rng = bd.ut_rng;
        wb.sub_ub = bd.ut_tile_ub;
        wb.scores = s->dev.post_score;
        wb.filter_bits = bd.filter_bits;
        wb.n_queries = bw.n_queries;
        wb.k = k;
        wb.sub_docs = bt->has_plan ? bt->plan_docs : ix->sub_docs;
        wb.q_leaves = bd.q_leaves;
        wb.q_plan_off = bd.q_plan_off;
        wb.plan_nodes = bd.plan_nodes;
        wb.max_leaves = bt->max_leaves;
        wb.n_sub = bd.n_tiles;
        wb.n_groups = (bd.n_tiles + kSubPerGroup - 1) / kSubPerGroup;
        wb.ms_frac = (float)ix->maxscore_pct / 100.0f;
        wb.thr_key = bd.thr_key;
        wb.topk_count = bd.topk_count;
        wb.lock = bd.lock;
        wb.topk_keys = bd.topk_keys;
        wb.work_counter = bd.work_counter;
        wb.stats = bd.stats;
        wb.match_count = bd.match_count;
        wb.q_cursor = bd.q_cursor;
        wb.q_saw = bd.q_saw;
        const int warps = kThreads / 32;
        // (plan batches: the matcher form of the kernel unless the staged plain-OR form applies — size for the larger)
        size_t wsmem = (size_t)warps * warp_kernel_smem_per_warp(wb.sub_docs, bt->matcher || (bt->has_plan && !bt->staged), prune,
                                                                 bt->has_plan ? bt->max_leaves : 1u);
        if (wsmem + 1024 > ix->smem_optin) return fail(ix, SLG_ERR_UNSUPPORTED, "sub_docs %u needs %zu B shared memory", ix->sub_docs, wsmem);
        uint32_t wper = (uint32_t)std::max<size_t>(1, (ix->smem_optin + 1024) / (wsmem + 1024));
        wper = std::min(wper, 8u);
        if (ix->ctas_per_sm) wper = std::min(wper, ix->ctas_per_sm);
        int grid = (int)std::min<uint64_t>((uint64_t)ix->n_sm * wper, ((uint64_t)wb.n_groups * wb.n_queries + warps - 1) / warps);
        rc = launch_warp(ix, bt->matcher, prune, bt->want_stats, bt->staged, bt->use_reg && bt->warp_cols, s->dev, wb, wsmem, grid,
                         bt->has_plan);
        if (rc) return rc;
        count_launch(ix);
        if (!bt->use_reg || !bt->n_heavy) ix->ctr.score_launches++;
      } else if (!bt->use_reg && !bt->use_warp) {
        int grid = (int)std::min<uint64_t>((uint64_t)ix->n_sm * per_sm, (uint64_t)bd.n_tiles * Q);
        rc = launch_score(ix, bt->matcher, prune, bt->want_stats, s->dev, bd, smem, grid, bt->has_plan);
        if (rc) return rc;
        count_launch(ix);
        ix->ctr.score_launches++;
      }
      SLG_CUDA(ix, cudaEventRecord(ix->ev[3], st));
    }
    HitDev *hits = bt->seg_hits.as<HitDev>() + (size_t)si * Q * k;
    uint32_t *cnts = bt->seg_counts.as<uint32_t>() + (size_t)si * Q;
    size_t fsmem = (size_t)(1u << (32 - __builtin_clz(std::max(k, 2u) - 1))) * 8;
    slg_finalize_kernel<<<Q, kThreads, fsmem, st>>>(bd, s->ord, hits, cnts);
    count_launch(ix);
    SLG_CUDA(ix, cudaGetLastError());
    if (ix->segs.size() > 1 && bt->U && s->doc_count) {
      // per-segment score time must be read before the events are reused
      SLG_CUDA(ix, cudaEventSynchronize(ix->ev[3]));
      float ms = 0;
      SLG_CUDA(ix, cudaEventElapsedTime(&ms, ix->ev[2], ix->ev[3]));
      ix->ctr.score_ms_total += ms;
      ix->ctr.last_score_ms = ms;
    }
    si++;
  }
  bt->n_segs_run = si;
  if (si > 1) {
    size_t msmem = (size_t)si * k * sizeof(HitDev);
    if (msmem > ix->smem_optin) return fail(ix, SLG_ERR_UNSUPPORTED, "merge of %u segments x k=%u does not fit shared memory", si, k);
    SLG_CUDA(ix, cudaFuncSetAttribute(slg_merge_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)msmem));
    slg_merge_kernel<<<Q, kThreads, msmem, st>>>(bt->seg_hits.as<HitDev>(), bt->seg_counts.as<uint32_t>(), si, Q, k,
                                                 bt->out_hits.as<HitDev>(), bt->out_counts.as<uint32_t>());
    count_launch(ix);
    SLG_CUDA(ix, cudaGetLastError());
  }
  SLG_CUDA(ix, cudaEventRecord(ix->ev[1], st));
  ix->ctr.last_posting_count = bt->posting_count;
  if (sync) {
    SLG_CUDA(ix, cudaStreamSynchronize(st));
    float ms = 0;
    SLG_CUDA(ix, cudaEventElapsedTime(&ms, ix->ev[0], ix->ev[1]));
    ix->ctr.last_batch_ms = ms;
    if (ix->segs.size() == 1 && bt->U) {
      SLG_CUDA(ix, cudaEventElapsedTime(&ms, ix->ev[2], ix->ev[3]));
      ix->ctr.score_ms_total += ms;
      ix->ctr.last_score_ms = ms;
    }
  }
  return SLG_OK;
}

int32_t slg_batch_enable_stats(slg_batch_t *bt, int32_t on) {
  if (!bt) return SLG_ERR_INVALID;
  bt->want_stats = on != 0;
  return SLG_OK;
}

int32_t slg_batch_device_results(slg_batch_t *bt, void **dev_hits, void **dev_counts) {
  if (!bt || !dev_hits || !dev_counts) return SLG_ERR_INVALID;
  bool merged = bt->n_segs_run > 1;
  *dev_hits = merged ? bt->out_hits.p : bt->seg_hits.p;
  *dev_counts = merged ? bt->out_counts.p : bt->seg_counts.p;
  return SLG_OK;
}

int32_t slg_batch_fetch(slg_batch_t *bt, slg_hit_t *out_hits, uint32_t *out_counts, slg_stats_t *out_stats) {
  if (!bt || !out_hits || !out_counts) return SLG_ERR_INVALID;
  slg_index *ix = bt->ix;
  SLG_CUDA(ix, cudaSetDevice(ix->device));
  cudaStream_t st = ix->stream;
  void *dh, *dc;
  slg_batch_device_results(bt, &dh, &dc);
  static_assert(sizeof(slg_hit_t) == sizeof(HitDev), "hit layout");
  size_t hb = (size_t)bt->Q * bt->k * sizeof(slg_hit_t), cb = (size_t)bt->Q * 4, sb = (size_t)bt->Q * 40;
  unsigned char *pin = static_cast<unsigned char *>(bt->pinned);
  SLG_CUDA(ix, cudaMemcpyAsync(pin, dh, hb, cudaMemcpyDeviceToHost, st));
  SLG_CUDA(ix, cudaMemcpyAsync(pin + hb, dc, cb, cudaMemcpyDeviceToHost, st));
  if (out_stats) SLG_CUDA(ix, cudaMemcpyAsync(pin + hb + cb, bt->stats.p, sb, cudaMemcpyDeviceToHost, st));
  SLG_CUDA(ix, cudaStreamSynchronize(st));
  ix->ctr.last_d2h_bytes = hb + cb + (out_stats ? sb : 0);
  std::memcpy(out_hits, pin, hb);
  std::memcpy(out_counts, pin + hb, cb);
  if (out_stats) {
    const unsigned long long *sv = reinterpret_cast<const unsigned long long *>(pin + hb + cb);
    for (uint32_t q = 0; q < bt->Q; q++) {
      out_stats[q].scored_docs = sv[q * 4 + 0];
      out_stats[q].postings_advanced = sv[q * 4 + 1];
      out_stats[q].blocks_skipped = sv[q * 4 + 2];
      out_stats[q].candidates_examined = sv[q * 4 + 3];
      out_stats[q].total_matches = sv[(size_t)bt->Q * 4 + q];
    }
  }
  return SLG_OK;
}

int32_t slg_batch_copy_results_device(slg_batch_t *bt, void *dst_hits, void *dst_counts) {
  if (!bt || !dst_hits || !dst_counts) return SLG_ERR_INVALID;
  slg_index *ix = bt->ix;
  SLG_CUDA(ix, cudaSetDevice(ix->device));
  void *dh, *dc;
  slg_batch_device_results(bt, &dh, &dc);
  SLG_CUDA(ix, cudaMemcpyAsync(dst_hits, dh, (size_t)bt->Q * bt->k * sizeof(HitDev), cudaMemcpyDeviceToDevice, ix->stream));
  SLG_CUDA(ix, cudaMemcpyAsync(dst_counts, dc, (size_t)bt->Q * 4, cudaMemcpyDeviceToDevice, ix->stream));
  return SLG_OK;
}

int32_t slg_batch_cursor_seen(slg_batch_t *bt, uint8_t *out_seen) {
  if (!bt || !out_seen) return SLG_ERR_INVALID;
  slg_index *ix = bt->ix;
  SLG_CUDA(ix, cudaSetDevice(ix->device));
  for (uint32_t q = 0; q < bt->Q; q++) out_seen[q] = 1;  // no cursor: saw_cursor starts true (api/reader.rs:2663)
  if (!bt->has_cursor) return SLG_OK;
  std::vector<uint32_t> saw(bt->Q);
  SLG_CUDA(ix, cudaMemcpyAsync(saw.data(), bt->cursor_saw.p, (size_t)bt->Q * 4, cudaMemcpyDeviceToHost, ix->stream));
  SLG_CUDA(ix, cudaStreamSynchronize(ix->stream));
  for (uint32_t q = 0; q < bt->Q; q++)
    if (bt->h_has_cursor[q]) out_seen[q] = saw[q] ? 1 : 0;
  return SLG_OK;
}

// PaginationCursor::encode / decode, api/reader.rs:630-691, and the generation check of decode_cursor, :821-841
int32_t slg_cursor_encode(uint32_t generation, uint32_t returned, const slg_hit_t *last_hit, char *out43) {
  if (!last_hit || !out43) return SLG_ERR_INVALID;
  uint32_t sb;
  std::memcpy(&sb, &last_hit->score, 4);
  unsigned char buf[21];
  buf[0] = 1;  // CURSOR_VERSION
  const uint32_t words[5] = {generation, sb, last_hit->segment_ord, last_hit->doc_id, returned};
  for (int w = 0; w < 5; w++)
    for (int b = 0; b < 4; b++) buf[1 + w * 4 + b] = (unsigned char)(words[w] >> (24 - 8 * b));
  static const char HEX[] = "0123456789abcdef";
  for (int i = 0; i < 21; i++) {
    out43[2 * i] = HEX[buf[i] >> 4];
    out43[2 * i + 1] = HEX[buf[i] & 15];
  }
  out43[42] = 0;
  return SLG_OK;
}

int32_t slg_cursor_decode(const char *raw, uint32_t manifest_generation, slg_hit_t *key, uint32_t *returned, char *err,
                          uint64_t err_len) {
  auto bail = [&](const char *fmt, auto... a) {
    if (err && err_len) std::snprintf(err, (size_t)err_len, fmt, a...);
    return (int32_t)SLG_ERR_INVALID;
  };
  if (!raw || !key || !returned) return bail("%s", "null argument");
  const size_t len = std::strlen(raw);
  if (len != 42) return bail("invalid cursor length: expected 42 hex chars, got %zu", len);
  unsigned char bytes[21];
  for (int i = 0; i < 21; i++) {
    int v = 0;
    for (int h = 0; h < 2; h++) {
      const char c = raw[2 * i + h];
      int d;
      if (c >= '0' && c <= '9') d = c - '0';
      else if (c >= 'a' && c <= 'f') d = c - 'a' + 10;
      else if (c >= 'A' && c <= 'F') d = c - 'A' + 10;  // u8::from_str_radix accepts both cases
      else return bail("decoding cursor at byte index %d", i);
      v = v * 16 + d;
    }
    bytes[i] = (unsigned char)v;
  }
  if (bytes[0] != 1) return bail("unsupported cursor version %u", (unsigned)bytes[0]);
  uint32_t words[5];
  for (int w = 0; w < 5; w++)
    words[w] = ((uint32_t)bytes[1 + w * 4] << 24) | ((uint32_t)bytes[2 + w * 4] << 16) | ((uint32_t)bytes[3 + w * 4] << 8) | bytes[4 + w * 4];
  if (words[4] > 50000u) return bail("cursor requests %u hits, which exceeds max supported 50000", words[4]);
  if (words[0] != manifest_generation)
    return bail("stale cursor for this index generation: expected %u, got %u", manifest_generation, words[0]);
  std::memcpy(&key->score, &words[1], 4);
  key->segment_ord = words[2];
  key->doc_id = words[3];
  *returned = words[4];
  return SLG_OK;
}

int32_t slg_batch_free(slg_batch_t *bt) {
  if (!bt) return SLG_OK;
  cudaSetDevice(bt->ix->device);
  cudaStreamSynchronize(bt->ix->stream);
  delete bt;
  return SLG_OK;
}

int32_t slg_search_batch(slg_index_t *ix, const slg_query_t *queries, uint32_t n_queries, uint32_t k, slg_exec_t exec,
                         uint32_t bmw_block_size, slg_hit_t *out_hits, uint32_t *out_counts, slg_stats_t *out_stats) {
  if (!ix) return SLG_ERR_INVALID;
  if (!out_hits || !out_counts) return fail(ix, SLG_ERR_INVALID, "output buffers are NULL");
  slg_batch_t *bt = nullptr;
  int32_t rc = slg_batch_prepare(ix, queries, n_queries, k, exec, bmw_block_size, &bt);
  if (rc) return rc;
  bt->want_stats = out_stats != nullptr;
  rc = slg_batch_run(bt, 0);
  if (rc == SLG_OK) rc = slg_batch_fetch(bt, out_hits, out_counts, out_stats);
  if (rc == SLG_OK && ix->segs.size() == 1 && bt->U) {
    float ms = 0;
    if (cudaEventElapsedTime(&ms, ix->ev[2], ix->ev[3]) == cudaSuccess) {
      ix->ctr.score_ms_total += ms;
      ix->ctr.last_score_ms = ms;
    }
    if (cudaEventElapsedTime(&ms, ix->ev[0], ix->ev[1]) == cudaSuccess) ix->ctr.last_batch_ms = ms;
  }
  slg_batch_free(bt);
  return rc;
}

int32_t slg_merge_gathered(slg_index_t *ix, const void *dev_hits, const void *dev_counts, uint32_t n_shards,
                           uint32_t n_queries, uint32_t k, slg_hit_t *out_hits, uint32_t *out_counts) {
  if (!ix || !dev_hits || !dev_counts || !out_hits || !out_counts || !n_shards || !n_queries || !k) return SLG_ERR_INVALID;
  SLG_CUDA(ix, cudaSetDevice(ix->device));
  cudaStream_t st = ix->stream;
  size_t msmem = (size_t)n_shards * k * sizeof(HitDev);
  if (msmem > ix->smem_optin) return fail(ix, SLG_ERR_UNSUPPORTED, "merge of %u shards x k=%u does not fit shared memory", n_shards, k);
  PoolScope pool_scope(st);  // per-call buffers from the stream-ordered pool
  DevBuf oh, oc;
  SLG_CUDA(ix, oh.alloc((size_t)n_queries * k * sizeof(HitDev)));
  SLG_CUDA(ix, oc.alloc((size_t)n_queries * 4));
  SLG_CUDA(ix, cudaFuncSetAttribute(slg_merge_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)msmem));
  slg_merge_kernel<<<n_queries, kThreads, msmem, st>>>(static_cast<const HitDev *>(dev_hits), static_cast<const uint32_t *>(dev_counts),
                                                       n_shards, n_queries, k, oh.as<HitDev>(), oc.as<uint32_t>());
  count_launch(ix);
  SLG_CUDA(ix, cudaGetLastError());
  SLG_CUDA(ix, cudaMemcpyAsync(out_hits, oh.p, (size_t)n_queries * k * sizeof(HitDev), cudaMemcpyDeviceToHost, st));
  SLG_CUDA(ix, cudaMemcpyAsync(out_counts, oc.p, (size_t)n_queries * 4, cudaMemcpyDeviceToHost, st));
  SLG_CUDA(ix, cudaStreamSynchronize(st));
  ix->ctr.last_d2h_bytes = (size_t)n_queries * k * sizeof(HitDev) + (size_t)n_queries * 4;
  return SLG_OK;
}

/* ---- vectors + rerank ---- */
int32_t slg_load_vectors(slg_index_t *ix, uint32_t segment_ord, uint32_t dim, const uint32_t *offsets, const float *values,
                         uint64_t n_rows, int32_t store_bf16) {
  if (!ix || !offsets || (!values && n_rows) || !dim) return SLG_ERR_INVALID;
  SLG_CUDA(ix, cudaSetDevice(ix->device));
  Segment *s = ix->find(segment_ord);
  if (!s) return fail(ix, SLG_ERR_INVALID, "no segment %u", segment_ord);
  if (dim % 8) return fail(ix, SLG_ERR_UNSUPPORTED, "vector dim must be a multiple of 8");
  cudaStream_t st = ix->stream;
  Vectors &v = s->vec;
  v.dim = dim;
  v.n_rows = n_rows;
  v.bf16 = store_bf16 != 0;
  SLG_CUDA(ix, v.offsets.alloc(std::max<size_t>(s->doc_count, 1) * 4));
  SLG_CUDA(ix, cudaMemcpyAsync(v.offsets.p, offsets, (size_t)s->doc_count * 4, cudaMemcpyHostToDevice, st));
  size_t n = (size_t)n_rows * dim;
  if (!v.bf16) {
    SLG_CUDA(ix, v.values.alloc(std::max<size_t>(n, 1) * 4));
    if (n) SLG_CUDA(ix, cudaMemcpyAsync(v.values.p, values, n * 4, cudaMemcpyHostToDevice, st));
  } else {
    DevBuf tmp;
    SLG_CUDA(ix, tmp.alloc(std::max<size_t>(n, 1) * 4));
    if (n) SLG_CUDA(ix, cudaMemcpyAsync(tmp.p, values, n * 4, cudaMemcpyHostToDevice, st));
    SLG_CUDA(ix, v.values.alloc(std::max<size_t>(n, 1) * 2));
    if (n) {
      slg_f32_to_bf16_kernel<<<(unsigned)std::min<size_t>((n + 255) / 256, 1u << 20), 256, 0, st>>>(tmp.as<float>(), v.values.as<__nv_bfloat16>(), n);
      count_launch(ix);
    }
    SLG_CUDA(ix, cudaStreamSynchronize(st));
  }
  SLG_CUDA(ix, cudaStreamSynchronize(st));
  return SLG_OK;
}

int32_t slg_rerank(slg_index_t *ix, const float *query_vecs, uint32_t n_queries, uint32_t dim, const slg_hit_t *cands,
                   const uint32_t *cand_counts, uint32_t cand_stride, float alpha, slg_metric_t metric, slg_hit_t *out_hits,
                   float *out_vector_scores) {
  if (!ix || !query_vecs || !cands || !cand_counts || !out_hits || !n_queries || !cand_stride) return SLG_ERR_INVALID;
  if (metric != SLG_METRIC_COSINE && metric != SLG_METRIC_L2) return fail(ix, SLG_ERR_INVALID, "unknown metric");
  SLG_CUDA(ix, cudaSetDevice(ix->device));
  cudaStream_t st = ix->stream;
  // segment table for the kernel
  std::vector<RerankSegDev> segs;
  for (auto &s : ix->segs) {
    RerankSegDev r{};
    r.segment_ord = s->ord;
    r.doc_count = s->doc_count;
    r.offsets = s->vec.offsets.as<uint32_t>();
    r.values = s->vec.values.p;
    r.bf16 = s->vec.bf16 ? 1 : 0;
    r.dim = s->vec.dim;
    if (s->vec.dim && s->vec.dim != dim) return fail(ix, SLG_ERR_INVALID, "query dim %u != stored dim %u", dim, s->vec.dim);
    segs.push_back(r);
  }
  if (cand_stride > kMaxRerankCands) return fail(ix, SLG_ERR_UNSUPPORTED, "more than %u candidates per query", kMaxRerankCands);
  PoolScope pool_scope(st);  // per-call buffers from the stream-ordered pool
  DevBuf d_segs, d_q, d_c, d_n, d_o, d_vs;
  size_t nh = (size_t)n_queries * cand_stride;
  SLG_CUDA(ix, d_segs.alloc(segs.size() * sizeof(RerankSegDev)));
  SLG_CUDA(ix, cudaMemcpyAsync(d_segs.p, segs.data(), segs.size() * sizeof(RerankSegDev), cudaMemcpyHostToDevice, st));
  SLG_CUDA(ix, d_q.alloc((size_t)n_queries * dim * 4));
  SLG_CUDA(ix, cudaMemcpyAsync(d_q.p, query_vecs, (size_t)n_queries * dim * 4, cudaMemcpyHostToDevice, st));
  SLG_CUDA(ix, d_c.alloc(nh * sizeof(HitDev)));
  SLG_CUDA(ix, cudaMemcpyAsync(d_c.p, cands, nh * sizeof(HitDev), cudaMemcpyHostToDevice, st));
  SLG_CUDA(ix, d_n.alloc((size_t)n_queries * 4));
  SLG_CUDA(ix, cudaMemcpyAsync(d_n.p, cand_counts, (size_t)n_queries * 4, cudaMemcpyHostToDevice, st));
  SLG_CUDA(ix, d_o.alloc(nh * sizeof(HitDev)));
  SLG_CUDA(ix, d_vs.alloc(nh * 4));
  size_t smem = (size_t)dim * 4 + (size_t)cand_stride * (sizeof(HitDev) + 4);
  if (smem > ix->smem_optin) return fail(ix, SLG_ERR_UNSUPPORTED, "rerank tile does not fit shared memory");
  SLG_CUDA(ix, cudaFuncSetAttribute(slg_rerank_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  slg_rerank_kernel<<<n_queries, 256, smem, st>>>(d_segs.as<RerankSegDev>(), (uint32_t)segs.size(), d_q.as<float>(), dim,
                                                  d_c.as<HitDev>(), d_n.as<uint32_t>(), cand_stride, alpha, (int)metric,
                                                  d_o.as<HitDev>(), d_vs.as<float>());
  count_launch(ix);
  SLG_CUDA(ix, cudaGetLastError());
  SLG_CUDA(ix, cudaMemcpyAsync(out_hits, d_o.p, nh * sizeof(HitDev), cudaMemcpyDeviceToHost, st));
  if (out_vector_scores) SLG_CUDA(ix, cudaMemcpyAsync(out_vector_scores, d_vs.p, nh * 4, cudaMemcpyDeviceToHost, st));
  SLG_CUDA(ix, cudaStreamSynchronize(st));
  return SLG_OK;
}

int32_t slg_selftest_div(slg_index_t *ix, uint64_t n, uint64_t seed, uint64_t *mismatches) {
  if (!ix || !mismatches) return SLG_ERR_INVALID;
  SLG_CUDA(ix, cudaSetDevice(ix->device));
  DevBuf d;
  SLG_CUDA(ix, d.alloc(8));
  SLG_CUDA(ix, cudaMemsetAsync(d.p, 0, 8, ix->stream));
  slg_selftest_div_kernel<<<ix->n_sm * 8, 256, 0, ix->stream>>>(n, seed, d.as<unsigned long long>());
  count_launch(ix);
  SLG_CUDA(ix, cudaGetLastError());
  unsigned long long v = 0;
  SLG_CUDA(ix, cudaMemcpyAsync(&v, d.p, 8, cudaMemcpyDeviceToHost, ix->stream));
  SLG_CUDA(ix, cudaStreamSynchronize(ix->stream));
  *mismatches = v;
  return SLG_OK;
}

int32_t slg_get_stream(const slg_index_t *ix, void **cuda_stream) {
  if (!ix || !cuda_stream) return SLG_ERR_INVALID;
  *cuda_stream = (void *)ix->stream;
  return SLG_OK;
}

int32_t slg_get_counters(const slg_index_t *ix, slg_counters_t *out) {
  if (!ix || !out) return SLG_ERR_INVALID;
  *out = ix->ctr;
  return SLG_OK;
}

}  // extern "C"
