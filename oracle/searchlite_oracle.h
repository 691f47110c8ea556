/*
 * searchlite_oracle.h — C ABI of the CPU oracle.
 *
 * TEST INFRASTRUCTURE ONLY.  This is a CPU restatement of the reference's
 * (davidkelley/searchlite) BM25 / WAND / BMW top-k path, used as the checker for
 * the CUDA engine.  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load it.  The product
 * (searchlite_b200/) never links, imports or calls anything in oracle/.
 *
 * Parity status: the reference is Rust and cannot be built in this image (no
 * cargo/rustc), so the oracle is pinned against the reference's own literal
 * test inputs and properties ("parity pinned on literals/properties only"; see tests/test_oracle_*.py):
 *   query/wand.rs:952-966, 969-1011, 1014-1021, 1024-1052; index/postings.rs:280-310;
 *   util/varint.rs:55-63; tests/pruning.rs:45-104 (property re-created);
 *   SURVEY.md §8c BMW counter-example.
 * There are no absolute BM25 golden numbers anywhere in the reference.
 *
 * All citations are relative to /root/reference/searchlite-core/src/.
 */
#ifndef SEARCHLITE_ORACLE_H
#define SEARCHLITE_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct slo_index slo_index_t;

/* api/types.rs:6-13 ExecutionStrategy (+ oracle-only variants) */
enum {
  SLO_EXEC_BM25 = 0,        /* brute_force, query/wand.rs:459-566 (hash-map accumulation)   */
  SLO_EXEC_WAND = 1,        /* wand_loop, term-wide bounds, query/wand.rs:659-903            */
  SLO_EXEC_BMW = 2,         /* wand_loop with block bounds — FAITHFUL, incl. the unsafe stop */
  SLO_EXEC_BM25_DENSE = 3   /* same result as BM25, dense accumulator ("fair" CPU port)      */
};

/* group roles of the flat matcher (api/reader.rs:1485-1565) */
enum { SLO_ROLE_SHOULD = 0, SLO_ROLE_MUST = 1, SLO_ROLE_MUST_NOT = 2 };

enum { SLO_TERM_SCORED = 1u };

typedef struct {
  uint32_t term_id; /* ordinal of the "field:term" key in this segment; UINT32_MAX = absent */
  float weight;     /* sum of group.boost*field.boost over duplicate keys, api/reader.rs:2971-2983 */
  uint32_t leaf;    /* ScorePlan leaf, query/planner.rs:284-306 */
  uint32_t group;   /* matcher term group this key belongs to */
  uint32_t flags;   /* SLO_TERM_SCORED if it contributes to the score */
} slo_term_t;

/* ScoreExpr (query/planner.rs:113-153) flattened to postfix: LEAF arg = leaf index;
 * SUM / DISMAX arg = number of children (the values on top of the evaluation stack). */
enum { SLO_PLAN_LEAF = 0, SLO_PLAN_SUM = 1, SLO_PLAN_DISMAX = 2 };
typedef struct {
  uint32_t op;
  uint32_t arg;
  float tie_breaker;
} slo_plan_node_t;

typedef struct {
  uint32_t n_terms;
  const slo_term_t *terms;
  uint32_t n_groups;
  const uint8_t *group_role; /* n_groups entries */
  uint32_t min_should;       /* resolved minimum_should_match */
  uint32_t leaf_count;       /* 0 => score_plan None: plain running sum */
  int32_t filter_id;         /* unused here (the root filter is a call argument); keeps the layout of slg_query_t */
  uint32_t n_plan_nodes;     /* 0 with leaf_count > 0 => Sum of leaves (query/planner.rs:354-360) */
  const slo_plan_node_t *plan; /* ScoreExpr in postfix order, query/planner.rs:113-153 */
  uint32_t has_cursor;       /* search-after cursor, api/reader.rs:3019-3028 */
  uint32_t cursor_segment_ord, cursor_doc_id;
  float cursor_score;
} slo_query_t;

/* Filter AST, api/types.rs:670-680, evaluated as query/filters.rs:84-149.  Prefix
 * encoding: a node is followed by its n_children sub-trees. */
enum {
  SLO_F_KEYWORD_EQ = 0,
  SLO_F_KEYWORD_IN = 1,
  SLO_F_I64_RANGE = 2,
  SLO_F_F64_RANGE = 3,
  SLO_F_AND = 4,
  SLO_F_OR = 5,
  SLO_F_NOT = 6
};

typedef struct {
  uint32_t op;
  int32_t column;       /* column handle from slo_index_add_*_column; -1 = unknown field */
  int64_t i_min, i_max; /* inclusive */
  double f_min, f_max;  /* inclusive */
  uint32_t n_children;  /* And / Or / Not */
  uint32_t value_begin, value_end; /* keyword values: range in the strings array */
} slo_filter_node_t;

typedef struct {
  uint32_t segment_ord;
  uint32_t doc_id;
  float score;
} slo_hit_t;

/* query/wand.rs:45-50 QueryStats, plus the accept counter of api/reader.rs:3029-3031 */
typedef struct {
  uint64_t scored_docs;
  uint64_t candidates_examined;
  uint64_t postings_advanced;
  uint64_t total_matches;
  uint64_t saw_cursor;       /* the cursor's own doc was met by accept (api/reader.rs:3022-3024) */
} slo_stats_t;

/* ---- scalar arithmetic (query/bm25.rs:1-6, query/wand.rs:269-303) ---- */
float slo_bm25(float tf, float df, float doc_len, float avgdl, float docs, float k1, float b);
float slo_score_tf(float tf, float df, float doc_len, float avgdl, float docs, float k1, float b, float weight);
float slo_upper_bound_tf(float tf, float df, float doc_len, float avgdl, float docs, float k1, float b, float weight);
/* idf term alone: ln((N-df+.5)/(df+.5)).max(0)+1 — what the device tables are built from */
float slo_idf(float df, float docs);

/* ---- codecs (util/varint.rs:5-49, index/postings.rs:78-212) ---- */
size_t slo_varint_write_u32(uint32_t v, uint8_t *out /* >=5 bytes */);
/* returns bytes consumed, 0 on error */
size_t slo_varint_read_u32(const uint8_t *buf, size_t len, uint32_t *out);
/* Encode one posting list as PostingsWriter::write_term does.  positions may be NULL
 * (keep_positions=false).  pos_offsets has n+1 entries when given.  Returns bytes
 * written; call with out=NULL to size. */
size_t slo_postings_encode(const uint32_t *docs, const uint32_t *tfs, size_t n, int keep_positions,
                           const uint32_t *pos_offsets, const uint32_t *positions, uint8_t *out, size_t cap);
/* Decode as PostingsReader::read_at.  Arrays sized by the caller from slo_postings_peek_df. */
int slo_postings_peek_df(const uint8_t *buf, size_t len, uint32_t *df, uint32_t *block_count);
int slo_postings_decode(const uint8_t *buf, size_t len, int keep_positions, uint32_t *docs, uint32_t *tfs,
                        float *max_tf, uint32_t *block_size, uint32_t *blk_max_doc, float *blk_max_tf,
                        uint32_t *n_blocks_out, size_t *consumed);

/* ---- index (one segment) ---- */
slo_index_t *slo_index_new(uint32_t segment_ord, uint32_t doc_count, float k1, float b);
void slo_index_free(slo_index_t *);
/* CSR postings, BORROWED (caller keeps arrays alive).  docs ascending per term. */
int slo_index_set_postings(slo_index_t *, uint64_t n_terms, const uint64_t *offsets, const uint32_t *docs,
                           const uint32_t *tfs);
/* `_len:<field>` i64 fast column → f32 doc lengths (api/reader.rs:3604-3621) and
 * avgdl = (sum as f32)/(doc_count as f32) (index/segment.rs:946-957).  present may be NULL. */
int slo_index_set_field_lengths(slo_index_t *, const int64_t *lens, const uint8_t *present, uint64_t total_tokens);
int slo_index_set_deleted(slo_index_t *, const uint32_t *docs, uint32_t n);
/* also build the reference's varint `.post` image per term so SLO "faithful" timing can re-decode per query */
/* positions of every posting (BORROWED; CSR over the postings): the image is then written with positions on */
int slo_index_set_positions(slo_index_t *, const uint64_t *pos_offsets, const uint32_t *positions);
int slo_index_build_post_image(slo_index_t *);
uint64_t slo_index_post_image_size(const slo_index_t *);
const uint8_t *slo_index_post_image(const slo_index_t *);
const uint64_t *slo_index_post_offsets(const slo_index_t *); /* n_terms+1 */
float slo_index_avgdl(const slo_index_t *);
float slo_index_live_docs(const slo_index_t *);
float slo_index_min_doc_len(const slo_index_t *);
/* fast-field columns (index/fastfields.rs:910-1039), COPIED. return column handle >=0 */
int32_t slo_index_add_i64_column(slo_index_t *, const int64_t *values, const uint8_t *present);
int32_t slo_index_add_f64_column(slo_index_t *, const double *values, const uint8_t *present);
/* ords: UINT32_MAX = missing */
int32_t slo_index_add_str_column(slo_index_t *, const char *const *dict, uint32_t n_dict, const uint32_t *ords);
/* list columns (Column::I64List / F64List / StrList): offsets[doc_count + 1], "any value" semantics (fastfields.rs:490-657) */
int32_t slo_index_add_i64_list_column(slo_index_t *, const uint32_t *offsets, const int64_t *values);
int32_t slo_index_add_f64_list_column(slo_index_t *, const uint32_t *offsets, const double *values);
int32_t slo_index_add_str_list_column(slo_index_t *, const char *const *dict, uint32_t n_dict, const uint32_t *offsets,
                                      const uint32_t *ords);

/* ---- search (api/reader.rs:2908-3128 search_segment + query/wand.rs:398-456) ----
 * k is the internal k (limit+1 semantics are the caller's).  block_size 0 => 128.
 * faithful_decode != 0: per query re-decode each term's varint list and rebuild the per-query
 * doc-length vector, as the reference does (requires slo_index_build_post_image).
 * out_hits has room for k entries, sorted score desc, doc asc (finalize_heap).  Returns the
 * number of hits, <0 on error. */
int32_t slo_search(const slo_index_t *, const slo_query_t *q, uint32_t k, int exec, uint32_t block_size,
                   const slo_filter_node_t *filter, uint32_t n_filter_nodes, const char *const *strings,
                   int faithful_decode, slo_hit_t *out_hits, slo_stats_t *stats);

int slo_max_threads(void);
/* batch, std::thread-parallel over queries when threads>1.  out_hits is n_queries*k, out_counts n_queries. */
int32_t slo_search_batch(const slo_index_t *, const slo_query_t *qs, uint32_t n_queries, uint32_t k, int exec,
                         uint32_t block_size, const slo_filter_node_t *filter, uint32_t n_filter_nodes,
                         const char *const *strings, int faithful_decode, int threads, slo_hit_t *out_hits,
                         uint32_t *out_counts, slo_stats_t *out_stats /* nullable, n_queries */);

/* merge per-segment hit lists as api/reader.rs:2777 (SortKey: score desc total_cmp, segment_ord asc,
 * doc_id asc, query/sort.rs:80-93) and truncate to limit. Returns count. */
/* ScorePlan::evaluate (query/planner.rs:133-164) on a postfix plan and explicit leaf scores */
float slo_plan_evaluate(const slo_plan_node_t *plan, uint32_t n_nodes, const float *leaves, uint32_t n_leaves);

uint32_t slo_merge_hits(const slo_hit_t *hits, uint32_t n, uint32_t limit, slo_hit_t *out);

/* evaluate the root filter for every doc into a bitmap (1 bit/doc, LSB first) — checker for the filter kernel */
int slo_filter_bitmap(const slo_index_t *, const slo_filter_node_t *filter, uint32_t n_filter_nodes,
                      const char *const *strings, uint32_t *bitmap_out /* ceil(doc_count/32) words */);

/* ---- phrases (query/phrase.rs:4-48) ---- */
/* matches_phrase on the position lists of one doc (lists in phrase order, ascending) */
int slo_matches_phrase_positions(uint32_t n_terms, const uint32_t *const *positions, const uint32_t *counts, uint32_t slop);
/* matches_phrase for every doc of a CSR segment with positions (pos_offsets: one entry per posting + 1) -> bitmap */
int slo_phrase_bitmap(uint32_t doc_count, uint64_t n_terms, const uint64_t *term_offsets, const uint32_t *docs,
                      const uint64_t *pos_offsets, const uint32_t *positions, const uint32_t *phrase_terms, uint32_t n_phrase,
                      uint32_t slop, uint32_t *bitmap_out);

/* ---- vectors (vectors/mod.rs:74-129, api/reader.rs:218-254) ---- */
enum { SLO_METRIC_COSINE = 0, SLO_METRIC_L2 = 1 };
void slo_normalize_in_place(float *v, size_t dim);
float slo_metric_similarity(int metric, const float *a, const float *b, size_t dim);
float slo_blend_scores(float bm25, float vector_score, float alpha, int higher_is_better);
/* single-clause compute_hybrid_score; has_vec=0 => missing_vector_score(metric) */
float slo_hybrid_score(float bm25_score, int has_vec, float vec_score, float alpha, int metric);
/* compute_hybrid_score for n clauses (vec_scores already boosted); returns the final score */
float slo_hybrid_score_clauses(float bm25_score, uint32_t n_clauses, const int *has_vec, const float *vec_scores, const float *alpha,
                               const int *metric, float *vector_sum_out, int *has_vector_out);
/* hybrid rescoring + re-sort of per-query BM25 candidates with exact similarities (api/reader.rs:2477-2537); one vector
 * store (offsets u32[doc_count], rows f32[n_rows][dim]) per listed segment_ord */
int slo_rerank_batch(uint32_t n_queries, uint32_t stride, const slo_hit_t *cands, const uint32_t *counts, uint32_t n_segs,
                     const uint32_t *seg_ords, const uint32_t *seg_doc_counts, const uint32_t *const *seg_offsets,
                     const float *const *seg_values, const uint64_t *seg_rows, uint32_t dim, uint32_t n_clauses,
                     const float *const *clause_qv, const float *alpha, const float *boost, const int *metric, slo_hit_t *out_hits,
                     uint32_t *out_counts, float *out_vs, int threads);
/* round f32 values to bf16 precision in place (the engine's storage option; round to nearest even) */
void slo_round_bf16(float *v, size_t n);

#ifdef __cplusplus
}
#endif
#endif
