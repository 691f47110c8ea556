// slg_warp_kernel.cuh — K2/K3, warp-autonomous variant with one accumulator slot per doc of a sub-tile: Bool
// queries, ScorePlans, per-query statistics and kernel choice 2 (the reference's summation order) for k <= 32
// and <= 8 terms per query.  Plain OR queries run on the posting-driven kernel of slg_items_kernel.cuh.
//
// Same arithmetic, same key order and same per-query global top-k protocol as
// slg_score_tiles_kernel, but the unit of cooperation is one WARP instead of one CTA, so no block
// barrier sits between a query's terms and a short posting range stalls 32 lanes, not 256:
//
//   work item   = (group of kSubPerGroup consecutive doc sub-tiles, query), handed out tile-major
//                 from a global counter, one item per warp at a time
//   accumulator = sub_docs f32 slots of shared memory private to the warp
//   candidates  = 64 keys of shared memory private to the warp; when more than 32 are pending the
//                 warp sorts them, keeps the best k and raises its local threshold (exact: nothing
//                 is ever dropped unsorted)
//   merge       = as in the CTA kernel: per-query lock, sort(local ∪ global)[:k], publish the new
//                 k-th key as the query's threshold
//
// Decode + score is split from accumulation (STAGED): the unit-weight BM25 contribution of every
// posting is computed once at segment load (seg.post_score, slg_score_postings_kernel: tf byte or
// wide tf, norm gather, the reference's arithmetic) into an f32 stream laid out like the posting
// array, and the accumulate loop of each (query, tile) then only streams (doc, score) pairs and
// adds them.  The per-query weight is applied at accumulate time (score_tf: base * weight,
// query/wand.rs:284-285), so results are bit-identical to computing the contribution in place.
//
// The kernel is bound by the shared-memory pipe (profiles/r1_v7_warp_kernel_summary.txt); two things keep that
// traffic down: one shared-memory instruction touches 32 consecutive postings (fewer bank conflicts on dense
// lists), and every lane tracks the largest value it wrote so that a sub-tile whose best score is below the
// query's threshold is cleared without being read back.
#pragma once
#include "slg_kernels.cuh"

namespace slg {

constexpr uint32_t kWarpMaxTerms = 8;   // per-query term slots of the padded QTerm table
constexpr uint32_t kWarpMaxK = 32;
constexpr uint32_t kSubPerGroup = 8;    // sub-tiles per work item
constexpr uint32_t kWarpCand = 64;

struct __align__(16) QTerm {  // one query term resolved against one segment (32 B)
  uint64_t base;      // term_start: first padded posting index
  uint64_t sc_base;   // element offset of the term's dense column in seg.cols, ~0 = the term has none
  uint32_t term;      // term id in the segment (a term streamed from its column, flags bit2: the column index)
  uint32_t uterm;     // row of the range / bound tables
  float weight;
  uint32_t flags;     // bit0 scored, bit1 valid, bit2 streamed from its dense column (items kernel), bits 8..15 group, 16..19 leaf
};

struct __align__(16) QHead {  // 16 B per query slot (processing order)
  uint32_t qi;        // original query index (state + output position)
  uint32_t nt;
  uint32_t masks;     // must | not<<8 | should<<16 | min_should<<24
  int32_t filter;
};

struct WarpBatchDev {
  const QTerm *qterms;   // [Q][kWarpMaxTerms]
  const QHead *qheads;   // [Q]
  const uint32_t *rng;   // [U][n_sub+1]
  const float *sub_ub;   // [U][n_sub]   (PRUNE)
  const float *scores;   // seg.post_score: resident unit-weight contributions (STAGED)
  const uint32_t *const *filter_bits;
  uint32_t n_queries, k, sub_docs, n_sub, n_groups;
  float ms_frac;         // MaxScore: the non-essential bounds may sum to at most this fraction of the k-th score
  unsigned long long *thr_key;
  uint32_t *topk_count, *lock;
  unsigned long long *topk_keys;
  uint32_t *work_counter;
  unsigned long long *stats;
  const unsigned long long *q_cursor;  // as in BatchDev
  uint32_t *q_saw;
  // ScorePlans (PLAN kernels): as in BatchDev; a term's leaf travels in QTerm.flags bits 16..19
  const uint8_t *q_leaves;
  const uint32_t *q_plan_off;
  const PlanNodeDev *plan_nodes;
  uint32_t max_leaves;
  unsigned long long *match_count;  // [Q] accepted docs (STATS)
  // k > kWarpMaxK (flat scan only): per query two candidate pools of pool_cap keys (0: posting scan, 1: column pass), each
  // with its own count and lock; slg_finalize_pools_kernel merges them.  nullptr: the query's sorted top-k in topk_keys.
  unsigned long long *pool_keys;    // [Q][2][pool_cap]
  uint32_t *pool_count, *pool_lock; // [Q][2]
  uint32_t pool_cap;
  // Sharded runs: the threshold board (slg_batch_set_threshold_board).  board[q] is THIS shard's slot for query q in
  // peer-accessible memory (NVLink / NVSwitch peer mappings): epoch << 32 | score bits of a k-th key (positive floats order
  // like their bits, and a newer epoch — the next batch — always wins, so the board is never cleared).  A shard that raises
  // a query's k-th key pushes it into every peer's board with a system-scope atomic max (fire and forget) and reads its own
  // board, a local access, whenever it picks the query up again — the best k-th score any shard has found is a valid lower
  // bound of the global one, so every shard prunes against (nearly) the global threshold without a kernel boundary, a
  // collective or the host in the loop.  nullptr: no exchange.
  unsigned long long *board;
  unsigned long long *peer_board[7];
  uint32_t n_peers, epoch;
};

constexpr uint32_t kMaxBoardPeers = 7;

// the query's k-th key as far as this shard knows: its own, or the bound a peer pushed (same epoch only; the doc half is
// dropped: across shards ties on the score are decided by segment_ord, so only the score transfers)
__device__ __forceinline__ unsigned long long load_threshold(const WarpBatchDev &wb, uint32_t qi) {
  unsigned long long t = ld_cg_u64(wb.thr_key + qi);
  if (wb.board) {
    const unsigned long long b = ld_cg_u64(wb.board + qi);
    if ((uint32_t)(b >> 32) == wb.epoch) t = max(t, b << 32);
  }
  return t;
}
// one lane: a raised k-th key goes to every peer's board
__device__ __forceinline__ void push_threshold(const WarpBatchDev &wb, uint32_t qi, unsigned long long key) {
  if (!wb.board) return;
  const unsigned long long v = ((unsigned long long)wb.epoch << 32) | (key >> 32);
  for (uint32_t p = 0; p < wb.n_peers; p++) atomicMax_system(wb.peer_board[p] + qi, v);
}

// resolve the batch's query terms against one segment (runs once per segment per batch).
// canonical: the term slots of a query are laid out in the DECLARED summation order of the column path
// (include/searchlite_gpu.h): terms without a dense column first, then the terms with one, each group in query
// order — every kernel that walks the slots in order then sums in that order.  use_cols: the terms with a column
// are flagged (bit2) and the items kernel streams them from seg.cols.
static __global__ void slg_build_qterms_kernel(SegmentDev seg, BatchDev bt, QTerm *qterms, QHead *qheads, bool canonical, bool use_cols) {
  const uint32_t slot = blockIdx.x * blockDim.x + threadIdx.x;
  if (slot >= bt.n_queries) return;
  const uint32_t qi = bt.q_order[slot];
  const uint32_t t0 = bt.q_term_off[qi], nt = bt.q_term_off[qi + 1] - t0;
  QHead h;
  h.qi = qi;
  h.nt = nt;
  h.masks = (uint32_t)bt.q_must[qi] | ((uint32_t)bt.q_not[qi] << 8) | ((uint32_t)bt.q_should[qi] << 16) |
            ((uint32_t)bt.q_min_should[qi] << 24);
  h.filter = bt.q_filter[qi];
  qheads[slot] = h;
  uint32_t out = 0;
  for (int pass = 0; pass < (canonical ? 2 : 1); pass++) {
    for (uint32_t t = 0; t < nt && t < kWarpMaxTerms; t++) {
      const uint32_t u = bt.qt_uterm[t0 + t];
      const uint32_t term = bt.ut_term[u];
      const bool scored = bt.qt_flags[t0 + t] & 1u;
      const bool has_col = seg.term_col && term < seg.n_terms && seg.term_col[term] >= 0 && scored;
      if (canonical && has_col != (pass == 1)) continue;
      QTerm r;
      r.base = term < seg.n_terms ? seg.term_start[term] : 0;  // seg.post_score is laid out like seg.post_doc
      r.sc_base = has_col ? (uint64_t)seg.term_col[term] * seg.col_stride : ~0ull;
      r.term = term;
      r.uterm = u;
      r.weight = bt.qt_weight[t0 + t];
      r.flags = (scored ? 1u : 0u) | 2u | ((uint32_t)bt.qt_group[t0 + t] << 8);
      if (bt.qt_leaf) r.flags |= (uint32_t)bt.qt_leaf[t0 + t] << 16;
      if (has_col && use_cols) {
        r.flags |= 4u;
        r.term = (uint32_t)seg.term_col[term];  // streamed terms are addressed by their column, never by their term id
      }
      qterms[(uint64_t)slot * kWarpMaxTerms + out++] = r;
    }
  }
  for (; out < kWarpMaxTerms; out++) {
    QTerm r;
    r.base = 0;
    r.sc_base = 0;
    r.term = 0;
    r.uterm = 0;
    r.weight = 0.0f;
    r.flags = 0;
    qterms[(uint64_t)slot * kWarpMaxTerms + out] = r;
  }
}

constexpr uint32_t kRbStride = 9;
__host__ __device__ inline size_t warp_kernel_smem_per_warp(uint32_t sub_docs, bool matcher, bool prune, uint32_t planes = 1) {
  // (every term is a multiple of 16 B)
  return (size_t)sub_docs * 4 * planes + kWarpCand * 8 + kWarpMaxTerms * sizeof(QTerm) + kWarpMaxTerms * kRbStride * 4 +
         (prune ? kWarpMaxTerms * 8 * 4 : 0) + (matcher ? sub_docs : 0);
}

// 64-key descending bitonic sort in the warp's shared buffer
__device__ __forceinline__ void warp_sort64_desc(unsigned long long *a, int lane) {
#pragma unroll 1
  for (uint32_t size = 2; size <= kWarpCand; size <<= 1) {
#pragma unroll 1
    for (uint32_t stride = size >> 1; stride > 0; stride >>= 1) {
      const uint32_t i = ((lane & ~(stride - 1)) << 1) | (lane & (stride - 1));
      const uint32_t j = i + stride;
      const unsigned long long x = a[i], y = a[j];
      const bool desc = (i & size) == 0;
      if (desc ? (x < y) : (x > y)) {
        a[i] = y;
        a[j] = x;
      }
      __syncwarp();
    }
  }
}

// postings [lo, hi) of one term from the resident (doc, score) streams into the warp's accumulator.
// Lane L takes postings L, L + 32, L + 64, L + 96 of every group of 128: one shared-memory instruction
// then touches 32 CONSECUTIVE postings, whose docs lie close together — for the dense head terms that
// carry most of the work they spread over the 32 banks far better than every fourth posting does.
// The next group's loads are issued before this group's read-modify-writes.
// wmax collects (as bits; scores are positive) the largest value this lane wrote: contributions are
// positive, so the largest value ever written to the accumulator is the largest final score of the
// sub-tile, and the scan for candidates can be skipped when it is below the threshold.
template <bool FIRST, bool UNIT_W>
__device__ __forceinline__ void accumulate_staged(const uint32_t *__restrict__ dptr, const float *__restrict__ sptr, uint32_t lo,
                                                  uint32_t hi, uint32_t tile_lo, float w, float *acc, int lane, uint32_t &wmax) {
  if (hi - lo <= 32u) {
    // short range: one posting per lane per step
    for (uint32_t i = lo + lane; i < hi; i += 32) {
      float v = __ldg(sptr + i);
      if (!UNIT_W) v = __fmul_rn(v, w);
      float *p = acc + (__ldg(dptr + i) - tile_lo);
      if (!FIRST) v = __fadd_rn(*p, v);
      *p = v;
      wmax = max(wmax, __float_as_uint(v));
    }
    return;
  }
  uint32_t i = lo + lane;
  uint32_t d[4];
  float s[4];
#pragma unroll
  for (int k = 0; k < 4; k++) {
    d[k] = tile_lo;
    s[k] = 0.0f;
    if (i + 32 * k < hi) {
      d[k] = __ldg(dptr + i + 32 * k);
      s[k] = __ldg(sptr + i + 32 * k);
    }
  }
  for (;;) {
    const uint32_t inext = i + 128;
    uint32_t dn[4];
    float sn[4];
#pragma unroll
    for (int k = 0; k < 4; k++) {
      dn[k] = tile_lo;
      sn[k] = 0.0f;
      if (inext + 32 * k < hi) {
        dn[k] = __ldg(dptr + inext + 32 * k);
        sn[k] = __ldg(sptr + inext + 32 * k);
      }
    }
#pragma unroll
    for (int k = 0; k < 4; k++) {
      if (i + 32 * k < hi) {
        float v = s[k];
        if (!UNIT_W) v = __fmul_rn(v, w);
        float *p = acc + (d[k] - tile_lo);
        if (!FIRST) v = __fadd_rn(*p, v);  // distinct docs inside a list: no aliasing between lanes
        *p = v;
        wmax = max(wmax, __float_as_uint(v));
      }
    }
    if (inext - lane >= hi) break;
    i = inext;
#pragma unroll
    for (int k = 0; k < 4; k++) {
      d[k] = dn[k];
      s[k] = sn[k];
    }
  }
}

// PRUNE on plain OR queries with resident scores adds MaxScore on top of the tile skip.  Per (query,
// sub-tile) the terms are ranked by their bound inside the sub-tile; the longest prefix whose bounds
// sum to less than the running k-th score is "non-essential": a doc that holds only such terms cannot
// enter the top k, so their postings are not scattered at all (the prefix is also capped at a fraction
// of the threshold, ms_frac: with loose bounds too many docs would need the exact rescoring below).
// A doc touched by the other terms is a
// candidate if its partial score plus the non-essential bounds can still reach the threshold; its
// exact score is then recomputed over ALL terms in slot order — dense-column lookup for column terms,
// binary search inside the sub-tile's posting range otherwise — so the result is bit-identical to
// the exhaustive run.  (TermState upper bounds: query/wand.rs:238-303; the reference's wand_loop
// prunes document-at-a-time with the same bounds, query/wand.rs:659-903.)
//
// PLAN (ScorePlans): the warp's accumulator holds wb.max_leaves planes of sub_docs floats; a scored
// term scatters into the plane of its leaf, one pass per sub-tile evaluates the query's postfix ScoreExpr per
// touched doc into plane 0 and clears the other planes, and the scan proceeds on plane 0.  MaxScore (partial
// sums) and the written-maximum shortcut do not apply to plans; the tile skip does (a plan never exceeds the
// sum of its terms' bounds).
template <bool MATCHER, bool PRUNE, bool STATS, bool STAGED, bool PLAN = false>
__global__ void __launch_bounds__(kThreads, 3) slg_score_warp_kernel(SegmentDev seg, WarpBatchDev wb) {
  constexpr bool MAXSCORE = PRUNE && !MATCHER && STAGED && !PLAN;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  constexpr int kWarps = kThreads / 32;
  const uint32_t sub_docs = wb.sub_docs;
  // per-warp layout: acc f32[sub_docs] | cand u64[64] | qt QTerm[8] | rb u32[8][9] | ub f32[8][8] (PRUNE) | gmask u8[sub_docs] (MATCHER)
  const uint32_t n_planes = PLAN ? wb.max_leaves : 1u;
  const size_t per_warp = warp_kernel_smem_per_warp(sub_docs, MATCHER, PRUNE, n_planes);
  unsigned char *mine = smem_raw + (size_t)warp * per_warp;
  float *acc = reinterpret_cast<float *>(mine);
  unsigned long long *cand = reinterpret_cast<unsigned long long *>(mine + (size_t)sub_docs * 4 * n_planes);
  QTerm *qt = reinterpret_cast<QTerm *>(cand + kWarpCand);
  uint32_t *rb = reinterpret_cast<uint32_t *>(qt + kWarpMaxTerms);   // [t][9], boundary j at rb[t*9 + j]
  float *ubs = reinterpret_cast<float *>(rb + kWarpMaxTerms * kRbStride);   // [t][8]
  uint8_t *gmask = reinterpret_cast<uint8_t *>(ubs + (PRUNE ? kWarpMaxTerms * 8 : 0));
  // MaxScore: up to 64 pending doc ids share the upper half of the candidate buffer, which is only
  // written by the sort inside push_keys — and that never runs while doc ids are parked there
  uint32_t *pend = reinterpret_cast<uint32_t *>(cand + 32);
  (void)pend;
  (void)kWarps;

  const uint32_t k = wb.k;
  const uint32_t total_items = wb.n_groups * wb.n_queries;
  const uint32_t lt_mask = (1u << lane) - 1u;

  for (uint32_t i = lane * 4; i < sub_docs * n_planes; i += 128) *reinterpret_cast<float4 *>(acc + i) = make_float4(0, 0, 0, 0);
  if (MATCHER)
    for (uint32_t i = lane * 4; i < sub_docs; i += 128) *reinterpret_cast<uint32_t *>(gmask + i) = 0u;
  __syncwarp();

  uint32_t item = 0;
  if (lane == 0) item = atomicAdd(wb.work_counter, 1u);
  item = __shfl_sync(0xFFFFFFFFu, item, 0);

  while (item < total_items) {
    uint32_t next_item = 0;
    if (lane == 0) next_item = atomicAdd(wb.work_counter, 1u);  // consumed at the end of this item

    const uint32_t tg = item / wb.n_queries;
    const uint32_t qslot = item - tg * wb.n_queries;
    const QHead head = wb.qheads[qslot];
    const uint32_t nt = head.nt;
    uint32_t n_leaves = 0, n_nodes = 0;  // PLAN: 0 leaves = this query has no plan (running sum in plane 0)
    const PlanNodeDev *nodes = nullptr;
    if (PLAN) {
      n_leaves = wb.q_leaves[head.qi];
      const uint32_t p0 = wb.q_plan_off[head.qi];
      n_nodes = wb.q_plan_off[head.qi + 1] - p0;
      nodes = wb.plan_nodes + p0;
    }
    const PlanInfo pinfo = plan_info(nodes, n_nodes, n_leaves);
    (void)pinfo;
    // stage the query's term records: 8 records x 32 B = 16 lanes x 16 B
    if (lane < 16) {
      const uint4 v = __ldg(reinterpret_cast<const uint4 *>(wb.qterms + (uint64_t)qslot * kWarpMaxTerms) + lane);
      reinterpret_cast<uint4 *>(qt)[lane] = v;
    }
    unsigned long long thr = ld_cg_u64(wb.thr_key + head.qi);
    __syncwarp();
    const uint32_t sub0 = tg * kSubPerGroup;
    // range boundaries sub0 .. sub0+8 of every term: lane = t*4 + c loads boundaries c, c+4, c+8
    {
      const uint32_t t = lane >> 2, c = lane & 3;
      if (t < nt) {
        const uint32_t *row = wb.rng + (uint64_t)qt[t].uterm * (wb.n_sub + 1);
        for (uint32_t j = c; j <= kSubPerGroup; j += 4) rb[t * kRbStride + j] = __ldg(row + min(sub0 + j, wb.n_sub));
        if (PRUNE) {
          const float *urow = wb.sub_ub + (uint64_t)qt[t].uterm * wb.n_sub;
          for (uint32_t j = c; j < kSubPerGroup; j += 4) ubs[t * 8 + j] = (sub0 + j < wb.n_sub) ? __ldg(urow + sub0 + j) : 0.0f;
        }
      }
    }
    __syncwarp();

    uint32_t cnt = 0;  // pending candidates in cand[]
    uint32_t n_touched = 0, n_post = 0, n_skipped = 0, n_match = 0;
    const uint32_t masks = head.masks;

#pragma unroll 1
    for (uint32_t j = 0; j < kSubPerGroup; j++) {
      const uint32_t sub = sub0 + j;
      if (sub >= wb.n_sub) break;
      const uint32_t tile_lo = sub * sub_docs;
      const uint32_t tile_n = min(sub_docs, seg.doc_count - tile_lo);
      // total scored postings (and bound) of this sub-tile: lanes 0..nt-1 hold one term each
      uint32_t mine_n = 0;
      float mine_ub = 0.0f;
      if (lane < (int)nt && (qt[lane].flags & 1u)) {
        mine_n = rb[lane * kRbStride + j + 1] - rb[lane * kRbStride + j];
        if (PRUNE) mine_ub = ubs[lane * 8 + j] * qt[lane].weight;
      }
      const uint32_t any = __ballot_sync(0xFFFFFFFFu, mine_n != 0u);
      if (any == 0u) continue;
      if (PRUNE) {
        float ub = mine_ub;
#pragma unroll
        for (int o = 4; o > 0; o >>= 1) ub += __shfl_xor_sync(0xFFFFFFFFu, ub, o);  // terms live in lanes 0..7
        ub = __shfl_sync(0xFFFFFFFFu, ub, 0);
        const float thr_score = __uint_as_float((uint32_t)(thr >> 32));
        if (thr != kThrInit && ub * 1.00001f < thr_score) {  // see slg_score_tiles_kernel
          n_skipped++;
          continue;
        }
      }
      // ---- MaxScore: the non-essential terms of this sub-tile ----
      uint32_t nmask = 0;
      float sum_n = 0.0f;
      if (MAXSCORE && thr != kThrInit) {
        const float thr_score = __uint_as_float((uint32_t)(thr >> 32));
        float pre = 0.0f;  // bounds of the terms ranked at or below this lane's term, ascending by bound
#pragma unroll
        for (int m = 0; m < (int)kWarpMaxTerms; m++) {
          const float u = __shfl_sync(0xFFFFFFFFu, mine_ub, m);
          if (u < mine_ub || (u == mine_ub && m <= lane)) pre += u;
        }
        const bool noness = mine_n != 0u && pre * 1.00001f < thr_score * wb.ms_frac;
        nmask = __ballot_sync(0xFFFFFFFFu, noness);
        float sn = noness ? pre : 0.0f;
#pragma unroll
        for (int o = 4; o > 0; o >>= 1) sn = fmaxf(sn, __shfl_xor_sync(0xFFFFFFFFu, sn, o));
        sum_n = __shfl_sync(0xFFFFFFFFu, sn, 0);
      }
      if (STATS) n_post += ((nmask >> lane) & 1u) ? 0u : mine_n;

      // ---- accumulate ----
      bool first = true;
      uint32_t wmax = 0;  // largest value written to the accumulator by this lane (bits)
#pragma unroll 1
      for (uint32_t t = 0; t < nt; t++) {
        const uint32_t lo = rb[t * kRbStride + j], hi = rb[t * kRbStride + j + 1];
        if (hi <= lo) continue;
        if (MAXSCORE && ((nmask >> t) & 1u)) continue;
        const QTerm q = qt[t];
        const bool scored = q.flags & 1u;
        float *plane = acc;
        if (PLAN && n_leaves) plane += (size_t)((q.flags >> 16) & 15u) * sub_docs;
        if (PLAN) first = false;  // planes: always add (0 + s == s)
        if (STAGED && !MATCHER) {
          const uint32_t *dptr = seg.post_doc + q.base;
          const float *sptr = wb.scores + q.base;
          if (q.weight == 1.0f) {
            if (first) accumulate_staged<true, true>(dptr, sptr, lo, hi, tile_lo, 1.0f, plane, lane, wmax);
            else accumulate_staged<false, true>(dptr, sptr, lo, hi, tile_lo, 1.0f, plane, lane, wmax);
          } else {
            if (first) accumulate_staged<true, false>(dptr, sptr, lo, hi, tile_lo, q.weight, plane, lane, wmax);
            else accumulate_staged<false, false>(dptr, sptr, lo, hi, tile_lo, q.weight, plane, lane, wmax);
          }
        } else {
          TermCtx tc;
          const uint64_t wide = seg.term_wide[q.term];
          tc.dptr = seg.post_doc + q.base;
          tc.fptr = seg.post_tf + q.base;
          tc.wptr = wide != ~0ull ? seg.tf_wide + wide : nullptr;
          tc.scored = scored;
          tc.nk = seg_nk(seg, q.term);
          tc.idf = seg.term_idf[q.term];
          tc.w = q.weight;
          tc.gbit = MATCHER ? (uint8_t)(1u << ((q.flags >> 8) & 7u)) : 0;
          if (first && scored && !MATCHER) accumulate_term<MATCHER, true, 32>(seg, tc, lo, hi, tile_lo, plane, gmask, lane);
          else accumulate_term<MATCHER, false, 32>(seg, tc, lo, hi, tile_lo, plane, gmask, lane);
          wmax = 0xFFFFFFFFu;  // not tracked on this path
        }
        if (scored) first = false;
        __syncwarp();
      }

      if (PLAN && n_leaves) {
        // a plan's score is at most the sum of its leaves (tie_breaker <= 1), and every leaf is at most the largest value
        // written to a plane: if n_leaves x that maximum cannot reach the threshold, the sub-tile is cleared unread
        if (STAGED && !MATCHER && !STATS && thr != kThrInit) {
          const float top = __uint_as_float(__reduce_max_sync(0xFFFFFFFFu, wmax));
          if (top * (float)n_leaves * 1.00001f < __uint_as_float((uint32_t)(thr >> 32))) {
            for (uint32_t l = 0; l < n_leaves; l++)
#pragma unroll 4
              for (uint32_t i0 = 0; i0 < tile_n; i0 += 128) *reinterpret_cast<float4 *>(acc + (size_t)l * sub_docs + i0 + lane * 4) = make_float4(0, 0, 0, 0);
            __syncwarp();
            continue;
          }
        }
        wmax = 0xFFFFFFFFu;  // the scan's own shortcut compares single written values: not valid for a plan
        {
          // ---- ScorePlan: leaves -> score in plane 0, other planes cleared (plan.evaluate, query/wand.rs:506) ----
#pragma unroll 1
          for (uint32_t i0 = 0; i0 < tile_n; i0 += 128) plan_combine_quad(pinfo, acc, sub_docs, i0 + lane * 4);
          __syncwarp();
        }
      }
      // ---- scan + clear; collect keys that beat the threshold ----
      uint32_t thr_hi = (uint32_t)(thr >> 32);
      if (MAXSCORE && nmask) {
        // a partial score below this cannot reach the threshold even with every non-essential bound added
        const float cut = __uint_as_float(thr_hi) * 0.99998f - sum_n * 1.00002f;
        thr_hi = cut > 0.0f ? __float_as_uint(cut) : 0u;
      }
      if (!STATS && !MATCHER && __reduce_max_sync(0xFFFFFFFFu, wmax) < thr_hi) {
        // no score of this sub-tile reaches the threshold: nothing to collect, only clear
#pragma unroll 4
        for (uint32_t i0 = 0; i0 < tile_n; i0 += 128) *reinterpret_cast<float4 *>(acc + i0 + lane * 4) = make_float4(0, 0, 0, 0);
        __syncwarp();
        continue;
      }
      // append one ballot round of keys to the warp's candidates; past 32 pending: sort, keep the best k
      auto push_keys = [&](bool pass, unsigned long long key) {
        const uint32_t bal = __ballot_sync(0xFFFFFFFFu, pass);
        if (!bal) return;
        if (pass) cand[cnt + __popc(bal & lt_mask)] = key;
        cnt += __popc(bal);
        __syncwarp();
        if (cnt > 32) {
          // keep the best k of everything seen so far in this item; raise the local threshold
          for (uint32_t z = cnt + lane; z < kWarpCand; z += 32) cand[z] = 0ull;
          __syncwarp();
          warp_sort64_desc(cand, lane);
          cnt = min(cnt, k);
          if (cnt == k) thr = max(thr, cand[k - 1]);
          __syncwarp();
        }
      };
      // MaxScore: exact score of one parked doc per lane — every term of the query at the doc, query
      // order: column lookup or binary search inside the sub-tile's posting range
      uint32_t npend = 0;
      auto rescore = [&](bool have, uint32_t doc) {
        float s = 0.0f;
        if (have) {
          for (uint32_t t = 0; t < nt; t++) {
            const QTerm &q = qt[t];
            if (!(q.flags & 1u)) continue;
            float c = 0.0f;
            if (q.sc_base != ~0ull) {
              c = __ldg(seg.cols + q.sc_base + doc);
            } else {
              const uint32_t *dp = seg.post_doc + q.base;
              const uint32_t end = rb[t * kRbStride + j + 1];
              uint32_t lo = rb[t * kRbStride + j], hi = end;
              while (lo < hi) {
                const uint32_t mid = (lo + hi) >> 1;
                if (__ldg(dp + mid) < doc) lo = mid + 1;
                else hi = mid;
              }
              if (lo < end && __ldg(dp + lo) == doc) c = __ldg(wb.scores + q.base + lo);
            }
            if (c != 0.0f) s = __fadd_rn(s, __fmul_rn(c, q.weight));
          }
        }
        const unsigned long long key = ((unsigned long long)__float_as_uint(s) << 32) | (unsigned long long)(0xFFFFFFFFu - doc);
        bool pass = have && key > thr;
        if (pass) pass = (seg.live_bits[doc >> 5] >> (doc & 31)) & 1u;
        if (pass && head.filter >= 0) pass = (wb.filter_bits[head.filter][doc >> 5] >> (doc & 31)) & 1u;
        if (pass) pass = cursor_accepts(wb.q_cursor, wb.q_saw, head.qi, key);
        push_keys(pass, key);
      };
      // every parked doc (at most 64): read them all before the first push may reuse the buffer
      auto rescore_pending = [&]() {
        const uint32_t n = npend;
        npend = 0;
        const uint32_t d0 = lane < (int)n ? pend[lane] : 0u;
        const uint32_t d1 = 32u + lane < n ? pend[32 + lane] : 0u;
        __syncwarp();
        rescore(lane < (int)n, d0);
        if (n > 32u) rescore(32u + lane < n, d1);
      };
#pragma unroll 1
      for (uint32_t i0 = 0; i0 < tile_n; i0 += 128) {
        const uint32_t i = i0 + lane * 4;  // (reads past tile_n stay inside the warp's buffer and are zero)
        const float4 v = *reinterpret_cast<const float4 *>(acc + i);
        const uint32_t b0 = __float_as_uint(v.x), b1 = __float_as_uint(v.y), b2 = __float_as_uint(v.z), b3 = __float_as_uint(v.w);
        const uint32_t m = max(max(b0, b1), max(b2, b3));
        uint32_t gm = 0;
        if (m != 0u) *reinterpret_cast<float4 *>(acc + i) = make_float4(0, 0, 0, 0);
        if (STATS) n_touched += (b0 != 0u) + (b1 != 0u) + (b2 != 0u) + (b3 != 0u);
        if (MATCHER) {
          gm = *reinterpret_cast<const uint32_t *>(gmask + i);
          if (gm) *reinterpret_cast<uint32_t *>(gmask + i) = 0u;
        }
        if (STATS && m != 0u) {
          // the accept closure for every scored doc (api/reader.rs:3009-3031): the match counter behind
          // total_hits_estimate.  Under MaxScore the scores seen here are partial and skipped terms' docs are
          // absent, so — as in the reference under pruning — the count is an estimate.
          const uint32_t sb[4] = {b0, b1, b2, b3};
#pragma unroll
          for (int e = 0; e < 4; e++) {
            if (sb[e] == 0u) continue;
            const uint32_t doc = tile_lo + i + e;
            bool pass = (seg.live_bits[doc >> 5] >> (doc & 31)) & 1u;
            if (pass && MATCHER) {
              const uint32_t mm = (gm >> (8 * e)) & 255u;
              const uint32_t must = masks & 255u, nots = (masks >> 8) & 255u, should = (masks >> 16) & 255u;
              pass = ((mm & must) == must) && ((mm & nots) == 0u) && (__popc(mm & should) >= (int)(masks >> 24));
            }
            if (pass && head.filter >= 0) pass = (wb.filter_bits[head.filter][doc >> 5] >> (doc & 31)) & 1u;
            if (pass && wb.q_cursor)
              pass = (((unsigned long long)sb[e] << 32) | (unsigned long long)(0xFFFFFFFFu - doc)) < __ldg(wb.q_cursor + head.qi);
            n_match += pass ? 1u : 0u;
          }
        }
        if (__any_sync(0xFFFFFFFFu, m >= thr_hi && m != 0u)) {
          const uint32_t bits[4] = {b0, b1, b2, b3};
          if (MAXSCORE && nmask) {
            // partial scores: park the docs that can still make it, rescore them 32 at a time
#pragma unroll
            for (int e = 0; e < 4; e++) {
              const bool pass = bits[e] >= thr_hi && bits[e] != 0u;
              const uint32_t bal = __ballot_sync(0xFFFFFFFFu, pass);
              if (pass) pend[npend + __popc(bal & lt_mask)] = tile_lo + i + e;
              npend += __popc(bal);
              __syncwarp();
              if (npend >= 32u) rescore_pending();
            }
          } else {
#pragma unroll
            for (int e = 0; e < 4; e++) {
              bool pass = bits[e] >= thr_hi && bits[e] != 0u;
              const uint32_t doc = tile_lo + i + e;
              const unsigned long long key = ((unsigned long long)bits[e] << 32) | (unsigned long long)(0xFFFFFFFFu - doc);
              if (pass) pass = key > thr;
              if (pass) pass = (seg.live_bits[doc >> 5] >> (doc & 31)) & 1u;
              if (pass && MATCHER) {
                const uint32_t mm = (gm >> (8 * e)) & 255u;
                const uint32_t must = masks & 255u, nots = (masks >> 8) & 255u, should = (masks >> 16) & 255u;
                pass = ((mm & must) == must) && ((mm & nots) == 0u) && (__popc(mm & should) >= (int)(masks >> 24));
              }
              if (pass && head.filter >= 0) pass = (wb.filter_bits[head.filter][doc >> 5] >> (doc & 31)) & 1u;
              if (pass) pass = cursor_accepts(wb.q_cursor, wb.q_saw, head.qi, key);
              push_keys(pass, key);
            }
          }
        }
      }
      if (MAXSCORE && npend) rescore_pending();
      __syncwarp();
    }

    if (STATS) {
      for (int o = 16; o > 0; o >>= 1) {
        n_touched += __shfl_xor_sync(0xFFFFFFFFu, n_touched, o);
        n_post += __shfl_xor_sync(0xFFFFFFFFu, n_post, o);
        n_match += __shfl_xor_sync(0xFFFFFFFFu, n_match, o);
      }
      if (lane == 0) {
        if (n_match) atomicAdd(wb.match_count + head.qi, (unsigned long long)n_match);
        if (n_touched) atomicAdd(wb.stats + (uint64_t)head.qi * 4 + 0, (unsigned long long)n_touched);
        if (n_post) atomicAdd(wb.stats + (uint64_t)head.qi * 4 + 1, (unsigned long long)n_post);
        if (n_skipped) atomicAdd(wb.stats + (uint64_t)head.qi * 4 + 2, (unsigned long long)n_skipped);
        if (cnt) atomicAdd(wb.stats + (uint64_t)head.qi * 4 + 3, (unsigned long long)cnt);
      }
    }

    // ---- merge into the query's global top-k (push_top_k, query/wand.rs:905-916) ----
    if (cnt > 0) {
      const unsigned long long thr_now = ld_cg_u64(wb.thr_key + head.qi);
      const bool useful = lane < (int)cnt && cand[lane] > thr_now;  // cnt <= 32 after every append
      if (__any_sync(0xFFFFFFFFu, useful)) {
        if (lane == 0) {
          while (atomicCAS(wb.lock + head.qi, 0u, 1u) != 0u) __nanosleep(64);
          __threadfence();
        }
        __syncwarp();
        const uint32_t ng = ld_cg_u32(wb.topk_count + head.qi);
        unsigned long long *gk = wb.topk_keys + (uint64_t)head.qi * k;
        if (lane < (int)ng) cand[cnt + lane] = ld_cg_u64(gk + lane);
        uint32_t total = cnt + ng;
        for (uint32_t z = total + lane; z < kWarpCand; z += 32) cand[z] = 0ull;
        __syncwarp();
        warp_sort64_desc(cand, lane);
        total = min(total, k);
        if (lane < (int)total) st_cg_u64(gk + lane, cand[lane]);
        __threadfence();
        __syncwarp();
        if (lane == 0) {
          st_cg_u32(wb.topk_count + head.qi, total);
          if (total == k) st_cg_u64(wb.thr_key + head.qi, cand[k - 1]);
          __threadfence();
          atomicExch(wb.lock + head.qi, 0u);
        }
        __syncwarp();
      }
    }
    item = __shfl_sync(0xFFFFFFFFu, next_item, 0);
  }
}

}  // namespace slg
