// slg_stream_kernel.cuh — K2/K5, exhaustive execution (`bm25`) of plain OR queries with k <= 32 and <= 8 terms
// per query on segments with resident scores: the automatic choice for the shape of BASELINE.json configs[1].
//
// brute_force (query/wand.rs:459-566) adds every posting of every query term to its doc's score and offers every
// touched doc to push_top_k.  Here that work is split by what a term looks like in HBM:
//
//   SPARSE PASS   slg_score_sparse_kernel — terms without a dense column.  One warp takes a (query, group of
//                 kSubPerGroup sub-tiles) item.  The (doc, score) runs of ALL the query's sparse terms inside a span of
//                 sub-tiles are brought into shared memory by bulk asynchronous copies (cp.async.bulk, completion on
//                 a warp-private mbarrier) — every byte of the item is in flight at once and no register is
//                 spent on the copy — and then, sub-tile by sub-tile, scattered from shared memory into the warp's
//                 accumulator A (slot order = declared summation order) and walked again: read A[doc], write 0,
//                 and a partial that can still reach the query's k-th score completes its score with one 4-byte
//                 gather per column term and is offered.  A query with a single sparse term needs no accumulator:
//                 its staged scores are compared as they are.
//   COLUMN PASS   slg_score_columns_kernel — terms with a doc-indexed f32 column (df >= N / dense_den).  Doc-block
//                 major: a CTA takes a block of kColBlock docs, brings that block of EVERY column the batch
//                 names into shared memory with one bulk asynchronous copy per column (double buffered: the next
//                 block lands while this one is scored) — each column block leaves HBM/L2 once per batch instead of
//                 once per query — and then scores all the batch's queries that have a column term against it, 32
//                 queries per warp step.  A query whose only column term is c sees v = c[d] * w for every doc of the
//                 block, and max(v) = max(c) * w (f32 multiplication by a positive weight is monotone), so its
//                 comparisons against the k-th score collapse into one against the block maximum the CTA took from
//                 the values it has just read; a query with several columns sums their blocks per doc, slot order,
//                 from shared memory.  Only where a doc can enter the top k is a block looked at per doc.  A doc
//                 that also sits in one of the query's sparse lists belongs to the sparse pass (its partial is in A
//                 there) and is left alone here.
//
// Every posting of every query term is read and takes part in the comparison that decides its doc; what is
// shared is the READ of a column block between the queries that name the column, and what is skipped is only the
// search for candidates that provably do not exist (as in slg_score_warp_kernel's written-maximum shortcut).  The
// sparse pass runs first so that the column pass starts with every query's k-th score already high.
//
// Float contract (include/searchlite_gpu.h): a doc's contributions are summed over the query's terms WITHOUT a
// column first, then over the terms WITH one, each group in query order — slg_build_qterms_kernel lays the slots
// out in that order, both passes add in slot order, so both produce the bits of brute_force on that permutation.
#pragma once
#include "slg_async.cuh"
#include "slg_items_kernel.cuh"

namespace slg {

constexpr uint32_t kColBlock = 256;      // docs per column-pass block (1 KB of every column)
constexpr int kSparseWarps = 4;          // warps per CTA of the sparse pass
constexpr int kColWarps = 8;             // warps per CTA of the column pass
constexpr uint32_t kColMaxSlots = 4096;  // used columns whose block maximum is kept (beyond: always looked at per doc)

struct __align__(16) ColQ {  // one query with >= 1 column term (32 B)
  uint32_t qslot, qi;
  int32_t filter;
  uint8_t nsp, ncol, unit_w, pad;  // slots [0, nsp) sparse, [nsp, nsp + ncol) columns; unit_w: every column weight is 1.0
  uint16_t slot[8];                // used-column slots of the column terms, slot order
};

struct StreamDev {
  ColQ *colq;            // [Q] the batch's queries with >= 1 column term
  uint32_t *n_colq;      // device counters
  uint32_t *ucol;        // [n_cols] slot -> column (the columns the batch names, densest first)
  uint32_t *n_ucol;
  uint32_t *col_slot;    // [n_cols] column -> slot (scratch of slg_colgroups_kernel)
  uint32_t *sparse_counter, *col_counter;  // work counters of the two passes
  uint32_t stage_cap;    // postings a warp of the sparse pass can stage per span (multiple of 4)
  uint32_t col_resident; // columns of a block the column pass keeps in shared memory (per buffer)
  uint32_t n_smax;       // block maxima kept in shared memory (min(columns of the segment, kColMaxSlots))
  unsigned long long *counters;  // [4] sparse postings visited, (query, column block) pairs looked at per doc, (query, column block) pairs scored, sparse items
};

// ---- the batch's column terms: used columns -> slots, queries with a column term -> ColQ (one CTA) -------------
static __global__ void __launch_bounds__(1024) slg_colgroups_kernel(SegmentDev seg, WarpBatchDev wb, StreamDev sd, uint32_t n_cols) {
  const uint32_t tid = threadIdx.x, nthr = blockDim.x;
  for (uint32_t c = tid; c < n_cols; c += nthr) sd.col_slot[c] = 0u;
  if (tid == 0) {
    *sd.n_colq = 0u;
    *sd.n_ucol = 0u;
  }
  __syncthreads();
  for (uint32_t slot = tid; slot < wb.n_queries; slot += nthr) {
    const uint32_t nt = wb.qheads[slot].nt;
    for (uint32_t t = 0; t < nt; t++) {
      const QTerm &q = wb.qterms[(uint64_t)slot * kWarpMaxTerms + t];
      if ((q.flags & 5u) == 5u) sd.col_slot[q.term] = 1u;
    }
  }
  __syncthreads();
  if (tid == 0) {  // slots in column order: column 0 is the densest term (n_cols is tens to a few thousand)
    uint32_t n = 0;
    for (uint32_t c = 0; c < n_cols; c++) {
      if (sd.col_slot[c]) {
        sd.ucol[n] = c;
        sd.col_slot[c] = n++;
      } else {
        sd.col_slot[c] = 0xFFFFFFFFu;
      }
    }
    *sd.n_ucol = n;
  }
  __syncthreads();
  for (uint32_t slot = tid; slot < wb.n_queries; slot += nthr) {
    const QHead h = wb.qheads[slot];
    ColQ r;
    r.qslot = slot;
    r.qi = h.qi;
    r.filter = h.filter;
    r.nsp = 0;
    r.ncol = 0;
    r.unit_w = 1;
    r.pad = 0;
    for (int i = 0; i < 8; i++) r.slot[i] = 0;
    for (uint32_t t = 0; t < h.nt; t++) {
      const QTerm &q = wb.qterms[(uint64_t)slot * kWarpMaxTerms + t];
      if (!(q.flags & 1u)) continue;
      if (q.flags & 4u) {
        r.slot[r.ncol++] = (uint16_t)sd.col_slot[q.term];
        if (q.weight != 1.0f) r.unit_w = 0;
      } else {
        r.nsp++;
      }
    }
    if (r.ncol) sd.colq[atomicAdd(sd.n_colq, 1u)] = r;
  }
}

// ---- the warp's candidate buffer: push_top_k (query/wand.rs:905-916) ------------------------------------------
struct WarpCand {
  unsigned long long *cand;  // [kWarpCand] shared
  unsigned long long thr;    // the query's k-th key as this warp knows it
  uint32_t cnt;
  uint32_t k;
  int lane;

  __device__ __forceinline__ void begin(unsigned long long *buf, unsigned long long thr0, uint32_t k_, int lane_) {
    cand = buf;
    thr = thr0;
    cnt = 0;
    k = k_;
    lane = lane_;
  }
  // append one ballot round of keys; past 32 pending: sort, keep the best k, raise the local threshold
  __device__ __forceinline__ void push(bool pass, unsigned long long key) {
    const uint32_t bal = __ballot_sync(0xFFFFFFFFu, pass);
    if (!bal) return;
    if (pass) cand[cnt + __popc(bal & ((1u << lane) - 1u))] = key;
    cnt += __popc(bal);
    __syncwarp();
    if (cnt > 32) {
      for (uint32_t z = cnt + lane; z < kWarpCand; z += 32) cand[z] = 0ull;
      __syncwarp();
      warp_sort64_desc(cand, lane);
      cnt = min(cnt, k);
      if (cnt == k) thr = max(thr, cand[k - 1]);
      __syncwarp();
    }
  }
  // accept (api/reader.rs:3009-3036) for a doc whose exact score is known
  __device__ __forceinline__ void offer(const SegmentDev &seg, const WarpBatchDev &wb, uint32_t qi, int32_t filter, bool have, uint32_t doc,
                                        float s) {
    const unsigned long long key = ((unsigned long long)__float_as_uint(s) << 32) | (unsigned long long)(0xFFFFFFFFu - doc);
    bool pass = have && key > thr && s != 0.0f;
    if (pass) pass = (seg.live_bits[doc >> 5] >> (doc & 31)) & 1u;
    if (pass && filter >= 0) pass = (wb.filter_bits[filter][doc >> 5] >> (doc & 31)) & 1u;
    if (pass) pass = cursor_accepts(wb.q_cursor, wb.q_saw, qi, key);
    push(pass, key);
  }
  // merge into the query's global top-k under its lock; returns the query's k-th key afterwards
  __device__ __forceinline__ void merge(const WarpBatchDev &wb, uint32_t qi) {
    if (cnt == 0) return;
    const unsigned long long thr_now = ld_cg_u64(wb.thr_key + qi);
    const bool useful = lane < (int)cnt && cand[lane] > thr_now;  // cnt <= 32 after every push
    if (__any_sync(0xFFFFFFFFu, useful)) {
      if (lane == 0) {
        while (atomicCAS(wb.lock + qi, 0u, 1u) != 0u) __nanosleep(64);
        __threadfence();
      }
      __syncwarp();
      const uint32_t ng = ld_cg_u32(wb.topk_count + qi);
      unsigned long long *gk = wb.topk_keys + (uint64_t)qi * k;
      if (lane < (int)ng) cand[cnt + lane] = ld_cg_u64(gk + lane);
      uint32_t total = cnt + ng;
      for (uint32_t z = total + lane; z < kWarpCand; z += 32) cand[z] = 0ull;
      __syncwarp();
      warp_sort64_desc(cand, lane);
      total = min(total, k);
      if (lane < (int)total) st_cg_u64(gk + lane, cand[lane]);
      __threadfence();
      __syncwarp();
      if (total == k) thr = max(thr, cand[k - 1]);
      if (lane == 0) {
        st_cg_u32(wb.topk_count + qi, total);
        if (total == k) st_cg_u64(wb.thr_key + qi, cand[k - 1]);
        __threadfence();
        atomicExch(wb.lock + qi, 0u);
      }
      __syncwarp();
    } else {
      thr = max(thr, thr_now);
    }
    cnt = 0;
  }
};

// shared memory of one warp of the sparse pass
__host__ __device__ inline size_t sparse_smem_per_warp(uint32_t sub_docs, uint32_t stage_cap) {
  // acc f32[sub_docs] | sdoc u32[cap] | ssc f32[cap] | cand u64[64] | qt QTerm[8] | rb u32[8][9] | runs u32[8][8] | rest f32[8] | soff u32[8] | slo u32[8] | bar u64 (+ pad)
  return (size_t)sub_docs * 4 + (size_t)stage_cap * 8 + kWarpCand * 8 + kWarpMaxTerms * sizeof(QTerm) + kWarpMaxTerms * kRbStride * 4 +
         kSubPerGroup * kWarpMaxTerms * 4 + kSubPerGroup * 4 + kWarpMaxTerms * 4 * 2 + 16;
}

// One sub-tile of the sparse pass.  Lane t < nsp holds the run [r0, r1) of sparse term t inside the sub-tile: indices into the
// staged stream (STAGED) or into the term's posting list in global memory (a sub-tile whose runs do not fit the staging area).
template <bool STAGED>
__device__ __forceinline__ void sparse_sub(const SegmentDev &seg, const WarpBatchDev &wb, const QHead &head, const QTerm *qt, const uint32_t r0,
                                           const uint32_t r1, const uint32_t colmask, const float rest, const uint32_t tile_lo, float *acc,
                                           const uint32_t *sdoc, const float *ssc, WarpCand &wc, const int lane) {
  const uint32_t ne = __ballot_sync(0xFFFFFFFFu, r1 > r0);
  if (!ne) return;
  auto cut_now = [&]() {
    if (wc.thr == kThrInit) return 0u;
    const float cf = __uint_as_float((uint32_t)(wc.thr >> 32)) * 0.99998f - rest * 1.00002f;
    return cf > 0.0f ? __float_as_uint(cf) : 0u;
  };
  // a doc of the sparse lists whose partial is v: add the column terms of exactly this doc, slot order
  auto complete = [&](bool pass, uint32_t doc, float v) {
    float s = v;
    if (pass)
      for (uint32_t cm = colmask; cm; cm &= cm - 1) {
        const uint32_t ct = __ffs(cm) - 1;
        const float c = __ldg(seg.cols + qt[ct].sc_base + doc);
        s = __fadd_rn(s, __fmul_rn(c, qt[ct].weight));
      }
    wc.offer(seg, wb, head.qi, head.filter, pass, doc, s);
  };
  uint32_t cut = cut_now();
  if ((ne & (ne - 1u)) == 0u) {
    // one term has postings here: its contributions are the partials, nothing to accumulate
    const uint32_t t = __ffs(ne) - 1;
    const uint32_t i0 = __shfl_sync(0xFFFFFFFFu, r0, t), i1 = __shfl_sync(0xFFFFFFFFu, r1, t);
    const float w = qt[t].weight;
    const uint32_t *dp = STAGED ? sdoc : seg.post_doc + qt[t].base;
    const float *sp = STAGED ? ssc : wb.scores + qt[t].base;
#pragma unroll 1
    for (uint32_t b = i0; b < i1; b += 32) {
      const uint32_t i = b + lane;
      const bool in = i < i1;
      float v = 0.0f;
      if (in) v = __fmul_rn(sp[i], w);
      const bool pass = in && __float_as_uint(v) >= cut;
      if (!__any_sync(0xFFFFFFFFu, pass)) continue;
      const uint32_t doc = pass ? dp[i] : 0u;
      complete(pass, doc, v);
      cut = cut_now();
    }
    return;
  }
  // ---- several terms: accumulate in slot order, then visit the same runs again: collect or just restore the zeros ----
  uint32_t wmax = 0;
  for (uint32_t m = ne; m; m &= m - 1) {
    const uint32_t t = __ffs(m) - 1;
    const uint32_t i0 = __shfl_sync(0xFFFFFFFFu, r0, t), i1 = __shfl_sync(0xFFFFFFFFu, r1, t);
    const float w = qt[t].weight;
    const uint32_t *dp = STAGED ? sdoc : seg.post_doc + qt[t].base;
    const float *sp = STAGED ? ssc : wb.scores + qt[t].base;
#pragma unroll 1
    for (uint32_t i = i0 + lane; i < i1; i += 32) {
      const uint32_t slot = dp[i] - tile_lo;
      const float v = __fadd_rn(acc[slot], __fmul_rn(sp[i], w));  // distinct docs inside a list: no aliasing between lanes
      acc[slot] = v;
      wmax = max(wmax, __float_as_uint(v));
    }
    __syncwarp();
  }
  const bool collect = __reduce_max_sync(0xFFFFFFFFu, wmax) >= cut;
  for (uint32_t m = ne; m; m &= m - 1) {
    const uint32_t t = __ffs(m) - 1;
    const uint32_t i0 = __shfl_sync(0xFFFFFFFFu, r0, t), i1 = __shfl_sync(0xFFFFFFFFu, r1, t);
    const uint32_t *dp = STAGED ? sdoc : seg.post_doc + qt[t].base;
    if (!collect) {
#pragma unroll 1
      for (uint32_t i = i0 + lane; i < i1; i += 32) acc[dp[i] - tile_lo] = 0.0f;
    } else {
#pragma unroll 1
      for (uint32_t b = i0; b < i1; b += 32) {
        const uint32_t i = b + lane;
        const bool in = i < i1;
        uint32_t slot = 0;
        float v = 0.0f;
        if (in) {
          slot = dp[i] - tile_lo;
          v = acc[slot];
          acc[slot] = 0.0f;
        }
        const bool pass = in && __float_as_uint(v) >= cut && v != 0.0f;  // v == 0: an earlier run already took this doc
        if (!__any_sync(0xFFFFFFFFu, pass)) continue;
        complete(pass, tile_lo + slot, v);
        cut = cut_now();
      }
    }
    __syncwarp();
  }
}

// ---- sparse pass -----------------------------------------------------------------------------------------------
template <bool UNUSED>
__global__ void __launch_bounds__(kSparseWarps * 32) slg_score_sparse_kernel(SegmentDev seg, WarpBatchDev wb, StreamDev sd) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t sub_docs = wb.sub_docs, cap = sd.stage_cap, k = wb.k;
  unsigned char *mine = smem_raw + (size_t)warp * sparse_smem_per_warp(sub_docs, cap);
  float *acc = reinterpret_cast<float *>(mine);
  uint32_t *sdoc = reinterpret_cast<uint32_t *>(acc + sub_docs);
  float *ssc = reinterpret_cast<float *>(sdoc + cap);
  unsigned long long *cand = reinterpret_cast<unsigned long long *>(ssc + cap);
  QTerm *qt = reinterpret_cast<QTerm *>(cand + kWarpCand);
  uint32_t *rb = reinterpret_cast<uint32_t *>(qt + kWarpMaxTerms);
  uint32_t *runs = rb + kWarpMaxTerms * kRbStride;                                   // [jj][t] = i0 | i1 << 16, staged indices
  float *rest = reinterpret_cast<float *>(runs + kSubPerGroup * kWarpMaxTerms);      // [jj] what the column terms can add
  uint32_t *soff = reinterpret_cast<uint32_t *>(rest + kSubPerGroup);
  uint32_t *slo = soff + kWarpMaxTerms;
  unsigned long long *bar = reinterpret_cast<unsigned long long *>(slo + kWarpMaxTerms);
  for (uint32_t i = lane * 4; i < sub_docs; i += 128) *reinterpret_cast<float4 *>(acc + i) = make_float4(0, 0, 0, 0);
  if (lane == 0) {
    mbar_init(bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncwarp();
  uint32_t parity = 0;
  const uint32_t total_items = wb.n_groups * wb.n_queries;
  uint32_t n_post = 0;  // per lane
  unsigned long long n_post_total = 0;
  uint32_t n_items = 0;
  WarpCand wc;

  uint32_t item = 0;
  if (lane == 0) item = atomicAdd(sd.sparse_counter, 1u);
  item = __shfl_sync(0xFFFFFFFFu, item, 0);
  while (item < total_items) {
    uint32_t next_item = 0;
    if (lane == 0) next_item = atomicAdd(sd.sparse_counter, 1u);
    const uint32_t tg = item / wb.n_queries, qslot = item - tg * wb.n_queries;
    const QHead head = wb.qheads[qslot];
    const uint32_t nt = head.nt;
    if (lane < 16) reinterpret_cast<uint4 *>(qt)[lane] = __ldg(reinterpret_cast<const uint4 *>(wb.qterms + (uint64_t)qslot * kWarpMaxTerms) + lane);
    const unsigned long long thr0 = ld_cg_u64(wb.thr_key + head.qi);
    __syncwarp();
    const uint32_t myflags = lane < (int)nt ? qt[lane].flags : 0u;
    const uint32_t spmask = __ballot_sync(0xFFFFFFFFu, (myflags & 5u) == 1u);  // canonical layout: the low nsp slots
    const uint32_t colmask = __ballot_sync(0xFFFFFFFFu, (myflags & 5u) == 5u);
    if (spmask) {
      n_items++;
      const uint32_t nsp = __popc(spmask);
      const uint32_t sub0 = tg * kSubPerGroup;
      const uint32_t jmax = min(kSubPerGroup, wb.n_sub - sub0);
      {
        // posting boundaries of the sparse terms: lane = t*4 + c
        const uint32_t t = lane >> 2, c = lane & 3;
        if (t < nsp) {
          const uint32_t *row = wb.rng + (uint64_t)qt[t].uterm * (wb.n_sub + 1);
          for (uint32_t j = c; j <= kSubPerGroup; j += 4) rb[t * kRbStride + j] = __ldg(row + min(sub0 + j, wb.n_sub));
        }
        // what the column terms can add to a partial inside each sub-tile: their exact maxima per 512 docs (seg.col_tmax)
        if (lane < (int)kSubPerGroup) {
          float r = 0.0f;
          if (lane < (int)jmax) {
            const uint32_t d0 = (sub0 + lane) * sub_docs, d1 = min(d0 + sub_docs, seg.doc_count) - 1u;
            for (uint32_t cm = colmask; cm; cm &= cm - 1) {
              const uint32_t ct = __ffs(cm) - 1;
              const float *tm = seg.col_tmax + (uint64_t)qt[ct].term * seg.tmax_stride;
              float b = 0.0f;
              for (uint32_t blk = d0 >> 9; blk <= (d1 >> 9); blk++) b = fmaxf(b, __ldg(tm + blk));
              r = __fadd_rn(r, __fmul_rn(b, qt[ct].weight));
            }
          }
          rest[lane] = r;
        }
      }
      __syncwarp();
      wc.begin(cand, thr0, k, lane);

      uint32_t j0 = 0;
      while (j0 < jmax) {
        // the longest span of sub-tiles [j0, j1) whose runs (rounded out to 16-byte pieces) fit the staging area
        uint32_t tot = 0;
        const uint32_t jc = j0 + 1 + lane;  // lanes 0..7 try j1 = j0+1 .. j0+8
        if (jc <= jmax)
          for (uint32_t t = 0; t < nsp; t++) {
            const uint32_t lo = rb[t * kRbStride + j0], hi = rb[t * kRbStride + jc];
            if (hi > lo) tot += ((hi + 3u) & ~3u) - (lo & ~3u);
          }
        const uint32_t fits = __ballot_sync(0xFFFFFFFFu, jc <= jmax && tot <= cap);
        if (!(fits & 1u)) {  // a single sub-tile does not fit: straight from global memory
          uint32_t r0 = 0, r1 = 0;
          if (lane < (int)nsp) {
            r0 = rb[lane * kRbStride + j0];
            r1 = rb[lane * kRbStride + j0 + 1];
            n_post += r1 - r0;
          }
          sparse_sub<false>(seg, wb, head, qt, r0, r1, colmask, rest[j0], (sub0 + j0) * sub_docs, acc, sdoc, ssc, wc, lane);
          j0++;
          continue;
        }
        const uint32_t nspan = __ffs(~fits) - 1;  // fits is a run of ones from bit 0 (tot grows with j1)
        const uint32_t j1 = j0 + nspan;
        const uint32_t span_tot = __shfl_sync(0xFFFFFFFFu, tot, nspan - 1);
        if (span_tot == 0) {
          j0 = j1;
          continue;
        }
        // ---- stage: lanes 0..7 lay the terms out, lanes 0..15 issue one bulk copy each (docs / scores per term) ----
        {
          uint32_t len = 0, lo_al = 0;
          if (lane < (int)nsp) {
            const uint32_t lo = rb[lane * kRbStride + j0], hi = rb[lane * kRbStride + j1];
            if (hi > lo) {
              lo_al = lo & ~3u;
              len = ((hi + 3u) & ~3u) - lo_al;
              n_post += hi - lo;
            }
          }
          uint32_t off = len;  // inclusive scan over lanes 0..7
#pragma unroll
          for (int o = 1; o < 8; o <<= 1) {
            const uint32_t n = __shfl_up_sync(0xFFFFFFFFu, off, o);
            if (lane >= o) off += n;
          }
          off -= len;
          if (lane < (int)kWarpMaxTerms) {
            soff[lane] = off;
            slo[lane] = lo_al;
          }
          __syncwarp();  // every lane is done reading the staging area of the previous span
          if (lane == 0) mbar_arrive_expect_tx(bar, span_tot * 8u);
          __syncwarp();
          const uint32_t t = (lane >> 1) & 7;
          const uint32_t tlen = __shfl_sync(0xFFFFFFFFu, len, t), tlo = __shfl_sync(0xFFFFFFFFu, lo_al, t), toff = __shfl_sync(0xFFFFFFFFu, off, t);
          if (lane < 16 && tlen) {
            if (lane & 1) bulk_copy_g2s(ssc + toff, wb.scores + qt[t].base + tlo, tlen * 4u, bar);
            else bulk_copy_g2s(sdoc + toff, seg.post_doc + qt[t].base + tlo, tlen * 4u, bar);
          }
          // run table of the span while the copies fly: entry e = jj*8 + t, two per lane
#pragma unroll
          for (int h = 0; h < 2; h++) {
            const uint32_t e = lane + 32 * h, et = e & 7, ej = e >> 3;
            uint32_t v = 0;
            if (et < nsp && j0 + ej < j1) {
              const uint32_t sh = soff[et] - slo[et];
              v = (rb[et * kRbStride + j0 + ej] + sh) | ((rb[et * kRbStride + j0 + ej + 1] + sh) << 16);
            }
            runs[e] = v;
          }
          __syncwarp();
          mbar_wait(bar, parity);
          parity ^= 1u;
        }
        for (uint32_t j = j0; j < j1; j++) {
          const uint32_t r = lane < (int)kWarpMaxTerms ? runs[(j - j0) * kWarpMaxTerms + lane] : 0u;
          sparse_sub<true>(seg, wb, head, qt, r & 0xFFFFu, r >> 16, colmask, rest[j], (sub0 + j) * sub_docs, acc, sdoc, ssc, wc, lane);
        }
        j0 = j1;
      }
      wc.merge(wb, head.qi);
    }
    if (n_post > 0x40000000u) {
      n_post_total += n_post;
      n_post = 0;
    }
    item = __shfl_sync(0xFFFFFFFFu, next_item, 0);
  }
  if (sd.counters) {
    n_post_total += n_post;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) n_post_total += __shfl_xor_sync(0xFFFFFFFFu, n_post_total, o);
    if (lane == 0) {
      if (n_post_total) atomicAdd(sd.counters + 0, n_post_total);
      if (n_items) atomicAdd(sd.counters + 3, (unsigned long long)n_items);
    }
  }
}

// ---- column pass -----------------------------------------------------------------------------------------------
// shared memory: buf f32[2][resident][kColBlock] | smax f32[n_smax] | cand u64[kColWarps][64] | bar u64[2]
__host__ __device__ inline size_t column_smem(uint32_t resident, uint32_t n_smax) {
  return (size_t)2 * resident * kColBlock * 4 + (((size_t)n_smax * 4 + 15) & ~(size_t)15) + (size_t)kColWarps * kWarpCand * 8 + 16;
}

template <bool PRUNE>
__global__ void __launch_bounds__(kColWarps * 32) slg_score_columns_kernel(SegmentDev seg, WarpBatchDev wb, StreamDev sd) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t n_used = *sd.n_ucol, n_colq = *sd.n_colq;
  const uint32_t resident = min(sd.col_resident, n_used), n_smax = min(n_used, sd.n_smax);
  float *buf = reinterpret_cast<float *>(smem_raw);
  float *smax = buf + (size_t)2 * sd.col_resident * kColBlock;
  unsigned long long *cand = reinterpret_cast<unsigned long long *>(reinterpret_cast<unsigned char *>(smax) + (((size_t)sd.n_smax * 4 + 15) & ~(size_t)15)) +
                             (size_t)warp * kWarpCand;
  unsigned long long *bar = reinterpret_cast<unsigned long long *>(smem_raw + column_smem(sd.col_resident, sd.n_smax) - 16);
  const uint32_t k = wb.k;
  const uint32_t n_blocks = (seg.doc_count + kColBlock - 1) / kColBlock;
  if (n_colq == 0 || n_used == 0) return;
  if (tid == 0) {
    mbar_init(bar, 1);
    mbar_init(bar + 1, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncthreads();
  unsigned long long n_tests = 0, n_looked = 0;
  WarpCand wc;
  // one bulk copy per resident column; the column stride keeps a whole block past doc_count in bounds and zero
  auto fetch = [&](uint32_t blk, uint32_t b) {
    if (warp == 0) {
      if (lane == 0) mbar_arrive_expect_tx(bar + b, resident * kColBlock * 4u);
      __syncwarp();
      for (uint32_t s = lane; s < resident; s += 32)
        bulk_copy_g2s(buf + ((size_t)b * sd.col_resident + s) * kColBlock, seg.cols + (uint64_t)sd.ucol[s] * seg.col_stride + (uint64_t)blk * kColBlock,
                      kColBlock * 4u, bar + b);
    }
  };
  uint32_t parity[2] = {0u, 0u};
  uint32_t blk = blockIdx.x;
  if (blk < n_blocks) fetch(blk, 0);
  uint32_t b = 0;
  for (; blk < n_blocks; blk += gridDim.x, b ^= 1u) {
    const uint32_t nxt = blk + gridDim.x;
    if (nxt < n_blocks) fetch(nxt, b ^ 1u);  // the other buffer was released by the barrier that ended the previous block
    mbar_wait(bar + b, parity[b]);
    parity[b] ^= 1u;
    const float *cb = buf + (size_t)b * sd.col_resident * kColBlock;
    const uint32_t d0 = blk * kColBlock;
    // block maximum of every used column (non-resident columns: from global memory)
    for (uint32_t s = warp; s < n_smax; s += kColWarps) {
      const float4 *p = s < resident ? reinterpret_cast<const float4 *>(cb + (size_t)s * kColBlock)
                                     : reinterpret_cast<const float4 *>(seg.cols + (uint64_t)sd.ucol[s] * seg.col_stride + d0);
      uint32_t top = 0u;
#pragma unroll
      for (uint32_t x = 0; x < kColBlock / 128; x++) {
        const float4 v = p[x * 32 + lane];
        top = max(top, max(max(__float_as_uint(v.x), __float_as_uint(v.y)), max(__float_as_uint(v.z), __float_as_uint(v.w))));
      }
      top = __reduce_max_sync(0xFFFFFFFFu, top);
      if (lane == 0) smax[s] = __uint_as_float(top);
    }
    __syncthreads();
    // ---- every query with a column term against this block: lane g = one query ----
    for (uint32_t c0 = warp * 32; c0 < n_colq; c0 += kColWarps * 32) {
      ColQ cq;
      cq.ncol = 0;
      unsigned long long thr = ~0ull;
      const bool live = c0 + lane < n_colq;
      if (live) {
        cq = sd.colq[c0 + lane];
        thr = ld_cg_u64(wb.thr_key + cq.qi);
      }
      bool look = false;
      if (live) {
        n_tests++;
        // one column: the maximum over the block of c * w is max(c) * w, the comparison of every doc collapses into one.
        // Several columns: the exhaustive execution sums them per doc; a pruned one first asks the sum of the maxima
        // (same slot order, every operation monotone: it dominates every doc's sum).
        float bound = 0.0f;
        bool known = true;
        if (PRUNE || cq.ncol == 1)
          for (uint32_t i = 0; i < cq.ncol; i++) {
            const uint32_t s = cq.slot[i];
            if (s >= n_smax) known = false;
            else bound = __fadd_rn(bound, __fmul_rn(smax[s], cq.unit_w ? 1.0f : __ldg(&wb.qterms[(uint64_t)cq.qslot * kWarpMaxTerms + cq.nsp + i].weight)));
          }
        else known = false;
        const uint32_t thr_bits = thr == kThrInit ? 0u : (uint32_t)(thr >> 32);
        look = !known || (bound != 0.0f && __float_as_uint(bound) >= thr_bits);
      }
      uint32_t hits = __ballot_sync(0xFFFFFFFFu, look);
      while (hits) {
        const int g = __ffs(hits) - 1;
        hits &= hits - 1;
        const uint32_t qslot = __shfl_sync(0xFFFFFFFFu, cq.qslot, g), qi = __shfl_sync(0xFFFFFFFFu, cq.qi, g);
        const uint32_t ncol = __shfl_sync(0xFFFFFFFFu, (uint32_t)cq.ncol, g), nsp = __shfl_sync(0xFFFFFFFFu, (uint32_t)cq.nsp, g);
        const uint32_t unit_w = __shfl_sync(0xFFFFFFFFu, (uint32_t)cq.unit_w, g);
        const int32_t filter = __shfl_sync(0xFFFFFFFFu, cq.filter, g);
        const unsigned long long qthr = __shfl_sync(0xFFFFFFFFu, thr, g);
        const uint32_t s01 = __shfl_sync(0xFFFFFFFFu, (uint32_t)cq.slot[0] | ((uint32_t)cq.slot[1] << 16), g);
        const uint32_t s23 = __shfl_sync(0xFFFFFFFFu, (uint32_t)cq.slot[2] | ((uint32_t)cq.slot[3] << 16), g);
        const uint32_t s45 = __shfl_sync(0xFFFFFFFFu, (uint32_t)cq.slot[4] | ((uint32_t)cq.slot[5] << 16), g);
        const uint32_t s67 = __shfl_sync(0xFFFFFFFFu, (uint32_t)cq.slot[6] | ((uint32_t)cq.slot[7] << 16), g);
        const QTerm *qts = wb.qterms + (uint64_t)qslot * kWarpMaxTerms;
        n_looked++;
        // v = sum of the query's columns over the block, slot order (the first product is the exact value of 0 + c * w)
        float4 v[kColBlock / 128];
#pragma unroll
        for (uint32_t x = 0; x < kColBlock / 128; x++) v[x] = make_float4(0, 0, 0, 0);
        for (uint32_t i = 0; i < ncol; i++) {
          const uint32_t pair = i < 2 ? s01 : (i < 4 ? s23 : (i < 6 ? s45 : s67));
          const uint32_t s = (i & 1u) ? pair >> 16 : pair & 0xFFFFu;
          const float w = unit_w ? 1.0f : __ldg(&qts[nsp + i].weight);
          const float4 *p = s < resident ? reinterpret_cast<const float4 *>(cb + (size_t)s * kColBlock)
                                         : reinterpret_cast<const float4 *>(seg.cols + (uint64_t)sd.ucol[s] * seg.col_stride + d0);
#pragma unroll
          for (uint32_t x = 0; x < kColBlock / 128; x++) {
            const float4 c = p[x * 32 + lane];
            v[x].x = __fadd_rn(v[x].x, __fmul_rn(c.x, w));
            v[x].y = __fadd_rn(v[x].y, __fmul_rn(c.y, w));
            v[x].z = __fadd_rn(v[x].z, __fmul_rn(c.z, w));
            v[x].w = __fadd_rn(v[x].w, __fmul_rn(c.w, w));
          }
        }
        uint32_t cut = qthr == kThrInit ? 0u : (uint32_t)(qthr >> 32);
        uint32_t mx = 0u;
#pragma unroll
        for (uint32_t x = 0; x < kColBlock / 128; x++)
          mx = max(mx, max(max(__float_as_uint(v[x].x), __float_as_uint(v[x].y)), max(__float_as_uint(v[x].z), __float_as_uint(v[x].w))));
        if (!__any_sync(0xFFFFFFFFu, mx >= cut && mx != 0u)) continue;
        // ---- per doc: the docs that can enter the top k, one per lane and round ----
        wc.begin(cand, qthr, k, lane);
        uint32_t todo = 0u;
#pragma unroll
        for (uint32_t x = 0; x < kColBlock / 128; x++) {
          const uint32_t bx[4] = {__float_as_uint(v[x].x), __float_as_uint(v[x].y), __float_as_uint(v[x].z), __float_as_uint(v[x].w)};
#pragma unroll
          for (int e = 0; e < 4; e++)
            if (bx[e] >= cut && bx[e] != 0u && d0 + x * 128 + lane * 4 + e < seg.doc_count) todo |= 1u << (x * 4 + e);
        }
        while (__any_sync(0xFFFFFFFFu, todo != 0u)) {
          {
            const uint32_t el = todo ? __ffs(todo) - 1 : 0u;
            const bool had = todo != 0u;
            todo &= todo - 1u;
            uint32_t bits = 0u;
#pragma unroll
            for (uint32_t x = 0; x < kColBlock / 128; x++) {
              if (el == x * 4 + 0) bits = __float_as_uint(v[x].x);
              if (el == x * 4 + 1) bits = __float_as_uint(v[x].y);
              if (el == x * 4 + 2) bits = __float_as_uint(v[x].z);
              if (el == x * 4 + 3) bits = __float_as_uint(v[x].w);
            }
            const uint32_t doc = d0 + (el >> 2) * 128 + lane * 4 + (el & 3u);
            cut = wc.thr == kThrInit ? 0u : (uint32_t)(wc.thr >> 32);
            bool pass = had && bits >= cut;
            if (!__any_sync(0xFFFFFFFFu, pass)) continue;
            if (pass && nsp) {
              // a doc of one of the query's sparse lists belongs to the sparse pass
              const uint32_t sub = doc / wb.sub_docs;
              for (uint32_t t = 0; t < nsp && pass; t++) {
                const uint32_t *row = wb.rng + (uint64_t)__ldg(&qts[t].uterm) * (wb.n_sub + 1);
                const uint32_t *dp = seg.post_doc + __ldg(&qts[t].base);
                uint32_t lo = __ldg(row + sub);
                const uint32_t end = __ldg(row + sub + 1);
                uint32_t hi = end;
                while (lo < hi) {
                  const uint32_t mid = (lo + hi) >> 1;
                  if (__ldg(dp + mid) < doc) lo = mid + 1;
                  else hi = mid;
                }
                if (lo < end && __ldg(dp + lo) == doc) pass = false;
              }
            }
            wc.offer(seg, wb, qi, filter, pass, doc, __uint_as_float(bits));
          }
        }
        wc.merge(wb, qi);
      }
    }
    __syncthreads();  // everyone is done with buffer b: the fetch two blocks ahead may overwrite it
  }
  if (sd.counters && lane == 0) {
    if (n_looked) atomicAdd(sd.counters + 1, n_looked);
    if (n_tests) atomicAdd(sd.counters + 2, n_tests);
  }
}

}  // namespace slg
