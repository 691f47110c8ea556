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
//   COLUMN PASS   slg_score_colgroups_kernel — terms with a doc-indexed f32 column (df >= N / dense_den).  Queries
//                 are grouped by their first column term (slg_colgroups_kernel); a warp takes a (chunk of <= 32
//                 queries of one group, range of 512-doc blocks) item and reads each block of the shared column
//                 ONCE for the whole chunk, 16 docs per lane in registers.  A query whose only column is the shared
//                 one sees v = c * w for every doc of the block, and max(v) = max(c) * w (f32 multiplication by a
//                 positive weight is monotone), so its 512 comparisons against the k-th score collapse into one
//                 against the block maximum taken from the streamed values; a query with further columns adds
//                 their blocks in slot order first.  Only where a doc can enter the top k is the block looked at
//                 per doc.  A doc that also sits in one of the query's sparse lists belongs to the sparse pass (its
//                 partial is in A there) and is left alone here.
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

constexpr uint32_t kColChunk = 32;       // queries per column-pass chunk: one per lane
constexpr uint32_t kColItemBlocks = 8;   // 512-doc blocks per column-pass item
constexpr int kSparseWarps = 4;          // warps per CTA of the sparse pass
constexpr int kColWarps = 8;             // warps per CTA of the column pass

struct __align__(16) ColQ {  // one query of a column group (32 B)
  uint32_t qslot, qi;
  float w;             // weight of the group's column in this query
  uint32_t ncol, nsp;  // column terms / sparse terms of the query (slots [0, nsp) sparse, [nsp, nsp + ncol) columns)
  int32_t filter;
  uint64_t pad;
};

struct __align__(16) ColChunk {  // <= kColChunk queries that share their first column (16 B)
  uint64_t sc_base;  // element offset of the column in seg.cols
  uint32_t begin, count;
};

struct StreamDev {
  ColQ *colq;            // [Q] queries with >= 1 column term, grouped by their first column
  ColChunk *chunks;      // [<= Q]
  uint32_t *n_chunks;    // device counter
  uint32_t *col_count;   // [n_cols + 1] scratch of slg_colgroups_kernel
  uint32_t *sparse_counter, *col_counter;  // work counters of the two passes
  uint32_t stage_cap;    // postings a warp of the sparse pass can stage per span (multiple of 4)
  unsigned long long *counters;  // [4] sparse postings visited, (query, column block) pairs looked at per doc, (query, column block) pairs streamed, sparse items
};

// ---- grouping of the batch's queries by their first column term (one CTA) ------------------------------------
static __global__ void __launch_bounds__(1024) slg_colgroups_kernel(SegmentDev seg, WarpBatchDev wb, StreamDev sd, uint32_t n_cols) {
  __shared__ uint32_t s_total;
  const uint32_t tid = threadIdx.x, nthr = blockDim.x;
  for (uint32_t c = tid; c <= n_cols; c += nthr) sd.col_count[c] = 0u;
  if (tid == 0) *sd.n_chunks = 0u;
  __syncthreads();
  // pass 1: histogram
  for (uint32_t slot = tid; slot < wb.n_queries; slot += nthr) {
    const QHead h = wb.qheads[slot];
    for (uint32_t t = 0; t < h.nt; t++) {
      const QTerm &q = wb.qterms[(uint64_t)slot * kWarpMaxTerms + t];
      if ((q.flags & 5u) == 5u) {
        atomicAdd(sd.col_count + (uint32_t)(q.sc_base / seg.col_stride), 1u);
        break;
      }
    }
  }
  __syncthreads();
  // exclusive scan (n_cols is small: tens to a few thousand) — thread 0, then the chunk list per column
  if (tid == 0) {
    uint32_t run = 0;
    for (uint32_t c = 0; c < n_cols; c++) {
      const uint32_t n = sd.col_count[c];
      sd.col_count[c] = run;
      run += n;
    }
    sd.col_count[n_cols] = run;
    s_total = run;
  }
  __syncthreads();
  for (uint32_t c = tid; c < n_cols; c += nthr) {
    const uint32_t b = sd.col_count[c], e = (c + 1 < n_cols) ? sd.col_count[c + 1] : s_total;
    // (col_count[c + 1] is still the untouched prefix here: pass 2 below only advances entries it owns)
    const uint32_t n = e - b;
    if (n) {
      const uint32_t nch = (n + kColChunk - 1) / kColChunk;
      const uint32_t at = atomicAdd(sd.n_chunks, nch);
      for (uint32_t i = 0; i < nch; i++) {
        ColChunk ch;
        ch.sc_base = (uint64_t)c * seg.col_stride;
        ch.begin = b + i * kColChunk;
        ch.count = min(kColChunk, n - i * kColChunk);
        sd.chunks[at + i] = ch;
      }
    }
  }
  __syncthreads();
  // pass 2: fill (the running prefix of a column doubles as its fill cursor; order inside a group is immaterial)
  for (uint32_t slot = tid; slot < wb.n_queries; slot += nthr) {
    const QHead h = wb.qheads[slot];
    uint32_t nsp = 0, ncol = 0, first = 0xFFFFFFFFu;
    for (uint32_t t = 0; t < h.nt; t++) {
      const QTerm &q = wb.qterms[(uint64_t)slot * kWarpMaxTerms + t];
      if (!(q.flags & 1u)) continue;
      if (q.flags & 4u) {
        if (first == 0xFFFFFFFFu) first = t;
        ncol++;
      } else {
        nsp++;
      }
    }
    if (first == 0xFFFFFFFFu) continue;
    const QTerm &q = wb.qterms[(uint64_t)slot * kWarpMaxTerms + first];
    const uint32_t pos = atomicAdd(sd.col_count + (uint32_t)(q.sc_base / seg.col_stride), 1u);
    ColQ r;
    r.qslot = slot;
    r.qi = h.qi;
    r.w = q.weight;
    r.ncol = ncol;
    r.nsp = nsp;
    r.filter = h.filter;
    r.pad = 0;
    sd.colq[pos] = r;
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
  // acc f32[sub_docs] | sdoc u32[cap] | ssc f32[cap] | cand u64[64] | qt QTerm[8] | rb u32[8][9] | ubs f32[8][8] | soff u32[8] | slo u32[8] | bar u64 (+ pad to 16)
  return (size_t)sub_docs * 4 + (size_t)stage_cap * 8 + kWarpCand * 8 + kWarpMaxTerms * sizeof(QTerm) + kWarpMaxTerms * kRbStride * 4 +
         kWarpMaxTerms * 8 * 4 + kWarpMaxTerms * 4 * 2 + 16;
}

// scatter [i0, i1) of the staged (doc, score) stream into the accumulator
template <bool FIRST>
__device__ __forceinline__ void scatter_staged(const uint32_t *sdoc, const float *ssc, uint32_t i0, uint32_t i1, uint32_t tile_lo, float w,
                                               float *acc, int lane, uint32_t &wmax) {
  uint32_t i = i0 + lane;
#pragma unroll 1
  for (; i + 32 < i1; i += 64) {  // two steps in flight
    const uint32_t d0 = sdoc[i], d1 = sdoc[i + 32];
    const float s0 = ssc[i], s1 = ssc[i + 32];
    float v0 = __fmul_rn(s0, w), v1 = __fmul_rn(s1, w);
    float *p0 = acc + (d0 - tile_lo), *p1 = acc + (d1 - tile_lo);
    if (!FIRST) {
      const float a0 = *p0, a1 = *p1;
      v0 = __fadd_rn(a0, v0);
      v1 = __fadd_rn(a1, v1);
    }
    *p0 = v0;
    *p1 = v1;
    wmax = max(wmax, max(__float_as_uint(v0), __float_as_uint(v1)));
  }
  if (i < i1) {
    float v = __fmul_rn(ssc[i], w);
    float *p = acc + (sdoc[i] - tile_lo);
    if (!FIRST) v = __fadd_rn(*p, v);
    *p = v;
    wmax = max(wmax, __float_as_uint(v));
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
  float *ubs = reinterpret_cast<float *>(rb + kWarpMaxTerms * kRbStride);
  uint32_t *soff = reinterpret_cast<uint32_t *>(ubs + kWarpMaxTerms * 8);
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
  unsigned long long n_post = 0;
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
    uint32_t myflags = lane < (int)nt ? qt[lane].flags : 0u;
    const uint32_t spmask = __ballot_sync(0xFFFFFFFFu, (myflags & 5u) == 1u);
    const uint32_t colmask = __ballot_sync(0xFFFFFFFFu, (myflags & 5u) == 5u);
    if (spmask) {
      n_items++;
      const uint32_t sub0 = tg * kSubPerGroup;
      {
        // posting boundaries of the sparse terms (lane = t*4 + c) and, per column term, the exact column maximum
        // inside each sub-tile (seg.col_tmax per 512 docs): what a scattered partial can still gain
        const uint32_t t = lane >> 2, c = lane & 3;
        if ((spmask >> t) & 1u) {
          const uint32_t *row = wb.rng + (uint64_t)qt[t].uterm * (wb.n_sub + 1);
          for (uint32_t j = c; j <= kSubPerGroup; j += 4) rb[t * kRbStride + j] = __ldg(row + min(sub0 + j, wb.n_sub));
        } else if ((colmask >> t) & 1u) {
          const float *tm = seg.col_tmax + (qt[t].sc_base / seg.col_stride) * seg.tmax_stride;
          for (uint32_t j = c; j < kSubPerGroup; j += 4) {
            float b = 0.0f;
            const uint32_t d0 = (sub0 + j) * sub_docs;
            if (sub0 + j < wb.n_sub) {
              const uint32_t d1 = min(d0 + sub_docs, seg.doc_count) - 1u;
              for (uint32_t blk = d0 >> 9; blk <= (d1 >> 9); blk++) b = fmaxf(b, __ldg(tm + blk));
            }
            ubs[t * 8 + j] = __fmul_rn(b, qt[t].weight) ;
          }
        }
      }
      __syncwarp();
      wc.begin(cand, thr0, k, lane);
      const uint32_t nsp = __popc(spmask);
      const uint32_t jmax = min(kSubPerGroup, wb.n_sub - sub0);

      // one sub-tile: scatter + walk from the staged stream (staged) or from global memory (a sub-tile whose runs do not fit)
      auto process_sub = [&](const uint32_t j, const bool staged) {
        const uint32_t tile_lo = (sub0 + j) * sub_docs;
        // bound of what a partial lacks: the column terms inside this sub-tile
        float rest = (lane < (int)nt && ((colmask >> lane) & 1u)) ? ubs[lane * 8 + j] : 0.0f;
#pragma unroll
        for (int o = 4; o > 0; o >>= 1) rest += __shfl_xor_sync(0xFFFFFFFFu, rest, o);
        rest = __shfl_sync(0xFFFFFFFFu, rest, 0);
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
        if (nsp == 1) {
          // one sparse term: its contributions are the partials
          const uint32_t t = __ffs(spmask) - 1;
          const float w = qt[t].weight;
          uint32_t i0 = rb[t * kRbStride + j], i1 = rb[t * kRbStride + j + 1];
          if (i1 <= i0) return;
          n_post += i1 - i0;
          const uint32_t *dp;
          const float *sp;
          if (staged) {
            const uint32_t sh = soff[t] - slo[t];
            i0 += sh;
            i1 += sh;
            dp = sdoc;
            sp = ssc;
          } else {
            dp = seg.post_doc + qt[t].base;
            sp = wb.scores + qt[t].base;
          }
          uint32_t cut = cut_now();
#pragma unroll 1
          for (uint32_t i = i0; i < i1; i += 32) {
            const bool in = i + lane < i1;
            float v = 0.0f;
            if (in) v = __fmul_rn(sp[i + lane], w);
            const bool pass = in && __float_as_uint(v) >= cut;
            if (!__any_sync(0xFFFFFFFFu, pass)) continue;
            const uint32_t doc = pass ? dp[i + lane] : 0u;
            complete(pass, doc, v);
            cut = cut_now();
          }
          return;
        }
        // ---- several sparse terms: scatter in slot order, then walk the same runs ----
        uint32_t wmax = 0;
        bool first = true, any = false;
        for (uint32_t m = spmask; m; m &= m - 1) {
          const uint32_t t = __ffs(m) - 1;
          uint32_t i0 = rb[t * kRbStride + j], i1 = rb[t * kRbStride + j + 1];
          if (i1 <= i0) continue;
          any = true;
          n_post += i1 - i0;
          const float w = qt[t].weight;
          if (staged) {
            const uint32_t sh = soff[t] - slo[t];
            if (first) scatter_staged<true>(sdoc, ssc, i0 + sh, i1 + sh, tile_lo, w, acc, lane, wmax);
            else scatter_staged<false>(sdoc, ssc, i0 + sh, i1 + sh, tile_lo, w, acc, lane, wmax);
          } else {
            const uint32_t *dptr = seg.post_doc + qt[t].base;
            const float *sptr = wb.scores + qt[t].base;
            if (first) accumulate_staged<true, false>(dptr, sptr, i0, i1, tile_lo, w, acc, lane, wmax);
            else accumulate_staged<false, false>(dptr, sptr, i0, i1, tile_lo, w, acc, lane, wmax);
          }
          first = false;
          __syncwarp();
        }
        if (!any) return;
        uint32_t cut = cut_now();
        const bool collect = __reduce_max_sync(0xFFFFFFFFu, wmax) >= cut;
        for (uint32_t m = spmask; m; m &= m - 1) {
          const uint32_t t = __ffs(m) - 1;
          uint32_t i0 = rb[t * kRbStride + j], i1 = rb[t * kRbStride + j + 1];
          if (i1 <= i0) continue;
          const uint32_t *dp = seg.post_doc + qt[t].base;
          if (staged) {
            const uint32_t sh = soff[t] - slo[t];
            i0 += sh;
            i1 += sh;
            dp = sdoc;
          }
#pragma unroll 1
          for (uint32_t i = i0; i < i1; i += 32) {
            const bool in = i + lane < i1;
            uint32_t slot = 0;
            if (in) slot = dp[i + lane] - tile_lo;
            if (!collect) {
              if (in) acc[slot] = 0.0f;
              continue;
            }
            float v = 0.0f;
            if (in) {
              v = acc[slot];
              acc[slot] = 0.0f;
            }
            const bool pass = in && __float_as_uint(v) >= cut && v != 0.0f;  // v == 0: an earlier run already took this doc
            if (!__any_sync(0xFFFFFFFFu, pass)) continue;
            complete(pass, tile_lo + slot, v);
            cut = cut_now();
          }
          __syncwarp();
        }
      };

      uint32_t j0 = 0;
      while (j0 < jmax) {
        // the longest span of sub-tiles [j0, j1) whose runs (rounded out to 16-byte pieces) fit the staging area
        uint32_t tot = 0;
        const uint32_t jc = j0 + 1 + lane;  // lanes 0..7 try j1 = j0+1 .. j0+8
        if (jc <= jmax)
          for (uint32_t m = spmask; m; m &= m - 1) {
            const uint32_t t = __ffs(m) - 1;
            const uint32_t lo = rb[t * kRbStride + j0], hi = rb[t * kRbStride + jc];
            if (hi > lo) tot += ((hi + 3u) & ~3u) - (lo & ~3u);
          }
        const uint32_t fits = __ballot_sync(0xFFFFFFFFu, jc <= jmax && tot <= cap);
        if (!(fits & 1u)) {  // a single sub-tile does not fit: straight from global memory
          process_sub(j0, false);
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
          if (lane < (int)kWarpMaxTerms && ((spmask >> lane) & 1u)) {
            const uint32_t lo = rb[lane * kRbStride + j0], hi = rb[lane * kRbStride + j1];
            if (hi > lo) {
              lo_al = lo & ~3u;
              len = ((hi + 3u) & ~3u) - lo_al;
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
          const uint32_t t = lane >> 1;
          const uint32_t tlen = __shfl_sync(0xFFFFFFFFu, len, t & 7), tlo = __shfl_sync(0xFFFFFFFFu, lo_al, t & 7),
                         toff = __shfl_sync(0xFFFFFFFFu, off, t & 7);
          if (lane < 16 && tlen) {
            if (lane & 1) bulk_copy_g2s(ssc + toff, wb.scores + qt[t].base + tlo, tlen * 4u, bar);
            else bulk_copy_g2s(sdoc + toff, seg.post_doc + qt[t].base + tlo, tlen * 4u, bar);
          }
          mbar_wait(bar, parity);
          parity ^= 1u;
        }
        for (uint32_t j = j0; j < j1; j++) process_sub(j, true);
        j0 = j1;
      }
      wc.merge(wb, head.qi);
    }
    item = __shfl_sync(0xFFFFFFFFu, next_item, 0);
  }
  if (sd.counters && lane == 0) {
    if (n_post) atomicAdd(sd.counters + 0, n_post);
    if (n_items) atomicAdd(sd.counters + 3, (unsigned long long)n_items);
  }
}

// ---- column pass -----------------------------------------------------------------------------------------------
template <bool UNUSED>
__global__ void __launch_bounds__(kColWarps * 32) slg_score_colgroups_kernel(SegmentDev seg, WarpBatchDev wb, StreamDev sd) {
  __shared__ __align__(16) unsigned long long s_cand[kColWarps][kWarpCand];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  unsigned long long *cand = s_cand[warp];
  const uint32_t k = wb.k;
  const uint32_t n_chunks = *sd.n_chunks;
  const uint32_t n_blocks = (seg.doc_count + 511u) >> 9;
  const uint32_t n_ranges = (n_blocks + kColItemBlocks - 1) / kColItemBlocks;
  const uint64_t total_items = (uint64_t)n_chunks * n_ranges;
  unsigned long long n_tests = 0, n_looked = 0;
  WarpCand wc;

  uint32_t item = 0;
  if (lane == 0) item = atomicAdd(sd.col_counter, 1u);
  item = __shfl_sync(0xFFFFFFFFu, item, 0);
  while (item < total_items) {
    uint32_t next_item = 0;
    if (lane == 0) next_item = atomicAdd(sd.col_counter, 1u);
    const uint32_t r = item / n_chunks, ci = item - r * n_chunks;  // range-major: the chunks of a column meet in L2
    const ColChunk ch = sd.chunks[ci];
    ColQ cq;
    cq.qslot = 0;
    cq.qi = 0;
    cq.w = 0.0f;
    cq.ncol = 0;
    cq.nsp = 0;
    cq.filter = -1;
    unsigned long long thr = ~0ull;
    if (lane < (int)ch.count) {
      cq = sd.colq[ch.begin + lane];
      thr = ld_cg_u64(wb.thr_key + cq.qi);
    }
    const float4 *col = reinterpret_cast<const float4 *>(seg.cols + ch.sc_base);
#pragma unroll 1
    for (uint32_t b = r * kColItemBlocks; b < min((r + 1) * kColItemBlocks, n_blocks); b++) {
      const uint32_t d0 = b << 9;
      float4 p[4];
#pragma unroll
      for (int x = 0; x < 4; x++) p[x] = __ldg(col + (d0 >> 2) + x * 32 + lane);
      uint32_t top = 0u;
#pragma unroll
      for (int x = 0; x < 4; x++)
        top = max(top, max(max(__float_as_uint(p[x].x), __float_as_uint(p[x].y)), max(__float_as_uint(p[x].z), __float_as_uint(p[x].w))));
      top = __reduce_max_sync(0xFFFFFFFFu, top);
      if (top == 0u) continue;  // the column is empty here
      // lane g = query g of the chunk.  One column: max over the block of c * w is max(c) * w.  More columns: the
      // other columns only add, so the block is summed per doc below.
      const uint32_t thr_bits = thr == kThrInit ? 0u : (uint32_t)(thr >> 32);
      const bool live = lane < (int)ch.count;
      const bool look = live && (cq.ncol > 1u || __float_as_uint(__fmul_rn(__uint_as_float(top), cq.w)) >= thr_bits);
      n_tests += live ? 1u : 0u;
      uint32_t hits = __ballot_sync(0xFFFFFFFFu, look);
      while (hits) {
        const int g = __ffs(hits) - 1;
        hits &= hits - 1;
        const uint32_t qslot = __shfl_sync(0xFFFFFFFFu, cq.qslot, g), qi = __shfl_sync(0xFFFFFFFFu, cq.qi, g);
        const uint32_t ncol = __shfl_sync(0xFFFFFFFFu, cq.ncol, g), nsp = __shfl_sync(0xFFFFFFFFu, cq.nsp, g);
        const int32_t filter = __shfl_sync(0xFFFFFFFFu, cq.filter, g);
        const float w = __shfl_sync(0xFFFFFFFFu, cq.w, g);
        const unsigned long long qthr = __shfl_sync(0xFFFFFFFFu, thr, g);
        const QTerm *qts = wb.qterms + (uint64_t)qslot * kWarpMaxTerms;
        n_looked++;
        float4 v[4];
#pragma unroll
        for (int x = 0; x < 4; x++) v[x] = make_float4(__fmul_rn(p[x].x, w), __fmul_rn(p[x].y, w), __fmul_rn(p[x].z, w), __fmul_rn(p[x].w, w));
        for (uint32_t t = nsp + 1; t < nsp + ncol; t++) {  // the query's further columns, slot order
          const float4 *c2 = reinterpret_cast<const float4 *>(seg.cols + __ldg(&qts[t].sc_base)) + (d0 >> 2) + lane;
          const float w2 = __ldg(&qts[t].weight);
          float4 c[4];
#pragma unroll
          for (int x = 0; x < 4; x++) c[x] = __ldg(c2 + x * 32);
#pragma unroll
          for (int x = 0; x < 4; x++) {
            v[x].x = __fadd_rn(v[x].x, __fmul_rn(c[x].x, w2));
            v[x].y = __fadd_rn(v[x].y, __fmul_rn(c[x].y, w2));
            v[x].z = __fadd_rn(v[x].z, __fmul_rn(c[x].z, w2));
            v[x].w = __fadd_rn(v[x].w, __fmul_rn(c[x].w, w2));
          }
        }
        uint32_t cut = qthr == kThrInit ? 0u : (uint32_t)(qthr >> 32);
        uint32_t mx = 0u;
#pragma unroll
        for (int x = 0; x < 4; x++)
          mx = max(mx, max(max(__float_as_uint(v[x].x), __float_as_uint(v[x].y)), max(__float_as_uint(v[x].z), __float_as_uint(v[x].w))));
        if (!__any_sync(0xFFFFFFFFu, mx >= cut && mx != 0u)) continue;
        // ---- per doc: the docs that can enter the top k ----
        wc.begin(cand, qthr, k, lane);
#pragma unroll 1
        for (int x = 0; x < 4; x++) {
          const uint32_t bits[4] = {__float_as_uint(v[x].x), __float_as_uint(v[x].y), __float_as_uint(v[x].z), __float_as_uint(v[x].w)};
#pragma unroll
          for (int e = 0; e < 4; e++) {
            const uint32_t doc = d0 + x * 128 + lane * 4 + e;
            cut = wc.thr == kThrInit ? 0u : (uint32_t)(wc.thr >> 32);
            bool pass = bits[e] >= cut && bits[e] != 0u && doc < seg.doc_count;
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
            wc.offer(seg, wb, qi, filter, pass, doc, __uint_as_float(bits[e]));
          }
        }
        wc.merge(wb, qi);
        if (lane == g) thr = max(thr, wc.thr);
      }
    }
    item = __shfl_sync(0xFFFFFFFFu, next_item, 0);
  }
  if (sd.counters && lane == 0) {
    if (n_looked) atomicAdd(sd.counters + 1, n_looked);
    if (n_tests) atomicAdd(sd.counters + 2, n_tests);
  }
}

}  // namespace slg
