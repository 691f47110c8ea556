// slg_warp_kernel.cuh — K2/K3, warp-autonomous variant (the default for k <= 32 and <= 8 terms
// per query).
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
#pragma once
#include "slg_kernels.cuh"

namespace slg {

constexpr uint32_t kWarpMaxTerms = 8;   // per-query term slots of the padded QTerm table
constexpr uint32_t kWarpMaxK = 32;
constexpr uint32_t kSubPerGroup = 8;    // sub-tiles per work item
constexpr uint32_t kWarpCand = 64;

struct __align__(16) QTerm {  // one query term resolved against one segment (32 B)
  uint64_t base;      // term_start: first padded posting index
  uint64_t wide;      // offset into tf_wide or ~0ull
  uint32_t uterm;     // row of the range / bound tables
  float idf;
  float weight;
  uint32_t flags;     // bit0 scored, bit1 valid, bits 8..15 group
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
  const uint32_t *const *filter_bits;
  uint32_t n_queries, k, sub_docs, n_sub, n_groups;
  unsigned long long *thr_key;
  uint32_t *topk_count, *lock;
  unsigned long long *topk_keys;
  uint32_t *work_counter;
  unsigned long long *stats;
};

// resolve the batch's query terms against one segment (runs once per segment per batch)
__global__ void slg_build_qterms_kernel(SegmentDev seg, BatchDev bt, QTerm *qterms, QHead *qheads) {
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
  for (uint32_t t = 0; t < kWarpMaxTerms; t++) {
    QTerm r;
    r.base = 0;
    r.wide = ~0ull;
    r.uterm = 0;
    r.idf = 0.0f;
    r.weight = 0.0f;
    r.flags = 0;
    if (t < nt) {
      const uint32_t u = bt.qt_uterm[t0 + t];
      const uint32_t term = bt.ut_term[u];
      r.base = seg.term_start[term];
      r.wide = seg.term_wide[term];
      r.uterm = u;
      r.idf = seg.term_idf[term];
      r.weight = bt.qt_weight[t0 + t];
      r.flags = (bt.qt_flags[t0 + t] & 1u) | 2u | ((uint32_t)bt.qt_group[t0 + t] << 8);
    }
    qterms[(uint64_t)slot * kWarpMaxTerms + t] = r;
  }
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

template <bool MATCHER, bool PRUNE, bool STATS>
__global__ void __launch_bounds__(kThreads) slg_score_warp_kernel(SegmentDev seg, WarpBatchDev wb) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  constexpr int kWarps = kThreads / 32;
  const uint32_t sub_docs = wb.sub_docs;
  // layout: [kWarps][sub_docs] f32 | [kWarps][64] u64 | [kWarps][sub_docs] u8
  float *acc = reinterpret_cast<float *>(smem_raw) + (size_t)warp * sub_docs;
  unsigned long long *cand =
      reinterpret_cast<unsigned long long *>(smem_raw + (size_t)kWarps * sub_docs * 4) + (size_t)warp * kWarpCand;
  uint8_t *gmask = smem_raw + (size_t)kWarps * sub_docs * 4 + (size_t)kWarps * kWarpCand * 8 + (size_t)warp * sub_docs;

  const uint32_t k = wb.k;
  const uint32_t total_items = wb.n_groups * wb.n_queries;
  const uint32_t lt_mask = (1u << lane) - 1u;

  for (uint32_t i = lane * 4; i < sub_docs; i += 128) *reinterpret_cast<float4 *>(acc + i) = make_float4(0, 0, 0, 0);
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
    // lanes 0..7: one query-term record each; every lane: the head
    const QHead head = wb.qheads[qslot];
    const uint32_t nt = head.nt;
    QTerm my;
    my.base = 0;
    my.wide = ~0ull;
    my.uterm = 0;
    my.idf = 0.0f;
    my.weight = 0.0f;
    my.flags = 0;
    if (lane < (int)kWarpMaxTerms) {
      const uint4 *src = reinterpret_cast<const uint4 *>(wb.qterms + (uint64_t)qslot * kWarpMaxTerms + lane);
      const uint4 a = __ldg(src), b = __ldg(src + 1);
      my.base = (uint64_t)a.x | ((uint64_t)a.y << 32);
      my.wide = (uint64_t)a.z | ((uint64_t)a.w << 32);
      my.uterm = b.x;
      my.idf = __uint_as_float(b.y);
      my.weight = __uint_as_float(b.z);
      my.flags = b.w;
    }
    unsigned long long thr = ld_cg_u64(wb.thr_key + head.qi);
    const uint32_t sub0 = tg * kSubPerGroup;
    // range boundaries: lane j (0..8) holds boundary sub0+j of term t in rb[t]
    uint32_t rb[kWarpMaxTerms];
    float ubv[kWarpMaxTerms];
#pragma unroll
    for (uint32_t t = 0; t < kWarpMaxTerms; t++) {
      const uint32_t u = __shfl_sync(0xFFFFFFFFu, my.uterm, t);
      rb[t] = 0;
      ubv[t] = 0.0f;
      if (t < nt) {
        const uint32_t b = min(sub0 + (uint32_t)lane, wb.n_sub);
        if (lane <= (int)kSubPerGroup) rb[t] = __ldg(wb.rng + (uint64_t)u * (wb.n_sub + 1) + b);
        if (PRUNE && lane < (int)kSubPerGroup && sub0 + lane < wb.n_sub) ubv[t] = __ldg(wb.sub_ub + (uint64_t)u * wb.n_sub + sub0 + lane);
      }
    }

    uint32_t cnt = 0;            // pending candidates in cand[]
    uint32_t n_touched = 0, n_post = 0, n_skipped = 0;
    const uint32_t masks = head.masks;

    for (uint32_t j = 0; j < kSubPerGroup; j++) {
      const uint32_t sub = sub0 + j;
      if (sub >= wb.n_sub) break;
      const uint32_t tile_lo = sub * sub_docs;
      const uint32_t tile_n = min(sub_docs, seg.doc_count - tile_lo);
      // ranges of every term in this sub-tile
      uint32_t tot = 0;
      float ub = 0.0f;
      uint32_t lo_t[kWarpMaxTerms], hi_t[kWarpMaxTerms];
#pragma unroll
      for (uint32_t t = 0; t < kWarpMaxTerms; t++) {
        lo_t[t] = __shfl_sync(0xFFFFFFFFu, rb[t], j);
        hi_t[t] = __shfl_sync(0xFFFFFFFFu, rb[t], j + 1);
        const uint32_t fl = __shfl_sync(0xFFFFFFFFu, my.flags, t);
        if (t < nt && (fl & 1u)) {
          tot += hi_t[t] - lo_t[t];
          if (PRUNE) ub += __shfl_sync(0xFFFFFFFFu, ubv[t], j) * __shfl_sync(0xFFFFFFFFu, my.weight, t);
        }
      }
      if (tot == 0) continue;
      if (PRUNE) {
        const float thr_score = __uint_as_float((uint32_t)(thr >> 32));
        if (thr != kThrInit && ub * 1.00001f < thr_score) {  // see slg_score_tiles_kernel
          n_skipped++;
          continue;
        }
      }
      if (STATS) n_post += tot;

      // ---- accumulate ----
      bool first = true;
#pragma unroll
      for (uint32_t t = 0; t < kWarpMaxTerms; t++) {
        if (t >= nt) break;
        const uint32_t lo = lo_t[t], hi = hi_t[t];
        TermCtx tc;
        const uint64_t base = __shfl_sync(0xFFFFFFFFu, my.base, t);
        const uint64_t wide = __shfl_sync(0xFFFFFFFFu, my.wide, t);
        const uint32_t fl = __shfl_sync(0xFFFFFFFFu, my.flags, t);
        tc.idf = __shfl_sync(0xFFFFFFFFu, my.idf, t);
        tc.w = __shfl_sync(0xFFFFFFFFu, my.weight, t);
        if (hi > lo) {
          tc.dptr = seg.post_doc + base;
          tc.fptr = seg.post_tf + base;
          tc.wptr = wide != ~0ull ? seg.tf_wide + wide : nullptr;
          tc.scored = fl & 1u;
          tc.gbit = MATCHER ? (uint8_t)(1u << ((fl >> 8) & 7u)) : 0;
          if (first && tc.scored && !MATCHER) accumulate_term<MATCHER, true, 32>(seg, tc, lo, hi, tile_lo, acc, gmask, lane);
          else accumulate_term<MATCHER, false, 32>(seg, tc, lo, hi, tile_lo, acc, gmask, lane);
          if (tc.scored) first = false;
          __syncwarp();
        }
      }

      // ---- scan + clear; collect keys that beat the threshold ----
      const uint32_t thr_hi = (uint32_t)(thr >> 32);
      for (uint32_t i0 = 0; i0 < tile_n; i0 += 128) {
        const uint32_t i = i0 + lane * 4;
        uint32_t b0 = 0, b1 = 0, b2 = 0, b3 = 0, gm = 0;
        if (i < tile_n) {
          const float4 v = *reinterpret_cast<const float4 *>(acc + i);
          b0 = __float_as_uint(v.x);
          b1 = __float_as_uint(v.y);
          b2 = __float_as_uint(v.z);
          b3 = __float_as_uint(v.w);
        }
        const uint32_t m = max(max(b0, b1), max(b2, b3));
        if (m != 0u) {
          *reinterpret_cast<float4 *>(acc + i) = make_float4(0, 0, 0, 0);
          if (STATS) n_touched += (b0 != 0u) + (b1 != 0u) + (b2 != 0u) + (b3 != 0u);
        }
        if (MATCHER && i < tile_n) {
          gm = *reinterpret_cast<const uint32_t *>(gmask + i);
          if (gm) *reinterpret_cast<uint32_t *>(gmask + i) = 0u;
        }
        if (__any_sync(0xFFFFFFFFu, m >= thr_hi && m != 0u)) {
          const uint32_t bits[4] = {b0, b1, b2, b3};
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
            const uint32_t bal = __ballot_sync(0xFFFFFFFFu, pass);
            if (bal) {
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
            }
          }
        }
      }
      __syncwarp();
    }

    if (STATS) {
      for (int o = 16; o > 0; o >>= 1) n_touched += __shfl_xor_sync(0xFFFFFFFFu, n_touched, o);
      if (lane == 0) {
        if (n_touched) atomicAdd(wb.stats + (uint64_t)head.qi * 4 + 0, (unsigned long long)n_touched);
        if (n_post) atomicAdd(wb.stats + (uint64_t)head.qi * 4 + 1, (unsigned long long)n_post);
        if (cnt) atomicAdd(wb.stats + (uint64_t)head.qi * 4 + 3, (unsigned long long)cnt);
      }
    }
    if (PRUNE && STATS && lane == 0 && n_skipped) atomicAdd(wb.stats + (uint64_t)head.qi * 4 + 2, (unsigned long long)n_skipped);

    // ---- merge into the query's global top-k (push_top_k, query/wand.rs:905-916) ----
    if (cnt > 0) {
      const unsigned long long thr_now = ld_cg_u64(wb.thr_key + head.qi);
      const bool useful = (lane < (int)cnt && cand[lane] > thr_now) || (lane + 32 < (int)cnt && cand[lane + 32] > thr_now);
      if (__any_sync(0xFFFFFFFFu, useful)) {
        if (cnt > 32) {  // (cannot happen: cnt <= 32 after every append) keep the merge buffer bounded
          for (uint32_t z = cnt + lane; z < kWarpCand; z += 32) cand[z] = 0ull;
          __syncwarp();
          warp_sort64_desc(cand, lane);
          cnt = min(cnt, k);
        }
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
