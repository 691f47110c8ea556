// slg_items_kernel.cuh — K2/K3/K5, posting-driven variant: the automatic choice for plain OR queries with
// k <= 32 and <= 8 terms per query on segments with resident scores (the shape of BASELINE.json configs[1]).
//
// Why another kernel.  slg_score_warp_kernel keeps one f32 slot per doc of a sub-tile and pays for the whole
// sub-tile whatever the query touches: a column fill or a clear (64 shared-memory wavefronts per 2048 docs) and a
// scan.  ncu (profiles/r1_v8_*) shows that pipe, L1TEX, as the bound.  At C2 a query's non-column terms touch
// about 5 % of a sub-tile's docs, so this kernel makes every step proportional to POSTINGS:
//
//   non-column terms   scattered into the warp's accumulator A (read-modify-write per posting, query order);
//                      afterwards the same posting runs are walked again: read A[doc], write 0, and only a
//                      partial that can still reach the query's k-th score becomes a candidate.  No scan, no
//                      clear: A is all zero again when the walk ends.
//   column terms       (df >= N / dense_den: a doc-indexed f32 column, seg.cols) are streamed through REGISTERS:
//                      v = sum of the query's column slices, compared against the k-th score on the fly.  They
//                      never touch shared memory.  A doc that also sits in a scattered list is left to the
//                      posting walk, which adds the column values of exactly that doc (one 4-byte gather each).
//
// Float contract (include/searchlite_gpu.h): a doc's contributions are summed over the query's terms WITHOUT a
// column first, then over the terms WITH one, each group in query order — brute_force (query/wand.rs:527-548)
// on that permutation of the query.  slg_build_qterms_kernel lays the term slots out in that order, so "slot
// order" below IS the declared order and every path (A + column gathers, register stream, exact rescoring)
// produces the same bits.
//
// PRUNE (execution wand / bmw; exact, byte-identical to bm25):
//   seeds        slg_seed_items_kernel: one warp per query scores the kSeedItems doc-range items with the largest
//                upper bound first, so every query has a tight k-th score before the sweep starts
//   item filter  slg_filter_items_kernel: every (query, sub-tile) whose bound  sum_t w_t * ub_t(sub-tile)  is below
//                that score is dropped; the survivors form a compact item list (tile-major)
//   MaxScore     per surviving sub-tile the terms with the smallest bounds whose sum stays below ms_frac of the
//                k-th score are "non-essential": not scattered, not streamed; a candidate's exact score is then
//                recomputed over ALL terms (binary search in the sub-tile's posting range / column gather)
//   block skip   an essential column is streamed in 512-doc blocks; a block whose exact column maxima
//                (seg.col_tmax) cannot reach the k-th score is not read
//   bounds       per (term, sub-tile): exact column maximum for column terms; for the others the maximum of the
//                resident scores of the 32-posting mini-blocks that overlap the sub-tile (seg.mb_max) — tighter
//                than TermState::block_upper_bound (query/wand.rs:238-251), which takes the block's max tf at
//                the segment's minimum doc length, and safe for the same reason (it dominates every posting).
// Without PRUNE (execution bm25) every posting of every term is visited — scattered or streamed — and the only
// shortcut is the one the warp kernel already had: when the largest value a sub-tile could hold is below the
// k-th score nobody looks for candidates there.
#pragma once
#include "slg_warp_kernel.cuh"

namespace slg {

constexpr uint32_t kSeedItems = 4;   // doc-range items (kSubPerGroup sub-tiles each) scored per query by the seed pass

struct ItemsDev {
  const uint2 *items;        // PRUNE sweep: (tg * n_queries + slot, sub-tile mask); nullptr = every (tg, slot), all sub-tiles
  const uint32_t *n_items;   // device counter next to the list
  uint8_t *done;             // [n_groups * n_queries] 1 = scored by the seed pass
  uint2 *items_out;          // filter kernel output
  uint32_t *n_items_out;
  uint32_t items_cap;
  unsigned long long *counters;  // [4] postings scattered, sub-tiles dropped, column blocks streamed, items (nullptr = not counted)
};

// ---- per-(term, sub-tile) bounds -----------------------------------------------------------------------------
// ub[u][j] = an upper bound of the unit-weight contribution of term u inside sub-tile j; 0 iff the term has no
// posting there.  Column terms: exact maximum of the column (col_tmax, per 512 docs).  Others: maximum of the
// mini-block maxima overlapping the posting range [rng[j], rng[j+1]).
static __global__ void slg_items_bounds_kernel(SegmentDev seg, const uint32_t *ut_term, uint32_t n_uterms, const uint32_t *rng,
                                        uint32_t sub_docs, uint32_t n_sub, float *ub) {
  const uint64_t gid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= (uint64_t)n_uterms * n_sub) return;
  const uint32_t u = (uint32_t)(gid / n_sub), j = (uint32_t)(gid % n_sub);
  const uint32_t term = ut_term[u];
  float b = 0.0f;
  if (term < seg.n_terms) {
    const int32_t col = seg.term_col ? seg.term_col[term] : -1;
    if (col >= 0) {
      const uint32_t d0 = j * sub_docs, d1 = min(d0 + sub_docs, seg.doc_count);
      const float *tm = seg.col_tmax + (uint64_t)col * seg.tmax_stride;
      for (uint32_t blk = d0 / 512u; blk <= (d1 - 1u) / 512u; blk++) b = fmaxf(b, tm[blk]);
    } else {
      const uint32_t lo = rng[(uint64_t)u * (n_sub + 1) + j], hi = rng[(uint64_t)u * (n_sub + 1) + j + 1];
      if (hi > lo) {
        const float *mb = seg.mb_max + (seg.term_start[term] >> 5);
        for (uint32_t m = lo >> 5; m <= (hi - 1u) >> 5; m++) b = fmaxf(b, mb[m]);
      }
    }
  }
  ub[(uint64_t)u * n_sub + j] = b;
}

// ---- the per-item body ---------------------------------------------------------------------------------------
// One warp, one (query slot, group of kSubPerGroup sub-tiles, mask of the sub-tiles to score).  Shared layout per
// warp as in slg_score_warp_kernel: acc f32[sub_docs] | cand u64[64] | qt QTerm[8] | rb u32[8][9] | ubs f32[8][8].
template <bool PRUNE>
__device__ __forceinline__ void score_item(const SegmentDev &seg, const WarpBatchDev &wb, const uint32_t qslot, const uint32_t tg,
                                           const uint32_t mask, float *acc, unsigned long long *cand, QTerm *qt, uint32_t *rb,
                                           float *ubs, const int lane, uint32_t (&st)[4]) {
  const uint32_t k = wb.k, sub_docs = wb.sub_docs;
  const uint32_t lt_mask = (1u << lane) - 1u;
  uint32_t *pend = reinterpret_cast<uint32_t *>(cand + 32);  // parked doc ids (see slg_score_warp_kernel)
  const QHead head = wb.qheads[qslot];
  const uint32_t nt = head.nt;
  if (lane < 16) {
    const uint4 v = __ldg(reinterpret_cast<const uint4 *>(wb.qterms + (uint64_t)qslot * kWarpMaxTerms) + lane);
    reinterpret_cast<uint4 *>(qt)[lane] = v;
  }
  unsigned long long thr = ld_cg_u64(wb.thr_key + head.qi);
  __syncwarp();
  const uint32_t sub0 = tg * kSubPerGroup;
  {
    // posting ranges (non-column terms) and bounds (all terms) of sub-tiles sub0 .. sub0+8: lane = t*4 + c
    const uint32_t t = lane >> 2, c = lane & 3;
    if (t < nt) {
      const QTerm q = qt[t];
      if (!(q.flags & 4u)) {
        const uint32_t *row = wb.rng + (uint64_t)q.uterm * (wb.n_sub + 1);
        for (uint32_t j = c; j <= kSubPerGroup; j += 4) rb[t * kRbStride + j] = __ldg(row + min(sub0 + j, wb.n_sub));
      }
      const float *urow = wb.sub_ub + (uint64_t)q.uterm * wb.n_sub;
      for (uint32_t j = c; j < kSubPerGroup; j += 4) ubs[t * 8 + j] = (sub0 + j < wb.n_sub) ? __ldg(urow + sub0 + j) : 0.0f;
    }
  }
  __syncwarp();

  uint32_t cnt = 0;  // pending candidates in cand[]
  uint32_t npend = 0;
  uint32_t n_post = 0, n_skipped = 0, n_streamed = 0;

  // append one ballot round of keys to the warp's candidates; past 32 pending: sort, keep the best k
  auto push_keys = [&](bool pass, unsigned long long key) {
    const uint32_t bal = __ballot_sync(0xFFFFFFFFu, pass);
    if (!bal) return;
    if (pass) cand[cnt + __popc(bal & lt_mask)] = key;
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
  };
  // accept (api/reader.rs:3009-3036) for a doc whose exact score is known, then the candidate buffer
  auto offer = [&](bool have, uint32_t doc, float s) {
    const unsigned long long key = ((unsigned long long)__float_as_uint(s) << 32) | (unsigned long long)(0xFFFFFFFFu - doc);
    bool pass = have && key > thr && s != 0.0f;
    if (pass) pass = (seg.live_bits[doc >> 5] >> (doc & 31)) & 1u;
    if (pass && head.filter >= 0) pass = (wb.filter_bits[head.filter][doc >> 5] >> (doc & 31)) & 1u;
    if (pass) pass = cursor_accepts(wb.q_cursor, wb.q_saw, head.qi, key);
    push_keys(pass, key);
  };

#pragma unroll 1
  for (uint32_t j = 0; j < kSubPerGroup; j++) {
    if (!((mask >> j) & 1u)) continue;
    const uint32_t sub = sub0 + j;
    if (sub >= wb.n_sub) break;
    const uint32_t tile_lo = sub * sub_docs;
    // lanes 0..nt-1 hold one term each
    float mine_ub = 0.0f;
    bool mine_col = false, mine_has = false;
    if (lane < (int)nt && (qt[lane].flags & 1u)) {
      mine_col = (qt[lane].flags & 4u) != 0u;
      mine_ub = ubs[lane * 8 + j] * qt[lane].weight;
      mine_has = mine_col ? (mine_ub > 0.0f) : (rb[lane * kRbStride + j + 1] > rb[lane * kRbStride + j]);
      if (!mine_has) mine_ub = 0.0f;
    }
    const uint32_t any = __ballot_sync(0xFFFFFFFFu, mine_has);
    if (any == 0u) continue;
    const uint32_t colmask = __ballot_sync(0xFFFFFFFFu, mine_has && mine_col);
    const uint32_t spmask = any & ~colmask;
    float ub_all = mine_ub;
#pragma unroll
    for (int o = 4; o > 0; o >>= 1) ub_all += __shfl_xor_sync(0xFFFFFFFFu, ub_all, o);  // terms live in lanes 0..7
    ub_all = __shfl_sync(0xFFFFFFFFu, ub_all, 0);
    bool have_thr = thr != kThrInit;
    float thr_score = __uint_as_float((uint32_t)(thr >> 32));
    if (PRUNE && have_thr && ub_all * 1.00001f < thr_score) {  // no doc of this sub-tile can beat the k-th key
      n_skipped++;
      continue;
    }
    // ---- MaxScore: the non-essential terms of this sub-tile (smallest bounds first, capped at ms_frac of the k-th score)
    uint32_t nmask = 0;
    float sum_n = 0.0f;
    if (PRUNE && have_thr) {
      float pre = 0.0f;
#pragma unroll
      for (int m = 0; m < (int)kWarpMaxTerms; m++) {
        const float u = __shfl_sync(0xFFFFFFFFu, mine_ub, m);
        if (u < mine_ub || (u == mine_ub && m <= lane)) pre += u;
      }
      const bool noness = mine_has && pre * 1.00001f < thr_score * wb.ms_frac;
      nmask = __ballot_sync(0xFFFFFFFFu, noness);
      float sn = noness ? pre : 0.0f;
#pragma unroll
      for (int o = 4; o > 0; o >>= 1) sn = fmaxf(sn, __shfl_xor_sync(0xFFFFFFFFu, sn, o));
      sum_n = __shfl_sync(0xFFFFFFFFu, sn, 0);
    }
    const uint32_t esp = spmask & ~nmask;   // scattered
    const uint32_t ecol = colmask & ~nmask; // streamed
    // bounds of what a scattered partial lacks: every column term + the non-essential non-column terms
    float rest_sp = (((colmask | (spmask & nmask)) >> lane) & 1u) ? mine_ub : 0.0f;
#pragma unroll
    for (int o = 4; o > 0; o >>= 1) rest_sp += __shfl_xor_sync(0xFFFFFFFFu, rest_sp, o);
    rest_sp = __shfl_sync(0xFFFFFFFFu, rest_sp, 0);
    const bool slow_sp = (spmask & nmask) != 0u;  // a scattered partial lacks a non-column term: exact rescoring
    const bool slow_col = nmask != 0u;            // a streamed value lacks some term: exact rescoring

    // exact score of one parked doc per lane: every term of the query at the doc, slot (= declared) order
    auto rescore = [&](bool have, uint32_t doc) {
      float s = 0.0f;
      if (have) {
        for (uint32_t t = 0; t < nt; t++) {
          const QTerm &q = qt[t];
          if (!(q.flags & 1u)) continue;
          float c = 0.0f;
          if (q.flags & 4u) {
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
      offer(have, doc, s);
    };
    auto rescore_pending = [&]() {
      const uint32_t n = npend;
      npend = 0;
      const uint32_t d0 = lane < (int)n ? pend[lane] : 0u;
      const uint32_t d1 = 32u + lane < n ? pend[32 + lane] : 0u;
      __syncwarp();
      rescore(lane < (int)n, d0);
      if (n > 32u) rescore(32u + lane < n, d1);
    };
    auto park = [&](bool pass, uint32_t doc) {
      const uint32_t bal = __ballot_sync(0xFFFFFFFFu, pass);
      if (!bal) return;
      if (pass) pend[npend + __popc(bal & lt_mask)] = doc;
      npend += __popc(bal);
      __syncwarp();
      if (npend >= 32u) rescore_pending();
    };

    // ---- 1. scatter the essential non-column terms, slot order ----
    uint32_t wmax = 0;
    {
      bool first = true;
      for (uint32_t m = esp; m; m &= m - 1) {
        const uint32_t t = __ffs(m) - 1;
        const uint32_t lo = rb[t * kRbStride + j], hi = rb[t * kRbStride + j + 1];
        const QTerm q = qt[t];
        const uint32_t *dptr = seg.post_doc + q.base;
        const float *sptr = wb.scores + q.base;
        n_post += hi - lo;
        if (q.weight == 1.0f) {
          if (first) accumulate_staged<true, true>(dptr, sptr, lo, hi, tile_lo, 1.0f, acc, lane, wmax);
          else accumulate_staged<false, true>(dptr, sptr, lo, hi, tile_lo, 1.0f, acc, lane, wmax);
        } else {
          if (first) accumulate_staged<true, false>(dptr, sptr, lo, hi, tile_lo, q.weight, acc, lane, wmax);
          else accumulate_staged<false, false>(dptr, sptr, lo, hi, tile_lo, q.weight, acc, lane, wmax);
        }
        first = false;
        __syncwarp();
      }
    }

    // ---- 2. stream the essential column terms through registers ----
    if (ecol) {
      // docs of no scattered list: score <= v + non-essential bounds
#pragma unroll 1
      for (uint32_t i0 = 0; i0 < sub_docs; i0 += 512) {
        if (tile_lo + i0 >= seg.doc_count) break;
        have_thr = thr != kThrInit;
        thr_score = __uint_as_float((uint32_t)(thr >> 32));
        if (PRUNE && have_thr) {
          // exact maxima of the streamed columns inside these (at most two) 512-doc blocks
          const uint32_t b0 = (tile_lo + i0) >> 9, b1 = (tile_lo + min(i0 + 384u, sub_docs - 128u)) >> 9;
          float bb = 0.0f;
          if (lane < (int)nt && ((ecol >> lane) & 1u)) {
            const float *tm = seg.col_tmax + (qt[lane].sc_base / seg.col_stride) * seg.tmax_stride;
            bb = fmaxf(__ldg(tm + b0), __ldg(tm + b1)) * qt[lane].weight;
          }
#pragma unroll
          for (int o = 4; o > 0; o >>= 1) bb += __shfl_xor_sync(0xFFFFFFFFu, bb, o);
          bb = __shfl_sync(0xFFFFFFFFu, bb, 0);
          if ((bb + sum_n) * 1.00001f < thr_score) continue;
        }
        n_streamed++;
        float4 v[4];
#pragma unroll
        for (int r = 0; r < 4; r++) v[r] = make_float4(0, 0, 0, 0);
        for (uint32_t m = ecol; m; m &= m - 1) {
          const uint32_t t = __ffs(m) - 1;
          const float4 *cp = reinterpret_cast<const float4 *>(seg.cols + qt[t].sc_base + tile_lo + i0) + lane;
          const float w = qt[t].weight;
          float4 c[4];
#pragma unroll
          for (int r = 0; r < 4; r++) c[r] = (i0 + r * 128 < sub_docs) ? __ldg(cp + r * 32) : make_float4(0, 0, 0, 0);
#pragma unroll
          for (int r = 0; r < 4; r++) {
            v[r].x = __fadd_rn(v[r].x, __fmul_rn(c[r].x, w));
            v[r].y = __fadd_rn(v[r].y, __fmul_rn(c[r].y, w));
            v[r].z = __fadd_rn(v[r].z, __fmul_rn(c[r].z, w));
            v[r].w = __fadd_rn(v[r].w, __fmul_rn(c[r].w, w));
          }
        }
        // a streamed value below this cannot reach the k-th score even with every non-essential bound added
        uint32_t cut = 0u;
        if (have_thr) {
          const float cf = thr_score * 0.99998f - sum_n * 1.00002f;
          cut = cf > 0.0f ? __float_as_uint(cf) : 0u;
        }
        uint32_t top = 0u;
#pragma unroll
        for (int r = 0; r < 4; r++)
          top = max(top, max(max(__float_as_uint(v[r].x), __float_as_uint(v[r].y)), max(__float_as_uint(v[r].z), __float_as_uint(v[r].w))));
        if (!__any_sync(0xFFFFFFFFu, top >= cut && top != 0u)) continue;
#pragma unroll 1
        for (int r = 0; r < 4; r++) {
          const uint32_t bits[4] = {__float_as_uint(v[r].x), __float_as_uint(v[r].y), __float_as_uint(v[r].z), __float_as_uint(v[r].w)};
          const uint32_t mx = max(max(bits[0], bits[1]), max(bits[2], bits[3]));
          if (!__any_sync(0xFFFFFFFFu, mx >= cut && mx != 0u)) continue;
#pragma unroll
          for (int e = 0; e < 4; e++) {
            const uint32_t slot = i0 + r * 128 + lane * 4 + e;
            bool pass = bits[e] >= cut && bits[e] != 0u && slot < sub_docs;
            // a doc of a scattered list is handled by the posting walk below (its partial is in A)
            if (pass && esp) pass = __float_as_uint(acc[slot]) == 0u;
            if (slow_col) park(pass, tile_lo + slot);
            else offer(pass, tile_lo + slot, __uint_as_float(bits[e]));
          }
        }
      }
      if (npend) rescore_pending();
    }

    // ---- 3. walk the scattered runs again: collect what can still qualify, leave A zero ----
    if (esp) {
      have_thr = thr != kThrInit;
      thr_score = __uint_as_float((uint32_t)(thr >> 32));
      uint32_t cut = 0u;
      if (have_thr) {
        const float cf = thr_score * 0.99998f - rest_sp * 1.00002f;
        cut = cf > 0.0f ? __float_as_uint(cf) : 0u;
      }
      const bool collect = __reduce_max_sync(0xFFFFFFFFu, wmax) >= cut;
      for (uint32_t m = esp; m; m &= m - 1) {
        const uint32_t t = __ffs(m) - 1;
        const uint32_t lo = rb[t * kRbStride + j], hi = rb[t * kRbStride + j + 1];
        const uint32_t *dptr = seg.post_doc + qt[t].base;
#pragma unroll 1
        for (uint32_t i = lo; i < hi; i += 32) {
          const bool in = i + lane < hi;
          uint32_t slot = 0;
          if (in) slot = __ldg(dptr + i + lane) - tile_lo;
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
          if (slow_sp) {
            park(pass, tile_lo + slot);
          } else {
            // every non-column term is in A: add the column terms of exactly this doc, slot order
            float s = v;
            if (pass)
              for (uint32_t cm = colmask; cm; cm &= cm - 1) {
                const uint32_t ct = __ffs(cm) - 1;
                const float c = __ldg(seg.cols + qt[ct].sc_base + tile_lo + slot);
                if (c != 0.0f) s = __fadd_rn(s, __fmul_rn(c, qt[ct].weight));
              }
            offer(pass, tile_lo + slot, s);
          }
        }
        __syncwarp();
      }
      if (npend) rescore_pending();
    }
    __syncwarp();
  }

  st[0] += n_post;      // postings scattered
  st[1] += n_skipped;   // sub-tiles dropped by the bound
  st[2] += n_streamed;  // 512-doc column blocks streamed
  st[3] += 1;           // items

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
}

__device__ __forceinline__ void items_smem_setup(unsigned char *smem_raw, const WarpBatchDev &wb, int warp, int lane, float *&acc,
                                                 unsigned long long *&cand, QTerm *&qt, uint32_t *&rb, float *&ubs) {
  const size_t per_warp = warp_kernel_smem_per_warp(wb.sub_docs, false, true, 1);
  unsigned char *mine = smem_raw + (size_t)warp * per_warp;
  acc = reinterpret_cast<float *>(mine);
  cand = reinterpret_cast<unsigned long long *>(mine + (size_t)wb.sub_docs * 4);
  qt = reinterpret_cast<QTerm *>(cand + kWarpCand);
  rb = reinterpret_cast<uint32_t *>(qt + kWarpMaxTerms);
  ubs = reinterpret_cast<float *>(rb + kWarpMaxTerms * kRbStride);
  for (uint32_t i = lane * 4; i < wb.sub_docs; i += 128) *reinterpret_cast<float4 *>(acc + i) = make_float4(0, 0, 0, 0);
  __syncwarp();
}

__device__ __forceinline__ void items_flush_stats(const ItemsDev &it, const uint32_t (&st)[4], int lane) {
  if (it.counters && lane == 0)
    for (int i = 0; i < 4; i++)
      if (st[i]) atomicAdd(it.counters + i, (unsigned long long)st[i]);
}

// The sweep: persistent warps take items from a global counter — every (tg, slot), tile-major, or the filtered list.
template <bool PRUNE>
__global__ void __launch_bounds__(kThreads, 3) slg_score_items_kernel(SegmentDev seg, WarpBatchDev wb, ItemsDev it) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float *acc;
  unsigned long long *cand;
  QTerm *qt;
  uint32_t *rb;
  float *ubs;
  items_smem_setup(smem_raw, wb, warp, lane, acc, cand, qt, rb, ubs);
  const uint32_t total_items = it.items ? *it.n_items : wb.n_groups * wb.n_queries;
  uint32_t st[4] = {0, 0, 0, 0};
  uint32_t item = 0;
  if (lane == 0) item = atomicAdd(wb.work_counter, 1u);
  item = __shfl_sync(0xFFFFFFFFu, item, 0);
  while (item < total_items) {
    uint32_t next_item = 0;
    if (lane == 0) next_item = atomicAdd(wb.work_counter, 1u);  // consumed at the end of this item
    uint32_t id = item, mask = 0xFFu;
    if (it.items) {
      const uint2 e = __ldg(it.items + item);
      id = e.x;
      mask = e.y;
    }
    const uint32_t tg = id / wb.n_queries;
    score_item<PRUNE>(seg, wb, id - tg * wb.n_queries, tg, mask, acc, cand, qt, rb, ubs, lane, st);
    item = __shfl_sync(0xFFFFFFFFu, next_item, 0);
  }
  items_flush_stats(it, st, lane);
}

// Seeds (PRUNE): one warp per query slot.  Every lane finds the group of sub-tiles with the largest bound among
// the groups tg == lane (mod 32); the kSeedItems best of those 32 are scored one after the other, so the query's
// k-th score is tight before the filter looks at it.  The scored items are marked in it.done.
template <int UNUSED>
__global__ void __launch_bounds__(kThreads, 3) slg_seed_items_kernel(SegmentDev seg, WarpBatchDev wb, ItemsDev it) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float *acc;
  unsigned long long *cand;
  QTerm *qt;
  uint32_t *rb;
  float *ubs;
  items_smem_setup(smem_raw, wb, warp, lane, acc, cand, qt, rb, ubs);
  const uint32_t n_warps = gridDim.x * (kThreads / 32);
  uint32_t st[4] = {0, 0, 0, 0};
  for (uint32_t qslot = blockIdx.x * (kThreads / 32) + warp; qslot < wb.n_queries; qslot += n_warps) {
    const uint32_t nt = wb.qheads[qslot].nt;
    // bound of a group = max over its sub-tiles of the query's bound there
    float best = 0.0f;
    uint32_t best_tg = 0xFFFFFFFFu;
    for (uint32_t tg = lane; tg < wb.n_groups; tg += 32) {
      float g = 0.0f;
      for (uint32_t j = 0; j < kSubPerGroup && tg * kSubPerGroup + j < wb.n_sub; j++) {
        float u = 0.0f;
        for (uint32_t t = 0; t < nt; t++) {
          const QTerm &q = wb.qterms[(uint64_t)qslot * kWarpMaxTerms + t];
          if (q.flags & 1u) u += __ldg(wb.sub_ub + (uint64_t)q.uterm * wb.n_sub + tg * kSubPerGroup + j) * q.weight;
        }
        g = fmaxf(g, u);
      }
      if (g > best) {
        best = g;
        best_tg = tg;
      }
    }
    for (uint32_t s = 0; s < kSeedItems; s++) {
      // the lane holding the largest remaining bound (ties: lowest lane)
      const uint32_t mx = __reduce_max_sync(0xFFFFFFFFu, __float_as_uint(best));
      if (mx == 0u) break;
      const uint32_t who = __ffs(__ballot_sync(0xFFFFFFFFu, __float_as_uint(best) == mx)) - 1;
      const uint32_t tg = __shfl_sync(0xFFFFFFFFu, best_tg, who);
      if (lane == (int)who) best = 0.0f;
      score_item<true>(seg, wb, qslot, tg, 0xFFu, acc, cand, qt, rb, ubs, lane, st);
      if (lane == 0) it.done[(uint64_t)tg * wb.n_queries + qslot] = 1;
      __syncwarp();
    }
  }
  items_flush_stats(it, st, lane);
}

// Item filter (PRUNE): thread per (tg, slot).  A sub-tile survives iff some scored term has a posting there and the
// query's bound can still reach its k-th score; (item, mask) pairs of the survivors are appended tile-major.
static __global__ void __launch_bounds__(256) slg_filter_items_kernel(WarpBatchDev wb, ItemsDev it) {
  const uint32_t total = wb.n_groups * wb.n_queries;
  const uint32_t id = blockIdx.x * blockDim.x + threadIdx.x;
  uint32_t mask = 0;
  if (id < total && !it.done[id]) {
    const uint32_t tg = id / wb.n_queries, qslot = id - tg * wb.n_queries;
    const QHead head = wb.qheads[qslot];
    const unsigned long long thr = wb.thr_key[head.qi];
    const float thr_score = __uint_as_float((uint32_t)(thr >> 32));
    float ub[kSubPerGroup];
#pragma unroll
    for (uint32_t j = 0; j < kSubPerGroup; j++) ub[j] = 0.0f;
    for (uint32_t t = 0; t < head.nt; t++) {
      const QTerm &q = wb.qterms[(uint64_t)qslot * kWarpMaxTerms + t];
      if (!(q.flags & 1u)) continue;
      const float *urow = wb.sub_ub + (uint64_t)q.uterm * wb.n_sub + (uint64_t)tg * kSubPerGroup;
#pragma unroll
      for (uint32_t j = 0; j < kSubPerGroup; j++)
        if (tg * kSubPerGroup + j < wb.n_sub) ub[j] += __ldg(urow + j) * q.weight;
    }
#pragma unroll
    for (uint32_t j = 0; j < kSubPerGroup; j++)
      if (ub[j] > 0.0f && (thr == kThrInit || ub[j] * 1.00001f >= thr_score)) mask |= 1u << j;
  }
  // warp-aggregated append
  const uint32_t bal = __ballot_sync(0xFFFFFFFFu, mask != 0u);
  if (bal) {
    const int lane = threadIdx.x & 31;
    uint32_t base = 0;
    if (lane == 0) base = atomicAdd(it.n_items_out, (uint32_t)__popc(bal));
    base = __shfl_sync(0xFFFFFFFFu, base, 0);
    if (mask) {
      const uint32_t pos = base + __popc(bal & ((1u << lane) - 1u));
      if (pos < it.items_cap) it.items_out[pos] = make_uint2(id, mask);
    }
  }
}

}  // namespace slg
