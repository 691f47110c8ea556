// slg_reg_kernel.cuh — K2/K3, register-tile variant: the default for plain OR queries
// (no matcher), k <= 32 and <= 8 terms per query — the shape of BASELINE.json configs[1].
//
// Why: a batch names the same head terms over and over (sum of df over term INSTANCES is ~17x the sum
// over UNIQUE terms), so the postings are served from L2, not DRAM, and the earlier kernels were
// bound first by shared-memory scatter wavefronts (profiles/r1_v4_warp_staged_summary.txt: L1TEX data
// pipe 86 %) and then by L2->SM bandwidth (every query re-reads its terms' bytes from L2).
//
//   resident scores   seg.post_score[i] = unit-weight BM25 contribution of posting i, computed once at
//                     segment load (same arithmetic as score_tf, query/wand.rs:269-286): a posting
//                     visit is (doc, score) -> add.
//   dense columns     a term with df >= doc_count / dense_den also has a doc-indexed f32 column
//                     (score or 0.0f).  Adding it to a tile is a 128-bit load + FADD per 4 docs, no
//                     scatter; x + 0.0f == x, so docs without the term are unaffected bit for bit.
//   tile ownership    one CTA owns a tile of TILE = 128*V docs at a time and sweeps ALL queries of the
//                     batch over it.  The tile's slices of the batch's most-used columns are staged in
//                     shared memory once per tile (H x TILE x 4 B) and every query that names one of
//                     them reads it from there; everything the CTA touches for the tile (remaining
//                     columns, posting streams, range table rows) sits in one narrow address range,
//                     which is what L1 can hold.
//   register tile     one warp = one (query, tile): lane L holds docs {128*i + 4*L .. +3 : i < V} in
//                     registers.  Sparse terms are scattered into a warp-private shared tile first,
//                     the tile is read once into the registers (and cleared), the columns are added in
//                     registers, and the registers are compared against the query's running k-th key.
//
// Summation order (the float contract of this kernel): the query's terms WITHOUT a column in query
// order, then the terms WITH a column in query order, one left fold.  That is brute_force
// (query/wand.rs:527-548) applied to a permutation of the query's terms; tests check it bit for bit
// against the oracle run on the permuted query and against the reference order under the 1e-5 rule.
//
// PRUNE (safe MaxScore-style block-max pruning, exact result): per (query, tile)
//   * sum of all terms' tile bounds < running k-th score          -> skip the tile;
//   * sum of the COLUMN terms' tile bounds < running k-th score   -> no doc without a sparse posting
//     can enter: only the docs of the sparse postings are completed (column values gathered per
//     candidate, same fold order) and the register pass is skipped;
//   * otherwise the full register pass.
#pragma once
#include "slg_warp_kernel.cuh"

namespace slg {

constexpr int kRegThreads = 512;  // 16 warps, one CTA per SM
constexpr int kRegWarps = kRegThreads / 32;

struct __align__(16) RTerm {  // 32 B, one query term resolved against one segment
  uint64_t base;     // first padded posting index (post_doc / post_score)
  uint64_t col_off;  // element offset of the term's column in seg.cols, ~0ull = no column
  uint32_t uterm;    // column of the transposed range / bound tables
  float weight;
  uint32_t hot;      // 1 + shared-memory slot of the column's staged slice, 0 = not staged
  uint32_t pad;
};

struct __align__(16) RHead {  // 16 B per query slot (processing order)
  uint32_t qi;       // original query index
  uint32_t ns;       // terms without a column (listed first)
  uint32_t nd;       // terms with a column
  int32_t filter;
};

struct RegBatchDev {
  const RTerm *rterms;      // [Q][kWarpMaxTerms]
  const RHead *rheads;      // [Q]
  const uint32_t *rng_t;    // [n_tiles+1][U]  transposed: row = tile boundary
  const float *ub_t;        // [n_tiles][U]    (PRUNE)
  const uint64_t *hot_cols; // [n_hot] element offset of each staged column in seg.cols
  const uint32_t *const *filter_bits;
  uint32_t n_queries, n_uterms, k, n_tiles, n_hot;
  unsigned long long *thr_key;
  uint32_t *topk_count, *lock;
  unsigned long long *topk_keys;
  uint32_t *work_counter;
  unsigned long long *stats;
};

// transposed plan tables: rng_t[j][u] = index of the first posting of unique term u with doc >= j*tile
// (plain lower_bound; replaces TermState::advance_to, query/wand.rs:205-232).  Column terms need no
// ranges unless bounds or statistics are wanted.
__global__ void slg_plan_ranges_t_kernel(SegmentDev seg, BatchDev bt, bool all_terms) {
  const uint64_t gid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= (uint64_t)bt.n_uterms * (bt.n_tiles + 1)) return;
  const uint32_t j = (uint32_t)(gid / bt.n_uterms), u = (uint32_t)(gid % bt.n_uterms);
  const uint32_t term = bt.ut_term[u];
  const uint32_t df = term < seg.n_terms ? seg.term_df[term] : 0u;
  uint32_t res = 0;
  if (df && (all_terms || !seg.term_col || seg.term_col[term] < 0)) {
    if (j == bt.n_tiles) {
      res = df;
    } else {
      const uint32_t target = j * bt.tile_docs;
      const uint32_t *d = seg.post_doc + seg.term_start[term];
      uint32_t lo = 0, hi = df;
      while (lo < hi) {
        const uint32_t mid = (lo + hi) >> 1;
        if (d[mid] < target) lo = mid + 1;
        else hi = mid;
      }
      res = lo;
    }
  }
  bt.ut_rng[gid] = res;
}

// ub_t[j][u]: unit-weight upper bound of unique term u inside tile j — max over the 128-posting
// blocks that overlap the tile of score_tf(block_max_tf, df, min_doc_len, ...) (query/wand.rs:238-251
// taken over the blocks that actually cover the doc range, which is what makes it safe)
__global__ void slg_plan_bounds_t_kernel(SegmentDev seg, BatchDev bt) {
  const uint64_t gid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= (uint64_t)bt.n_uterms * bt.n_tiles) return;  // gid = tile * U + u
  const uint32_t u = (uint32_t)(gid % bt.n_uterms);
  const uint32_t term = bt.ut_term[u];
  const uint32_t lo = bt.ut_rng[gid], hi = bt.ut_rng[gid + bt.n_uterms];
  float ub = 0.0f;
  if (hi > lo && term < seg.n_terms) {
    const uint32_t b0 = lo / kBlock, b1 = (hi - 1) / kBlock;
    const float *bm = seg.blk_max_tf + seg.term_blk[term];
    float mtf = 0.0f;
    for (uint32_t b = b0; b <= b1; b++) mtf = fmaxf(mtf, bm[b]);
    if (mtf > 0.0f) ub = bm25_contrib(mtf, seg.term_idf[term], seg.k1p1, seg.min_nk, 1.0f);
  }
  bt.ut_tile_ub[gid] = ub;
}

// canonical term order of the register kernel: sparse (no column) first, then column terms, both in
// query order.  hot_slot[u] = 1 + shared-memory slot of unique term u's column, or 0.  Runs once per
// segment per batch.
__global__ void slg_build_rterms_kernel(SegmentDev seg, BatchDev bt, const uint32_t *hot_slot, RTerm *rterms, RHead *rheads) {
  const uint32_t slot = blockIdx.x * blockDim.x + threadIdx.x;
  if (slot >= bt.n_queries) return;
  const uint32_t qi = bt.q_order[slot];
  const uint32_t t0 = bt.q_term_off[qi], nt = bt.q_term_off[qi + 1] - t0;
  uint32_t n_out = 0, ns = 0;
  RTerm *out = rterms + (uint64_t)slot * kWarpMaxTerms;
  for (int pass = 0; pass < 2; pass++) {
    for (uint32_t t = 0; t < nt && t < kWarpMaxTerms; t++) {
      const uint32_t u = bt.qt_uterm[t0 + t];
      const uint32_t term = bt.ut_term[u];
      const bool present = term < seg.n_terms;
      const int32_t col = present && seg.term_col ? seg.term_col[term] : -1;
      if ((col >= 0) != (pass == 1)) continue;
      RTerm r;
      r.base = present ? seg.term_start[term] : 0;
      r.col_off = col >= 0 ? (uint64_t)col * seg.col_stride : ~0ull;
      r.uterm = u;
      r.weight = bt.qt_weight[t0 + t];
      r.hot = col >= 0 && hot_slot ? hot_slot[u] : 0u;
      r.pad = 0;
      out[n_out++] = r;
    }
    if (pass == 0) ns = n_out;
  }
  for (uint32_t t = n_out; t < kWarpMaxTerms; t++) {
    RTerm r;
    r.base = 0;
    r.col_off = ~0ull;
    r.uterm = 0;
    r.weight = 0.0f;
    r.hot = r.pad = 0;
    out[t] = r;
  }
  RHead h;
  h.qi = qi;
  h.ns = ns;
  h.nd = n_out - ns;
  h.filter = bt.q_filter[qi];
  rheads[slot] = h;
}

template <int V>
__host__ __device__ constexpr size_t reg_kernel_smem_per_warp() {
  return (size_t)128 * V * 4 + kWarpCand * 8;  // M f32[128*V] | cand u64[64]
}

// append the keys of one ballot round to the warp's candidate buffer; when more than 32 are pending,
// sort, keep the best k and raise the local threshold (exact: nothing is dropped unsorted)
__device__ __forceinline__ void reg_push(bool pass, unsigned long long key, unsigned long long *cand, uint32_t &cnt,
                                         unsigned long long &thr, uint32_t k, int lane, uint32_t lt_mask) {
  const uint32_t bal = __ballot_sync(0xFFFFFFFFu, pass);
  if (bal == 0u) return;
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
}

__device__ __forceinline__ uint64_t shfl_u64(uint64_t v, int src) {
  const uint32_t lo = __shfl_sync(0xFFFFFFFFu, (uint32_t)v, src), hi = __shfl_sync(0xFFFFFFFFu, (uint32_t)(v >> 32), src);
  return ((uint64_t)hi << 32) | lo;
}

template <int V, bool PRUNE, bool STATS>
__global__ void __launch_bounds__(kRegThreads, 1) slg_score_reg_kernel(SegmentDev seg, RegBatchDev rb) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  constexpr uint32_t TILE = 128u * V;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  // layout: hot f32[n_hot][TILE] | per warp: M f32[TILE], cand u64[64]
  float *hot = reinterpret_cast<float *>(smem_raw);
  unsigned char *mine = smem_raw + (size_t)rb.n_hot * TILE * 4 + (size_t)warp * reg_kernel_smem_per_warp<V>();
  float *M = reinterpret_cast<float *>(mine);
  unsigned long long *cand = reinterpret_cast<unsigned long long *>(mine + (size_t)TILE * 4);
  __shared__ uint32_t s_tile, s_next_q;

  const uint32_t k = rb.k, U = rb.n_uterms;
  const uint32_t lt_mask = (1u << lane) - 1u;

  for (uint32_t i = lane * 4; i < TILE; i += 128) *reinterpret_cast<float4 *>(M + i) = make_float4(0, 0, 0, 0);

  for (;;) {
    __syncthreads();  // every warp is done with the previous tile's staged slices and counters
    if (threadIdx.x == 0) {
      s_tile = atomicAdd(rb.work_counter, 1u);
      s_next_q = 0;
    }
    __syncthreads();
    const uint32_t tile = s_tile;
    if (tile >= rb.n_tiles) break;
    const uint32_t tile_lo = tile * TILE;
    // ---- stage this tile's slices of the batch's hottest columns ----
    for (uint32_t i = threadIdx.x; i < rb.n_hot * (TILE / 4); i += kRegThreads) {
      const uint32_t h = i / (TILE / 4), o = i - h * (TILE / 4);
      const float4 v = ldg_stream_f4(reinterpret_cast<const float4 *>(seg.cols + rb.hot_cols[h] + tile_lo) + o);
      reinterpret_cast<float4 *>(hot)[i] = v;
    }
    __syncthreads();
    const uint32_t *rng0 = rb.rng_t + (uint64_t)tile * U;
    const float *ub0 = rb.ub_t + (uint64_t)tile * U;

    // ---- sweep the batch: one (query, tile) per warp at a time ----
    for (;;) {
      uint32_t qslot = 0;
      if (lane == 0) qslot = atomicAdd(&s_next_q, 1u);
      qslot = __shfl_sync(0xFFFFFFFFu, qslot, 0);
      if (qslot >= rb.n_queries) break;
      const RHead head = rb.rheads[qslot];
      const uint32_t ns = head.ns, nt = head.ns + head.nd;
      if (nt == 0) continue;
      // lane t < nt holds term t
      RTerm mt;
      mt.base = 0;
      mt.col_off = 0;
      mt.uterm = 0;
      mt.weight = 0.0f;
      mt.hot = 0;
      uint32_t lo = 0, hi = 0;
      float ub = 0.0f;
      if (lane < (int)nt) {
        const uint4 *src = reinterpret_cast<const uint4 *>(rb.rterms + (uint64_t)qslot * kWarpMaxTerms + lane);
        const uint4 a = __ldg(src), b = __ldg(src + 1);
        mt.base = ((uint64_t)a.y << 32) | a.x;
        mt.col_off = ((uint64_t)a.w << 32) | a.z;
        mt.uterm = b.x;
        mt.weight = __uint_as_float(b.y);
        mt.hot = b.z;
        if (STATS || lane < (int)ns) {
          lo = __ldg(rng0 + mt.uterm);
          hi = __ldg(rng0 + U + mt.uterm);
        }
        if (PRUNE) ub = __fmul_rn(__ldg(ub0 + mt.uterm), mt.weight);
      }
      unsigned long long thr = ld_cg_u64(rb.thr_key + head.qi);
      const uint32_t sparse_any = __ballot_sync(0xFFFFFFFFu, lane < (int)ns && hi > lo);
      if (head.nd == 0 && sparse_any == 0u) continue;

      uint32_t cnt = 0;
      uint32_t n_touched = 0, n_post = 0, n_skipped = 0;
      bool sparse_driven = false;
      float ub_d = 0.0f;
      bool skip = false;
      if (PRUNE) {
        float us = (lane < (int)ns) ? ub : 0.0f;
        float ud = (lane >= (int)ns) ? ub : 0.0f;
#pragma unroll
        for (int o = 4; o > 0; o >>= 1) {  // terms live in lanes 0..7
          us += __shfl_xor_sync(0xFFFFFFFFu, us, o);
          ud += __shfl_xor_sync(0xFFFFFFFFu, ud, o);
        }
        us = __shfl_sync(0xFFFFFFFFu, us, 0);
        ub_d = __shfl_sync(0xFFFFFFFFu, ud, 0);
        const float thr_score = __uint_as_float((uint32_t)(thr >> 32));
        if (thr != kThrInit) {
          // float sums are not exact: widen the bound before comparing (strict <: equal scores can
          // still win on doc id)
          skip = (us + ub_d) * 1.00001f < thr_score;
          sparse_driven = ub_d * 1.00001f < thr_score;
        }
      }
      if (skip) {
        if (STATS && lane == 0) atomicAdd(rb.stats + (uint64_t)head.qi * 4 + 2, 1ull);
        continue;
      }
      if (STATS) n_post = (sparse_driven && lane >= (int)ns) ? 0u : hi - lo;

      // ---- sparse terms: scatter (doc, score) into the warp's shared tile, query order ----
      if (sparse_any) {
        bool first = true;
#pragma unroll 1
        for (uint32_t t = 0; t < ns; t++) {
          if (!((sparse_any >> t) & 1u)) continue;
          const uint64_t base = shfl_u64(mt.base, t);
          const uint32_t tlo = __shfl_sync(0xFFFFFFFFu, lo, t), thi = __shfl_sync(0xFFFFFFFFu, hi, t);
          const float w = __shfl_sync(0xFFFFFFFFu, mt.weight, t);
          const uint32_t *dptr = seg.post_doc + base;
          const float *sptr = seg.post_score + base;
          if (w == 1.0f) {
            if (first) accumulate_staged<true, true>(dptr, sptr, tlo, thi, tile_lo, 1.0f, M, lane);
            else accumulate_staged<false, true>(dptr, sptr, tlo, thi, tile_lo, 1.0f, M, lane);
          } else {
            if (first) accumulate_staged<true, false>(dptr, sptr, tlo, thi, tile_lo, w, M, lane);
            else accumulate_staged<false, false>(dptr, sptr, tlo, thi, tile_lo, w, M, lane);
          }
          first = false;
          __syncwarp();
        }
      }

      if (PRUNE && sparse_driven) {
        // ---- only docs with a sparse posting can enter: complete those, clear the tile ----
        const float thr_score0 = __uint_as_float((uint32_t)(thr >> 32));
#pragma unroll 1
        for (uint32_t t = 0; t < ns; t++) {
          if (!((sparse_any >> t) & 1u)) continue;
          const uint32_t *dptr = seg.post_doc + shfl_u64(mt.base, t);
          const uint32_t tlo = __shfl_sync(0xFFFFFFFFu, lo, t), thi = __shfl_sync(0xFFFFFFFFu, hi, t);
#pragma unroll 1
          for (uint32_t i0 = tlo; i0 < thi; i0 += 32) {
            const uint32_t i = i0 + lane;
            bool pass = false;
            uint32_t doc = 0;
            float s = 0.0f;
            if (i < thi) {
              doc = __ldg(dptr + i);
              s = M[doc - tile_lo];
              M[doc - tile_lo] = 0.0f;  // a doc of two lists is completed once: the second visit reads 0
              pass = s != 0.0f && (s + ub_d) * 1.00001f >= thr_score0;
            }
            if (STATS) n_touched += s != 0.0f;
            if (__any_sync(0xFFFFFFFFu, pass)) {
#pragma unroll 1
              for (uint32_t dt = ns; dt < nt; dt++) {
                const uint64_t coff = shfl_u64(mt.col_off, dt);
                const float w = __shfl_sync(0xFFFFFFFFu, mt.weight, dt);
                if (pass) s = __fadd_rn(s, __fmul_rn(__ldg(seg.cols + coff + doc), w));
              }
              const unsigned long long key = ((unsigned long long)__float_as_uint(s) << 32) | (unsigned long long)(0xFFFFFFFFu - doc);
              if (pass) pass = key > thr;
              if (pass) pass = (seg.live_bits[doc >> 5] >> (doc & 31)) & 1u;
              if (pass && head.filter >= 0) pass = (rb.filter_bits[head.filter][doc >> 5] >> (doc & 31)) & 1u;
              reg_push(pass, key, cand, cnt, thr, k, lane, lt_mask);
            }
          }
          __syncwarp();
        }
      } else {
        // ---- registers <- shared tile (then cleared), or zero ----
        float4 R[V];
        if (sparse_any) {
#pragma unroll
          for (int i = 0; i < V; i++) {
            R[i] = *reinterpret_cast<const float4 *>(M + i * 128 + lane * 4);
            const uint32_t m = max(max(__float_as_uint(R[i].x), __float_as_uint(R[i].y)), max(__float_as_uint(R[i].z), __float_as_uint(R[i].w)));
            if (m != 0u) *reinterpret_cast<float4 *>(M + i * 128 + lane * 4) = make_float4(0, 0, 0, 0);
          }
        } else {
#pragma unroll
          for (int i = 0; i < V; i++) R[i] = make_float4(0, 0, 0, 0);
        }

        // ---- column terms in query order: staged slice (shared) or global, adds in registers ----
#pragma unroll 1
        for (uint32_t dt = ns; dt < nt; dt++) {
          const uint32_t hs = __shfl_sync(0xFFFFFFFFu, mt.hot, dt);
          const float w = __shfl_sync(0xFFFFFFFFu, mt.weight, dt);
          const uint64_t coff = shfl_u64(mt.col_off, dt);
          float4 c[V];
          if (hs) {
            const float4 *cp = reinterpret_cast<const float4 *>(hot + (size_t)(hs - 1) * TILE) + lane;
#pragma unroll
            for (int i = 0; i < V; i++) c[i] = cp[i * 32];
          } else {
            const float4 *cp = reinterpret_cast<const float4 *>(seg.cols + coff + tile_lo) + lane;
#pragma unroll
            for (int i = 0; i < V; i++) c[i] = __ldg(cp + i * 32);
          }
          if (w == 1.0f) {
#pragma unroll
            for (int i = 0; i < V; i++) {
              R[i].x = __fadd_rn(R[i].x, c[i].x);
              R[i].y = __fadd_rn(R[i].y, c[i].y);
              R[i].z = __fadd_rn(R[i].z, c[i].z);
              R[i].w = __fadd_rn(R[i].w, c[i].w);
            }
          } else {
#pragma unroll
            for (int i = 0; i < V; i++) {
              R[i].x = __fadd_rn(R[i].x, __fmul_rn(c[i].x, w));
              R[i].y = __fadd_rn(R[i].y, __fmul_rn(c[i].y, w));
              R[i].z = __fadd_rn(R[i].z, __fmul_rn(c[i].z, w));
              R[i].w = __fadd_rn(R[i].w, __fmul_rn(c[i].w, w));
            }
          }
        }

        // ---- compare the registers with the running k-th key ----
        const uint32_t thr_hi = (uint32_t)(thr >> 32);
        uint32_t mx = 0;
#pragma unroll
        for (int i = 0; i < V; i++) {
          const uint32_t b0 = __float_as_uint(R[i].x), b1 = __float_as_uint(R[i].y), b2 = __float_as_uint(R[i].z), b3 = __float_as_uint(R[i].w);
          mx = max(mx, max(max(b0, b1), max(b2, b3)));
          if (STATS) n_touched += (b0 != 0u) + (b1 != 0u) + (b2 != 0u) + (b3 != 0u);
        }
        if (__any_sync(0xFFFFFFFFu, mx >= thr_hi && mx != 0u)) {
          // rare after warm-up: park the registers in the (now zero) shared tile and walk it
#pragma unroll
          for (int i = 0; i < V; i++) *reinterpret_cast<float4 *>(M + i * 128 + lane * 4) = R[i];
          __syncwarp();
#pragma unroll 1
          for (uint32_t i0 = 0; i0 < TILE; i0 += 128) {
            const uint32_t i = i0 + lane * 4;
            const float4 v = *reinterpret_cast<const float4 *>(M + i);
            const uint32_t bits[4] = {__float_as_uint(v.x), __float_as_uint(v.y), __float_as_uint(v.z), __float_as_uint(v.w)};
            const uint32_t m = max(max(bits[0], bits[1]), max(bits[2], bits[3]));
            if (m != 0u) *reinterpret_cast<float4 *>(M + i) = make_float4(0, 0, 0, 0);
            if (__any_sync(0xFFFFFFFFu, m >= (uint32_t)(thr >> 32) && m != 0u)) {
#pragma unroll
              for (int e = 0; e < 4; e++) {
                const uint32_t doc = tile_lo + i + e;
                const unsigned long long key = ((unsigned long long)bits[e] << 32) | (unsigned long long)(0xFFFFFFFFu - doc);
                bool pass = bits[e] != 0u && key > thr && doc < seg.doc_count;
                if (pass) pass = (seg.live_bits[doc >> 5] >> (doc & 31)) & 1u;
                if (pass && head.filter >= 0) pass = (rb.filter_bits[head.filter][doc >> 5] >> (doc & 31)) & 1u;
                reg_push(pass, key, cand, cnt, thr, k, lane, lt_mask);
              }
            }
          }
          __syncwarp();
        }
      }

      if (STATS) {
        for (int o = 16; o > 0; o >>= 1) {
          n_touched += __shfl_xor_sync(0xFFFFFFFFu, n_touched, o);
          n_post += __shfl_xor_sync(0xFFFFFFFFu, n_post, o);
        }
        if (lane == 0) {
          if (n_touched) atomicAdd(rb.stats + (uint64_t)head.qi * 4 + 0, (unsigned long long)n_touched);
          if (n_post) atomicAdd(rb.stats + (uint64_t)head.qi * 4 + 1, (unsigned long long)n_post);
          if (cnt) atomicAdd(rb.stats + (uint64_t)head.qi * 4 + 3, (unsigned long long)cnt);
        }
        (void)n_skipped;
      }

      // ---- merge into the query's global top-k (push_top_k, query/wand.rs:905-916) ----
      if (cnt > 0) {
        const unsigned long long thr_now = ld_cg_u64(rb.thr_key + head.qi);
        const bool useful = lane < (int)cnt && cand[lane] > thr_now;  // cnt <= 32 after every append
        if (__any_sync(0xFFFFFFFFu, useful)) {
          if (lane == 0) {
            while (atomicCAS(rb.lock + head.qi, 0u, 1u) != 0u) __nanosleep(64);
            __threadfence();
          }
          __syncwarp();
          const uint32_t ng = ld_cg_u32(rb.topk_count + head.qi);
          unsigned long long *gk = rb.topk_keys + (uint64_t)head.qi * k;
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
            st_cg_u32(rb.topk_count + head.qi, total);
            if (total == k) st_cg_u64(rb.thr_key + head.qi, cand[k - 1]);
            __threadfence();
            atomicExch(rb.lock + head.qi, 0u);
          }
          __syncwarp();
        }
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// residency: unit-weight contribution of every posting (one CTA of 128 threads per 128-posting
// block, like slg_transcode_csr_kernel) and the dense columns
__global__ void __launch_bounds__(128) slg_score_postings_kernel(SegmentDev seg, uint32_t n_blocks, float *post_score) {
  const uint32_t blk = blockIdx.x;
  if (blk >= n_blocks) return;
  __shared__ uint32_t s_term;
  if (threadIdx.x == 0) {
    uint64_t lo = 0, hi = seg.n_terms;  // last term with term_blk[t] <= blk
    while (lo + 1 < hi) {
      const uint64_t mid = (lo + hi) >> 1;
      if (seg.term_blk[mid] <= blk) lo = mid;
      else hi = mid;
    }
    s_term = (uint32_t)lo;
  }
  __syncthreads();
  const uint32_t term = s_term;
  const uint32_t i = (blk - seg.term_blk[term]) * kBlock + threadIdx.x;
  const uint32_t df = seg.term_df[term];
  if (i >= df) return;
  const uint64_t base = seg.term_start[term];
  const uint32_t doc = seg.post_doc[base + i];
  uint32_t tf = seg.post_tf[base + i];
  const uint64_t wide = seg.term_wide[term];
  if (tf == 255u && wide != ~0ull) tf = seg.tf_wide[wide + i];
  post_score[base + i] = bm25_contrib_fast(tf, seg.term_idf[term], seg.k1p1, seg.nk[doc], 1.0f);
}

// grid (chunks, n_cols): column c holds the scores of term col_terms[c] at their doc slots
__global__ void slg_fill_columns_kernel(SegmentDev seg, const uint32_t *col_terms, uint32_t n_cols, float *cols) {
  const uint32_t c = blockIdx.y;
  if (c >= n_cols) return;
  const uint32_t term = col_terms[c];
  const uint32_t df = seg.term_df[term];
  const uint64_t base = seg.term_start[term];
  float *col = cols + (uint64_t)c * seg.col_stride;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < df; i += gridDim.x * blockDim.x)
    col[seg.post_doc[base + i]] = seg.post_score[base + i];
}

}  // namespace slg
