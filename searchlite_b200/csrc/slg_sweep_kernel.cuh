// slg_sweep_kernel.cuh — K2/K3, tile-sweep variant with register accumulators (kernel choice 3 with
// option heavy_kernel = 1) for plain OR queries (no matcher), k <= 32 and <= 8 terms per query — the
// shape of BASELINE.json configs[1].  Correct on every parity test, not (yet) the fastest path: see
// profiles/r1_sweep_experiments.txt for the six variants measured in round 1 and what limits them.
//
// Why this shape.  A batch names the same head terms over and over (at C2 the sum of df over term
// INSTANCES is ~17x the sum over UNIQUE terms, and 83 % of all posting visits belong to the ~90 terms
// with df >= N/8).  The scatter kernel (slg_score_warp_kernel) is bound by the shared-memory pipe:
// every visit is a read-modify-write and every (query, sub-tile) ends with a clear.
//
//   resident scores   seg.post_score / sw.post_pair: unit-weight BM25 contribution of every posting,
//                     computed once at segment load with score_tf's arithmetic (query/wand.rs:269-286).
//   dense columns     a term with df >= doc_count / dense_den also has a doc-indexed f32 column
//                     (score or +0.0f).  Adding a column to a tile is a 128-bit load + 4 FADD per 4
//                     docs; x + 0.0f == x, so docs without the term are unaffected bit for bit.
//   chunks            the swept queries are sorted by their first column term and cut into chunks of
//                     kSweepChunk; one CTA takes one (chunk, range of tiles) unit at a time.  The
//                     chunk's most-used columns are bulk-copied (cp.async.bulk + mbarrier) into
//                     shared memory one block of kSweepBlockDocs docs ahead, double buffered.
//   item records      before the sweep, slg_sweep_records_kernel resolves every query's sparse terms
//                     against every tile (posting ranges from the range table, a compacted list of
//                     the non-empty ones, a bitmap of where each term starts in their concatenation,
//                     the tile's upper bound for the pruned modes) into an 80-byte record per
//                     (query, tile).  All per-(query, tile) bookkeeping is scalar, embarrassingly
//                     parallel code there; the sweep streams the records.
//   register tile     one warp = one (query, tile) at a time: lane L holds docs {128*i + 4*L .. +3 :
//                     i < V} in registers.  The column terms are added in registers (staged slice from
//                     shared memory, conflict-free 128-bit reads) and compared against the query's
//                     k-th score; only if the query also has sparse postings in the tile is the
//                     register tile parked in a warp-private shared tile, the (doc, score) postings
//                     added to it and the touched docs re-checked.  No accumulator scan, no clearing.
//   cp.async rings    inside a block every warp walks its own interleaved share of the (query, tile)
//                     items; the record of item g + 4 and the first 64 postings of item g + 2 are in
//                     flight (warp-private cp.async rings) while item g is computed.
//   seed pass         the same kernel first runs the first tiles alone, which gives every query a
//                     useful threshold before many CTAs start merging into the same top-k lists.
//
// Queries for which a sweep over every tile would be wasted work (no column term and few postings)
// are "light": they go to slg_score_warp_kernel in the same batch run.
//
// Summation order (the float contract of this kernel, shared with slg_score_warp_kernel<COLS>): the
// query's terms WITH a column in query order, then the terms WITHOUT one in query order, one left
// fold.  That is brute_force (query/wand.rs:527-548) applied to a permutation of the query's terms;
// tests check it bit for bit against the oracle run on the permuted query and against the reference
// order under the 1e-5 rule.
//
// PRUNE (safe, exact result): a (query, tile) item is skipped iff
//   sum over sparse terms with postings in the tile of w * term-wide bound (query/wand.rs:289-303)
//   + sum over column terms of w * max of the column inside the tile          (exact tile maximum)
// is below the query's running k-th score (strictly; widened by 1e-5 for float summation order).
#pragma once
#include "slg_warp_kernel.cuh"

namespace slg {

constexpr int kSweepThreads = 512;  // 16 warps, one CTA per SM
constexpr int kSweepWarps = kSweepThreads / 32;
constexpr uint32_t kSweepChunk = 64;       // query slots per chunk
constexpr uint32_t kSweepBlockDocs = 4096; // docs per staged block of column slices (SB = 4096 / TILE tiles)
constexpr uint32_t kSweepSlotWords = 32;   // static description of one query slot
constexpr uint32_t kSweepRecWords = 20;    // per-(slot, tile) record
constexpr uint32_t kSweepMaxSlots = 16384; // query slots per launch
constexpr uint32_t kSweepTileGroup = 8;    // tiles resolved by one thread of the record kernel

// static slot description, 32 words:
//   [0..7]   row of the tile-range table per term position (sparse terms, and column terms when
//            statistics are wanted); positions: sparse terms first, then column terms, query order
//   [8..15]  sparse: first padded posting index of the term (u32); column: column index
//   [16,17]  per column term c one byte: index of its slice among the chunk's staged columns, or
//            255 = read the column from global memory
//   [18] query index  [19] filter id  [20] ns | ncol << 4 | any_weight << 8
// per-(slot, tile) record, 20 words — the sparse terms of the slot resolved against the tile:
//   [0..3]  B: 128-bit map over the concatenated postings of the non-empty sparse terms (query order);
//           bit j set = the r-th (r >= 1) non-empty term starts at position j.  rank(j) = popc(B[0..j])
//   [4] tot | ncol << 16 | nne << 20 | ns << 24   [5] upper bound of the tile for this query (PRUNE)
//   [8..15] adj_r = first posting of the r-th non-empty sparse term - its position in the concatenation
//   [16] 3 bits per r: term position of the r-th non-empty sparse term (weights)
//   [17] postings of all terms of the query inside the tile (statistics)

struct SweepDev {
  const uint4 *sstat;           // [n_slots][8]
  const float *weights;         // [n_slots][8] term positions as in sstat
  const float *ubw;             // [n_slots][8] PRUNE: weight * term-wide bound (sparse) or weight (column)
  const uint32_t *rng;          // [rows][n_tiles + 1] first posting with doc >= tile * TILE
  const uint2 *post_pair;       // [n_post_padded] (doc, unit-weight score bits) of every posting
  const float *col_tmax;        // [n_cols][tmax_stride] column maxima per 512 docs (PRUNE)
  const uint32_t *chunk_cols;   // [n_chunks][kSweepStage] staged column of the chunk or ~0
  const uint32_t *const *filter_bits;
  uint32_t *records;            // [n_slots][n_tiles][kSweepRecWords]
  uint32_t n_slots, k, n_tiles, n_chunks;
  uint32_t tile_begin, tile_end;  // tiles of this launch
  uint32_t part_tiles;            // tiles per unit of work (chunk x tile range)
  uint32_t tmax_stride;
  unsigned long long *thr_key;
  uint32_t *topk_count, *lock;
  unsigned long long *topk_keys;
  uint32_t *work_counter;
  unsigned long long *stats;
};

// column slices staged per chunk: what fits next to the warps' private tiles
template <int V>
__host__ __device__ constexpr uint32_t sweep_stage() { return V >= 8 ? 3u : 4u; }
constexpr uint32_t kSweepStage = 4;  // row length of the chunk_cols table (>= sweep_stage<V>())

template <int V>
__host__ __device__ constexpr size_t sweep_smem_per_warp() {
  // M f32[128*V] | cand u64[64] | record ring u32[8][24] | posting ring u32[4][128]
  return (size_t)128 * V * 4 + kWarpCand * 8 + 768 + 2048;
}
template <int V>
__host__ __device__ constexpr size_t sweep_smem_bytes() {
  // slices f32[2][stage][kSweepBlockDocs] | mbarrier[2] | query state u32[kSweepChunk][8] | per-warp areas
  return (size_t)2 * sweep_stage<V>() * kSweepBlockDocs * 4 + 16 + (size_t)kSweepChunk * 32 + (size_t)kSweepWarps * sweep_smem_per_warp<V>();
}

// ---- mbarrier + bulk-copy (TMA) primitives for the slice staging ----
__device__ __forceinline__ void mbar_init(void *bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(void *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(void *bar, uint32_t parity) {
  const uint32_t a = (uint32_t)__cvta_generic_to_shared(bar);
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(a),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void bulk_copy_g2s(void *smem, const void *gmem, uint32_t bytes, void *bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   (uint32_t)__cvta_generic_to_shared(smem)),
               "l"(gmem), "r"(bytes), "r"((uint32_t)__cvta_generic_to_shared(bar))
               : "memory");
}

__device__ __forceinline__ void cp_async_16(void *smem, const void *gmem) {
  const uint32_t s = (uint32_t)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_4(void *smem, const void *gmem) {
  const uint32_t s = (uint32_t)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(s), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_8(void *smem, const void *gmem) {
  const uint32_t s = (uint32_t)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(s), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

// append the keys of one ballot round to the warp's candidate buffer; when more than 32 are pending,
// sort, keep the best k and raise the local threshold (exact: nothing is dropped unsorted)
__device__ __forceinline__ void sweep_push(bool pass, unsigned long long key, unsigned long long *cand, uint32_t &cnt,
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

// Slow path of one (query, tile): M holds the final scores of the tile.  The keys that beat the
// warp's current k-th key go to its local candidate list (at most 32 entries between calls; every key
// that can still be in the query's top k is in it).  state = {cnt, thr}.
struct SweepLocal {
  unsigned long long thr;
  uint32_t cnt;
};
template <int V>
__device__ __noinline__ SweepLocal sweep_collect(const float *M, unsigned long long *cand, uint32_t tile_lo, int32_t filter,
                                                  SweepLocal st, const SegmentDev &seg, const SweepDev &sw, int lane) {
  constexpr uint32_t TILE = 128u * V;
  const uint32_t k = sw.k;
  const uint32_t lt_mask = (1u << lane) - 1u;
  unsigned long long thr = st.thr;
  uint32_t cnt = st.cnt;
#pragma unroll 1
  for (uint32_t i0 = 0; i0 < TILE; i0 += 128) {
    const uint32_t i = i0 + lane * 4;
    const float4 v = *reinterpret_cast<const float4 *>(M + i);
    const uint32_t bits[4] = {__float_as_uint(v.x), __float_as_uint(v.y), __float_as_uint(v.z), __float_as_uint(v.w)};
    const uint32_t m = max(max(bits[0], bits[1]), max(bits[2], bits[3]));
    if (__any_sync(0xFFFFFFFFu, m >= (uint32_t)(thr >> 32) && m != 0u)) {
#pragma unroll
      for (int e = 0; e < 4; e++) {
        const uint32_t doc = tile_lo + i + e;
        const unsigned long long key = ((unsigned long long)bits[e] << 32) | (unsigned long long)(0xFFFFFFFFu - doc);
        bool pass = bits[e] != 0u && key > thr && doc < seg.doc_count;
        if (pass) pass = (seg.live_bits[doc >> 5] >> (doc & 31)) & 1u;
        if (pass && filter >= 0) pass = (sw.filter_bits[filter][doc >> 5] >> (doc & 31)) & 1u;
        sweep_push(pass, key, cand, cnt, thr, k, lane, lt_mask);
      }
    }
  }
  __syncwarp();
  st.thr = thr;
  st.cnt = cnt;
  return st;
}

// merge the warp's local candidates into the query's global top-k (push_top_k, query/wand.rs:905-916)
__device__ __noinline__ void sweep_merge_global(unsigned long long *cand, uint32_t cnt, uint32_t qi, const SweepDev &sw, int lane) {
  const uint32_t k = sw.k;
  const unsigned long long thr_now = ld_cg_u64(sw.thr_key + qi);
  const bool useful = lane < (int)cnt && cand[lane] > thr_now;  // cnt <= 32 after every append
  if (!__any_sync(0xFFFFFFFFu, useful)) return;
  if (lane == 0) {
    while (atomicCAS(sw.lock + qi, 0u, 1u) != 0u) __nanosleep(64);
    __threadfence();
  }
  __syncwarp();
  const uint32_t ng = ld_cg_u32(sw.topk_count + qi);
  unsigned long long *gk = sw.topk_keys + (uint64_t)qi * k;
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
    st_cg_u32(sw.topk_count + qi, total);
    if (total == k) st_cg_u64(sw.thr_key + qi, cand[k - 1]);
    __threadfence();
    atomicExch(sw.lock + qi, 0u);
  }
  __syncwarp();
}

// Sparse postings of an item with more than 128 of them in the tile (no columns, or unusually long
// lists): term by term straight from the range table, 32 postings at a time.  M holds the parked
// register tile.
template <bool WEIGHTS>
__device__ __noinline__ void sweep_scatter_long(float *M, uint32_t slot, uint32_t tile, uint32_t tile_lo, const SegmentDev &seg,
                                                const SweepDev &sw, int lane) {
  const uint32_t *sst = reinterpret_cast<const uint32_t *>(sw.sstat + (size_t)slot * 8);
  const uint32_t ns = __ldg(sst + 20) & 15u;
  uint32_t lo = 0, hi = 0, base = 0;
  float w = 1.0f;
  if ((uint32_t)lane < ns) {
    const uint32_t *p = sw.rng + (uint64_t)__ldg(sst + lane) * (sw.n_tiles + 1) + tile;
    lo = __ldg(p);
    hi = __ldg(p + 1);
    base = __ldg(sst + 8 + lane);
    if (WEIGHTS) w = __ldg(sw.weights + (size_t)slot * 8 + lane);
  }
#pragma unroll 1
  for (uint32_t t = 0; t < ns; t++) {
    const uint32_t start = __shfl_sync(0xFFFFFFFFu, base + lo, t), cnt = __shfl_sync(0xFFFFFFFFu, hi - lo, t);
    const float wt = __shfl_sync(0xFFFFFFFFu, w, t);
#pragma unroll 1
    for (uint32_t b = 0; b < cnt; b += 32) {
      const uint32_t j = b + lane;
      if (j < cnt) {
        const uint32_t doc = __ldg(seg.post_doc + start + j);
        float s = __ldg(seg.post_score + start + j);
        if (WEIGHTS) s = __fmul_rn(s, wt);
        M[doc - tile_lo] = __fadd_rn(M[doc - tile_lo], s);
      }
      __syncwarp();
    }
  }
}

// Slow path of one (query, tile): M holds the final scores of the tile.  Collect the keys that beat
// the query's current k-th key and merge them into the query's global top-k (push_top_k,
// query/wand.rs:905-916).  Returns the score bits of the query's k-th key afterwards (0 = none yet).
template <int V>
__device__ __noinline__ uint32_t sweep_collect_merge(const float *M, unsigned long long *cand, uint32_t tile_lo, uint32_t qi, int32_t filter,
                                                     const SegmentDev &seg, const SweepDev &sw, int lane, uint32_t *n_cand) {
  SweepLocal st;
  st.thr = ld_cg_u64(sw.thr_key + qi);
  st.cnt = 0;
  st = sweep_collect<V>(M, cand, tile_lo, filter, st, seg, sw, lane);
  *n_cand = st.cnt;
  if (st.cnt) sweep_merge_global(cand, st.cnt, qi, sw, lane);
  const unsigned long long t = ld_cg_u64(sw.thr_key + qi);
  return t == kThrInit ? 0u : (uint32_t)(t >> 32);
}

// One CTA = one chunk of kSweepChunk query slots over a range of tiles.  The slots of a chunk are
// neighbours in the column-sorted slot order, so a handful of column slices — bulk-copied (TMA) into
// shared memory one block of kSweepBlockDocs docs ahead — serve most of their column terms.  Inside a
// block every warp walks its own interleaved share of the (query, tile) items, so each warp sees a
// mix of cheap and expensive queries and the one barrier per block costs little.  Records and
// postings of the items arrive through warp-private cp.async rings, four and two items ahead.
template <int V, bool PRUNE, bool STATS, bool WEIGHTS>
__global__ void __launch_bounds__(kSweepThreads, 1) slg_score_sweep_kernel(const SegmentDev seg, const SweepDev sw) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  constexpr uint32_t TILE = 128u * V;
  constexpr uint32_t FULL = 0xFFFFFFFFu;
  constexpr uint32_t STAGE = sweep_stage<V>();
  constexpr uint32_t BT = kSweepBlockDocs / TILE;           // tiles per staged block
  constexpr uint32_t IPW = kSweepChunk * BT / kSweepWarps;  // items per warp and block
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float *slices = reinterpret_cast<float *>(smem_raw);  // [2][STAGE][kSweepBlockDocs]
  unsigned long long *mbar = reinterpret_cast<unsigned long long *>(smem_raw + (size_t)2 * STAGE * kSweepBlockDocs * 4);  // [2]
  uint32_t *q_state = reinterpret_cast<uint32_t *>(mbar + 2);  // [kSweepChunk][8]: cc lo, cc hi, ns | ncol << 4, thr, qi, filter, -, -
  unsigned char *mine = reinterpret_cast<unsigned char *>(q_state + kSweepChunk * 8) + (size_t)warp * sweep_smem_per_warp<V>();
  float *M = reinterpret_cast<float *>(mine);
  unsigned long long *cand = reinterpret_cast<unsigned long long *>(mine + (size_t)TILE * 4);
  uint32_t *recring = reinterpret_cast<uint32_t *>(mine + (size_t)TILE * 4 + kWarpCand * 8);  // [8][24]
  uint2 *postring = reinterpret_cast<uint2 *>(recring + 192);                                  // [4][2][32]
  __shared__ uint32_t s_unit;

  const uint32_t lt_mask = (1u << lane) - 1u;
  const uint32_t le_mask = lt_mask | (1u << lane);
  const uint2 *__restrict__ post_pair = sw.post_pair;
  const uint32_t n_parts = (sw.tile_end - sw.tile_begin + sw.part_tiles - 1) / sw.part_tiles;
  const uint32_t n_units = sw.n_chunks * n_parts;

  if (threadIdx.x == 0) {
    mbar_init(mbar, 1);
    mbar_init(mbar + 1, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  uint32_t phase0 = 0, phase1 = 0;  // parity of the next completion of each slice buffer

  for (;;) {
    __syncthreads();
    if (threadIdx.x == 0) s_unit = atomicAdd(sw.work_counter, 1u);
    __syncthreads();
    const uint32_t unit = s_unit;
    if (unit >= n_units) break;
    const uint32_t chunk = unit % sw.n_chunks, part = unit / sw.n_chunks;
    const uint32_t t0 = sw.tile_begin + part * sw.part_tiles, t1 = min(t0 + sw.part_tiles, sw.tile_end);
    const uint32_t n_blocks = (t1 - t0 + BT - 1) / BT;
    const uint32_t slot0 = chunk * kSweepChunk;
    const uint32_t nq = min(kSweepChunk, sw.n_slots - slot0);

    // ---- the chunk's staged columns: one bulk copy per column and block, issued by thread 0 ----
    auto issue_block = [&](uint32_t j) {  // block j of the unit -> buffer j & 1
      const uint32_t buf = j & 1u;
      uint32_t n_stage = 0;
#pragma unroll
      for (uint32_t i = 0; i < STAGE; i++)
        if (__ldg(sw.chunk_cols + (size_t)chunk * kSweepStage + i) != 0xFFFFFFFFu) n_stage = i + 1;
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      mbar_arrive_expect_tx(mbar + buf, n_stage * kSweepBlockDocs * 4);
      for (uint32_t i = 0; i < n_stage; i++) {
        const uint32_t c = __ldg(sw.chunk_cols + (size_t)chunk * kSweepStage + i);
        bulk_copy_g2s(slices + ((size_t)buf * STAGE + i) * kSweepBlockDocs,
                      seg.cols + (uint64_t)c * seg.col_stride + (uint64_t)(t0 + j * BT) * TILE, kSweepBlockDocs * 4, mbar + buf);
      }
    };
    if (threadIdx.x == 0) issue_block(0);
    // ---- the chunk's queries ----
    if (threadIdx.x < kSweepChunk) {
      uint32_t *qs = q_state + threadIdx.x * 8;
      if (threadIdx.x < nq) {
        const uint32_t *sst = reinterpret_cast<const uint32_t *>(sw.sstat + (size_t)(slot0 + threadIdx.x) * 8);
        const uint32_t qi = __ldg(sst + 18);
        const unsigned long long t = ld_cg_u64(sw.thr_key + qi);
        qs[0] = __ldg(sst + 16);
        qs[1] = __ldg(sst + 17);
        qs[2] = __ldg(sst + 20) & 255u;
        qs[3] = t == kThrInit ? 0u : (uint32_t)(t >> 32);
        qs[4] = qi;
        qs[5] = __ldg(sst + 19);
      } else {
        qs[2] = 0u;
      }
    }
    __syncthreads();

    // item g of this warp: block g / IPW; inside the block item i = warp + 16 * (g % IPW) of the
    // (query, tile) grid, tile-minor: this warp keeps one tile of the block and walks the queries
    auto item_q = [&](uint32_t g) -> uint32_t { return ((uint32_t)warp + kSweepWarps * (g % IPW)) / BT; };
    auto item_tile = [&](uint32_t g) -> uint32_t { return t0 + (g / IPW) * BT + ((uint32_t)warp + kSweepWarps * (g % IPW)) % BT; };
    const uint32_t n_items = n_blocks * IPW;
    const uint32_t *__restrict__ recs = sw.records + (size_t)slot0 * sw.n_tiles * kSweepRecWords;

    auto copy_record = [&](uint32_t g) {
      if (g >= n_items) return;
      const uint32_t q = item_q(g), tile = item_tile(g);
      if (q < nq && tile < t1 && lane < (int)(kSweepRecWords / 4))
        cp_async_16(recring + (g & 7u) * 24 + lane * 4, recs + ((size_t)q * sw.n_tiles + tile) * kSweepRecWords + lane * 4);
    };
    // rank of position 32 * rho + lane among the starts of the concatenated non-empty terms
    auto owner_rank = [&](const uint4 &b, int rho) -> uint32_t {
      uint32_t r = 0;
      if (rho > 0) r += __popc(b.x);
      if (rho > 1) r += __popc(b.y);
      if (rho > 2) r += __popc(b.z);
      const uint32_t bw = rho == 0 ? b.x : (rho == 1 ? b.y : (rho == 2 ? b.z : b.w));
      return r + __popc(bw & le_mask);
    };
    // rounds 0 and 1 of an item's postings: asynchronous copies into the posting ring
    auto copy_postings = [&](uint32_t g) {
      if (g >= n_items) return;
      if (item_q(g) >= nq || item_tile(g) >= t1) return;
      const uint32_t *rr = recring + (g & 7u) * 24;
      const uint4 b = *reinterpret_cast<const uint4 *>(rr);
      const uint32_t tot = rr[4] & 0xFFFFu;
      if (tot == 0u || tot > 128u) return;
      const uint32_t a = lane < 8 ? rr[8 + lane] : 0u;
      uint2 *pr = postring + (g & 3u) * 64;
      const uint32_t idx0 = __shfl_sync(FULL, a, owner_rank(b, 0)) + lane;
      if ((uint32_t)lane < tot) cp_async_8(pr + lane, post_pair + idx0);
      if (tot > 32u) {
        const uint32_t idx1 = __shfl_sync(FULL, a, owner_rank(b, 1)) + 32u + lane;
        if (32u + lane < tot) cp_async_8(pr + 32 + lane, post_pair + idx1);
      }
    };

    // ---- prologue of the warp's rings: records of items 0..3, postings of items 0 and 1 ----
    copy_record(0);
    copy_record(1);
    copy_record(2);
    copy_record(3);
    cp_async_commit();
    cp_async_wait<0>();
    __syncwarp();
    copy_postings(0);
    copy_postings(1);
    cp_async_commit();
    cp_async_wait<0>();
    __syncwarp();

    const float *blk = slices + lane * 4;
#pragma unroll 1
    for (uint32_t g = 0; g < n_items; g++) {
      if ((g % IPW) == 0u) {
        // ---- next block: every warp is done with the block before, so its buffer may be refilled ----
        const uint32_t j = g / IPW;
        __syncthreads();
        if (threadIdx.x == 0 && j + 1 < n_blocks) issue_block(j + 1);
        if (threadIdx.x < nq) {  // what other ranges of these queries have found meanwhile
          const unsigned long long t = ld_cg_u64(sw.thr_key + q_state[threadIdx.x * 8 + 4]);
          if (t != kThrInit) atomicMax(q_state + threadIdx.x * 8 + 3, (uint32_t)(t >> 32));
        }
        if ((j & 1u) == 0u) {
          mbar_wait(mbar, phase0);
          phase0 ^= 1u;
        } else {
          mbar_wait(mbar + 1, phase1);
          phase1 ^= 1u;
        }
        blk = slices + (size_t)(j & 1u) * STAGE * kSweepBlockDocs + lane * 4;
      }
      // ring upkeep: everything issued two items ago has landed (record g + 2, postings g)
      cp_async_wait<1>();
      __syncwarp();
      copy_record(g + 4);
      copy_postings(g + 2);
      cp_async_commit();

      const uint32_t q = item_q(g), t = item_tile(g);
      if (q >= nq || t >= t1) continue;
      // ---- the item (query q, tile t) ----
      const uint32_t tile_lo = t * TILE;
      const uint32_t *rr = recring + (g & 7u) * 24;
      const uint4 qs = *reinterpret_cast<const uint4 *>(q_state + q * 8);
      const uint32_t ns = qs.z & 15u, ncol = qs.z >> 4;
      const uint32_t thr_hi = qs.w;
      const uint4 b = *reinterpret_cast<const uint4 *>(rr);
      const uint32_t tot = rr[4] & 0xFFFFu;
      if ((tot | ncol) == 0u) continue;
      if (PRUNE && thr_hi != 0u && __uint_as_float(rr[5]) * 1.00001f < __uint_as_float(thr_hi)) {
        if (STATS && lane == 0) atomicAdd(sw.stats + (uint64_t)q_state[q * 8 + 4] * 4 + 2, 1ull);
        continue;
      }
      // rounds 2 and 3 travel while the columns are added
      uint2 p2 = make_uint2(tile_lo, 0u), p3 = make_uint2(tile_lo, 0u);
      if (tot > 64u && tot <= 128u) {
        const uint32_t a = lane < 8 ? rr[8 + lane] : 0u;
        uint32_t idx = __shfl_sync(FULL, a, owner_rank(b, 2)) + 64u + lane;
        if (64u + lane < tot) p2 = __ldg(post_pair + idx);
        if (tot > 96u) {
          idx = __shfl_sync(FULL, a, owner_rank(b, 3)) + 96u + lane;
          if (96u + lane < tot) p3 = __ldg(post_pair + idx);
        }
      }
      float4 R[V];
#pragma unroll
      for (int v = 0; v < V; v++) R[v] = make_float4(0, 0, 0, 0);
      // ---- column terms in query order: staged slice (shared) or global, adds in registers ----
      const uint64_t cc = ((uint64_t)qs.y << 32) | qs.x;
      const uint32_t in_blk = ((t - t0) % BT) * TILE;
#pragma unroll 1
      for (uint32_t c = 0; c < ncol; c++) {
        const uint32_t code = (uint32_t)(cc >> (8 * c)) & 255u;
        float w = 1.0f;
        if (WEIGHTS) w = __ldg(sw.weights + (size_t)(slot0 + q) * 8 + ns + c);
        if (code < STAGE) {
          const float4 *cp = reinterpret_cast<const float4 *>(blk + (size_t)code * kSweepBlockDocs + in_blk);
#pragma unroll
          for (int v = 0; v < V; v++) {
            float4 cv = cp[v * 32];
            if (WEIGHTS) {
              cv.x = __fmul_rn(cv.x, w);
              cv.y = __fmul_rn(cv.y, w);
              cv.z = __fmul_rn(cv.z, w);
              cv.w = __fmul_rn(cv.w, w);
            }
            R[v].x = __fadd_rn(R[v].x, cv.x);
            R[v].y = __fadd_rn(R[v].y, cv.y);
            R[v].z = __fadd_rn(R[v].z, cv.z);
            R[v].w = __fadd_rn(R[v].w, cv.w);
          }
        } else {
          const uint32_t col = __ldg(reinterpret_cast<const uint32_t *>(sw.sstat + (size_t)(slot0 + q) * 8) + 8 + ns + c);
          const float4 *cp = reinterpret_cast<const float4 *>(seg.cols + (uint64_t)col * seg.col_stride + tile_lo) + lane;
#pragma unroll
          for (int h = 0; h < V; h += 4) {
            float4 cv[4];
#pragma unroll
            for (int v = 0; v < 4; v++) cv[v] = __ldg(cp + (h + v) * 32);
#pragma unroll
            for (int v = 0; v < 4; v++) {
              if (WEIGHTS) {
                cv[v].x = __fmul_rn(cv[v].x, w);
                cv[v].y = __fmul_rn(cv[v].y, w);
                cv[v].z = __fmul_rn(cv[v].z, w);
                cv[v].w = __fmul_rn(cv[v].w, w);
              }
              R[h + v].x = __fadd_rn(R[h + v].x, cv[v].x);
              R[h + v].y = __fadd_rn(R[h + v].y, cv[v].y);
              R[h + v].z = __fadd_rn(R[h + v].z, cv[v].z);
              R[h + v].w = __fadd_rn(R[h + v].w, cv[v].w);
            }
          }
        }
      }
      // ---- compare the registers with the running k-th score ----
      uint32_t mx = 0;
#pragma unroll
      for (int v = 0; v < V; v++) {
        const uint32_t b0 = __float_as_uint(R[v].x), b1 = __float_as_uint(R[v].y), b2 = __float_as_uint(R[v].z), b3 = __float_as_uint(R[v].w);
        mx = max(mx, max(max(b0, b1), max(b2, b3)));
      }
      bool hit = mx >= thr_hi && mx != 0u;
      if (tot) {
        // ---- sparse terms: park the register tile, add the (doc, score) postings in query order ----
#pragma unroll
        for (int v = 0; v < V; v++) *reinterpret_cast<float4 *>(M + v * 128 + lane * 4) = R[v];
        __syncwarp();
        if (tot <= 128u) {
          const uint2 *pr = postring + (g & 3u) * 64;
          const uint2 p0 = pr[lane], p1 = pr[32 + lane];
          float wst = 1.0f;
          uint32_t posmap = 0;
          if (WEIGHTS) {
            wst = lane < 8 ? __ldg(sw.weights + (size_t)(slot0 + q) * 8 + lane) : 1.0f;
            posmap = rr[16];
          }
          // one round: lanes of one term hit distinct docs; terms are applied one after the other
          auto round = [&](int rho, uint2 p) {
            if (tot <= 32u * rho) return;
            const bool valid = 32u * rho + lane < tot;
            float *slot_p = M + (valid ? p.x - tile_lo : 0u);
            float s = __uint_as_float(p.y);
            const uint32_t bw = rho == 0 ? b.x : (rho == 1 ? b.y : (rho == 2 ? b.z : b.w));
            if (WEIGHTS) s = __fmul_rn(s, __shfl_sync(FULL, wst, (posmap >> (3 * owner_rank(b, rho))) & 7u));
            if (bw == 0u) {
              if (valid) *slot_p = __fadd_rn(*slot_p, s);
            } else {
              const uint32_t myr = __popc(bw & le_mask);  // rank inside the round
              const uint32_t r_hi = __popc(bw);
              for (uint32_t r = 0; r <= r_hi; r++) {
                if (valid && myr == r) *slot_p = __fadd_rn(*slot_p, s);
                __syncwarp();
              }
            }
            __syncwarp();
          };
          round(0, p0);
          round(1, p1);
          round(2, p2);
          round(3, p3);
          // the touched docs against the threshold
          auto recheck = [&](int rho, uint2 p) {
            if (tot <= 32u * rho) return;
            const uint32_t v = (32u * rho + lane < tot) ? __float_as_uint(M[p.x - tile_lo]) : 0u;
            hit = hit || (v >= thr_hi && v != 0u);
          };
          recheck(0, p0);
          recheck(1, p1);
          recheck(2, p2);
          recheck(3, p3);
        } else {
          sweep_scatter_long<WEIGHTS>(M, slot0 + q, t, tile_lo, seg, sw, lane);
          uint32_t m2 = 0;
#pragma unroll
          for (int v = 0; v < V; v++) {
            const float4 x = *reinterpret_cast<const float4 *>(M + v * 128 + lane * 4);
            m2 = max(m2, max(max(__float_as_uint(x.x), __float_as_uint(x.y)), max(__float_as_uint(x.z), __float_as_uint(x.w))));
          }
          hit = hit || (m2 >= thr_hi && m2 != 0u);
        }
      }
      uint32_t n_cand = 0;
      if (__any_sync(FULL, hit)) {
        // rare after warm-up: walk the tile's final scores in shared memory
        if (!tot) {
#pragma unroll
          for (int v = 0; v < V; v++) *reinterpret_cast<float4 *>(M + v * 128 + lane * 4) = R[v];
        }
        __syncwarp();
        const uint32_t thr_now = sweep_collect_merge<V>(M, cand, tile_lo, q_state[q * 8 + 4], (int32_t)q_state[q * 8 + 5], seg, sw, lane, &n_cand);
        if (lane == 0) atomicMax(q_state + q * 8 + 3, thr_now);
        __syncwarp();
      }
      if (STATS) {
        uint32_t n_touched = 0;
#pragma unroll
        for (int v = 0; v < V; v++) {
          float4 x = R[v];
          if (tot) x = *reinterpret_cast<const float4 *>(M + v * 128 + lane * 4);
          n_touched += (x.x != 0.0f) + (x.y != 0.0f) + (x.z != 0.0f) + (x.w != 0.0f);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) n_touched += __shfl_xor_sync(FULL, n_touched, o);
        if (lane == 0) {
          const uint32_t qi = q_state[q * 8 + 4];
          if (n_touched) atomicAdd(sw.stats + (uint64_t)qi * 4 + 0, (unsigned long long)n_touched);
          if (rr[17]) atomicAdd(sw.stats + (uint64_t)qi * 4 + 1, (unsigned long long)rr[17]);
          if (n_cand) atomicAdd(sw.stats + (uint64_t)qi * 4 + 3, (unsigned long long)n_cand);
        }
        __syncwarp();
      }
    }
    cp_async_wait<0>();
  }
}

// ------------------------------------------------------------------------------------------------
// records[slot][tile]: every query slot resolved against every tile (layout above).  One thread per
// (slot, group of kSweepTileGroup consecutive tiles): the static description is read once and each
// term's range row is walked over consecutive entries.
template <bool PRUNE>
__global__ void __launch_bounds__(128) slg_sweep_records_kernel(SweepDev sw, uint32_t tile_v, bool stats) {
  const uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= sw.n_slots) return;
  const uint32_t tile0 = blockIdx.y * kSweepTileGroup;
  const uint4 *st = sw.sstat + (size_t)s * (kSweepSlotWords / 4);
  const uint4 r0 = __ldg(st), r1 = __ldg(st + 1), b0 = __ldg(st + 2), b1 = __ldg(st + 3), m0 = __ldg(st + 4);
  const uint4 m1v = __ldg(st + 5);
  const uint32_t m1 = m1v.x;
  const uint32_t rows[8] = {r0.x, r0.y, r0.z, r0.w, r1.x, r1.y, r1.z, r1.w};
  const uint32_t bases[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
  const uint32_t ns = m1 & 15u, ncol = (m1 >> 4) & 15u;
  const uint32_t rng_stride = sw.n_tiles + 1;
  float ub[8];
#pragma unroll
  for (int t = 0; t < 8; t++) ub[t] = (PRUNE && (uint32_t)t < ns + ncol) ? __ldg(sw.ubw + (size_t)s * 8 + t) : 0.0f;
  uint32_t prev[8];
#pragma unroll
  for (int t = 0; t < 8; t++) {
    prev[t] = 0;
    if ((uint32_t)t < ns || (stats && (uint32_t)t < ns + ncol)) prev[t] = __ldg(sw.rng + (uint64_t)rows[t] * rng_stride + tile0);
  }
#pragma unroll 1
  for (uint32_t tile = tile0; tile < tile0 + kSweepTileGroup && tile < sw.n_tiles; tile++) {
    uint32_t rec[kSweepRecWords];
#pragma unroll
    for (uint32_t i = 0; i < kSweepRecWords; i++) rec[i] = 0u;
    uint32_t E = 0, nne = 0, n_post = 0, posmap = 0;
    uint32_t B[4] = {0u, 0u, 0u, 0u};
    float bound = 0.0f;
#pragma unroll
    for (int t = 0; t < 8; t++) {
      uint32_t next = 0;
      if ((uint32_t)t < ns || (stats && (uint32_t)t < ns + ncol)) next = __ldg(sw.rng + (uint64_t)rows[t] * rng_stride + tile + 1);
      const uint32_t lo = prev[t], span = next - lo;
      prev[t] = next;
      n_post += span;
      const uint32_t cnt = (uint32_t)t < ns ? span : 0u;
      if (cnt) {
        const uint32_t adj = bases[t] + lo - E;
        // rec[8 + nne] = adj without dynamic register indexing
#pragma unroll
        for (int r = 0; r < 8; r++)
          if (nne == (uint32_t)r) rec[8 + r] = adj;
        if (nne && E < 128u) {
#pragma unroll
          for (int wd = 0; wd < 4; wd++)
            if ((E >> 5) == (uint32_t)wd) B[wd] |= 1u << (E & 31u);
        }
        posmap |= (uint32_t)t << (3 * nne);
        if (PRUNE) bound += ub[t];
        nne++;
        E += cnt;
      }
      if (PRUNE && (uint32_t)t >= ns && (uint32_t)t < ns + ncol) {
        const float *tm = sw.col_tmax + (uint64_t)bases[t] * sw.tmax_stride + (uint64_t)tile * (tile_v / 4);
        float tmax = 0.0f;
        for (uint32_t j = 0; j < tile_v / 4; j++) tmax = fmaxf(tmax, __ldg(tm + j));
        bound += __fmul_rn(tmax, ub[t]);
      }
    }
    rec[0] = B[0];
    rec[1] = B[1];
    rec[2] = B[2];
    rec[3] = B[3];
    rec[4] = min(E, 0xFFFFu) | (ncol << 16) | (nne << 20) | (ns << 24);
    rec[5] = __float_as_uint(bound);
    rec[16] = posmap;
    rec[17] = n_post;
    uint4 *out = reinterpret_cast<uint4 *>(sw.records + ((size_t)s * sw.n_tiles + tile) * kSweepRecWords);
#pragma unroll
    for (uint32_t i = 0; i < kSweepRecWords / 4; i++) out[i] = make_uint4(rec[4 * i], rec[4 * i + 1], rec[4 * i + 2], rec[4 * i + 3]);
  }
}

// ------------------------------------------------------------------------------------------------
// Per-batch tables of the sweep.

// rng[r][j] = index of the first posting of row r's term with doc >= j * tile_docs, j = 0..n_tiles
// (replaces the cursor movement of TermState::advance_to, query/wand.rs:205-232).  Short lists are
// walked once (posting i fills the boundaries between its predecessor's tile and its own); long
// lists take one binary search per boundary.  grid = (rows, chunks).
__global__ void __launch_bounds__(256) slg_sweep_plan_kernel(SegmentDev seg, const uint32_t *ut_term, const uint32_t *row_u,
                                                              uint32_t n_rows, uint32_t tile_docs, uint32_t n_tiles,
                                                              bool column_rows, uint32_t *rng) {
  const uint32_t r = blockIdx.x;
  if (r >= n_rows) return;
  const uint32_t term = ut_term[row_u ? row_u[r] : r];
  uint32_t *out = rng + (uint64_t)r * (n_tiles + 1);
  const uint32_t step = gridDim.y * blockDim.x;
  const uint32_t first = blockIdx.y * blockDim.x + threadIdx.x;
  if (term >= seg.n_terms) {  // a key this segment does not hold: empty list
    for (uint32_t j = first; j <= n_tiles; j += step) out[j] = 0u;
    return;
  }
  if (!column_rows && seg.term_col && seg.term_col[term] >= 0) return;
  const uint32_t df = seg.term_df[term];
  const uint32_t *d = seg.post_doc + seg.term_start[term];
  if ((uint64_t)df <= 8ull * (n_tiles + 1)) {
    for (uint32_t i = first; i <= df; i += step) {
      const uint32_t a = i ? d[i - 1] / tile_docs + 1 : 0u;
      const uint32_t b = i < df ? d[i] / tile_docs : n_tiles;
      for (uint32_t j = a; j <= b; j++) out[j] = i;
    }
  } else {
    for (uint32_t j = first; j <= n_tiles; j += step) {
      const uint64_t target = (uint64_t)j * tile_docs;
      uint32_t lo = 0, hi = df;
      while (lo < hi) {
        const uint32_t mid = (lo + hi) >> 1;
        if (d[mid] < target) lo = mid + 1;
        else hi = mid;
      }
      out[j] = lo;
    }
  }
}

// Static slot descriptions (layout above): sparse terms first, then column terms, both in query
// order.  u_row[u] = row of the range table; chunk_cols = the columns staged for each chunk of
// kSweepChunk slots.  Runs once per segment per batch.
__global__ void slg_build_sweep_kernel(SegmentDev seg, BatchDev bt, uint32_t n_slots, const uint32_t *u_row, const uint32_t *chunk_cols,
                                       uint4 *sstat, float *weights, float *ubw) {
  const uint32_t slot = blockIdx.x * blockDim.x + threadIdx.x;
  if (slot >= n_slots) return;
  const uint32_t qi = bt.q_order[slot];
  const uint32_t t0 = bt.q_term_off[qi], nt = bt.q_term_off[qi + 1] - t0;
  uint32_t *out = reinterpret_cast<uint32_t *>(sstat + (uint64_t)slot * (kSweepSlotWords / 4));
  for (uint32_t i = 0; i < kSweepSlotWords; i++) out[i] = 0u;
  for (uint32_t i = 0; i < 8; i++) {
    weights[(uint64_t)slot * 8 + i] = 1.0f;
    ubw[(uint64_t)slot * 8 + i] = 0.0f;
  }
  uint32_t p = 0, ns = 0, ncol = 0, anyw = 0;
  for (int pass = 0; pass < 2; pass++) {
    for (uint32_t t = 0; t < nt && t < kWarpMaxTerms; t++) {
      const uint32_t u = bt.qt_uterm[t0 + t];
      const uint32_t term = bt.ut_term[u];
      if (term >= seg.n_terms) continue;  // seg.postings(key) == None
      if (seg.term_df[term] == 0) continue;
      const int32_t col = seg.term_col ? seg.term_col[term] : -1;
      if ((col >= 0) != (pass == 1)) continue;
      const float w = bt.qt_weight[t0 + t];
      float ub = w;
      out[p] = u_row[u];
      if (col >= 0) {
        out[8 + p] = (uint32_t)col;
        uint32_t code = 255u;
        for (uint32_t i = 0; i < kSweepStage; i++)
          if (chunk_cols[(uint64_t)(slot / kSweepChunk) * kSweepStage + i] == (uint32_t)col) code = i;
        out[16 + (ncol >> 2)] |= code << (8 * (ncol & 3));
        ncol++;
      } else {
        out[8 + p] = (uint32_t)seg.term_start[term];
        const float mtf = seg.term_max_tf[term];
        ub = mtf > 0.0f ? __fmul_rn(bm25_contrib(mtf, seg.term_idf[term], seg.k1p1, seg_min_nk(seg, term), 1.0f), w) : 0.0f;
        ns++;
      }
      if (w != 1.0f) anyw = 1;
      weights[(uint64_t)slot * 8 + p] = w;
      ubw[(uint64_t)slot * 8 + p] = ub;
      p++;
    }
  }
  out[18] = qi;
  out[19] = (uint32_t)bt.q_filter[qi];
  out[20] = ns | (ncol << 4) | (anyw << 8);
}

// ------------------------------------------------------------------------------------------------
// residency: unit-weight contribution of every posting (one CTA of 128 threads per 128-posting
// block, like slg_transcode_csr_kernel), the dense columns and their per-512-doc maxima
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
  post_score[base + i] = bm25_contrib_fast(tf, seg.term_idf[term], seg.k1p1, seg_nk(seg, term)[doc], 1.0f);
}

// (doc, score bits) of every posting side by side: one 8-byte copy per posting in the sweep
__global__ void slg_pair_postings_kernel(const uint32_t *post_doc, const float *post_score, uint64_t n, uint2 *post_pair) {
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x)
    post_pair[i] = make_uint2(post_doc[i], __float_as_uint(post_score[i]));
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

// one warp per (512-doc slice, column): the exact maximum contribution inside the slice
__global__ void slg_column_tmax_kernel(const float *cols, uint64_t col_stride, uint32_t n_cols, uint32_t tmax_stride, float *tmax) {
  const uint64_t w = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (w >= (uint64_t)n_cols * tmax_stride) return;
  const uint32_t c = (uint32_t)(w / tmax_stride), j = (uint32_t)(w % tmax_stride);
  const float4 *p = reinterpret_cast<const float4 *>(cols + (uint64_t)c * col_stride + (uint64_t)j * 512) + lane;
  float m = 0.0f;
#pragma unroll
  for (int i = 0; i < 4; i++) {
    const float4 v = p[i * 32];
    m = fmaxf(m, fmaxf(fmaxf(v.x, v.y), fmaxf(v.z, v.w)));
  }
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xFFFFFFFFu, m, o));
  if (lane == 0) tmax[w] = m;
}

}  // namespace slg
