// slg_sweep_kernel.cuh — K2/K3, tile-sweep variant with register accumulators: the default for plain
// OR queries (no matcher), k <= 32 and <= 8 terms per query — the shape of BASELINE.json configs[1].
//
// Why this shape.  A batch names the same head terms over and over (at C2 the sum of df over term
// INSTANCES is ~17x the sum over UNIQUE terms, and 83 % of all posting visits belong to the ~90 terms
// with df >= N/8).  Scatter kernels (slg_score_warp_kernel) spend their time on instruction issue:
// every visit is a shared-memory read-modify-write and every (query, tile) ends with a scan of the
// whole accumulator (profiles/r1_v3_warp_kernel_summary.txt: 68 warp instructions per 32 visits).
//
//   resident scores   seg.post_score[i] = unit-weight BM25 contribution of posting i, computed once at
//                     segment load with score_tf's arithmetic (query/wand.rs:269-286).
//   dense columns     a term with df >= doc_count / dense_den also has a doc-indexed f32 column
//                     (score or +0.0f).  Adding a column to a tile is a 128-bit load + 4 FADD per 4
//                     docs; x + 0.0f == x, so docs without the term are unaffected bit for bit.
//   tile ownership    one CTA owns a tile of TILE = 128*V docs at a time and sweeps ALL heavy queries
//                     of the batch over it.  The tile's slices of the batch's most-used columns are
//                     staged in shared memory once per tile and every query that names one of them
//                     reads it from there (L2 -> SM traffic drops from 8 B per visit to one column
//                     slice per tile per CTA).
//   register tile     one warp = one (query, tile): lane L holds docs {128*i + 4*L .. +3 : i < V} in
//                     registers.  The query's terms without a column are scattered into a
//                     warp-private shared tile first (posting-driven, a few postings per tile), the
//                     tile is read once into the registers (and cleared), the columns are added in
//                     registers, and the registers are compared against the query's running k-th
//                     score — no accumulator scan, no shared-memory traffic for the dense part.
//   software pipeline every per-item input (query record, tile ranges of the sparse terms, the first
//                     32 postings) is loaded one to three items ahead by the same warp, which walks a
//                     static, rotated sequence of query slots; nothing in an item waits on L2.
//   seed pass         the same kernel first runs tiles [0, seed_tiles) with the query slots split
//                     across CTAs (no two warps share a query), which gives every query a useful
//                     threshold before 148 CTAs start merging into the same top-k lists.
//
// Queries for which a sweep over every tile would be wasted work (no column term and few postings)
// are "light": they go to slg_score_warp_kernel in the same batch run.
//
// Summation order (the float contract of this kernel): the query's terms WITHOUT a column in query
// order, then the terms WITH a column in query order, one left fold.  That is brute_force
// (query/wand.rs:527-548) applied to a permutation of the query's terms; tests check it bit for bit
// against the oracle run on the permuted query and against the reference order under the 1e-5 rule.
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
constexpr uint32_t kSweepRec = 9;   // uint4 per query slot: 8 term records + head
constexpr uint32_t kSweepMaxSlots = 8192;  // query slots per launch (threshold cache in shared memory)

// term record (uint4): x | y << 32 = base, z = row of the tile-range table, w = code
//   base  sparse: first padded posting index (post_doc / post_score); column: element offset in seg.cols
//   code  bits 0..1 kind (0 none, 1 sparse, 2 column), bits 2..9 1 + shared-memory slot of the
//         column's staged slice (0 = read the column from global), bits 10..30 column index,
//         bit 31 weight != 1
// head (uint4): x = query index, y = filter id, z = ns | nt << 8 | any_weight << 16
constexpr uint32_t kSweepWBit = 0x80000000u;

struct SweepDev {
  const uint4 *recs;            // [n_slots][kSweepRec]
  const float *weights;         // [n_slots][8]
  const float *ubw;             // [n_slots][8] PRUNE: weight * term-wide bound (sparse) or weight (column)
  const uint32_t *slot_qi;      // [n_slots]
  const uint32_t *rng;          // [rows][n_tiles + 1] first posting with doc >= tile * TILE
  const float *col_tmax;        // [n_cols][tmax_stride] column maxima per 512 docs (PRUNE)
  const uint64_t *hot_cols;     // [n_hot] element offset of each staged column in seg.cols
  const uint32_t *const *filter_bits;
  uint32_t n_slots, k, n_tiles, n_hot;
  uint32_t tile_begin, tile_end;  // tiles of this launch
  uint32_t seed;                  // 1: every CTA walks all tiles of the launch over its own share of the slots
  uint32_t tmax_stride;
  unsigned long long *thr_key;
  uint32_t *topk_count, *lock;
  unsigned long long *topk_keys;
  uint32_t *work_counter;
  unsigned long long *stats;
};

template <int V>
__host__ __device__ constexpr size_t sweep_smem_per_warp() {
  return (size_t)128 * V * 4 + kWarpCand * 8;  // M f32[128*V] | cand u64[64]
}

__device__ __forceinline__ uint64_t shfl_u64(uint64_t v, int src) {
  const uint32_t lo = __shfl_sync(0xFFFFFFFFu, (uint32_t)v, src), hi = __shfl_sync(0xFFFFFFFFu, (uint32_t)(v >> 32), src);
  return ((uint64_t)hi << 32) | lo;
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

// Slow path of one (query, tile): the register tile has been parked in M.  Collect the keys that
// beat the query's current k-th key, clear M, merge into the query's global top-k (push_top_k,
// query/wand.rs:905-916).  Returns the score bits of the best threshold now known.
template <int V>
__device__ __noinline__ uint2 sweep_collect(float *M, unsigned long long *cand, uint32_t tile_lo, uint32_t qi, int32_t filter,
                                            const SegmentDev &seg, const SweepDev &sw, int lane) {
  constexpr uint32_t TILE = 128u * V;
  const uint32_t k = sw.k;
  const uint32_t lt_mask = (1u << lane) - 1u;
  unsigned long long thr = ld_cg_u64(sw.thr_key + qi);
  uint32_t cnt = 0;
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
        if (pass && filter >= 0) pass = (sw.filter_bits[filter][doc >> 5] >> (doc & 31)) & 1u;
        sweep_push(pass, key, cand, cnt, thr, k, lane, lt_mask);
      }
    }
  }
  __syncwarp();
  const uint32_t n_cand = cnt;
  if (cnt > 0) {
    const unsigned long long thr_now = ld_cg_u64(sw.thr_key + qi);
    thr = max(thr, thr_now);
    const bool useful = lane < (int)cnt && cand[lane] > thr_now;  // cnt <= 32 after every append
    if (__any_sync(0xFFFFFFFFu, useful)) {
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
      if (total == k) thr = max(thr, cand[k - 1]);
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
  }
  return make_uint2(thr == kThrInit ? 0u : (uint32_t)(thr >> 32), n_cand);
}

template <int V, bool PRUNE, bool STATS>
__global__ void __launch_bounds__(kSweepThreads, 1) slg_score_sweep_kernel(const SegmentDev seg, const SweepDev sw) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  constexpr uint32_t TILE = 128u * V;
  constexpr uint32_t FULL = 0xFFFFFFFFu;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  // layout: hot f32[n_hot][TILE] | thr_s u32[n_slots rounded to 4] | per warp: M f32[TILE], cand u64[64]
  float *hot = reinterpret_cast<float *>(smem_raw);
  uint32_t *thr_s = reinterpret_cast<uint32_t *>(smem_raw + (size_t)sw.n_hot * TILE * 4);
  unsigned char *mine = reinterpret_cast<unsigned char *>(thr_s + ((sw.n_slots + 3u) & ~3u)) + (size_t)warp * sweep_smem_per_warp<V>();
  float *M = reinterpret_cast<float *>(mine);
  unsigned long long *cand = reinterpret_cast<unsigned long long *>(mine + (size_t)TILE * 4);
  __shared__ uint32_t s_tile;

  const uint32_t lt_mask = (1u << lane) - 1u;
  const uint32_t rng_stride = sw.n_tiles + 1;
  (void)lt_mask;

  for (uint32_t i = lane * 4; i < TILE; i += 128) *reinterpret_cast<float4 *>(M + i) = make_float4(0, 0, 0, 0);

  // query slots of this CTA: all of them, or (seed pass) its own share
  uint32_t slot_lo = 0, slot_n = sw.n_slots;
  if (sw.seed) {
    const uint32_t per = (sw.n_slots + gridDim.x - 1) / gridDim.x;
    slot_lo = min(blockIdx.x * per, sw.n_slots);
    slot_n = min(per, sw.n_slots - slot_lo);
  }
  if (slot_n == 0) return;
  const int n_iter = slot_n > (uint32_t)warp ? (int)((slot_n - warp + kSweepWarps - 1) / kSweepWarps) : 0;

  for (uint32_t it = 0;; it++) {
    __syncthreads();  // every warp is done with the previous tile's staged slices and thresholds
    uint32_t tile;
    if (sw.seed) {
      tile = sw.tile_begin + it;
    } else {
      if (threadIdx.x == 0) s_tile = sw.tile_begin + atomicAdd(sw.work_counter, 1u);
      __syncthreads();
      tile = s_tile;
    }
    if (tile >= sw.tile_end) break;
    const uint32_t tile_lo = tile * TILE;
    // ---- stage this tile's slices of the batch's hottest columns, refresh the threshold cache ----
    for (uint32_t i = threadIdx.x; i < sw.n_hot * (TILE / 4); i += kSweepThreads) {
      const uint32_t h = i / (TILE / 4), o = i - h * (TILE / 4);
      reinterpret_cast<float4 *>(hot)[i] = ldg_stream_f4(reinterpret_cast<const float4 *>(seg.cols + sw.hot_cols[h] + tile_lo) + o);
    }
    for (uint32_t s = threadIdx.x; s < slot_n; s += kSweepThreads) {
      const unsigned long long t = ld_cg_u64(sw.thr_key + sw.slot_qi[slot_lo + s]);
      thr_s[slot_lo + s] = t == kThrInit ? 0u : (uint32_t)(t >> 32);
    }
    __syncthreads();

    // this warp walks slots rot + warp, rot + warp + 16, ... (mod slot_n): CTAs on different tiles
    // are on different queries at any moment, so their top-k merges do not pile up on one lock
    const uint32_t rot = sw.seed ? 0u : (uint32_t)(((uint64_t)tile * 2654435761ull >> 9) % slot_n);
    auto slot_at = [&](int p) -> uint32_t {
      uint32_t s = rot + (uint32_t)warp + (uint32_t)kSweepWarps * (uint32_t)p;
      if (s >= slot_n) s -= slot_n;
      return slot_lo + s;
    };

    // ---- software pipeline over this warp's items: D (record) -> R (ranges) -> P (postings) -> X ----
    uint4 recD = make_uint4(0, 0, 0, 0), recR = recD, recP = recD, recX = recD;
    uint32_t lo2 = 0, hi2 = 0, lo1 = 0, hi1 = 0;
    float w2 = 1.0f, w1 = 1.0f, w0 = 1.0f;
    float ub2 = 0.0f, ub1 = 0.0f, tmax2 = 0.0f, tmax1 = 0.0f;
    uint32_t tot0 = 0, E0 = 0, own0 = 0, pd0 = 0, npost0 = 0;
    uint64_t st0 = 0;
    float ps0 = 0.0f, bound0 = 0.0f;
    (void)ub2; (void)ub1; (void)tmax2; (void)tmax1; (void)npost0; (void)bound0;

#pragma unroll 1
    for (int i = -3; i < n_iter; i++) {
      // ---- P: item i + 1 — counts, owners, first 32 postings ----
      uint32_t totN = 0, EN = 0, ownN = 0, pdN = 0, npostN = 0;
      uint64_t stN = 0;
      float psN = 0.0f, boundN = 0.0f;
      if (i + 1 >= 0 && i + 1 < n_iter) {
        const uint32_t kind = lane < 8 ? (recP.w & 3u) : 0u;
        const uint32_t cnt = kind == 1u ? hi1 - lo1 : 0u;
        uint32_t inc = cnt;
#pragma unroll
        for (int o = 1; o < 8; o <<= 1) {
          const uint32_t v = __shfl_up_sync(FULL, inc, o);
          if (lane >= o) inc += v;
        }
        totN = __shfl_sync(FULL, inc, 7);
        EN = inc - cnt;
        stN = (((uint64_t)recP.y << 32) | recP.x) + lo1 - EN;
        if (totN) {
          // owner of position j: the last term slot s in 0..7 whose exclusive prefix E_s <= j
          uint32_t o = 0;
          uint32_t e = __shfl_sync(FULL, EN, 4);
          if ((uint32_t)lane >= e) o = 4;
          e = __shfl_sync(FULL, EN, o + 2);
          if ((uint32_t)lane >= e) o += 2;
          e = __shfl_sync(FULL, EN, o + 1);
          if ((uint32_t)lane >= e) o += 1;
          ownN = o;
          const uint64_t idx = shfl_u64(stN, o) + lane;
          if ((uint32_t)lane < totN) {
            pdN = __ldg(seg.post_doc + idx);
            psN = __ldg(seg.post_score + idx);
          }
        }
        if (PRUNE) {
          float b = kind == 1u ? (cnt ? ub1 : 0.0f) : (kind == 2u ? __fmul_rn(tmax1, ub1) : 0.0f);
#pragma unroll
          for (int o = 4; o > 0; o >>= 1) b += __shfl_xor_sync(FULL, b, o);  // terms live in lanes 0..7
          boundN = __shfl_sync(FULL, b, 0);
        }
        if (STATS) npostN = kind ? hi1 - lo1 : 0u;
      }
      // ---- R: item i + 2 — tile ranges of the sparse terms, weights, bounds ----
      if (i + 2 >= 0 && i + 2 < n_iter) {
        lo2 = hi2 = 0;
        w2 = 1.0f;
        if (lane < 8) {
          const uint32_t kind = recR.w & 3u;
          if (kind == 1u || (STATS && kind == 2u)) {
            const uint32_t *p = sw.rng + (uint64_t)recR.z * rng_stride + tile;
            lo2 = __ldg(p);
            hi2 = __ldg(p + 1);
          }
          if ((recR.w & kSweepWBit) || PRUNE) {
            const uint32_t s = slot_at(i + 2);
            if (recR.w & kSweepWBit) w2 = __ldg(sw.weights + (uint64_t)s * 8 + lane);
            if (PRUNE) {
              ub2 = kind ? __ldg(sw.ubw + (uint64_t)s * 8 + lane) : 0.0f;
              tmax2 = 0.0f;
              if (kind == 2u) {
                const float *tm = sw.col_tmax + (uint64_t)((recR.w >> 10) & 0x1FFFFFu) * sw.tmax_stride + (uint64_t)tile * (V / 4);
#pragma unroll
                for (int j = 0; j < V / 4; j++) tmax2 = fmaxf(tmax2, __ldg(tm + j));
              }
            }
          }
        }
      }
      // ---- D: item i + 3 — query record ----
      if (i + 3 < n_iter) {
        recD = make_uint4(0, 0, 0, 0);
        if (lane < (int)kSweepRec) recD = __ldg(sw.recs + (uint64_t)slot_at(i + 3) * kSweepRec + lane);
      }

      // ---- X: item i ----
      if (i >= 0) {
        const uint32_t slot = slot_at(i);
        const uint32_t hd = __shfl_sync(FULL, recX.z, 8);
        const uint32_t ns = hd & 255u, nt = (hd >> 8) & 255u;
        const bool anyw = (hd >> 16) != 0u;
        const uint32_t thr_hi = thr_s[slot];
        bool skip = false;
        if (PRUNE) skip = thr_hi != 0u && bound0 * 1.00001f < __uint_as_float(thr_hi);
        if (!skip && (tot0 != 0u || nt > ns)) {
          float4 R[V];
          if (tot0) {
            // ---- sparse terms: scatter (doc, score) into the warp's shared tile, query order ----
            uint32_t own = own0, doc = pd0;
            float s = ps0;
            uint32_t j0 = 0;
#pragma unroll 1
            for (;;) {
              const bool valid = j0 + lane < tot0;
              if (anyw) s = __fmul_rn(s, __shfl_sync(FULL, w0, own));
              const uint32_t local = doc - tile_lo;
              const uint32_t o_first = __shfl_sync(FULL, own, 0);
              if (!__any_sync(FULL, valid && own != o_first)) {
                // one term in this round: its docs are distinct
                if (valid) M[local] = __fadd_rn(M[local], s);
              } else {
                // several terms: lanes that hit the same doc add in lane (= term) order
                const uint32_t grp = __match_any_sync(FULL, valid ? local : (0x80000000u | (uint32_t)lane));
                const uint32_t rank = __popc(grp & lt_mask);
                const uint32_t maxr = __reduce_max_sync(FULL, rank);
                for (uint32_t r = 0; r <= maxr; r++) {
                  if (valid && rank == r) M[local] = __fadd_rn(M[local], s);
                  __syncwarp();
                }
              }
              j0 += 32;
              if (j0 >= tot0) break;
              __syncwarp();
              const uint32_t j = j0 + lane;
              uint32_t o = 0;
              uint32_t e = __shfl_sync(FULL, E0, 4);
              if (j >= e) o = 4;
              e = __shfl_sync(FULL, E0, o + 2);
              if (j >= e) o += 2;
              e = __shfl_sync(FULL, E0, o + 1);
              if (j >= e) o += 1;
              own = o;
              const uint64_t idx = shfl_u64(st0, o) + j;
              doc = tile_lo;
              s = 0.0f;
              if (j < tot0) {
                doc = __ldg(seg.post_doc + idx);
                s = __ldg(seg.post_score + idx);
              }
            }
            __syncwarp();
#pragma unroll
            for (int v = 0; v < V; v++) {
              R[v] = *reinterpret_cast<const float4 *>(M + v * 128 + lane * 4);
              *reinterpret_cast<float4 *>(M + v * 128 + lane * 4) = make_float4(0, 0, 0, 0);
            }
          } else {
#pragma unroll
            for (int v = 0; v < V; v++) R[v] = make_float4(0, 0, 0, 0);
          }

          // ---- column terms in query order: staged slice (shared) or global, adds in registers ----
#pragma unroll 1
          for (uint32_t dt = ns; dt < nt; dt++) {
            const uint32_t code = __shfl_sync(FULL, recX.w, dt);
            const uint32_t hs = (code >> 2) & 255u;
            float4 c[V];
            if (hs) {
              const float4 *cp = reinterpret_cast<const float4 *>(hot + (size_t)(hs - 1) * TILE) + lane;
#pragma unroll
              for (int v = 0; v < V; v++) c[v] = cp[v * 32];
            } else {
              const uint64_t coff = ((uint64_t)__shfl_sync(FULL, recX.y, dt) << 32) | __shfl_sync(FULL, recX.x, dt);
              const float4 *cp = reinterpret_cast<const float4 *>(seg.cols + coff + tile_lo) + lane;
#pragma unroll
              for (int v = 0; v < V; v++) c[v] = __ldg(cp + v * 32);
            }
            if (code & kSweepWBit) {
              const float w = __shfl_sync(FULL, w0, dt);
#pragma unroll
              for (int v = 0; v < V; v++) {
                c[v].x = __fmul_rn(c[v].x, w);
                c[v].y = __fmul_rn(c[v].y, w);
                c[v].z = __fmul_rn(c[v].z, w);
                c[v].w = __fmul_rn(c[v].w, w);
              }
            }
#pragma unroll
            for (int v = 0; v < V; v++) {
              R[v].x = __fadd_rn(R[v].x, c[v].x);
              R[v].y = __fadd_rn(R[v].y, c[v].y);
              R[v].z = __fadd_rn(R[v].z, c[v].z);
              R[v].w = __fadd_rn(R[v].w, c[v].w);
            }
          }

          // ---- compare the registers with the running k-th score ----
          uint32_t mx = 0, n_touched = 0;
#pragma unroll
          for (int v = 0; v < V; v++) {
            const uint32_t b0 = __float_as_uint(R[v].x), b1 = __float_as_uint(R[v].y), b2 = __float_as_uint(R[v].z), b3 = __float_as_uint(R[v].w);
            mx = max(mx, max(max(b0, b1), max(b2, b3)));
            if (STATS) n_touched += (b0 != 0u) + (b1 != 0u) + (b2 != 0u) + (b3 != 0u);
          }
          uint32_t n_cand = 0;
          if (__any_sync(FULL, mx >= thr_hi && mx != 0u)) {
            // rare after warm-up: park the registers in the (zero) shared tile and walk it
#pragma unroll
            for (int v = 0; v < V; v++) *reinterpret_cast<float4 *>(M + v * 128 + lane * 4) = R[v];
            __syncwarp();
            const uint32_t qi = __shfl_sync(FULL, recX.x, 8);
            const int32_t filter = (int32_t)__shfl_sync(FULL, recX.y, 8);
            const uint2 t = sweep_collect<V>(M, cand, tile_lo, qi, filter, seg, sw, lane);
            n_cand = t.y;
            if (lane == 0 && t.x > thr_hi) thr_s[slot] = t.x;
            __syncwarp();
          }
          if (STATS) {
            uint32_t n_post = npost0;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
              n_touched += __shfl_xor_sync(FULL, n_touched, o);
              n_post += __shfl_xor_sync(FULL, n_post, o);
            }
            const uint32_t qi = __shfl_sync(FULL, recX.x, 8);
            if (lane == 0) {
              if (n_touched) atomicAdd(sw.stats + (uint64_t)qi * 4 + 0, (unsigned long long)n_touched);
              if (n_post) atomicAdd(sw.stats + (uint64_t)qi * 4 + 1, (unsigned long long)n_post);
              if (n_cand) atomicAdd(sw.stats + (uint64_t)qi * 4 + 3, (unsigned long long)n_cand);
            }
          }
        } else if (STATS && skip) {
          const uint32_t qi = __shfl_sync(FULL, recX.x, 8);
          if (lane == 0) atomicAdd(sw.stats + (uint64_t)qi * 4 + 2, 1ull);
        }
      }

      // ---- rotate the pipeline ----
      recX = recP;
      recP = recR;
      recR = recD;
      lo1 = lo2;
      hi1 = hi2;
      w0 = w1;
      w1 = w2;
      if (PRUNE) {
        ub1 = ub2;
        tmax1 = tmax2;
        bound0 = boundN;
      }
      tot0 = totN;
      E0 = EN;
      own0 = ownN;
      pd0 = pdN;
      ps0 = psN;
      st0 = stN;
      if (STATS) npost0 = npostN;
    }
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
  const uint32_t term = ut_term[row_u[r]];
  if (term >= seg.n_terms) return;  // the record builder drops terms this segment does not hold
  if (!column_rows && seg.term_col && seg.term_col[term] >= 0) return;
  const uint32_t df = seg.term_df[term];
  const uint32_t *d = seg.post_doc + seg.term_start[term];
  uint32_t *out = rng + (uint64_t)r * (n_tiles + 1);
  const uint32_t step = gridDim.y * blockDim.x;
  const uint32_t first = blockIdx.y * blockDim.x + threadIdx.x;
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

// Query records in the kernel's canonical term order: sparse (no column) first, then column terms,
// both in query order.  hot_slot[u] = 1 + shared-memory slot of unique term u's column, or 0;
// u_row[u] = row of the range table.  Runs once per segment per batch.
__global__ void slg_build_sweep_kernel(SegmentDev seg, BatchDev bt, uint32_t n_slots, const uint32_t *hot_slot, const uint32_t *u_row,
                                       uint4 *recs, float *weights, float *ubw, uint32_t *slot_qi) {
  const uint32_t slot = blockIdx.x * blockDim.x + threadIdx.x;
  if (slot >= n_slots) return;
  const uint32_t qi = bt.q_order[slot];
  const uint32_t t0 = bt.q_term_off[qi], nt = bt.q_term_off[qi + 1] - t0;
  uint32_t n_out = 0, ns = 0, anyw = 0;
  uint4 *out = recs + (uint64_t)slot * kSweepRec;
  for (int pass = 0; pass < 2; pass++) {
    for (uint32_t t = 0; t < nt && t < kWarpMaxTerms; t++) {
      const uint32_t u = bt.qt_uterm[t0 + t];
      const uint32_t term = bt.ut_term[u];
      if (term >= seg.n_terms) continue;  // seg.postings(key) == None
      const uint32_t df = seg.term_df[term];
      if (df == 0) continue;
      const int32_t col = seg.term_col ? seg.term_col[term] : -1;
      if ((col >= 0) != (pass == 1)) continue;
      const float w = bt.qt_weight[t0 + t];
      uint64_t base;
      uint32_t code;
      float ub = 0.0f;
      if (col >= 0) {
        base = (uint64_t)col * seg.col_stride;
        code = 2u | ((hot_slot ? hot_slot[u] : 0u) << 2) | ((uint32_t)col << 10);
        ub = w;
      } else {
        base = seg.term_start[term];
        code = 1u;
        const float mtf = seg.term_max_tf[term];
        if (mtf > 0.0f) ub = __fmul_rn(bm25_contrib(mtf, seg.term_idf[term], seg.k1p1, seg.min_nk, 1.0f), w);
      }
      if (w != 1.0f) {
        code |= kSweepWBit;
        anyw = 1;
      }
      out[n_out] = make_uint4((uint32_t)base, (uint32_t)(base >> 32), u_row[u], code);
      weights[(uint64_t)slot * 8 + n_out] = w;
      ubw[(uint64_t)slot * 8 + n_out] = ub;
      n_out++;
    }
    if (pass == 0) ns = n_out;
  }
  for (uint32_t t = n_out; t < kWarpMaxTerms; t++) {
    out[t] = make_uint4(0, 0, 0, 0);
    weights[(uint64_t)slot * 8 + t] = 1.0f;
    ubw[(uint64_t)slot * 8 + t] = 0.0f;
  }
  out[8] = make_uint4(qi, (uint32_t)bt.q_filter[qi], ns | (n_out << 8) | (anyw << 16), 0u);
  slot_qi[slot] = qi;
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
