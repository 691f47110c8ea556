// slg_rerank.cuh — K7: exact vector rerank of BM25 candidates (what gpu::rerank should have been;
// searchlite-core/src/gpu/rerank.rs:3-5 is an identity stub).
//
// Per (query, candidate hit): gather the candidate's vector row (VectorStore::vector, vectors/mod.rs:63-71) and, per
// vector clause, similarity = dot (Cosine, vectors pre-normalised; NaN -> 0) or -sqrt(sum (x-y)^2) (L2)
// (metric_similarity, vectors/mod.rs:107-120), times the clause's boost (api/reader.rs:2421); then
// compute_hybrid_score (api/reader.rs:226-254): per clause alpha >= 1 -> bm25, alpha <= 0 -> vec, else
// alpha*bm25 + (1-alpha)*vec, a missing vector scoring -1 (cosine) / f32::MIN (L2) (api/reader.rs:218-223); the
// clause values are summed from 0.0 in clause order and divided by the clause count.  Candidates are then re-ordered
// by (score desc under total_cmp, segment_ord asc, doc_id asc) — SortKey order, query/sort.rs:80-93.
//
// Summation order.  The reference's dot / squared distance is a sequential f32 fold over the dimensions with separate
// multiply and add (Rust neither reassociates nor contracts).  slg_rerank_scores_kernel keeps exactly that order: a warp
// takes 32 candidates, stages 128 bytes of each of the 32 rows per step in shared memory with coalesced 16-byte loads
// (8 lanes per row), and then every lane folds ITS OWN candidate's 32 (f32) or 64 (bf16) values in dimension order with
// explicit round-to-nearest multiplies and adds (dimensions that are not a multiple of 32 / 64 take a scalar walk of the
// same order).  f32 rows therefore give the reference's bits; bf16 rows (this build's
// storage option) give the reference's arithmetic on the once-rounded rows.  The dependent add chain costs nothing: the
// kernel is bound by the row gather from HBM (1.5 KB per candidate at 768-d bf16), not by the 2 flops per 2 bytes.
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "slg_kernels.cuh"

namespace slg {

constexpr uint32_t kMaxRerankCands = 2048;
constexpr uint32_t kMaxRerankClauses = 8;  // MAX_VECTOR_CLAUSES, api/reader.rs:134
constexpr uint32_t kRerankRowWords = 36;   // shared-memory words per staged row: 32 of data + 4 of padding (16-byte accesses stay conflict-free)

struct RerankSegDev {
  uint32_t segment_ord, doc_count;
  const uint32_t *offsets;
  const void *values;
  int32_t bf16;
  uint32_t dim;
  uint64_t n_rows;
};

struct RerankClausesDev {
  const float *qv[kMaxRerankClauses];  // [n_queries][dim] per clause
  float alpha[kMaxRerankClauses];
  float boost[kMaxRerankClauses];
  int32_t metric[kMaxRerankClauses];   // 0 cosine, 1 l2
  uint32_t n;
};

static __global__ void slg_f32_to_bf16_kernel(const float *in, __nv_bfloat16 *out, size_t n) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    out[i] = __float2bfloat16_rn(in[i]);
}

// load-time check of a vector store: every offset names a row of the store (or is the "no vector" mark)
static __global__ void slg_check_vector_offsets_kernel(const uint32_t *offsets, uint32_t doc_count, uint64_t n_rows, uint32_t *bad) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < doc_count; i += (size_t)gridDim.x * blockDim.x) {
    const uint32_t o = offsets[i];
    if (o != 0xFFFFFFFFu && (uint64_t)o >= n_rows) atomicAdd(bad, 1u);
  }
}

// hybrid score of every candidate of one hit block.  grid = (n_queries, splits); CTA (q, s) takes the 32-candidate groups
// g = s * 8 + warp, + 8 * splits, ...  NC = compiled clause capacity (clauses.n <= NC).
// out_score[q][c] = final hybrid score, out_vsum[q][c] = sum of the clause similarities (NaN bits 0x7FC00001 = no vector)
template <int NC, bool BF16>
static __global__ void __launch_bounds__(256) slg_rerank_scores_kernel(const RerankSegDev *segs, uint32_t n_segs, RerankClausesDev cl, uint32_t dim,
                                                                 const HitDev *cands, const uint32_t *cand_counts, uint32_t stride,
                                                                 float *out_score, float *out_vsum) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float *qv = reinterpret_cast<float *>(smem_raw);                                   // [NC][dim]
  uint32_t *tiles = reinterpret_cast<uint32_t *>(smem_raw + (size_t)NC * dim * 4);   // [8 warps][32 rows][kRerankRowWords]
  const uint32_t qi = blockIdx.x;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t n = min(cand_counts[qi], stride);
#pragma unroll
  for (int c = 0; c < NC; c++)
    if ((uint32_t)c < cl.n)
      for (uint32_t i = tid; i < dim; i += 256) qv[(size_t)c * dim + i] = cl.qv[c][(size_t)qi * dim + i];
  __syncthreads();
  uint32_t *tile = tiles + (size_t)warp * 32 * kRerankRowWords;
  const uint32_t row_bytes = dim * (BF16 ? 2u : 4u);
  const bool staged = row_bytes % 128 == 0;  // rows of whole 128-byte steps (dim a multiple of 32 f32 / 64 bf16; cudaMalloc aligns the store)
  const uint32_t n_chunks = staged ? row_bytes / 128 : 0;
  const int sub = lane >> 3, part = lane & 7;
  for (uint32_t g = blockIdx.y * 8 + warp; g * 32 < n; g += gridDim.y * 8) {
    const uint32_t c = g * 32 + lane;
    HitDev h{};
    const unsigned char *rowp = nullptr;
    if (c < n) {
      h = cands[(size_t)qi * stride + c];
      for (uint32_t s = 0; s < n_segs; s++) {
        if (segs[s].segment_ord != h.segment_ord || segs[s].dim != dim || h.doc_id >= segs[s].doc_count) continue;
        const uint32_t row = segs[s].offsets[h.doc_id];
        if (row != 0xFFFFFFFFu && (uint64_t)row < segs[s].n_rows) rowp = static_cast<const unsigned char *>(segs[s].values) + (size_t)row * row_bytes;
      }
    }
    const unsigned long long my_row = reinterpret_cast<unsigned long long>(rowp);
    // row pointers of the 8 rows this lane helps to load (rows 4 j + sub)
    unsigned long long rp[8];
#pragma unroll
    for (int j = 0; j < 8; j++) rp[j] = __shfl_sync(0xFFFFFFFFu, my_row, 4 * j + sub);
    float acc[NC];
#pragma unroll
    for (int cc = 0; cc < NC; cc++) acc[cc] = 0.0f;
    if (!staged) {
      // any other dimension (the reference accepts every dim): each lane walks its own row with scalar loads, same order
      if (rowp) {
        for (uint32_t d = 0; d < dim; d++) {
          float y;
          if (BF16) y = __uint_as_float((uint32_t)reinterpret_cast<const uint16_t *>(rowp)[d] << 16);
          else y = reinterpret_cast<const float *>(rowp)[d];
#pragma unroll
          for (int cc = 0; cc < NC; cc++) {
            if ((uint32_t)cc >= cl.n) continue;
            const float x = qv[(size_t)cc * dim + d];
            if (cl.metric[cc] == 0) {
              acc[cc] = __fadd_rn(acc[cc], __fmul_rn(x, y));
            } else {
              const float df = __fsub_rn(x, y);
              acc[cc] = __fadd_rn(acc[cc], __fmul_rn(df, df));
            }
          }
        }
      }
    }
    uint4 nxt[8];
#pragma unroll
    for (int j = 0; j < 8; j++)
      nxt[j] = (staged && rp[j]) ? ldg_nc_u4(reinterpret_cast<const uint32_t *>(rp[j] + part * 16)) : make_uint4(0, 0, 0, 0);
    for (uint32_t ch = 0; ch < n_chunks; ch++) {
      __syncwarp();
#pragma unroll
      for (int j = 0; j < 8; j++) *reinterpret_cast<uint4 *>(tile + (4 * j + sub) * kRerankRowWords + part * 4) = nxt[j];
      if (ch + 1 < n_chunks) {
#pragma unroll
        for (int j = 0; j < 8; j++)
          nxt[j] = rp[j] ? ldg_nc_u4(reinterpret_cast<const uint32_t *>(rp[j] + (size_t)(ch + 1) * 128 + part * 16)) : make_uint4(0, 0, 0, 0);
      }
      __syncwarp();
      const uint32_t *mine = tile + lane * kRerankRowWords;
      const uint32_t d0 = ch * (BF16 ? 64u : 32u);
#pragma unroll
      for (int w4 = 0; w4 < 8; w4++) {
        const uint4 raw = *reinterpret_cast<const uint4 *>(mine + w4 * 4);
        const uint32_t rw[4] = {raw.x, raw.y, raw.z, raw.w};
        if (BF16) {
#pragma unroll
          for (int cc = 0; cc < NC; cc++) {
            if ((uint32_t)cc >= cl.n) continue;
            const float4 qa = *reinterpret_cast<const float4 *>(qv + (size_t)cc * dim + d0 + w4 * 8);
            const float4 qb = *reinterpret_cast<const float4 *>(qv + (size_t)cc * dim + d0 + w4 * 8 + 4);
            const float qs[8] = {qa.x, qa.y, qa.z, qa.w, qb.x, qb.y, qb.z, qb.w};
#pragma unroll
            for (int e = 0; e < 8; e++) {
              const uint32_t word = rw[e >> 1];
              const float y = __uint_as_float((e & 1) ? (word & 0xFFFF0000u) : (word << 16));
              if (cl.metric[cc] == 0) {
                acc[cc] = __fadd_rn(acc[cc], __fmul_rn(qs[e], y));
              } else {
                const float d = __fsub_rn(qs[e], y);
                acc[cc] = __fadd_rn(acc[cc], __fmul_rn(d, d));
              }
            }
          }
        } else {
#pragma unroll
          for (int cc = 0; cc < NC; cc++) {
            if ((uint32_t)cc >= cl.n) continue;
            const float4 qa = *reinterpret_cast<const float4 *>(qv + (size_t)cc * dim + d0 + w4 * 4);
            const float qs[4] = {qa.x, qa.y, qa.z, qa.w};
#pragma unroll
            for (int e = 0; e < 4; e++) {
              const float y = __uint_as_float(rw[e]);
              if (cl.metric[cc] == 0) {
                acc[cc] = __fadd_rn(acc[cc], __fmul_rn(qs[e], y));
              } else {
                const float d = __fsub_rn(qs[e], y);
                acc[cc] = __fadd_rn(acc[cc], __fmul_rn(d, d));
              }
            }
          }
        }
      }
    }
    if (c < n) {
      // compute_hybrid_score, api/reader.rs:226-254
      const bool has = rowp != nullptr;
      float blended_sum = 0.0f, vector_sum = 0.0f;
#pragma unroll
      for (int cc = 0; cc < NC; cc++) {
        if ((uint32_t)cc >= cl.n) continue;
        float vs;
        if (has) {
          vs = cl.metric[cc] == 0 ? (isnan(acc[cc]) ? 0.0f : acc[cc]) : -__fsqrt_rn(acc[cc]);
          vs = __fmul_rn(vs, cl.boost[cc]);
          vector_sum = __fadd_rn(vector_sum, vs);
        } else {
          vs = cl.metric[cc] == 0 ? -1.0f : -3.402823466e+38f;
        }
        const float a = cl.alpha[cc];
        float blended;
        if (a >= 1.0f) blended = h.score;
        else if (a <= 0.0f) blended = vs;
        else blended = __fadd_rn(__fmul_rn(a, h.score), __fmul_rn(__fsub_rn(1.0f, a), vs));
        blended_sum = __fadd_rn(blended_sum, blended);
      }
      const float denom = (float)max(cl.n, 1u);
      out_score[(size_t)qi * stride + c] = __fdiv_rn(blended_sum, denom);
      out_vsum[(size_t)qi * stride + c] = has ? vector_sum : __uint_as_float(0x7FC00001u);
    }
  }
}

// comparator-driven bitonic sort of candidate indices in shared memory (n2 = power of two >= n; indices >= n sort last)
__device__ __forceinline__ bool rerank_index_before(const HitDev *hits, uint32_t n, uint32_t a, uint32_t b) {
  if (a >= n || b >= n) return a < n && b >= n ? true : (a >= n && b >= n ? a < b : false);
  return hit_before(hits[a], hits[b]);
}

// One CTA per query: replace the candidates' scores by their hybrid scores, drop vector-less candidates of an
// all-vector plan (api/reader.rs:2474-2476), sort in SortKey order and write hits, vector scores and the new count in place.
static __global__ void __launch_bounds__(256) slg_rerank_sort_kernel(HitDev *hits_io, uint32_t *counts_io, uint32_t stride, const float *score,
                                                               const float *vsum, int all_vector_only, float *out_vs) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const uint32_t qi = blockIdx.x;
  const int tid = threadIdx.x;
  const uint32_t n_in = min(counts_io[qi], stride);
  uint32_t n2 = 32;
  while (n2 < n_in) n2 <<= 1;
  HitDev *hits = reinterpret_cast<HitDev *>(smem_raw);                 // [stride]
  float *vss = reinterpret_cast<float *>(hits + stride);               // [stride]
  uint32_t *idx = reinterpret_cast<uint32_t *>(vss + stride);          // [n2]
  __shared__ uint32_t s_n;
  if (tid == 0) s_n = 0;
  __syncthreads();
  // compact (stable order is irrelevant: the sort follows)
  for (uint32_t base = 0; base < n_in; base += 256) {
    const uint32_t i = base + tid;
    bool keep = false;
    HitDev h{};
    float v = 0.0f;
    if (i < n_in) {
      h = hits_io[(size_t)qi * stride + i];
      h.score = score[(size_t)qi * stride + i];
      v = vsum[(size_t)qi * stride + i];
      keep = !(all_vector_only && __float_as_uint(v) == 0x7FC00001u);
    }
    if (keep) {
      const uint32_t slot = atomicAdd(&s_n, 1u);
      hits[slot] = h;
      vss[slot] = __float_as_uint(v) == 0x7FC00001u ? 0.0f : v;  // (RankedHit.vector_score is None without a vector)
    }
  }
  __syncthreads();
  const uint32_t n = s_n;
  for (uint32_t i = tid; i < n2; i += 256) idx[i] = i;
  __syncthreads();
  for (uint32_t size = 2; size <= n2; size <<= 1) {
    for (uint32_t st = size >> 1; st > 0; st >>= 1) {
      for (uint32_t t = tid; t < n2 / 2; t += 256) {
        const uint32_t lo = ((t / st) * (st << 1)) + (t % st), hi = lo + st;
        const bool up = ((lo & size) == 0);
        const uint32_t a = idx[lo], b = idx[hi];
        const bool a_first = rerank_index_before(hits, n, a, b);
        if (a_first != up) {
          idx[lo] = b;
          idx[hi] = a;
        }
      }
      __syncthreads();
    }
  }
  for (uint32_t i = tid; i < stride; i += 256) {
    HitDev h;
    float v = 0.0f;
    if (i < n) {
      h = hits[idx[i]];
      v = vss[idx[i]];
    } else {
      h.segment_ord = 0xFFFFFFFFu;
      h.doc_id = 0xFFFFFFFFu;
      h.score = 0.0f;
    }
    hits_io[(size_t)qi * stride + i] = h;
    out_vs[(size_t)qi * stride + i] = v;
  }
  if (tid == 0) counts_io[qi] = n;
}

}  // namespace slg
