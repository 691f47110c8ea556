// slg_rerank.cuh — K7: exact vector rerank of BM25 candidates (what gpu::rerank should have been;
// searchlite-core/src/gpu/rerank.rs:3-5 is an identity stub).
//
// Per query: gather each candidate's vector row (VectorStore::vector, vectors/mod.rs:63-71),
// similarity = dot (Cosine, vectors pre-normalised; NaN -> 0) or -sqrt(sum (x-y)^2) (L2)
// (metric_similarity, vectors/mod.rs:107-120), blend as compute_hybrid_score for one clause
// (api/reader.rs:226-254): alpha >= 1 -> bm25, alpha <= 0 -> vec, else alpha*bm25 + (1-alpha)*vec;
// a missing vector scores -1 (cosine) / f32::MIN (L2) (api/reader.rs:218-223).  Candidates are then
// re-ordered by (score desc, segment_ord asc, doc_id asc).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "slg_kernels.cuh"

namespace slg {

constexpr uint32_t kMaxRerankCands = 2048;

struct RerankSegDev {
  uint32_t segment_ord, doc_count;
  const uint32_t *offsets;
  const void *values;
  int32_t bf16;
  uint32_t dim;
};

static __global__ void slg_f32_to_bf16_kernel(const float *in, __nv_bfloat16 *out, size_t n) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    out[i] = __float2bfloat16_rn(in[i]);
}

static __global__ void __launch_bounds__(256) slg_rerank_kernel(const RerankSegDev *segs, uint32_t n_segs, const float *query_vecs,
                                                          uint32_t dim, const HitDev *cands, const uint32_t *cand_counts,
                                                          uint32_t stride, float alpha, int metric, HitDev *out_hits,
                                                          float *out_vs) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float *qv = reinterpret_cast<float *>(smem_raw);
  HitDev *hits = reinterpret_cast<HitDev *>(smem_raw + (size_t)dim * 4);
  float *vss = reinterpret_cast<float *>(hits + stride);
  const uint32_t qi = blockIdx.x;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t n = min(cand_counts[qi], stride);
  for (uint32_t i = tid; i < dim; i += 256) qv[i] = query_vecs[(size_t)qi * dim + i];
  __syncthreads();
  const float missing = metric == 0 ? -1.0f : -3.402823466e+38f;
  for (uint32_t c = warp; c < n; c += 8) {
    HitDev h = cands[(size_t)qi * stride + c];
    const RerankSegDev *sg = nullptr;
    for (uint32_t s = 0; s < n_segs; s++)
      if (segs[s].segment_ord == h.segment_ord) sg = &segs[s];
    uint32_t row = 0xFFFFFFFFu;
    if (sg && sg->dim == dim && h.doc_id < sg->doc_count) row = sg->offsets[h.doc_id];
    float vs = missing;
    bool has = row != 0xFFFFFFFFu;
    if (has) {
      float acc = 0.0f;
      if (sg->bf16) {
        const __nv_bfloat16 *v = static_cast<const __nv_bfloat16 *>(sg->values) + (size_t)row * dim;
        for (uint32_t i = lane * 8; i < dim; i += 256) {
          const uint4 raw = *reinterpret_cast<const uint4 *>(v + i);
          const __nv_bfloat162 *p2 = reinterpret_cast<const __nv_bfloat162 *>(&raw);
#pragma unroll
          for (int j = 0; j < 4; j++) {
            const float2 f = __bfloat1622float2(p2[j]);
            const float a0 = qv[i + 2 * j], a1 = qv[i + 2 * j + 1];
            if (metric == 0) {
              acc = fmaf(a0, f.x, acc);
              acc = fmaf(a1, f.y, acc);
            } else {
              const float d0 = a0 - f.x, d1 = a1 - f.y;
              acc = fmaf(d0, d0, acc);
              acc = fmaf(d1, d1, acc);
            }
          }
        }
      } else {
        const float *v = static_cast<const float *>(sg->values) + (size_t)row * dim;
        for (uint32_t i = lane * 4; i < dim; i += 128) {
          const float4 f = *reinterpret_cast<const float4 *>(v + i);
          const float4 a = *reinterpret_cast<const float4 *>(qv + i);
          if (metric == 0) {
            acc = fmaf(a.x, f.x, acc);
            acc = fmaf(a.y, f.y, acc);
            acc = fmaf(a.z, f.z, acc);
            acc = fmaf(a.w, f.w, acc);
          } else {
            const float d0 = a.x - f.x, d1 = a.y - f.y, d2 = a.z - f.z, d3 = a.w - f.w;
            acc = fmaf(d0, d0, acc);
            acc = fmaf(d1, d1, acc);
            acc = fmaf(d2, d2, acc);
            acc = fmaf(d3, d3, acc);
          }
        }
      }
      for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xFFFFFFFFu, acc, o);
      if (metric == 0) vs = isnan(acc) ? 0.0f : acc;
      else vs = -sqrtf(acc);
    }
    if (lane == 0) {
      float blended;
      if (alpha >= 1.0f) blended = h.score;
      else if (alpha <= 0.0f) blended = vs;
      else blended = __fadd_rn(__fmul_rn(alpha, h.score), __fmul_rn(__fsub_rn(1.0f, alpha), vs));
      h.score = blended;
      hits[c] = h;
      vss[c] = has ? vs : missing;
    }
  }
  __syncthreads();
  for (uint32_t i = tid; i < n; i += 256) {
    const HitDev h = hits[i];
    uint32_t rank = 0;
    for (uint32_t j = 0; j < n; j++) rank += (j != i) && hit_before(hits[j], h);
    out_hits[(size_t)qi * stride + rank] = h;
    out_vs[(size_t)qi * stride + rank] = vss[i];
  }
  for (uint32_t i = n + tid; i < stride; i += 256) {
    HitDev h;
    h.segment_ord = 0xFFFFFFFFu;
    h.doc_id = 0xFFFFFFFFu;
    h.score = 0.0f;
    out_hits[(size_t)qi * stride + i] = h;
    out_vs[(size_t)qi * stride + i] = 0.0f;
  }
}

}  // namespace slg
