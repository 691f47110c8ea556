// slg_rerank.cu — vector stores and the hybrid rerank entry points of libsearchlite_gpu.so (kernels: slg_rerank.cuh).
//
//   slg_load_vectors / slg_load_vectors_bf16   VectorStore residency (index/segment.rs:1030-1053), host or device source
//   slg_rerank / slg_rerank_clauses            candidates handed over by the caller (host memory)
//   slg_rerank_batch                           the pipeline form: the device-resident top-k of the batch's last run is
//                                              rescored and re-sorted per segment in place, then merged again — what
//                                              IndexReader::search does between search_segment and the final sort when a
//                                              vector plan is present (api/reader.rs:2752-2773, merge_vector_hits :2477-2537)
//   slg_batch_fetch_vector_scores              RankedHit.vector_score of the fetched hits
#include "slg_host.h"
#include "slg_kernels.cuh"
#include "slg_rerank.cuh"

using namespace slg;

namespace {

struct ClauseSet {
  RerankClausesDev dev{};
  int all_vector_only = 1;
  DevBuf qv;  // [n][Q][dim]
};

int32_t check_clauses(slg_index *ix, const slg_vector_clause_t *clauses, uint32_t n_clauses) {
  if (!clauses || !n_clauses) return fail(ix, SLG_ERR_INVALID, "no vector clauses");
  // "too many vector clauses: got {}, max supported {}", api/reader.rs:2089-2095
  if (n_clauses > kMaxRerankClauses) return fail(ix, SLG_ERR_INVALID, "too many vector clauses: got %u, max supported %u", n_clauses, kMaxRerankClauses);
  for (uint32_t c = 0; c < n_clauses; c++) {
    const slg_vector_clause_t &cl = clauses[c];
    if (!cl.query_vecs) return fail(ix, SLG_ERR_INVALID, "vector clause %u has no query vectors", c);
    if (cl.metric != SLG_METRIC_COSINE && cl.metric != SLG_METRIC_L2) return fail(ix, SLG_ERR_INVALID, "unknown metric");
    // api/reader.rs:2127-2129, :2159-2161
    if (!(cl.alpha >= 0.0f && cl.alpha <= 1.0f)) return fail(ix, SLG_ERR_INVALID, "vector alpha must be a finite value between 0 and 1 inclusive");
    if (!(cl.boost >= 0.0f) || std::isinf(cl.boost)) return fail(ix, SLG_ERR_INVALID, "vector boost must be finite and non-negative");
  }
  return SLG_OK;
}

// query vectors of every clause -> device (one allocation); fills the kernel's clause table
int32_t upload_clauses(slg_index *ix, const slg_vector_clause_t *clauses, uint32_t n_clauses, uint32_t n_queries, uint32_t dim, ClauseSet &cs) {
  cudaStream_t st = ix->stream;
  const size_t per = (size_t)n_queries * dim * 4;
  SLG_CUDA(ix, cs.qv.alloc(per * n_clauses));
  cs.dev.n = n_clauses;
  for (uint32_t c = 0; c < n_clauses; c++) {
    float *dst = reinterpret_cast<float *>(cs.qv.as<unsigned char>() + per * c);
    SLG_CUDA(ix, cudaMemcpyAsync(dst, clauses[c].query_vecs, per, cudaMemcpyDefault, st));
    cs.dev.qv[c] = dst;
    cs.dev.alpha[c] = clauses[c].alpha;
    cs.dev.boost[c] = clauses[c].boost;
    cs.dev.metric[c] = clauses[c].metric == SLG_METRIC_COSINE ? 0 : 1;
    if (clauses[c].alpha > 0.0f) cs.all_vector_only = 0;  // plan.clauses.iter().all(|c| c.alpha <= 0.0), api/reader.rs:2472
  }
  ix->ctr.last_h2d_bytes += per * n_clauses;
  return SLG_OK;
}

int32_t segment_table(slg_index *ix, uint32_t dim, DevBuf &d_segs, uint32_t *n_segs, bool *bf16) {
  std::vector<RerankSegDev> segs;
  bool any = false, all_bf16 = true, any_bf16 = false;
  for (auto &s : ix->segs) {
    RerankSegDev r{};
    r.segment_ord = s->ord;
    r.doc_count = s->doc_count;
    r.offsets = s->vec.offsets.as<uint32_t>();
    r.values = s->vec.values.p;
    r.bf16 = s->vec.bf16 ? 1 : 0;
    r.dim = s->vec.dim;
    r.n_rows = s->vec.n_rows;
    if (s->vec.dim && s->vec.dim != dim)
      return fail(ix, SLG_ERR_INVALID, "vector field expects dimension %u, got %u", s->vec.dim, dim);  // api/reader.rs:2111-2117
    if (s->vec.dim) {
      any = true;
      all_bf16 = all_bf16 && s->vec.bf16;
      any_bf16 = any_bf16 || s->vec.bf16;
    }
    segs.push_back(r);
  }
  if (any && any_bf16 && !all_bf16) return fail(ix, SLG_ERR_UNSUPPORTED, "the segments of one handle must store their vectors in one format");
  *bf16 = any && all_bf16;
  *n_segs = (uint32_t)segs.size();
  SLG_CUDA(ix, d_segs.alloc(std::max<size_t>(segs.size(), 1) * sizeof(RerankSegDev)));
  if (!segs.empty()) SLG_CUDA(ix, cudaMemcpyAsync(d_segs.p, segs.data(), segs.size() * sizeof(RerankSegDev), cudaMemcpyHostToDevice, ix->stream));
  return SLG_OK;
}

template <int NC>
cudaError_t launch_scores_nc(bool bf16, dim3 grid, size_t smem, cudaStream_t st, const RerankSegDev *segs, uint32_t n_segs, const RerankClausesDev &cl,
                             uint32_t dim, const HitDev *cands, const uint32_t *counts, uint32_t stride, float *score, float *vsum) {
  cudaError_t e;
  if (bf16) {
    e = cudaFuncSetAttribute(slg_rerank_scores_kernel<NC, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    slg_rerank_scores_kernel<NC, true><<<grid, 256, smem, st>>>(segs, n_segs, cl, dim, cands, counts, stride, score, vsum);
  } else {
    e = cudaFuncSetAttribute(slg_rerank_scores_kernel<NC, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    slg_rerank_scores_kernel<NC, false><<<grid, 256, smem, st>>>(segs, n_segs, cl, dim, cands, counts, stride, score, vsum);
  }
  return cudaGetLastError();
}

// scores + sort of one hit block (in place); `score` / `vsum` are scratch of n_queries * stride floats each
int32_t rerank_block(slg_index *ix, const DevBuf &d_segs, uint32_t n_segs, bool bf16, const ClauseSet &cs, uint32_t n_queries, uint32_t dim,
                     HitDev *hits, uint32_t *counts, uint32_t stride, float *score, float *vsum, float *out_vs) {
  cudaStream_t st = ix->stream;
  const uint32_t nc = cs.dev.n <= 1 ? 1 : cs.dev.n <= 2 ? 2 : cs.dev.n <= 4 ? 4 : 8;
  const size_t smem = (size_t)nc * dim * 4 + (size_t)8 * 32 * kRerankRowWords * 4;
  if (smem > ix->smem_optin) return fail(ix, SLG_ERR_UNSUPPORTED, "%u clauses of dimension %u do not fit shared memory", cs.dev.n, dim);
  const uint32_t groups = (stride + 255) / 256;  // 8 warps x 32 candidates per pass of a CTA
  uint32_t splits = (uint32_t)std::max(1, (ix->n_sm * 4 + (int)n_queries - 1) / (int)n_queries);
  splits = std::min(splits, std::max(groups, 1u));
  const dim3 grid(n_queries, splits);
  cudaError_t e;
  const RerankSegDev *sg = d_segs.as<RerankSegDev>();
  switch (nc) {
    case 1: e = launch_scores_nc<1>(bf16, grid, smem, st, sg, n_segs, cs.dev, dim, hits, counts, stride, score, vsum); break;
    case 2: e = launch_scores_nc<2>(bf16, grid, smem, st, sg, n_segs, cs.dev, dim, hits, counts, stride, score, vsum); break;
    case 4: e = launch_scores_nc<4>(bf16, grid, smem, st, sg, n_segs, cs.dev, dim, hits, counts, stride, score, vsum); break;
    default: e = launch_scores_nc<8>(bf16, grid, smem, st, sg, n_segs, cs.dev, dim, hits, counts, stride, score, vsum); break;
  }
  SLG_CUDA(ix, e);
  count_launch(ix);
  uint32_t n2 = 32;
  while (n2 < stride) n2 <<= 1;
  const size_t ssmem = (size_t)stride * (sizeof(HitDev) + 4) + (size_t)n2 * 4;
  if (ssmem > ix->smem_optin) return fail(ix, SLG_ERR_UNSUPPORTED, "%u candidates per query do not fit shared memory", stride);
  SLG_CUDA(ix, cudaFuncSetAttribute(slg_rerank_sort_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ssmem));
  slg_rerank_sort_kernel<<<n_queries, 256, ssmem, st>>>(hits, counts, stride, score, vsum, cs.all_vector_only, out_vs);
  SLG_CUDA(ix, cudaGetLastError());
  count_launch(ix);
  return SLG_OK;
}

int32_t check_dim(slg_index *ix, uint32_t dim, bool) {
  if (!dim) return fail(ix, SLG_ERR_INVALID, "vector dimension 0");
  return SLG_OK;
}

int32_t store_vectors(slg_index *ix, uint32_t segment_ord, uint32_t dim, const uint32_t *offsets, const void *values, bool values_bf16,
                      uint64_t n_rows, bool store_bf16) {
  if (!ix || !offsets || (!values && n_rows) || !dim) return SLG_ERR_INVALID;
  SLG_CUDA(ix, cudaSetDevice(ix->device));
  Segment *s = ix->find(segment_ord);
  if (!s) return fail(ix, SLG_ERR_INVALID, "no segment %u", segment_ord);
  if (n_rows > 0xFFFFFFFEull) return fail(ix, SLG_ERR_INVALID, "more vector rows than a u32 offset can name");
  if (n_rows && (uint64_t)dim > (std::numeric_limits<size_t>::max() / 4) / n_rows) return fail(ix, SLG_ERR_INVALID, "vector store size overflows");
  cudaStream_t st = ix->stream;
  Vectors v;
  v.dim = dim;
  v.n_rows = n_rows;
  v.bf16 = store_bf16;
  SLG_CUDA(ix, v.offsets.alloc(std::max<size_t>(s->doc_count, 1) * 4));
  SLG_CUDA(ix, cudaMemcpyAsync(v.offsets.p, offsets, (size_t)s->doc_count * 4, cudaMemcpyDefault, st));
  // VectorStore::vector returns None for an offset past the store (values.get(start..end)); a store whose offsets point
  // outside its rows is a corrupt file: refuse it at load rather than index with it
  {
    PoolScope pool_scope(st);
    DevBuf bad;
    SLG_CUDA(ix, bad.alloc(4));
    SLG_CUDA(ix, cudaMemsetAsync(bad.p, 0, 4, st));
    if (s->doc_count) {
      slg_check_vector_offsets_kernel<<<(unsigned)std::min<size_t>(((size_t)s->doc_count + 255) / 256, 4096), 256, 0, st>>>(
          v.offsets.as<uint32_t>(), s->doc_count, n_rows, bad.as<uint32_t>());
      count_launch(ix);
    }
    uint32_t n_bad = 0;
    SLG_CUDA(ix, cudaMemcpyAsync(&n_bad, bad.p, 4, cudaMemcpyDeviceToHost, st));
    SLG_CUDA(ix, cudaStreamSynchronize(st));
    if (n_bad) return fail(ix, SLG_ERR_INVALID, "vector store of segment %u: %u offsets point past its %llu rows", segment_ord, n_bad, (unsigned long long)n_rows);
  }
  const size_t n = (size_t)n_rows * dim;
  if (values_bf16 && !store_bf16) return fail(ix, SLG_ERR_INVALID, "bf16 input rows are stored as bf16");
  if (values_bf16 || !store_bf16) {
    const size_t esz = store_bf16 ? 2 : 4;
    SLG_CUDA(ix, v.values.alloc(std::max<size_t>(n, 1) * esz));
    if (n) SLG_CUDA(ix, cudaMemcpyAsync(v.values.p, values, n * esz, cudaMemcpyDefault, st));
  } else {
    // f32 in, bf16 resident: converted in slices so that the f32 staging never exceeds 1 GiB
    SLG_CUDA(ix, v.values.alloc(std::max<size_t>(n, 1) * 2));
    const size_t slice = (size_t)1 << 28;  // elements
    DevBuf tmp;
    SLG_CUDA(ix, tmp.alloc(std::min(std::max<size_t>(n, 1), slice) * 4));
    for (size_t o = 0; o < n; o += slice) {
      const size_t m = std::min(slice, n - o);
      SLG_CUDA(ix, cudaMemcpyAsync(tmp.p, static_cast<const float *>(values) + o, m * 4, cudaMemcpyDefault, st));
      slg_f32_to_bf16_kernel<<<(unsigned)std::min<size_t>((m + 255) / 256, 1u << 20), 256, 0, st>>>(tmp.as<float>(), v.values.as<__nv_bfloat16>() + o, m);
      count_launch(ix);
    }
  }
  SLG_CUDA(ix, cudaStreamSynchronize(st));
  s->vec = std::move(v);
  return SLG_OK;
}

}  // namespace

extern "C" {

int32_t slg_load_vectors(slg_index_t *ix, uint32_t segment_ord, uint32_t dim, const uint32_t *offsets, const float *values, uint64_t n_rows,
                         int32_t store_bf16) {
  return store_vectors(ix, segment_ord, dim, offsets, values, false, n_rows, store_bf16 != 0);
}

int32_t slg_load_vectors_bf16(slg_index_t *ix, uint32_t segment_ord, uint32_t dim, const uint32_t *offsets, const uint16_t *values_bf16,
                              uint64_t n_rows) {
  return store_vectors(ix, segment_ord, dim, offsets, values_bf16, true, n_rows, true);
}

int32_t slg_rerank_clauses(slg_index_t *ix, const slg_vector_clause_t *clauses, uint32_t n_clauses, uint32_t n_queries, uint32_t dim,
                           const slg_hit_t *cands, const uint32_t *cand_counts, uint32_t cand_stride, slg_hit_t *out_hits, uint32_t *out_counts,
                           float *out_vector_scores) {
  if (!ix || !cands || !cand_counts || !out_hits || !n_queries || !cand_stride) return SLG_ERR_INVALID;
  int32_t rc = check_clauses(ix, clauses, n_clauses);
  if (rc) return rc;
  SLG_CUDA(ix, cudaSetDevice(ix->device));
  cudaStream_t st = ix->stream;
  if (cand_stride > kMaxRerankCands) return fail(ix, SLG_ERR_UNSUPPORTED, "more than %u candidates per query", kMaxRerankCands);
  PoolScope pool_scope(st);  // per-call buffers from the stream-ordered pool
  DevBuf d_segs, d_c, d_n, d_s, d_v, d_o;
  uint32_t n_segs = 0;
  bool bf16 = false;
  if ((rc = segment_table(ix, dim, d_segs, &n_segs, &bf16))) return rc;
  if ((rc = check_dim(ix, dim, bf16))) return rc;
  ClauseSet cs;
  ix->ctr.last_h2d_bytes = 0;
  if ((rc = upload_clauses(ix, clauses, n_clauses, n_queries, dim, cs))) return rc;
  const size_t nh = (size_t)n_queries * cand_stride;
  SLG_CUDA(ix, d_c.alloc(nh * sizeof(HitDev)));
  SLG_CUDA(ix, cudaMemcpyAsync(d_c.p, cands, nh * sizeof(HitDev), cudaMemcpyDefault, st));
  SLG_CUDA(ix, d_n.alloc((size_t)n_queries * 4));
  SLG_CUDA(ix, cudaMemcpyAsync(d_n.p, cand_counts, (size_t)n_queries * 4, cudaMemcpyDefault, st));
  ix->ctr.last_h2d_bytes += nh * sizeof(HitDev) + (size_t)n_queries * 4;
  SLG_CUDA(ix, d_s.alloc(nh * 4));
  SLG_CUDA(ix, d_v.alloc(nh * 4));
  SLG_CUDA(ix, d_o.alloc(nh * 4));
  if ((rc = rerank_block(ix, d_segs, n_segs, bf16, cs, n_queries, dim, d_c.as<HitDev>(), d_n.as<uint32_t>(), cand_stride, d_s.as<float>(),
                         d_v.as<float>(), d_o.as<float>())))
    return rc;
  SLG_CUDA(ix, cudaMemcpyAsync(out_hits, d_c.p, nh * sizeof(HitDev), cudaMemcpyDeviceToHost, st));
  if (out_counts) SLG_CUDA(ix, cudaMemcpyAsync(out_counts, d_n.p, (size_t)n_queries * 4, cudaMemcpyDeviceToHost, st));
  if (out_vector_scores) SLG_CUDA(ix, cudaMemcpyAsync(out_vector_scores, d_o.p, nh * 4, cudaMemcpyDeviceToHost, st));
  SLG_CUDA(ix, cudaStreamSynchronize(st));
  ix->ctr.last_d2h_bytes = nh * sizeof(HitDev) + (out_counts ? (size_t)n_queries * 4 : 0) + (out_vector_scores ? nh * 4 : 0);
  return SLG_OK;
}

int32_t slg_rerank(slg_index_t *ix, const float *query_vecs, uint32_t n_queries, uint32_t dim, const slg_hit_t *cands,
                   const uint32_t *cand_counts, uint32_t cand_stride, float alpha, slg_metric_t metric, slg_hit_t *out_hits,
                   float *out_vector_scores) {
  if (!ix || !query_vecs) return SLG_ERR_INVALID;
  slg_vector_clause_t cl{};
  cl.query_vecs = query_vecs;
  cl.alpha = alpha;
  cl.boost = 1.0f;
  cl.metric = metric;
  return slg_rerank_clauses(ix, &cl, 1, n_queries, dim, cands, cand_counts, cand_stride, out_hits, nullptr, out_vector_scores);
}

int32_t slg_rerank_batch(slg_batch_t *bt, const slg_vector_clause_t *clauses, uint32_t n_clauses, uint32_t dim, int32_t sync) {
  if (!bt) return SLG_ERR_INVALID;
  slg_index *ix = bt->ix;
  int32_t rc = check_clauses(ix, clauses, n_clauses);
  if (rc) return rc;
  if (!bt->n_segs_run) return fail(ix, SLG_ERR_INVALID, "slg_rerank_batch follows slg_batch_run");
  if (bt->reranked) return fail(ix, SLG_ERR_INVALID, "the batch's last run was already reranked (run it again first)");
  if (bt->has_cursor) return fail(ix, SLG_ERR_UNSUPPORTED, "search-after cursors over hybrid scores are not built");
  if (bt->k > kMaxRerankCands) return fail(ix, SLG_ERR_UNSUPPORTED, "more than %u candidates per query", kMaxRerankCands);
  SLG_CUDA(ix, cudaSetDevice(ix->device));
  cudaStream_t st = ix->stream;
  PoolScope pool_scope(st);
  DevBuf d_segs, d_s, d_v;
  uint32_t n_segs = 0;
  bool bf16 = false;
  if ((rc = segment_table(ix, dim, d_segs, &n_segs, &bf16))) return rc;
  if ((rc = check_dim(ix, dim, bf16))) return rc;
  ClauseSet cs;
  if ((rc = upload_clauses(ix, clauses, n_clauses, bt->Q, dim, cs))) return rc;
  const size_t nh = (size_t)bt->Q * bt->k;
  SLG_CUDA(ix, d_s.alloc(nh * 4));
  SLG_CUDA(ix, d_v.alloc(nh * 4));
  SLG_CUDA(ix, cudaEventRecord(ix->ev[4], st));
  const size_t hb = nh * sizeof(HitDev), cb = (size_t)bt->Q * 4;
  // every segment's own top-k is rescored (the reference hands the concatenated per-segment lists to merge_vector_hits)
  for (uint32_t si = 0; si < bt->n_segs_run; si++) {
    unsigned char *blk = bt->results + (size_t)si * bt->result_stride;
    if ((rc = rerank_block(ix, d_segs, n_segs, bf16, cs, bt->Q, dim, reinterpret_cast<HitDev *>(blk), reinterpret_cast<uint32_t *>(blk + hb), bt->k,
                           d_s.as<float>(), d_v.as<float>(), reinterpret_cast<float *>(blk + hb + cb))))
      return rc;
  }
  if (bt->n_segs_run > 1) {
    const uint32_t S = bt->n_segs_run;
    const size_t msmem = (size_t)S * bt->k * sizeof(HitDev);
    if (msmem > ix->smem_optin || S > kMaxMergeLists) return fail(ix, SLG_ERR_UNSUPPORTED, "merge of %u segments x k=%u does not fit shared memory", S, bt->k);
    SLG_CUDA(ix, cudaFuncSetAttribute(slg_merge_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)msmem));
    unsigned char *mo = bt->results + (size_t)S * bt->result_stride;
    slg_merge_kernel<<<bt->Q, kThreads, msmem, st>>>(reinterpret_cast<const uint32_t *>(bt->results), S, bt->Q, bt->k, (uint32_t)(bt->result_stride / 4),
                                                     reinterpret_cast<HitDev *>(mo), reinterpret_cast<uint32_t *>(mo + hb), (uint32_t)((hb + cb) / 4),
                                                     reinterpret_cast<float *>(mo + hb + cb));
    count_launch(ix);
    SLG_CUDA(ix, cudaGetLastError());
  }
  SLG_CUDA(ix, cudaEventRecord(ix->ev[5], st));
  bt->reranked = true;
  if (sync) {
    SLG_CUDA(ix, cudaStreamSynchronize(st));
    float ms = 0;
    if (cudaEventElapsedTime(&ms, ix->ev[4], ix->ev[5]) == cudaSuccess) {
      ix->ctr.last_rerank_ms = ms;
      ix->ctr.rerank_ms_total += ms;
      ix->ctr.rerank_launches++;
    }
  }
  return SLG_OK;
}

int32_t slg_batch_fetch_vector_scores(slg_batch_t *bt, float *out_vector_scores) {
  if (!bt || !out_vector_scores) return SLG_ERR_INVALID;
  slg_index *ix = bt->ix;
  if (!bt->reranked) return fail(ix, SLG_ERR_INVALID, "the batch was not reranked");
  SLG_CUDA(ix, cudaSetDevice(ix->device));
  const size_t nh = (size_t)bt->Q * bt->k;
  const unsigned char *blk = bt->results + (bt->n_segs_run > 1 ? (size_t)bt->n_segs_run * bt->result_stride : 0);
  SLG_CUDA(ix, cudaMemcpyAsync(out_vector_scores, blk + nh * sizeof(HitDev) + (size_t)bt->Q * 4, nh * 4, cudaMemcpyDeviceToHost, ix->stream));
  SLG_CUDA(ix, cudaStreamSynchronize(ix->stream));
  ix->ctr.last_d2h_bytes += nh * 4;
  return SLG_OK;
}

}  // extern "C"
