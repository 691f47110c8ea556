// slg_launch_warp.cu — instantiations of slg_score_warp_kernel (warp per (doc-range group, query), one accumulator
// slot per doc: Bool queries, ScorePlans, statistics, the reference's summation order)
#include "slg_launch.h"

namespace slg {
namespace {
template <bool M, bool P, bool S, bool G, bool PL>
cudaError_t go(const SegmentDev &sd, const WarpBatchDev &wb, size_t smem, int grid, cudaStream_t st) {
  auto kern = slg_score_warp_kernel<M, P, S, G, PL>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  kern<<<grid, kThreads, smem, st>>>(sd, wb);
  return cudaGetLastError();
}
template <bool M, bool G, bool PL>
cudaError_t go2(bool prune, bool stats, const SegmentDev &sd, const WarpBatchDev &wb, size_t smem, int grid, cudaStream_t st) {
  if (prune) return stats ? go<M, true, true, G, PL>(sd, wb, smem, grid, st) : go<M, true, false, G, PL>(sd, wb, smem, grid, st);
  return stats ? go<M, false, true, G, PL>(sd, wb, smem, grid, st) : go<M, false, false, G, PL>(sd, wb, smem, grid, st);
}
}  // namespace

// staged: the (doc, score) stream form (plain OR queries, resident scores); otherwise postings are scored in place
// and group masks are kept when the batch has a matcher
cudaError_t launch_score_warp(bool matcher, bool prune, bool stats, bool staged, bool plan, const SegmentDev &sd, const WarpBatchDev &wb,
                              size_t smem, int grid, cudaStream_t st) {
  if (plan) {  // ScorePlans: resident scores for plain OR queries, in-place scoring + group masks otherwise
    if (staged && !matcher) return go2<false, true, true>(prune, stats, sd, wb, smem, grid, st);
    return go2<true, false, true>(prune, stats, sd, wb, smem, grid, st);
  }
  if (staged && !matcher) return go2<false, true, false>(prune, stats, sd, wb, smem, grid, st);
  return matcher ? go2<true, false, false>(prune, stats, sd, wb, smem, grid, st) : go2<false, false, false>(prune, stats, sd, wb, smem, grid, st);
}
}  // namespace slg
