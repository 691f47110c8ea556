// slg_launch_tiles.cu — instantiations of slg_score_tiles_kernel (CTA per (doc tile, query): general k, Bool, plans)
#include "slg_launch.h"

namespace slg {
namespace {
template <bool M, bool P, bool S, bool PL>
cudaError_t go(const SegmentDev &sd, const BatchDev &bd, size_t smem, int grid, cudaStream_t st) {
  auto kern = slg_score_tiles_kernel<M, P, S, PL>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  kern<<<grid, kThreads, smem, st>>>(sd, bd);
  return cudaGetLastError();
}
template <bool M, bool PL>
cudaError_t go2(bool prune, bool stats, const SegmentDev &sd, const BatchDev &bd, size_t smem, int grid, cudaStream_t st) {
  if (prune) return stats ? go<M, true, true, PL>(sd, bd, smem, grid, st) : go<M, true, false, PL>(sd, bd, smem, grid, st);
  return stats ? go<M, false, true, PL>(sd, bd, smem, grid, st) : go<M, false, false, PL>(sd, bd, smem, grid, st);
}
}  // namespace

cudaError_t launch_score_tiles(bool matcher, bool prune, bool stats, bool plan, const SegmentDev &sd, const BatchDev &bd, size_t smem,
                               int grid, cudaStream_t st) {
  if (plan) return matcher ? go2<true, true>(prune, stats, sd, bd, smem, grid, st) : go2<false, true>(prune, stats, sd, bd, smem, grid, st);
  return matcher ? go2<true, false>(prune, stats, sd, bd, smem, grid, st) : go2<false, false>(prune, stats, sd, bd, smem, grid, st);
}
}  // namespace slg
