// rollout_scalar.cu -- the one-drone-per-thread time-sliced fp32 rollout (metrics only, per-rollout missions), see rollout_impl.cuh.
#include "rollout_impl.cuh"

namespace uavb {

void launch_rollout_sliced_scalar(bool mc, int grid, size_t smem, cudaStream_t st, const RolloutDev<float>& p, const SliceSched& sch) {
  if (mc) rollout_sliced_scalar_kernel<true><<<grid, kScalarThreads, smem, st>>>(p, sch);
  else rollout_sliced_scalar_kernel<false><<<grid, kScalarThreads, smem, st>>>(p, sch);
}

}  // namespace uavb
