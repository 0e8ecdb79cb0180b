// rollout_sliced.cu -- instantiations of the time-sliced persistent rollout kernel (metrics-only fp32, the production path).
#include "rollout_impl.cuh"

namespace uavb {

void launch_rollout_sliced(bool mc, bool table, int grid, size_t smem, cudaStream_t st, const RolloutDev<float>& p, const SliceSched& sch) {
  if (table) {
    if (mc) rollout_sliced_kernel<true, true, false><<<grid, kRolloutThreads, smem, st>>>(p, sch);
    else rollout_sliced_kernel<false, true, false><<<grid, kRolloutThreads, smem, st>>>(p, sch);
  } else {
    if (mc) rollout_sliced_kernel<true, false, false><<<grid, kRolloutThreads, smem, st>>>(p, sch);
    else rollout_sliced_kernel<false, false, false><<<grid, kRolloutThreads, smem, st>>>(p, sch);
  }
}

}  // namespace uavb
