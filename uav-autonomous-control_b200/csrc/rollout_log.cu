// rollout_log.cu -- instantiations of the time-sliced fp32 rollout kernel with a state log (HBM-bound at full rate).
#include "rollout_impl.cuh"

namespace uavb {

void launch_rollout_sliced_log(bool mc, bool table, int grid, size_t smem, cudaStream_t st, const RolloutDev<float>& p, const SliceSched& sch) {
  if (table) {
    if (mc) rollout_sliced_kernel<true, true, true><<<grid, kRolloutThreadsLog, smem, st>>>(p, sch);
    else rollout_sliced_kernel<false, true, true><<<grid, kRolloutThreadsLog, smem, st>>>(p, sch);
  } else {
    if (mc) rollout_sliced_kernel<true, false, true><<<grid, kRolloutThreadsLog, smem, st>>>(p, sch);
    else rollout_sliced_kernel<false, false, true><<<grid, kRolloutThreadsLog, smem, st>>>(p, sch);
  }
}

}  // namespace uavb
