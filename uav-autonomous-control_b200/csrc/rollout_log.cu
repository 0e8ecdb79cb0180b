// rollout_log.cu -- instantiations of the one-shot fp32 rollout kernel with a state log (HBM-bound at full rate).
#include "rollout_impl.cuh"

namespace uavb {

void launch_rollout_log_f32(bool mc, int grid, size_t smem, cudaStream_t st, const RolloutDev<float>& p) {
  if (mc) rollout_kernel<float, true, true><<<grid, kRolloutThreads, smem, st>>>(p);
  else rollout_kernel<float, true, false><<<grid, kRolloutThreads, smem, st>>>(p);
}

}  // namespace uavb
