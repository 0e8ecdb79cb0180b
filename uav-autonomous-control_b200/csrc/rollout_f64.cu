// rollout_f64.cu -- instantiations of the fp64 validation build of the rollout kernel (same source, every state variable
// and operation in double; used by the parity tests to separate algorithmic from rounding differences).
#include "rollout_impl.cuh"

namespace uavb {

void launch_rollout_f64(bool log, bool mc, int grid, size_t smem, cudaStream_t st, const RolloutDev<double>& p) {
  if (log && mc) rollout_kernel<double, true, true><<<grid, kRolloutThreadsF64, smem, st>>>(p);
  else if (log) rollout_kernel<double, true, false><<<grid, kRolloutThreadsF64, smem, st>>>(p);
  else if (mc) rollout_kernel<double, false, true><<<grid, kRolloutThreadsF64, smem, st>>>(p);
  else rollout_kernel<double, false, false><<<grid, kRolloutThreadsF64, smem, st>>>(p);
}

}  // namespace uavb
