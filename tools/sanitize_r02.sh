#!/bin/bash
# GPU box: compute-sanitizer over the kernels added in round 2 (correction loop, TMA writers, trajectory list, ground floor).
mkdir -p gpurun_out
S=/usr/local/cuda/bin/compute-sanitizer
run() { tool=$1; shift; echo "== $tool $*"; timeout 900 $S --tool $tool --error-exitcode 9 python -m pytest "$@" -m gpu -x -q 2>&1 | tail -4; }
{
run memcheck tests/test_correction_gpu.py -k "golden_and_limits or fused"
run memcheck tests/test_minsnap_gpu.py tests/test_dropin_gpu.py -k "every_spline or ragged or sampled or constraint_system or obstacle_correction"
run memcheck tests/test_rollout_gpu.py -k "tensor_stores or actual_trajectory or ground_floor or state_log_layout"
run racecheck tests/test_minsnap_gpu.py tests/test_correction_gpu.py -k "every_spline or sampled_tables or golden_and_limits"
run racecheck tests/test_rollout_gpu.py -k "state_log_layout"
run synccheck tests/test_minsnap_gpu.py tests/test_rollout_gpu.py -k "every_spline or sampled_table_and or state_log_layout"
run initcheck tests/test_correction_gpu.py tests/test_rollout_gpu.py -k "golden_and_limits or actual_trajectory"
} > gpurun_out/r02_sanitizer.log 2>&1
cat gpurun_out/r02_sanitizer.log
