#!/bin/bash
# GPU box: compute-sanitizer over the slice work of K2 v14 (per-rollout constants cached by the first slice, slices of decreasing length).
mkdir -p gpurun_out
S=/usr/local/cuda/bin/compute-sanitizer
run() { tool=$1; shift; echo "== $tool $*"; timeout 1200 $S --tool $tool --error-exitcode 9 python -m pytest "$@" -m gpu -x -q 2>&1 | tail -4; }
{
run memcheck tests/test_rollout_gpu.py -k "default_slicing or state_log_is_identical or time_sliced_schedule"
run initcheck tests/test_rollout_gpu.py -k "default_slicing or state_log_is_identical"
run racecheck tests/test_rollout_gpu.py -k "state_log_is_identical"
} > gpurun_out/r02_sanitizer_slices.log 2>&1
cat gpurun_out/r02_sanitizer_slices.log
