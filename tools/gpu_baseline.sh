#!/bin/bash
# GPU box: full GPU test suite, the bench line (both arms), an ncu launch list of a short bench run and a full capture of K2.
#   tools/gpu_baseline.sh <tag>
tag=${1:-r02}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q -s > gpurun_out/pytest_gpu_$tag.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu_$tag.log
python bench.py > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; echo "bench rc=$?"
python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/${tag}_bench_ref.json 2>> gpurun_out/${tag}_bench.err; echo "ref rc=$?"
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_launches.csv python bench.py --steps 2 --warmup 1 --no-extras --no-cpu-baseline > gpurun_out/${tag}_launches.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:rollout_sliced -s 1 -c 1 -f -o gpurun_out/${tag}_k2 python tools/k2_probe.py 100000 2 > gpurun_out/${tag}_k2_ncu.log 2>&1; echo "ncu k2 rc=$?"
head -c 1500 gpurun_out/${tag}_bench.json
