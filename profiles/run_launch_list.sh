#!/bin/bash
# ncu launch list (cold-cache, serialised per-launch device times) of ONE bench step (one 448x1024 pair,
# 12 iterations), profiled inside a cudaProfilerStart/Stop range after the warm-up.
# usage (under gpurun): bash profiles/run_launch_list.sh <tag>
tag=${1:-r01}
mkdir -p gpurun_out
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file gpurun_out/launches_$tag.csv python bench.py --ncu > gpurun_out/bench_under_ncu_$tag.log 2>&1
echo "ncu rc=$?"
python profiles/summarize_launches.py gpurun_out/launches_$tag.csv > gpurun_out/launch_summary_$tag.txt 2>&1
head -45 gpurun_out/launch_summary_$tag.txt
