#!/bin/bash
# ncu launch list (cold-cache, serialised per-launch device times) of one bench step.
# usage (under gpurun): bash profiles/run_launch_list.sh <tag>
tag=${1:-r01}
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_$tag.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/bench_under_ncu_$tag.log 2>&1
python profiles/summarize_launches.py gpurun_out/launches_$tag.csv > gpurun_out/launch_summary_$tag.txt 2>&1
tail -40 gpurun_out/launch_summary_$tag.txt
