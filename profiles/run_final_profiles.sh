#!/bin/bash
# One GPU-box pass that regenerates every artefact under profiles/ for a round (run under gpurun):
#   bash profiles/run_final_profiles.sh r01
# -> gpurun_out/<tag>_*: bench line, ncu launch list + summary, `ncu --set full` captures of the
#    aggregator P.V kernel and the correlation build, per-kernel CUDA-event timings.
tag=${1:-r01}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/${tag}_gpu.txt 2>&1
python bench.py --steps 20 --warmup 3 > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${tag}_bench_reference.json 2>> gpurun_out/${tag}_bench.err
timeout 120 python profiles/kernel_only.py pv,pv_f2,corr,lse,lse_f2,lookup0,lookup,gru_zr,heads 20 > gpurun_out/${tag}_kernel_times.txt 2>&1
bash profiles/run_launch_list.sh ${tag} > /dev/null 2>&1
# model warm-up (iters=1) launches attn_pv twice and scores_kernel six times before the kernel under test
CRAFT_B200_NO_GRAPH=1 timeout 300 ncu --set full --clock-control none --import-source on -k regex:attn_pv_kernel -s 2 -c 1 -f \
    -o gpurun_out/${tag}_pv python profiles/kernel_only.py pv 2 > gpurun_out/${tag}_ncu_pv.log 2>&1
python profiles/ncu_key_metrics.py gpurun_out/${tag}_pv.ncu-rep gpurun_out/${tag}_pv_traffic.json > gpurun_out/${tag}_pv_ncu_full.txt 2>&1
CRAFT_B200_NO_GRAPH=1 timeout 300 ncu --set full --clock-control none --import-source on -k regex:scores_kernel -s 6 -c 1 -f \
    -o gpurun_out/${tag}_corr python profiles/kernel_only.py corr 2 > gpurun_out/${tag}_ncu_corr.log 2>&1
python profiles/ncu_key_metrics.py gpurun_out/${tag}_corr.ncu-rep > gpurun_out/${tag}_corr_ncu_full.txt 2>&1
cat gpurun_out/${tag}_bench.json gpurun_out/${tag}_kernel_times.txt gpurun_out/${tag}_pv_ncu_full.txt
