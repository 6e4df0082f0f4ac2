#!/bin/bash
# One GPU-box pass that regenerates the round-2 artefacts under profiles/ (run under gpurun):
#   bash profiles/run_final_profiles_r02.sh
# -> gpurun_out/r02_*: bench lines (default config at 3 lanes and at 1 lane, kitti, gma), the reference arm, the ncu
#    launch list + summary of one graph-replayed step, `ncu --set full` captures (key metrics + SASS stall summary as
#    text; the .ncu-rep files are deleted again: gpurun_out/ must stay under 64 MiB) of the aggregator P.V kernel, the
#    correlation build, the LSE kernel, the SepConvGRU z/r GEMM, the all-level lookup and the mode finalize,
#    per-kernel CUDA-event timings, the kernel timeline of one replay, the attn_pv and GEMM phase timelines.
tag=r02
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/${tag}_gpu.txt 2>&1
python bench.py --steps 24 --warmup 4 --gpu-reference > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
python bench.py --steps 24 --warmup 4 --lanes 1 --no-cpu-baseline > gpurun_out/${tag}_bench_lanes1.json 2>> gpurun_out/${tag}_bench.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${tag}_bench_reference.json 2>> gpurun_out/${tag}_bench.err
python bench.py --config kitti --steps 12 --warmup 4 --no-cpu-baseline > gpurun_out/${tag}_bench_kitti.json 2>> gpurun_out/${tag}_bench.err
python bench.py --config gma --steps 12 --warmup 4 --no-cpu-baseline > gpurun_out/${tag}_bench_gma.json 2>> gpurun_out/${tag}_bench.err
timeout 200 python profiles/kernel_only.py pv,pv_f2,corr,lse,lse_f2,lookup_all,finalize,gru_zr,heads 20 > gpurun_out/${tag}_kernel_times.txt 2>&1
bash profiles/run_launch_list.sh ${tag} > /dev/null 2>&1
cp gpurun_out/launches_${tag}.csv gpurun_out/${tag}_launches.csv; cp gpurun_out/launch_summary_${tag}.txt gpurun_out/${tag}_launch_summary_graph.txt
rm -f gpurun_out/launches_${tag}.csv
timeout 300 python profiles/graph_timeline.py ${tag} > /dev/null 2>&1
for k in pv corr lse gru_zr lookup_all finalize; do
  case $k in pv) rx=attn_pv_kernel; n=1;; corr) rx=scores_kernel; n=1;; lse) rx=scores_kernel; n=1;; gru_zr) rx=shift_gemm_kernel; n=2;; lookup_all) rx=corr_lookup_kernel; n=1;; finalize) rx=modes_finalize_kernel; n=1;; esac
  KO_PROFILE=1 CRAFT_B200_NO_GRAPH=1 timeout 400 ncu --set full --clock-control none --import-source on --profile-from-start off \
      -k regex:$rx -c $n -f -o gpurun_out/${tag}_$k python profiles/kernel_only.py $k 1 > gpurun_out/${tag}_ncu_$k.log 2>&1
  python profiles/ncu_key_metrics.py gpurun_out/${tag}_$k.ncu-rep gpurun_out/${tag}_${k}_traffic.json > gpurun_out/${tag}_${k}_ncu_full.txt 2>&1
  ncu -i gpurun_out/${tag}_$k.ncu-rep --page source --csv --print-source sass > gpurun_out/_sass.csv 2>/dev/null
  python profiles/ncu_sass_stalls.py gpurun_out/_sass.csv 30 > gpurun_out/${tag}_${k}_sass_stalls.txt 2>&1
  rm -f gpurun_out/${tag}_$k.ncu-rep gpurun_out/_sass.csv
done
CRAFT_PV_TRACE=gpurun_out/${tag}_pv_trace_raw.txt timeout 120 python profiles/kernel_only.py pv 1 > /dev/null 2>&1
python profiles/pv_trace_report.py gpurun_out/${tag}_pv_trace_raw.txt > gpurun_out/${tag}_pv_timeline.txt 2>&1
CRAFT_B200_NO_GRAPH=1 CRAFT_GEMM_TRACE=gpurun_out/${tag}_gemm_trace_raw.txt timeout 200 python profiles/kernel_only.py iter_gemms 3 > /dev/null 2>&1
python profiles/gemm_trace_report.py gpurun_out/${tag}_gemm_trace_raw.txt > gpurun_out/${tag}_gemm_trace.txt 2>&1
rm -f gpurun_out/${tag}_gemm_trace_raw.txt gpurun_out/${tag}_pv_trace_raw.txt
cat gpurun_out/${tag}_kernel_times.txt; head -30 gpurun_out/${tag}_launch_summary_graph.txt; cat gpurun_out/${tag}_pv_ncu_full.txt gpurun_out/${tag}_corr_ncu_full.txt
tail -5 gpurun_out/${tag}_bench.err
du -sh gpurun_out
