t=r02o
CRAFT_PV_TRACE=gpurun_out/${t}_pv_trace_raw.txt timeout 120 python profiles/kernel_only.py pv 1 > /dev/null 2>&1
python profiles/pv_trace_report.py gpurun_out/${t}_pv_trace_raw.txt > gpurun_out/${t}_pv_timeline.txt 2>&1
CRAFT_GEMM_ASHARE=1 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/${t}_bench_ashare1.json 2> gpurun_out/${t}_bench.err
CRAFT_GEMM_ASHARE=1 CRAFT_B200_NO_GRAPH=1 CRAFT_GEMM_TRACE=gpurun_out/${t}_gemm_trace_raw.txt timeout 200 python profiles/kernel_only.py iter_gemms 3 > gpurun_out/${t}_gemm_trace.log 2>&1
python profiles/gemm_trace_report.py gpurun_out/${t}_gemm_trace_raw.txt > gpurun_out/${t}_gemm_trace_ashare1.txt 2>&1
rm -f gpurun_out/${t}_gemm_trace_raw.txt
cut -c1-200 gpurun_out/${t}_bench_ashare1.json; echo
cat gpurun_out/${t}_gemm_trace_ashare1.txt
cat gpurun_out/${t}_pv_timeline.txt
