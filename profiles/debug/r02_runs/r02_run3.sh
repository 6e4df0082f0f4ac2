python -m pytest tests -m gpu -q 2>&1 | tail -40 > gpurun_out/r02c_tests_all.txt
timeout 200 python profiles/kernel_only.py pv,pv_f2,corr,lse,lse_f2,lookup0,lookup,gru_zr,heads 20 > gpurun_out/r02c_kernel_times.txt 2>&1
CRAFT_PV_TRACE=gpurun_out/r02c_pv_trace.txt timeout 120 python profiles/kernel_only.py pv 1 > /dev/null 2>&1
python profiles/pv_trace_report.py gpurun_out/r02c_pv_trace.txt > gpurun_out/r02c_pv_timeline.txt 2>&1
python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r02c_bench.json 2> gpurun_out/r02c_bench.err
tail -5 gpurun_out/r02c_tests_all.txt; cat gpurun_out/r02c_kernel_times.txt; cut -c1-300 gpurun_out/r02c_bench.json
