# level 0 as fp16 key blocks (new default) vs on-demand; GEMM phase trace
t=r02m
python -m pytest tests/test_gpu_kernels.py -q -x -k "corr_lookup or corr_build" 2>&1 | tail -5 > gpurun_out/${t}_kernel_tests.txt
python -m pytest tests/test_gpu_e2e.py -q -x 2>&1 | tail -8 > gpurun_out/${t}_e2e_tests.txt
python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/${t}_bench_h16.json 2> gpurun_out/${t}_bench.err
CRAFT_B200_LEVEL0=ondemand python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/${t}_bench_ondemand.json 2>> gpurun_out/${t}_bench.err
timeout 200 python profiles/kernel_only.py corr,lookup_all,lookup0,lookup,iter_gemms 20 > gpurun_out/${t}_kernel_times.txt 2>&1
CRAFT_B200_NO_GRAPH=1 CRAFT_GEMM_TRACE=gpurun_out/${t}_gemm_trace_raw.txt timeout 200 python profiles/kernel_only.py iter_gemms 3 > gpurun_out/${t}_gemm_trace.log 2>&1
python profiles/gemm_trace_report.py gpurun_out/${t}_gemm_trace_raw.txt > gpurun_out/${t}_gemm_trace.txt 2>&1
rm -f gpurun_out/${t}_gemm_trace_raw.txt
cat gpurun_out/${t}_kernel_tests.txt gpurun_out/${t}_e2e_tests.txt gpurun_out/${t}_kernel_times.txt
for f in h16 ondemand; do cut -c1-200 gpurun_out/${t}_bench_$f.json; echo; done
cat gpurun_out/${t}_gemm_trace.txt
