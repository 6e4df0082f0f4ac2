t=r02s
timeout 300 python -m pytest tests/test_gpu_kernels.py -q -x -k "attn_lse_pv" 2>&1 | tail -5 > gpurun_out/${t}_tests.txt
timeout 200 python profiles/kernel_only.py pv,pv_f2 20 > gpurun_out/${t}_kernel_times.txt 2>&1
timeout 600 python -m pytest tests/test_gpu_e2e.py tests/test_gpu_modules.py -q -x 2>&1 | tail -5 >> gpurun_out/${t}_tests.txt
python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/${t}_bench.json 2> gpurun_out/${t}_bench.err
CRAFT_PV_TRACE=gpurun_out/${t}_pv_trace_raw.txt timeout 120 python profiles/kernel_only.py pv 1 > /dev/null 2>&1
python profiles/pv_trace_report.py gpurun_out/${t}_pv_trace_raw.txt > gpurun_out/${t}_pv_timeline.txt 2>&1
cat gpurun_out/${t}_tests.txt gpurun_out/${t}_kernel_times.txt
cut -c1-200 gpurun_out/${t}_bench.json; echo
grep -A 30 "SM g0" gpurun_out/${t}_pv_timeline.txt | head -34; tail -1 gpurun_out/${t}_pv_timeline.txt
