t=r02aw
timeout 900 python -m pytest tests/test_gpu_kernels.py -q -x -k "gemm" --tb=short 2>&1 | tail -6 > gpurun_out/${t}_tests.txt
timeout 900 python -m pytest tests/test_gpu_e2e.py tests/test_gpu_modules.py -q -x --tb=short 2>&1 | tail -6 >> gpurun_out/${t}_tests.txt
for v in 1 2; do
python bench.py --steps 24 --warmup 4 --no-cpu-baseline > gpurun_out/${t}_bench_$v.json 2>> gpurun_out/${t}_bench.err
python - <<PY
import json
d=json.load(open('gpurun_out/${t}_bench_$v.json')); print('run $v', round(d['value'],1), round(d['ms_per_step'],3), round(d['e2e']['value'],1), round(d['config']['single_lane']['value'],1), d['gpu_launches']//24)
print([(k['kernel'][:24], round(k['us_per_launch'],1)) for k in d['roofline']['kernels'] if 'gemm' in k['kernel'].lower() or 'GEMM' in k['kernel']])
PY
done
CRAFT_B200_NO_GRAPH=1 CRAFT_GEMM_TRACE=gpurun_out/${t}_gemm_trace_raw.txt timeout 200 python profiles/kernel_only.py iter_gemms 3 > /dev/null 2>&1
python profiles/gemm_trace_report.py gpurun_out/${t}_gemm_trace_raw.txt > gpurun_out/${t}_gemm_trace.txt 2>&1
rm -f gpurun_out/${t}_gemm_trace_raw.txt
cat gpurun_out/${t}_tests.txt
grep -A1 "epi=1 T=5\|epi=2 T=5" gpurun_out/${t}_gemm_trace.txt | head -8
tail -3 gpurun_out/${t}_bench.err
