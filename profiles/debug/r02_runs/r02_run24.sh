t=r02x
timeout 600 python -m pytest tests/test_gpu_kernels.py -q -x -k "attn_lse_pv_finalize or gma or pv" --timeout=200 2>&1 | tail -5 > gpurun_out/${t}_tests.txt
timeout 600 python -m pytest tests/test_gpu_e2e.py -q -x --timeout=300 2>&1 | tail -5 >> gpurun_out/${t}_tests.txt
python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/${t}_bench.json 2> gpurun_out/${t}_bench.err
CRAFT_PV_BULK=0 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/${t}_bench_nobulk.json 2>> gpurun_out/${t}_bench.err
timeout 200 python profiles/kernel_only.py pv,finalize 20 > gpurun_out/${t}_kernel_times.txt 2>&1
CRAFT_PV_BULK=0 timeout 200 python profiles/kernel_only.py pv 20 >> gpurun_out/${t}_kernel_times.txt 2>&1
CRAFT_PV_TRACE=gpurun_out/${t}_pv_trace_raw.txt timeout 120 python profiles/kernel_only.py pv 1 > /dev/null 2>&1
python profiles/pv_trace_report.py gpurun_out/${t}_pv_trace_raw.txt > gpurun_out/${t}_pv_timeline.txt 2>&1
cat gpurun_out/${t}_tests.txt gpurun_out/${t}_kernel_times.txt
python - <<'PY'
import json
for f in ('r02x_bench.json','r02x_bench_nobulk.json'):
    d=json.load(open('gpurun_out/'+f)); print(f, d['value'], d['ms_per_step'], d['e2e']['value'])
PY
tail -3 gpurun_out/${t}_bench.err
