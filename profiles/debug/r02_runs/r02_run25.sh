t=r02y
timeout 600 python -m pytest tests/test_gpu_kernels.py -q -x -k "instnorm_one_launch or encoder" --timeout=200 2>&1 | tail -15 > gpurun_out/${t}_tests.txt
timeout 600 python -m pytest tests/test_gpu_e2e.py -q -x --timeout=300 2>&1 | tail -5 >> gpurun_out/${t}_tests.txt
python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/${t}_bench.json 2> gpurun_out/${t}_bench.err
CRAFT_B200_FUSED_IN=0 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/${t}_bench_in3.json 2>> gpurun_out/${t}_bench.err
bash profiles/run_launch_list.sh ${t} > /dev/null 2>&1
cat gpurun_out/${t}_tests.txt
python - <<'PY'
import json
for f in ('r02y_bench.json','r02y_bench_in3.json'):
    d=json.load(open('gpurun_out/'+f)); print(f, d['value'], d['ms_per_step'], d['e2e']['value'])
PY
tail -3 gpurun_out/${t}_bench.err
head -40 gpurun_out/launch_summary_${t}.txt
