t=r02w
timeout 300 python -m pytest tests/test_gpu_e2e.py -q -x -k "pair_stream" 2>&1 | tail -3 > gpurun_out/${t}_tests.txt
python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/${t}_bench.json 2> gpurun_out/${t}_bench.err
python bench.py --steps 20 --warmup 3 --no-cpu-baseline --e2e-serial > gpurun_out/${t}_bench_serial.json 2>> gpurun_out/${t}_bench.err
cat gpurun_out/${t}_tests.txt
python - <<'PY'
import json
for f in ('r02w_bench.json','r02w_bench_serial.json'):
    d=json.load(open('gpurun_out/'+f)); print(f, d['value'], d['ms_per_step'], d['e2e']['value'])
PY
