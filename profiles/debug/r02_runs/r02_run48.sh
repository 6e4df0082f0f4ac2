t=r02au
timeout 600 python -m pytest tests/test_gpu_kernels.py -q -x -k "encoder or s2d or conv3x3 or instnorm" --tb=short 2>&1 | tail -4 > gpurun_out/${t}_tests.txt
timeout 600 python -m pytest tests/test_gpu_e2e.py -q -x --tb=short 2>&1 | tail -4 >> gpurun_out/${t}_tests.txt
timeout 300 python profiles/conv64_probe.py 2>&1 | tail -3 | cut -c1-40,200-400
for v in 1 2; do
python bench.py --steps 24 --warmup 4 --no-cpu-baseline > gpurun_out/${t}_bench_$v.json 2>> gpurun_out/${t}_bench.err
python - <<PY
import json
d=json.load(open('gpurun_out/${t}_bench_$v.json')); print('run $v', round(d['value'],1), round(d['ms_per_step'],3), round(d['e2e']['value'],1), round(d['config']['single_lane']['value'],1), d['gpu_launches']//24)
PY
done
cat gpurun_out/${t}_tests.txt
tail -3 gpurun_out/${t}_bench.err
