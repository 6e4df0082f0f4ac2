timeout 600 python -m pytest tests/test_gpu_e2e.py -q -x -k "pair_stream" --tb=short 2>&1 | tail -3
for K in 20 24 20 24 20 24 30; do
python bench.py --steps $K --warmup 3 --no-cpu-baseline > gpurun_out/r02ba_bench_$K.json 2>> gpurun_out/r02ba_bench.err
python - <<PY
import json
d=json.load(open('gpurun_out/r02ba_bench_$K.json')); print('K=$K lanes', d['config']['lanes'], round(d['value'],1), round(d['ms_per_step'],3), round(d['e2e']['value'],1), round(d['config']['single_lane']['value'],1))
PY
done
tail -2 gpurun_out/r02ba_bench.err
