for K in 20 24 30 16; do
python bench.py --steps $K --warmup 3 --no-cpu-baseline > gpurun_out/r02ay_bench_$K.json 2>> gpurun_out/r02ay_bench.err
python - <<PY
import json
d=json.load(open('gpurun_out/r02ay_bench_$K.json')); print('K=$K lanes', d['config']['lanes'], round(d['value'],1), round(d['ms_per_step'],3), round(d['e2e']['value'],1), round(d['config']['single_lane']['value'],1))
PY
done
python bench.py --steps 20 --warmup 3 --no-cpu-baseline --lanes 3 > gpurun_out/r02ay_bench_20_l3.json 2>> gpurun_out/r02ay_bench.err
python - <<PY
import json
d=json.load(open('gpurun_out/r02ay_bench_20_l3.json')); print('K=20 lanes', d['config']['lanes'], round(d['value'],1), round(d['ms_per_step'],3), round(d['e2e']['value'],1))
PY
tail -2 gpurun_out/r02ay_bench.err
