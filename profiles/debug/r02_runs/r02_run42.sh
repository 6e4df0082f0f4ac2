t=r02ap
timeout 600 python -m pytest tests/test_gpu_kernels.py -q -x -k "encoder or s2d" --tb=short 2>&1 | tail -8 > gpurun_out/${t}_tests.txt
timeout 600 python -m pytest tests/test_gpu_e2e.py -q -x --tb=short 2>&1 | tail -8 >> gpurun_out/${t}_tests.txt
for v in 1 0 1 0; do
CRAFT_B200_FOLD_BN=$v python bench.py --steps 24 --warmup 4 --no-cpu-baseline > gpurun_out/${t}_bench_fold$v.json 2>> gpurun_out/${t}_bench.err
python - <<PY
import json
d=json.load(open('gpurun_out/${t}_bench_fold$v.json')); print('fold_bn=$v', round(d['value'],1), round(d['ms_per_step'],3), round(d['e2e']['value'],1), d['config']['single_lane']['value'], d['gpu_launches']//24)
PY
done
cat gpurun_out/${t}_tests.txt
tail -3 gpurun_out/${t}_bench.err
