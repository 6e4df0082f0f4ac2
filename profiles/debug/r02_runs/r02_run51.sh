timeout 600 python -m pytest tests/test_gpu_kernels.py -q -x -k "conv3x3 or instnorm or encoder" --tb=short 2>&1 | tail -3
timeout 600 python -m pytest tests/test_gpu_e2e.py -q -x --tb=short 2>&1 | tail -3
python bench.py --steps 24 --warmup 4 --no-cpu-baseline > gpurun_out/r02ax_bench.json 2> gpurun_out/r02ax_bench.err
python - <<PY
import json
d=json.load(open('gpurun_out/r02ax_bench.json')); print(round(d['value'],1), round(d['ms_per_step'],3), round(d['e2e']['value'],1), round(d['config']['single_lane']['value'],1))
PY
tail -2 gpurun_out/r02ax_bench.err
