t=r02v
timeout 300 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_modules.py -q -x -k "corr or lookup" 2>&1 | tail -3 > gpurun_out/${t}_tests.txt
python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/${t}_bench.json 2> gpurun_out/${t}_bench.err
timeout 600 python -m pytest tests/test_gpu_e2e.py -q -x -k "flow_matches or level0" 2>&1 | tail -3 >> gpurun_out/${t}_tests.txt
cat gpurun_out/${t}_tests.txt
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02v_bench.json'))
print(d['value'], d['ms_per_step'])
for k in d['roofline']['kernels']:
    print('%-80s x%-3d %7.2f us' % (k['kernel'][:80], k['launches_per_pair'], k['us_per_launch']))
PY
