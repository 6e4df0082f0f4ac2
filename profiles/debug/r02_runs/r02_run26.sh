t=r02z
timeout 600 python -m pytest tests/test_gpu_kernels.py -q -x -k "instnorm_one_launch or encoder" --timeout=200 2>&1 | tail -5 > gpurun_out/${t}_tests.txt
CRAFT_B200_IN_COOP=0 timeout 600 python -m pytest tests/test_gpu_kernels.py -q -x -k "instnorm_one_launch or encoder" --timeout=200 2>&1 | tail -5 >> gpurun_out/${t}_tests.txt
python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/${t}_bench_coop1.json 2> gpurun_out/${t}_bench.err
CRAFT_B200_IN_COOP=0 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/${t}_bench_coop0.json 2>> gpurun_out/${t}_bench.err
CRAFT_B200_FUSED_IN=0 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/${t}_bench_in3.json 2>> gpurun_out/${t}_bench.err
python profiles/graph_timeline.py ${t}_coop1 > /dev/null 2>&1
CRAFT_B200_IN_COOP=0 python profiles/graph_timeline.py ${t}_coop0 > /dev/null 2>&1
CRAFT_B200_FUSED_IN=0 python profiles/graph_timeline.py ${t}_in3 > /dev/null 2>&1
cat gpurun_out/${t}_tests.txt
python - <<'PY'
import json
for f in ('coop1','coop0','in3'):
    d=json.load(open('gpurun_out/r02z_bench_%s.json'%f)); print(f, d['value'], d['ms_per_step'], d['e2e']['value'])
PY
tail -3 gpurun_out/${t}_bench.err
head -3 gpurun_out/graph_timeline_${t}_*.txt
