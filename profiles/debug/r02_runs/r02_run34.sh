t=r02ah
timeout 900 python -m pytest tests/test_gpu_kernels.py -q -x -k "small_kernels or gemm_conv or corr_build" --timeout=300 2>&1 | tail -5 > gpurun_out/${t}_tests.txt
timeout 600 python -m pytest tests/test_gpu_e2e.py -q -x --timeout=300 2>&1 | tail -5 >> gpurun_out/${t}_tests.txt
python bench.py --steps 24 --warmup 4 --no-cpu-baseline --lanes 3 > gpurun_out/${t}_bench.json 2>> gpurun_out/${t}_bench.err
python bench.py --steps 24 --warmup 4 --no-cpu-baseline --lanes 3 > gpurun_out/${t}_bench_b.json 2>> gpurun_out/${t}_bench.err
timeout 300 python profiles/graph_timeline.py ${t} > /dev/null 2>&1
cat gpurun_out/${t}_tests.txt
python - <<'PY'
import json
for L in ('', '_b'):
    d=json.load(open('gpurun_out/r02ah_bench%s.json'%L)); print(L, round(d['value'],1), round(d['ms_per_step'],3), round(d['e2e']['value'],1), d['config']['single_lane'])
PY
grep -n "convf1" gpurun_out/graph_timeline_${t}.txt
tail -3 gpurun_out/${t}_bench.err
