t=r02ae
timeout 600 python -m pytest tests/test_gpu_e2e.py -q -x -k "pair_stream" --timeout=300 2>&1 | tail -5 > gpurun_out/${t}_tests.txt
for L in 1 2 3 4; do
python bench.py --steps 24 --warmup 4 --no-cpu-baseline --lanes $L > gpurun_out/${t}_bench_l$L.json 2>> gpurun_out/${t}_bench.err
done
cat gpurun_out/${t}_tests.txt
python - <<'PY'
import json
for L in (1,2,3,4):
    d=json.load(open('gpurun_out/r02ae_bench_l%d.json'%L)); print(L, round(d['value'],1), round(d['ms_per_step'],3), round(d['e2e']['value'],1), d['config']['single_lane'], d['gpu_launches'])
PY
tail -3 gpurun_out/${t}_bench.err
