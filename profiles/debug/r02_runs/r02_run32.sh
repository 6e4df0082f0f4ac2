t=r02af
CRAFT_B200_FUSED_IN=1 python bench.py --steps 24 --warmup 4 --no-cpu-baseline --lanes 3 > gpurun_out/${t}_bench_in1.json 2>> gpurun_out/${t}_bench.err
python bench.py --steps 24 --warmup 4 --no-cpu-baseline --lanes 3 > gpurun_out/${t}_bench_in0.json 2>> gpurun_out/${t}_bench.err
CRAFT_B200_FUSED_IN=1 python bench.py --steps 24 --warmup 4 --no-cpu-baseline --lanes 3 > gpurun_out/${t}_bench_in1b.json 2>> gpurun_out/${t}_bench.err
python bench.py --steps 24 --warmup 4 --no-cpu-baseline --lanes 3 > gpurun_out/${t}_bench_in0b.json 2>> gpurun_out/${t}_bench.err
python - <<'PY'
import json
for L in ('in1','in0','in1b','in0b'):
    d=json.load(open('gpurun_out/r02af_bench_%s.json'%L)); print(L, round(d['value'],1), round(d['ms_per_step'],3), round(d['e2e']['value'],1))
PY
tail -3 gpurun_out/${t}_bench.err
