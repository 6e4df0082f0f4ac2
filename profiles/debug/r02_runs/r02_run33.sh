t=r02ag
for P in 0 1 2; do echo "POLY=$P" >> gpurun_out/${t}_poly.txt; CRAFT_PV_POLY=$P timeout 200 python profiles/kernel_only.py pv 20 >> gpurun_out/${t}_poly.txt 2>&1; done
CRAFT_PV_POLY=1 python bench.py --steps 24 --warmup 4 --no-cpu-baseline --lanes 3 > gpurun_out/${t}_bench_poly1.json 2>> gpurun_out/${t}_bench.err
python bench.py --steps 24 --warmup 4 --no-cpu-baseline --lanes 3 > gpurun_out/${t}_bench_poly0.json 2>> gpurun_out/${t}_bench.err
python bench.py --config kitti --steps 12 --warmup 4 --no-cpu-baseline > gpurun_out/${t}_bench_kitti.json 2>> gpurun_out/${t}_bench.err
python bench.py --config gma --steps 12 --warmup 4 --no-cpu-baseline > gpurun_out/${t}_bench_gma.json 2>> gpurun_out/${t}_bench.err
cat gpurun_out/${t}_poly.txt
python - <<'PY'
import json
for L in ('poly1','poly0','kitti','gma'):
    d=json.load(open('gpurun_out/r02ag_bench_%s.json'%L)); print(L, round(d['value'],1), round(d['ms_per_step'],3), round(d['e2e']['value'],1), d['config']['single_lane'])
PY
tail -3 gpurun_out/${t}_bench.err
