tag=r02
python bench.py --steps 24 --warmup 4 --gpu-reference > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
python bench.py --steps 24 --warmup 4 --lanes 1 --no-cpu-baseline > gpurun_out/${tag}_bench_lanes1.json 2>> gpurun_out/${tag}_bench.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${tag}_bench_reference.json 2>> gpurun_out/${tag}_bench.err
python bench.py --config kitti --steps 12 --warmup 4 --no-cpu-baseline > gpurun_out/${tag}_bench_kitti.json 2>> gpurun_out/${tag}_bench.err
python bench.py --config gma --steps 12 --warmup 4 --no-cpu-baseline > gpurun_out/${tag}_bench_gma.json 2>> gpurun_out/${tag}_bench.err
python bench.py > gpurun_out/${tag}_bench_noflags.json 2>> gpurun_out/${tag}_bench.err
python - <<PY
import json
for f in ("r02_bench","r02_bench_lanes1","r02_bench_kitti","r02_bench_gma","r02_bench_reference","r02_bench_noflags"):
    d=json.loads(open("gpurun_out/%s.json"%f).readline()); print(f, round(d["value"],2), round(d.get("ms_per_step"),3), d.get("e2e",{}).get("value"), d.get("steps"))
PY
