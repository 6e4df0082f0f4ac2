python -m pytest tests -m gpu -q 2>&1 | tail -40 > gpurun_out/r02d_tests_all.txt
timeout 300 python profiles/kernel_only.py pv,pv_f2,corr,lse,lse_f2,lookup0 20 > gpurun_out/r02d_kernel_times.txt 2>&1
python bench.py --steps 20 --warmup 3 --gpu-reference > gpurun_out/r02d_bench.json 2> gpurun_out/r02d_bench.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02d_bench_reference.json 2>> gpurun_out/r02d_bench.err
python bench.py --config kitti --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02d_bench_kitti.json 2>> gpurun_out/r02d_bench.err
python bench.py --config gma --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02d_bench_gma.json 2>> gpurun_out/r02d_bench.err
tail -8 gpurun_out/r02d_tests_all.txt; cat gpurun_out/r02d_kernel_times.txt; tail -5 gpurun_out/r02d_bench.err
python - <<'PY'
import json
for f in ("r02d_bench","r02d_bench_reference","r02d_bench_kitti","r02d_bench_gma"):
    try:
        d=json.loads(open("gpurun_out/%s.json"%f).read().strip().splitlines()[-1])
        print(f, round(d["value"],3), d.get("ms_per_step"), d.get("e2e",{}).get("value"), d.get("cpu_baseline"), d.get("gpu_reference"))
        if d.get("roofline"):
            for r in d["roofline"]["kernels"]: print("   %-80s %8.1f us  %8.1f %s  frac %.3f" % (r["kernel"][:80], r["us_per_launch"], r["achieved"], r["unit"], r["frac"]))
    except Exception as e: print(f, "ERR", e)
PY
