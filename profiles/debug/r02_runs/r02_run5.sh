python -m pytest tests -m gpu -q 2>&1 | tail -40 > gpurun_out/r02e_tests_all.txt
python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r02e_bench.json 2> gpurun_out/r02e_bench.err
CRAFT_B200_PRECISION=fp16 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r02e_bench_fp16.json 2>> gpurun_out/r02e_bench.err
CRAFT_B200_PRECISION=fp32-parity python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r02e_bench_fp32parity.json 2>> gpurun_out/r02e_bench.err
tail -8 gpurun_out/r02e_tests_all.txt; tail -5 gpurun_out/r02e_bench.err
grep ":fp\|:tiers" gpurun_out/e2e_parity.jsonl | tail -14
python - <<'PY'
import json
for f in ("r02e_bench","r02e_bench_fp16","r02e_bench_fp32parity"):
    try:
        d=json.loads(open("gpurun_out/%s.json"%f).read().strip().splitlines()[-1])
        print(f, round(d["value"],3), d.get("ms_per_step"), d.get("e2e",{}).get("value"))
        if d.get("roofline") and f=="r02e_bench":
            for r in d["roofline"]["kernels"]: print("   %-80s %8.1f us  %8.1f %s  frac %.3f" % (r["kernel"][:80], r["us_per_launch"], r["achieved"], r["unit"], r["frac"]))
    except Exception as e: print(f, "ERR", e)
PY
