t=r02aa
timeout 900 python -m pytest tests/test_gpu_kernels.py -q -x -k "corr_build or attn_lse or clamp or lookup" --timeout=300 2>&1 | tail -8 > gpurun_out/${t}_tests.txt
timeout 900 python -m pytest tests/test_gpu_e2e.py tests/test_gpu_modules.py -q --timeout=300 2>&1 | tail -8 >> gpurun_out/${t}_tests.txt
python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/${t}_bench.json 2> gpurun_out/${t}_bench.err
timeout 200 python profiles/kernel_only.py corr,lse,lse_f2 20 > gpurun_out/${t}_kernel_times.txt 2>&1
cat gpurun_out/${t}_tests.txt gpurun_out/${t}_kernel_times.txt
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02aa_bench.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'])
print([(k['kernel'][:22], round(k['us_per_launch'],1)) for k in d['roofline']['kernels']])
PY
tail -3 gpurun_out/${t}_bench.err
