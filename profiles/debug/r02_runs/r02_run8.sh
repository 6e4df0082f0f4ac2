for cap in 0 144 192 256 400; do echo "CAP=$cap" >> gpurun_out/r02h_lookup0.txt; CRAFT_LOOKUP0_CAP=$cap timeout 100 python profiles/kernel_only.py lookup0 20 >> gpurun_out/r02h_lookup0.txt 2>&1; done
python -m pytest tests -m gpu -q 2>&1 | tail -15 > gpurun_out/r02h_tests_all.txt
CRAFT_LOOKUP0_CAP=0 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r02h_bench_cap0.json 2> gpurun_out/r02h_bench.err
python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r02h_bench.json 2>> gpurun_out/r02h_bench.err
cat gpurun_out/r02h_lookup0.txt | grep -v Warn; tail -6 gpurun_out/r02h_tests_all.txt; cut -c1-200 gpurun_out/r02h_bench_cap0.json; echo; cut -c1-200 gpurun_out/r02h_bench.json
