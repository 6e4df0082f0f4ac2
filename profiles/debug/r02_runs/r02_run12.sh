python -m pytest tests/test_gpu_kernels.py -q -k "gemm" 2>&1 | tail -5 > gpurun_out/r02l_gemm_tests.txt
python profiles/gemm_sweep2.py > gpurun_out/r02l_sweep_default.txt 2>&1
CRAFT_GEMM_ASHARE=1 python profiles/gemm_sweep2.py > gpurun_out/r02l_sweep_ashare.txt 2>&1
CRAFT_GEMM_ASHARE=1 python -m pytest tests/test_gpu_e2e.py -q -k "flow_matches" 2>&1 | tail -4 > gpurun_out/r02l_ashare_e2e.txt
CRAFT_GEMM_ASHARE=1 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r02l_bench_ashare.json 2> gpurun_out/r02l_bench.err
python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r02l_bench.json 2>> gpurun_out/r02l_bench.err
python bench.py --config train --steps 5 --warmup 3 > gpurun_out/r02l_train_n1.json 2>> gpurun_out/r02l_bench.err
python bench.py --config train --steps 5 --warmup 3 --dropout-prob 0 > gpurun_out/r02l_train_n1_nodrop.json 2>> gpurun_out/r02l_bench.err
tail -3 gpurun_out/r02l_gemm_tests.txt; paste -d'|' <(cut -c1-95 gpurun_out/r02l_sweep_default.txt) <(cut -c52-95 gpurun_out/r02l_sweep_ashare.txt); tail -3 gpurun_out/r02l_ashare_e2e.txt
for f in bench_ashare bench train_n1 train_n1_nodrop; do cut -c1-190 gpurun_out/r02l_$f.json; echo; done
