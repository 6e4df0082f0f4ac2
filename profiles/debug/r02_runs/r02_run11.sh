timeout 200 python profiles/kernel_only.py pv,pv_f2,lse,lse_f2,corr 20 > gpurun_out/r02k_kernel_times.txt 2>&1
python -m pytest tests -m gpu -q -x 2>&1 | tail -6 > gpurun_out/r02k_tests.txt
python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r02k_bench.json 2> gpurun_out/r02k_bench.err
timeout 300 python profiles/debug/train_profile.py > gpurun_out/r02k_train_profile_dropout.txt 2>&1
timeout 300 python profiles/debug/train_profile.py 0 > gpurun_out/r02k_train_profile_kernels.txt 2>&1
grep -v Warn gpurun_out/r02k_kernel_times.txt; tail -3 gpurun_out/r02k_tests.txt; cut -c1-200 gpurun_out/r02k_bench.json; echo; grep -v Warn gpurun_out/r02k_train_profile_dropout.txt | cut -c1-200 | head -34; grep -v Warn gpurun_out/r02k_train_profile_kernels.txt | cut -c1-200 | head -34
