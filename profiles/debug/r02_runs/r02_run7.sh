python -m pytest tests/test_gpu_training.py -q -k "kernel_forward or batch_of_two" 2>&1 | tail -30 > gpurun_out/r02g_train_tests.txt
python -m pytest tests/test_gpu_kernels.py -q -k "lookup0" 2>&1 | tail -12 > gpurun_out/r02g_lookup0_tests.txt
timeout 200 python profiles/kernel_only.py lookup0,pv 20 > gpurun_out/r02g_kernel_times.txt 2>&1
timeout 300 python profiles/debug/attn_sparsity_probe.py > gpurun_out/r02g_sparsity.txt 2>&1
python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r02g_bench.json 2> gpurun_out/r02g_bench.err
cat gpurun_out/train_grad_report.txt; tail -25 gpurun_out/r02g_train_tests.txt | cut -c1-300; tail -5 gpurun_out/r02g_lookup0_tests.txt; cat gpurun_out/r02g_kernel_times.txt; grep -v Warn gpurun_out/r02g_sparsity.txt | tail -4; cut -c1-200 gpurun_out/r02g_bench.json
