python -m pytest tests/test_gpu_training.py -q 2>&1 | tail -40 > gpurun_out/r02f_train_tests.txt
python -m pytest tests -m gpu -q --deselect tests/test_gpu_training.py 2>&1 | tail -15 > gpurun_out/r02f_tests_all.txt
python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r02f_bench.json 2> gpurun_out/r02f_bench.err
cat gpurun_out/r02f_train_tests.txt | tail -40; tail -4 gpurun_out/r02f_tests_all.txt; cat gpurun_out/train_grad_report.txt; cut -c1-400 gpurun_out/r02f_bench.json
