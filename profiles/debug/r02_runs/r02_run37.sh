t=r02ak
timeout 600 python -m pytest tests/test_gpu_modules.py -q -k "data_parallel" --tb=short 2>&1 | tail -40 > gpurun_out/${t}_dp.txt
cat gpurun_out/${t}_dp.txt
