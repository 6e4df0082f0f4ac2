t=r02ao
timeout 600 python -m pytest tests/test_gpu_modules.py -q -k "data_parallel or follow_their_input" --tb=short 2>&1 | tail -6 > gpurun_out/${t}_dp.txt
timeout 600 python -m pytest tests/test_gpu_kernels.py -q -k "attn_lse_pv_finalize" --tb=short 2>&1 | tail -6 >> gpurun_out/${t}_dp.txt
cat gpurun_out/${t}_dp.txt
