timeout 600 python -m pytest tests/test_gpu_kernels.py -q -x -k "corr_build or attn_lse or lookup" --tb=short 2>&1 | tail -3
timeout 200 python profiles/kernel_only.py corr,lse,lse_f2 20 2>&1 | tail -3
timeout 600 python -m pytest tests/test_gpu_e2e.py -q -x --tb=short 2>&1 | tail -3
