timeout 300 python -m pytest tests/test_gpu_kernels.py -q -x -k "conv3x3_c64" --tb=short 2>&1 | tail -4
timeout 300 python profiles/conv64_probe.py 2>&1 | tail -4
