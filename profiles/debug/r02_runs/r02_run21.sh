t=r02u
timeout 300 python -m pytest tests/test_gpu_kernels.py -q -x -k "attn_lse_pv" 2>&1 | tail -3 > gpurun_out/${t}_tests.txt
timeout 200 python profiles/kernel_only.py pv,pv_f2 20 > gpurun_out/${t}_kernel_times.txt 2>&1
timeout 900 python -m pytest tests -q -m gpu 2>&1 | tail -6 >> gpurun_out/${t}_tests.txt
python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/${t}_bench.json 2> gpurun_out/${t}_bench.err
cat gpurun_out/${t}_tests.txt gpurun_out/${t}_kernel_times.txt
cut -c1-200 gpurun_out/${t}_bench.json; echo
