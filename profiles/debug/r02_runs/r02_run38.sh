t=r02al
for mode in fused_one_stream plain fused; do
echo "== CRAFT_B200_DP_ENCODER=$mode" >> gpurun_out/${t}_dp.txt
CRAFT_B200_DP_ENCODER=$mode timeout 600 python -m pytest tests/test_gpu_modules.py -q -k "data_parallel" --tb=line 2>&1 | tail -6 >> gpurun_out/${t}_dp.txt
done
cat gpurun_out/${t}_dp.txt
