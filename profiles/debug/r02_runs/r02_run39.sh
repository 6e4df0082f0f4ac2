t=r02am
for d in build/bisect/4aa2c96 build/bisect/dcf7bb7 .; do
for rep in 1 2; do
echo "== $d rep $rep" >> gpurun_out/${t}_dp.txt
(cd $d && timeout 300 python -m pytest tests/test_gpu_modules.py -q -k "data_parallel" --tb=line -p no:cacheprovider 2>&1 | tail -3) >> gpurun_out/${t}_dp.txt
done
done
cat gpurun_out/${t}_dp.txt
