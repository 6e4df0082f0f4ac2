t=r02an
for d in build/bisect/75c729f build/bisect/c0e3ced build/bisect/9b1e954; do
echo "== $d" >> gpurun_out/${t}_dp.txt
(cd $d && timeout 300 python -m pytest tests/test_gpu_modules.py -q -k "data_parallel" --tb=line -p no:cacheprovider 2>&1 | tail -3) >> gpurun_out/${t}_dp.txt
done
cat gpurun_out/${t}_dp.txt
