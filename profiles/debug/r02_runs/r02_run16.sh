t=r02p
python -m pytest tests -q -m gpu -x 2>&1 | tail -8 > gpurun_out/${t}_tests.txt
python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/${t}_bench.json 2> gpurun_out/${t}_bench.err
bash profiles/run_launch_list.sh ${t} > /dev/null 2>&1
cat gpurun_out/${t}_tests.txt
cut -c1-200 gpurun_out/${t}_bench.json; echo
head -60 gpurun_out/launch_summary_${t}.txt
