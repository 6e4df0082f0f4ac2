t=r02q
python -m pytest tests -q -m gpu 2>&1 | tail -12 > gpurun_out/${t}_tests.txt
python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/${t}_bench.json 2> gpurun_out/${t}_bench.err
timeout 120 python profiles/kernel_only.py pv 20 > gpurun_out/${t}_kernel_times.txt 2>&1
bash profiles/run_launch_list.sh ${t} > /dev/null 2>&1
cat gpurun_out/${t}_tests.txt gpurun_out/${t}_kernel_times.txt
cut -c1-200 gpurun_out/${t}_bench.json; echo
head -40 gpurun_out/launch_summary_${t}.txt
