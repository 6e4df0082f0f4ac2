# level 0 as fp16 deltas + fp32 half-block means; encoders -> token packer directly (NHWC)
t=r02n
python -m pytest tests -q -m gpu -x 2>&1 | tail -8 > gpurun_out/${t}_tests.txt
python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/${t}_bench.json 2> gpurun_out/${t}_bench.err
CRAFT_B200_ENCODER_NHWC=0 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/${t}_bench_nchw.json 2>> gpurun_out/${t}_bench.err
timeout 200 python profiles/kernel_only.py corr,lookup_all 20 > gpurun_out/${t}_kernel_times.txt 2>&1
cat gpurun_out/${t}_tests.txt gpurun_out/${t}_kernel_times.txt
for f in bench bench_nchw; do cut -c1-200 gpurun_out/${t}_$f.json; echo; done
grep -h "level0_modes\|clip0" gpurun_out/e2e_parity.jsonl | cut -c1-300
