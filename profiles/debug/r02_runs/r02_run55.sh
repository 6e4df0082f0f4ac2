t=r02bb
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
$TR --nproc-per-node 4 --master-port 29631 bench.py --gpus 4 --steps 20 --warmup 3 > gpurun_out/${t}_sintel_n4.json 2> gpurun_out/${t}_n4.err
$TR --nproc-per-node 4 --master-port 29632 bench.py --impl reference --gpus 4 --steps 2 --warmup 1 > gpurun_out/${t}_ref_n4.json 2>> gpurun_out/${t}_n4.err
timeout 600 python -m pytest tests/test_gpu_modules.py tests/test_gpu_training.py -q -k "data_parallel or follow_their_input or ddp" 2>&1 | tail -4 > gpurun_out/${t}_multi_tests.txt
for f in sintel_n4 ref_n4; do echo $f; grep -v NCCL gpurun_out/${t}_$f.json | cut -c1-300; done
cat gpurun_out/${t}_multi_tests.txt
tail -3 gpurun_out/${t}_n4.err
