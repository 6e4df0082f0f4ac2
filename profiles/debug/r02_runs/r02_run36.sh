t=r02aj
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
$TR --nproc-per-node 2 --master-port 29621 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/${t}_sintel_n2.json 2> gpurun_out/${t}_n2.err
$TR --nproc-per-node 2 --master-port 29622 bench.py --gpus 2 --config kitti --steps 10 --warmup 3 > gpurun_out/${t}_kitti_n2.json 2>> gpurun_out/${t}_n2.err
$TR --nproc-per-node 2 --master-port 29623 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > gpurun_out/${t}_ref_n2.json 2>> gpurun_out/${t}_n2.err
timeout 600 python -m pytest tests/test_gpu_modules.py -q -k "data_parallel or follow_their_input" 2>&1 | tail -4 > gpurun_out/${t}_multi_tests.txt
for f in sintel_n2 kitti_n2 ref_n2; do echo $f; cut -c1-260 gpurun_out/${t}_$f.json; done
cat gpurun_out/${t}_multi_tests.txt
tail -5 gpurun_out/${t}_n2.err
