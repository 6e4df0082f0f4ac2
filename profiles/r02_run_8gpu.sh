python -m pytest tests/test_gpu_modules.py -q -k "data_parallel or follow_their_input" 2>&1 | tail -8 > gpurun_out/r02j_multi_tests.txt
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
for n in 2 4 8; do $TR --nproc-per-node $n --master-port 2960$n bench.py --gpus $n --config kitti --steps 10 --warmup 3 > gpurun_out/r02j_kitti_n$n.json 2> gpurun_out/r02j_kitti_n$n.err; done
$TR --nproc-per-node 8 --master-port 29611 bench.py --gpus 8 --steps 20 --warmup 3 > gpurun_out/r02j_sintel_n8.json 2> gpurun_out/r02j_sintel_n8.err
$TR --nproc-per-node 8 --master-port 29612 bench.py --gpus 8 --config train --steps 5 --warmup 3 > gpurun_out/r02j_train_n8.json 2> gpurun_out/r02j_train_n8.err
$TR --nproc-per-node 8 --master-port 29613 bench.py --gpus 8 --config train --steps 5 --warmup 3 --dropout-prob 0 > gpurun_out/r02j_train_n8_nodrop.json 2>> gpurun_out/r02j_train_n8.err
tail -4 gpurun_out/r02j_multi_tests.txt
for f in kitti_n2 kitti_n4 kitti_n8 sintel_n8 train_n8 train_n8_nodrop; do echo $f; cut -c1-230 gpurun_out/r02j_$f.json; done
tail -3 gpurun_out/r02j_train_n8.err
