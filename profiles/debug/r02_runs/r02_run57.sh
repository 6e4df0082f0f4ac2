t=r02bc
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
$TR --nproc-per-node 8 --master-port 29641 bench.py --gpus 8 --steps 20 --warmup 3 > gpurun_out/${t}_sintel_n8.json 2> gpurun_out/${t}_n8.err
$TR --nproc-per-node 8 --master-port 29642 bench.py --gpus 8 --config kitti --steps 12 --warmup 3 > gpurun_out/${t}_kitti_n8.json 2>> gpurun_out/${t}_n8.err
for f in sintel_n8 kitti_n8; do echo $f; grep -v NCCL gpurun_out/${t}_$f.json | cut -c1-260; done
tail -3 gpurun_out/${t}_n8.err
