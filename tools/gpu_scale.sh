#!/bin/bash
# weak-scaling record: Franka 30000 envs per GPU on N = 1, 2, 4, 8 GPUs of one box (the launch the driver uses)
O=gpurun_out/scale
mkdir -p $O
python bench.py --gpus 1 --steps 500 --warmup 50 --no-cpu-baseline > $O/scale_1.json 2>$O/scale_1.err
for N in 2 4 8; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$N bench.py --gpus $N --steps 500 --warmup 50 --no-cpu-baseline > $O/scale_$N.json 2>$O/scale_$N.err
done
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29520 bench.py --gpus 8 --workload go2 --steps 500 --warmup 300 --no-cpu-baseline > $O/scale_go2_8.json 2>$O/scale_go2_8.err
for f in $O/scale_*.json; do tail -1 $f | cut -c1-150; done
