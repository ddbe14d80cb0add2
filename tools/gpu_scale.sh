#!/bin/bash
# weak-scaling record on one box (the launch the driver uses): Franka 30000 / Go2 4096 envs per GPU on 8 GPUs, and the mixed
# workload of BASELINE.json configs[4] (65536 envs in total, even ranks Franka, odd ranks Go2) on 2 / 4 / 8 GPUs.
# Usage: gpurun --gpus 8 -- bash tools/gpu_scale.sh <tag>
TAG=${1:-scale}
O=gpurun_out/$TAG
mkdir -p $O
run() {  # N workload steps warmup
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 400)) bench.py --gpus $1 --workload $2 --steps $3 --warmup $4 --no-cpu-baseline > $O/scale_$2_$1.json 2> $O/scale_$2_$1.err
  tail -1 $O/scale_$2_$1.json | cut -c1-170
}
run 8 franka 500 100
run 8 go2 500 300
run 8 mixed 500 300
run 4 mixed 500 300
run 2 mixed 500 300
