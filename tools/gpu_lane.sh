#!/bin/bash
# GPU session for the lane-per-environment kernel: parity suite with GSR_GROUP_SIZE=1, timing sweep, ncu capture.
set -u
TAG=${1:-lane}
O=gpurun_out/$TAG
mkdir -p $O
( GSR_GROUP_SIZE=1 timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_gpu_lane.log 2>&1; echo "pytest exit $?" >> $O/pytest_gpu_lane.log )
tail -15 $O/pytest_gpu_lane.log
for W in 1 2; do
  for D in 1 0; do
    GSR_GROUP_SIZE=1 GSR_WARPS=$W GSR_DIAGNOSTICS=$D timeout 300 python tools/sweep_one.py franka 30000 2>&1 | grep "##"
  done
done
GSR_GROUP_SIZE=1 GSR_WARPS=1 timeout 300 python tools/sweep_one.py go2 4096 2>&1 | grep "##"
GSR_GROUP_SIZE=1 timeout 600 python bench.py --steps 500 --warmup 50 > $O/bench_franka_lane.json 2> $O/bench_franka_lane.err; tail -1 $O/bench_franka_lane.json
GSR_GROUP_SIZE=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_step -s 10 -c 1 -f -o $O/prof_franka_lane \
     python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $O/ncu_full_franka_lane.log 2>&1
ls -la $O
