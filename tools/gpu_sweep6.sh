#!/bin/bash
for W in 1 2 4 8; do
  GSR_WARPS=$W timeout 600 python bench.py --workload go2 --steps 1000 --warmup 300 --no-cpu-baseline 2>&1 | tail -1 | cut -c1-120 | sed "s/^/go2 warps=$W /"
done
for W in 1 4; do
  GSR_WARPS=$W timeout 600 python bench.py --steps 1000 --warmup 100 --no-cpu-baseline 2>&1 | tail -1 | cut -c1-120 | sed "s/^/franka warps=$W /"
done
GSR_REORDER_PERIOD=0 GSR_WARPS=1 timeout 600 python bench.py --steps 1000 --warmup 100 --no-cpu-baseline 2>&1 | tail -1 | cut -c1-120 | sed "s/^/franka warps=1 noreorder /"
