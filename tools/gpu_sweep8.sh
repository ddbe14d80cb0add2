#!/bin/bash
for W in 6 7 8; do
  GSR_WARPS=$W timeout 600 python bench.py --workload go2 --steps 1000 --warmup 300 --no-cpu-baseline 2>&1 | tail -1 | cut -c1-120 | sed "s/^/go2 warps=$W /"
done
for R in 32 48 64; do
  GSR_SMEM_ROWS=$R GSR_WARPS=8 timeout 600 python bench.py --workload go2 --steps 1000 --warmup 300 --no-cpu-baseline 2>&1 | tail -1 | cut -c1-120 | sed "s/^/go2 rows=$R /"
done
GSR_SMEM_ROWS=48 GSR_WARPS=9 timeout 600 python bench.py --workload go2 --steps 1000 --warmup 300 --no-cpu-baseline 2>&1 | tail -1 | cut -c1-120 | sed "s/^/go2 rows=48 warps=9 /"
GSR_SMEM_ROWS=32 GSR_WARPS=10 timeout 600 python bench.py --workload go2 --steps 1000 --warmup 300 --no-cpu-baseline 2>&1 | tail -1 | cut -c1-120 | sed "s/^/go2 rows=32 warps=10 /"
