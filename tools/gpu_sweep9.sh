#!/bin/bash
for R in 24 32 40; do for W in 7 8; do
  GSR_SMEM_ROWS=$R GSR_WARPS=$W timeout 600 python bench.py --workload go2 --steps 1000 --warmup 300 --no-cpu-baseline 2>&1 | tail -1 | cut -c1-120 | sed "s/^/go2 rows=$R warps=$W /"
done; done
