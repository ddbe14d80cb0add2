#!/bin/bash
for G in 8 16; do for W in 1 2; do
  GSR_GROUP_SIZE=$G GSR_WARPS=$W timeout 600 python bench.py --steps 1000 --warmup 100 --no-cpu-baseline 2>&1 | tail -1 | cut -c1-120 | sed "s/^/franka G=$G warps=$W /"
done; done
