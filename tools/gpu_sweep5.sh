#!/bin/bash
# Franka 30000: resident environments per SM (on-chip row capacity x warps per CTA) against the wave quantisation of the launch
for R in 24 16 8; do for W in 1 2 4; do
  GSR_SMEM_ROWS=$R GSR_WARPS=$W timeout 300 python tools/sweep_one.py franka 30000 2>&1 | grep "##" | sed "s/^/rows=$R warps=$W /"
done; done
